"""Build the C++ clients of the host mirror (tnl_lbm_b200/host) against liblbmx.so into examples/bin/ (git-ignored; the
binaries travel with the working tree to the GPU box):

  * examples/abi_minimal.c                     -- the C ABI from plain C99
  * examples/channel3d.cpp                     -- this repository's own solver in the reference's style (A-B and A-A builds)
  * examples/box3d.cpp                         -- the bench workload (periodic box, body force) as an ordinary solver: the drop-in rate
  * /root/reference/sim_NSE/sim_1.cu,          -- the reference's UNMODIFIED solver sources, when the reference tree is present;
    /root/reference/sim_NSE/sim_2.cu, sim_3.cu,   their third-party includes (argparse, fmt, spdlog, magic_enum) are satisfied by the
    /root/reference/sim_2D/sim2d_1.cu, sim2d_3.cu stand-ins under tests/solver_shims (the reference fetches the real ones with CMake)
"""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "bin")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
BASE = [CXX, "-std=c++17", "-O2", f"-I{ROOT}/tnl_lbm_b200/host", f"-I{ROOT}/include"]
LINK = [f"-L{ROOT}/tnl_lbm_b200", "-llbmx", "-Wl,-rpath,$ORIGIN/../../tnl_lbm_b200"]


def build(reference: str = "/root/reference") -> list[str]:
    os.makedirs(BIN, exist_ok=True)
    built = []
    jobs = [("channel3d", [os.path.join(ROOT, "examples", "channel3d.cpp")], []),
            ("channel3d_aa", [os.path.join(ROOT, "examples", "channel3d.cpp")], ["-DAA_PATTERN"]),
            ("box3d", [os.path.join(ROOT, "examples", "box3d.cpp")], []),
            ("box3d_aa", [os.path.join(ROOT, "examples", "box3d.cpp")], ["-DAA_PATTERN"])]
    shims = [f"-I{ROOT}/tests/solver_shims"]
    for name, rel in (("ref_sim_1", "sim_NSE/sim_1.cu"), ("ref_sim_2", "sim_NSE/sim_2.cu"), ("ref_sim_3", "sim_NSE/sim_3.cu"), ("ref_sim2d_1", "sim_2D/sim2d_1.cu"), ("ref_sim2d_2", "sim_2D/sim2d_2.cu"), ("ref_sim2d_3", "sim_2D/sim2d_3.cu")):
        src = os.path.join(reference, rel)
        if os.path.exists(src):
            for pat in ("AB", "AA"):
                jobs.append((f"{name}_{pat.lower()}", ["-x", "c++", src], shims + [f"-D{pat}_PATTERN"]))
    # plain C (C99, -pedantic): the header is a C header, and this is the call sequence a binding in any language reproduces
    c_out = os.path.join(BIN, "abi_minimal")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT}/include", os.path.join(ROOT, "examples", "abi_minimal.c"), "-o", c_out] + LINK + ["-lm"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building abi_minimal failed:\n{r.stderr}")
    built.append(c_out)
    for name, src, extra in jobs:
        out = os.path.join(BIN, name)
        cmd = BASE + extra + src + ["-o", out] + LINK
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building {name} failed:\n{' '.join(cmd)}\n{r.stderr}")
        built.append(out)
    return built


if __name__ == "__main__":
    for b in build():
        print(b)
