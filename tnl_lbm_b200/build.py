"""Build liblbmx.so (hand-written CUDA for sm_100a + the C ABI of include/lbmx.h) in-tree with nvcc.

    python -m tnl_lbm_b200.build [--force] [--jobs N]

One object per (lattice, operator, precision) kernel family so that the heavy unrolled kernels compile in parallel.
The result, tnl_lbm_b200/liblbmx.so, is git-ignored but travels with the working tree to the GPU box.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "liblbmx.so")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-ccbin", HOST_CXX, "-I/usr/include"]

FAMILIES = [
    ("d3q27_cum", "D3Q27", "K_CUM"),
    ("d3q27_srt", "D3Q27", "K_SRT"),
    ("d3q27_bgk", "D3Q27", "K_BGK"),
    ("d3q27_bgkgal", "D3Q27", "K_BGK_GAL"),
    ("d3q27_cumhp", "D3Q27", "K_CUM_HP_RHO"),
    ("d3q27_mrt", "D3Q27", "K_MRT"),
    ("d3q27_clbm", "D3Q27", "K_CLBM"),
    ("d3q27_srtmf", "D3Q27", "K_SRT_MF"),
    ("d3q27_cum2017", "D3Q27", "K_CUM_2017"),
    ("d3q27_cumaa", "D3Q27", "K_CUM_AALIAS"),
    ("d3q27_cum2017aa", "D3Q27", "K_CUM_2017_AALIAS"),
    ("d3q27_kbcn1", "D3Q27", "K_KBC_N1"),
    ("d3q27_kbcn2", "D3Q27", "K_KBC_N2"),
    ("d3q27_kbcn3", "D3Q27", "K_KBC_N3"),
    ("d3q27_kbcn4", "D3Q27", "K_KBC_N4"),
    ("d3q27_kbcc1", "D3Q27", "K_KBC_C1"),
    ("d3q27_kbcc2", "D3Q27", "K_KBC_C2"),
    ("d3q27_kbcc3", "D3Q27", "K_KBC_C3"),
    ("d3q27_kbcc4", "D3Q27", "K_KBC_C4"),
    ("d3q19_srt", "D3Q19", "K_SRT"),
    ("d3q19_mrt", "D3Q19", "K_MRT"),
    ("d2q9_srt", "D2Q9", "K_SRT"),
    ("d2q9_clbm", "D2Q9", "K_CLBM"),
]


def _digest(names_filter) -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")) and names_filter(name):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(" ".join(COMMON + ARCH).encode())
    return h.hexdigest()


def _sources_digest() -> str:
    return _digest(lambda name: True)


def _kernel_digest() -> str:
    """What the kernel-family objects (inst.cu compiled once per family) depend on: everything but engine.cu."""
    return _digest(lambda name: name != "engine.cu")


def _run(cmd, log):
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"build step failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return r.stdout + r.stderr


def build(force: bool = False, jobs: int | None = None, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _sources_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found: liblbmx.so cannot be built (there is no CPU fallback)")
    tasks = []
    for fam, lat, kind in FAMILIES:
        for real in ("float", "double"):
            obj = os.path.join(OBJ, f"inst_{fam}_{real}.o")
            cmd = [NVCC, *ARCH, *COMMON, "-Xptxas", "-v", f"-DLBMX_FAMILY={fam}", f"-DLBMX_LAT={lat}", f"-DLBMX_KIND={kind}", f"-DLBMX_REAL={real}",
                   "-c", os.path.join(CSRC, "inst.cu"), "-o", obj]
            tasks.append((cmd, obj))
    # every reference family once more in parity arithmetic: reference association (collide_strict.cuh) and no FMA contraction
    for fam, lat, kind in FAMILIES:
        if lat == "D3Q19":
            continue
        for real in ("float", "double"):
            obj = os.path.join(OBJ, f"inst_{fam}_{real}_strict.o")
            cmd = [NVCC, *ARCH, *COMMON, "-fmad=false", "-DLBMX_STRICT=1", "-Xptxas", "-v", f"-DLBMX_FAMILY={fam}_strict", f"-DLBMX_LAT={lat}",
                   f"-DLBMX_KIND={kind}", f"-DLBMX_REAL={real}", "-c", os.path.join(CSRC, "inst.cu"), "-o", obj]
            tasks.append((cmd, obj))
    eng = os.path.join(OBJ, "engine.o")
    tasks.append(([NVCC, *ARCH, *COMMON, "-c", os.path.join(CSRC, "engine.cu"), "-o", eng], eng))
    # kernel-family objects are reused when only engine.cu (host code) changed
    kstamp = os.path.join(OBJ, "kernel_digest.txt")
    kdigest = _kernel_digest()
    reuse = not force and os.path.exists(kstamp) and open(kstamp).read() == kdigest
    all_objs = [obj for _, obj in tasks]
    if reuse:
        tasks = [(cmd, obj) for cmd, obj in tasks if obj == eng or not os.path.exists(obj)]
    jobs = jobs or min(len(tasks), os.cpu_count() or 4)
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        futs = [ex.submit(_run, cmd, obj + ".log") for cmd, obj in tasks]
        for f in futs:
            out = f.result()
            if verbose:
                print(out)
    _run([NVCC, *ARCH, "-shared", "-ccbin", HOST_CXX, "-o", LIB, *all_objs, "-cudart", "static", "-ldl"], os.path.join(OBJ, "link.log"))
    with open(stamp, "w") as f:
        f.write(digest)
    with open(kstamp, "w") as f:
        f.write(kdigest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.jobs, a.verbose))
