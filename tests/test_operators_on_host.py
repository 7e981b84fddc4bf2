"""The per-cell operators of the CUDA engine (tnl_lbm_b200/csrc/collide*.cuh) compiled for the HOST with g++ -ffp-contract=off and run
against the CPU restatement on a periodic box: in parity arithmetic (LBMX_STRICT=1) every operator must be bit-identical -- the same
check the GPU tests make, available without a GPU.  tests/host_harness/host_strict_check.cpp is the harness."""
import os
import re
import subprocess
import tempfile

import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not O.available("port"), reason="oracle port not built")


def _build_and_run(strict: bool):
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "check")
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", f"-DLBMX_STRICT={1 if strict else 0}", os.path.join(ROOT, "tests", "host_harness", "host_strict_check.cpp"), "-o", exe,
               f"-L{ROOT}/oracle", "-loracle_port", f"-Wl,-rpath,{ROOT}/oracle"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    rows = re.findall(r"coll (\d+) eq (\d+) prec (\d+): (\d+) of (\d+) differ, max abs (\S+)", r.stdout)
    assert len(rows) >= 28 and "(D2Q9)" in r.stdout, r.stdout
    return r.returncode, rows


def test_parity_arithmetic_operators_are_bit_identical_on_the_host():
    rc, rows = _build_and_run(strict=True)
    bad = [row for row in rows if int(row[3]) != 0]
    assert rc == 0 and not bad, bad


def test_default_arithmetic_operators_agree_to_rounding_on_the_host():
    _, rows = _build_and_run(strict=False)
    for coll, eq, prec, ndiff, total, mx in rows:
        tol = 1e-6 if int(prec) == O.F32 else 1e-14  # one step from populations of order 1e-1: a few ulp
        assert float(mx) <= tol, (coll, eq, prec, mx)
