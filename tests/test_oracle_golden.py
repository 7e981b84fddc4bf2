"""The CPU restatement against the committed golden vectors (produced by the reference's own code, see
tests/golden/make_golden.py).  Bit for bit: SHA-256 of the full arrays, plus the stored samples."""
import hashlib
import json
import os

import numpy as np
import pytest

import golden_cases as gc
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))

pytestmark = pytest.mark.skipif(not O.available("port"), reason="oracle port not built (python -c 'import __graft_entry__ as g; g.build()')")


def test_manifest_covers_registry():
    assert sorted(MANIFEST) == sorted(c.name for c in gc.CASES)
    for c in gc.CASES:
        assert os.path.exists(os.path.join(GOLD, c.name + ".npz")), c.name


@pytest.mark.parametrize("name", [c.name for c in gc.CASES])
def test_port_reproduces_golden(name):
    case = gc.BY_NAME[name]
    df, mac = gc.run_case(case, "port")
    z = np.load(os.path.join(GOLD, name + ".npz"))
    stride = int(z["stride"])
    assert np.array_equal(gc.sample(df, stride), z["df_sample"])
    assert np.array_equal(gc.sample(mac, stride), z["macro_sample"])
    assert hashlib.sha256(df.tobytes()).hexdigest() == MANIFEST[name]["df_sha256"]
    assert hashlib.sha256(mac.tobytes()).hexdigest() == MANIFEST[name]["macro_sha256"]


def test_port_is_thread_count_invariant():
    case = gc.BY_NAME["cum_f64_aa_box"]
    a, _ = gc.run_case(case, "port", nthreads=1)
    b, _ = gc.run_case(case, "port", nthreads=4)
    assert np.array_equal(a, b)


def test_mass_and_momentum_budget():
    """Periodic box with a body force: mass conserved, momentum grows by F per cell per step (col_cum.h:341-345 forcing convention)."""
    case = gc.BY_NAME["cum_f64_ab_box"]
    d = case.desc
    orc = O.Oracle(d, "port")
    df0 = gc.initial_df(case, orc)
    df, _ = gc.run_case(case, "port")
    n = d.X * d.Y * d.Z
    assert abs(df.sum() - df0.sum()) / n < 1e-13
    import lbm_cases as lc
    jx0 = float((df0 * lc.C27[:, 0][:, None, None, None]).sum())
    jx = float((df * lc.C27[:, 0][:, None, None, None]).sum())
    assert abs((jx - jx0) / n - case.nsteps * case.params.fx) < 1e-12
