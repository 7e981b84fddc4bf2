"""The host-side decomposition + halo plan (lbmx_decompose_x / lbmx_halo_plan, no GPU needed) pinned on the CPU oracle:
a lattice cut into N ghosted x-slabs whose planes are exchanged according to the plan must reproduce the undivided run
bit for bit -- under A-B and under both parities of A-A."""
import numpy as np
import pytest

import golden_cases as gc
import lbm_cases as lc
from oracle import oracle as O
from slab_emulation import run_slabs_oracle

pytestmark = pytest.mark.skipif(not O.available("port"), reason="oracle port not built")


def box_case(streaming, nsteps, coll=O.CUM, eq=O.EQ_INV_CUM, prec=O.F64, X=12):
    d = O.Desc(coll=coll, eq=eq, streaming=streaming, precision=prec, X=X, Y=7, Z=6)
    return gc.Case("duct", d, O.Params(lbmViscosity=0.01, fx=2e-5, fy=-1e-5, fz=3e-5), lc.map_duct_slab_safe, nsteps, "noisy")


@pytest.mark.parametrize("nslabs", [2, 3, 4])
@pytest.mark.parametrize("streaming", [O.AB, O.AA])
def test_n_slabs_equal_one_domain(streaming, nslabs):
    """Periodic-x duct on which the 1-process wrap rule and the ghost-plane rule describe the same lattice
    (lbm_cases.map_duct_slab_safe).  A fully periodic box is NOT such a case for the reference: with nproc > 1 its periodic
    cells stop wrapping in y and z as well (kernels.h:24-28, "TODO: use nproc_y and nproc_z") and step out of the array; the
    engine wraps there, and tests/test_gpu_multi.py checks it against the undivided run instead."""
    case = box_case(streaming, 7)
    ref_df, ref_mac = gc.run_case(case, "port")
    df, mac = run_slabs_oracle(case, nslabs)
    assert np.array_equal(df, ref_df)
    assert np.array_equal(mac, ref_mac)


@pytest.mark.parametrize("streaming", [O.AB, O.AA])
def test_slab_count_invariance_with_walls(streaming):
    """Body-force duct whose walls touch the periodic x faces (sim_NSE/sim_2.cu:115-139): defined only under the ghost-plane
    rule; 1 slab with self-exchange == 2 == 4 slabs."""
    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=streaming, X=8, Y=8, Z=8)
    case = gc.Case("duct", d, O.Params(lbmViscosity=5e-3, fx=1e-5), lc.map_duct_periodic_x, 9, "noisy")
    runs = [run_slabs_oracle(case, n) for n in (1, 2, 4)]
    for df, mac in runs[1:]:
        assert np.array_equal(df, runs[0][0])
        assert np.array_equal(mac, runs[0][1])
    assert np.isfinite(runs[0][0]).all()


@pytest.mark.skipif(not (O.available("reference", O.AB) and O.available("reference", O.AA)), reason="oracle/_ref not built")
@pytest.mark.parametrize("streaming", [O.AB, O.AA])
def test_plan_against_the_reference_build(streaming):
    """Same check with the reference's own per-cell code doing the slab updates."""
    case = box_case(streaming, 4, X=8)
    ref_df, ref_mac = gc.run_case(case, "reference")
    df, mac = run_slabs_oracle(case, 2, kind="reference")
    assert np.array_equal(df, ref_df) and np.array_equal(mac, ref_mac)


def test_d2q9_slabs():
    d = O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AB, X=12, Y=9, Z=1)
    case = gc.Case("duct2d", d, O.Params(lbmViscosity=0.02, fx=1e-5, fy=2e-5), lc.map_duct_slab_safe, 6, "noisy")
    ref_df, ref_mac = gc.run_case(case, "port")
    df, mac = run_slabs_oracle(case, 3)
    assert np.array_equal(df, ref_df) and np.array_equal(mac, ref_mac)
