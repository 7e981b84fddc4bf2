"""k_bulk_tma (tnl_lbm_b200/csrc/kernels_tma.cuh): the A-A bulk kernels whose populations travel as TMA tensor boxes.

The arithmetic per cell is the plain kernel's, so every case must be BIT-IDENTICAL to the same engine with LBMX_TMA=0, and within
the north-star tolerance of the CPU oracle.  Lattice shapes are chosen to hit every tile geometry (one row per CTA, several rows per
CTA, a ragged last row group), the wrapped y / z / x faces, ghosted slabs, tiles whose non-bulk cells may be rewritten (GEO_WALL /
GEO_NOTHING ducts) and tiles that must fall back to per-lane stores (every other cell type)."""
import os

import numpy as np
import pytest

import golden_cases as gc
import lbm_cases as lc
from engine_runner import run_case_engine
from oracle import oracle as O
from tnl_lbm_b200 import binding as B

pytestmark = pytest.mark.gpu

TOL = {O.F64: 1e-12, O.F32: 1e-5}


def _zoo(d):
    return lc.map_random_aa(d, seed=21, frac_special=0.15)


def _duct(d):
    return gc._duct_all_periodic_faces(d)


P_BOX = O.Params(lbmViscosity=1e-3, fx=1e-6, fy=2e-6, fz=-1e-6)
P_ZOO = gc._p3()

CASES = [
    # name, desc, params, map, steps, init
    gc.Case("tma_cum_f64_box_rows4", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=5, Y=32, Z=10), P_BOX, lc.map_periodic, 7, "smooth"),
    gc.Case("tma_cum_f64_box_rows8", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=4, Y=48, Z=9), P_BOX, lc.map_periodic, 6, "noisy"),
    gc.Case("tma_cum_f64_box_row1", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=3, Y=256, Z=3), P_BOX, lc.map_periodic, 6, "noisy"),
    gc.Case("tma_cum_f64_duct", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=6, Y=32, Z=12), O.Params(lbmViscosity=5e-3, fx=1e-5), _duct, 9, "uniform"),
    gc.Case("tma_cum_f64_zoo", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=7, Y=32, Z=9), P_ZOO, _zoo, 5, "noisy"),
    gc.Case("tma_srt_f32_zoo", O.Desc(coll=O.SRT, eq=O.EQ_STD, streaming=O.AA, precision=O.F32, X=6, Y=64, Z=5), P_ZOO, _zoo, 5, "noisy"),
    gc.Case("tma_bgk_f32_box", O.Desc(coll=O.BGK, eq=O.EQ_STD, streaming=O.AA, precision=O.F32, X=4, Y=128, Z=4), P_BOX, lc.map_periodic, 6, "smooth"),
    gc.Case("tma_mrt_f64_duct", O.Desc(coll=O.MRT_LES, eq=O.EQ_STD, streaming=O.AA, X=5, Y=16, Z=11), O.Params(lbmViscosity=5e-3), _duct, 6, "noisy"),
    gc.Case("tma_kbcn4_f64_box", O.Desc(coll=O.KBC_N4, eq=O.EQ_ENTROPIC, streaming=O.AA, X=4, Y=32, Z=6), P_BOX, lc.map_periodic, 5, "smooth"),
    gc.Case("tma_d2q9_srt_f64_zoo", O.Desc(lattice=O.D2Q9, coll=O.SRT, eq=O.EQ_STD, streaming=O.AA, X=19, Y=64, Z=1), gc._p2(), _zoo, 6, "noisy"),
    gc.Case("tma_d2q9_clbm_f32_box", O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AA, precision=O.F32, X=9, Y=96, Z=1), gc._p2(), lc.map_periodic, 8, "smooth"),
]


def _tile_kernels_expected(case):
    """The tile kernels know fluid cells only: maps whose obstacle (GEO_WALL away from the faces) or GEO_NOTHING cells belong to the
    bulk kernel (kernels.cuh: cell_in_boundary_list) run on the plain kernels whatever LBMX_TMA says."""
    m = case.make_map(case.desc)
    return bool(np.all((m != 1) & (m != (7 if case.desc.lattice == O.D2Q9 else 8))))


def _run(case, mode, **kw):
    old = os.environ.get("LBMX_TMA")
    os.environ["LBMX_TMA"] = mode
    try:
        return run_case_engine(case, **kw)
    finally:
        if old is None:
            del os.environ["LBMX_TMA"]
        else:
            os.environ["LBMX_TMA"] = old


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
@pytest.mark.parametrize("mode", ["both", "odd", "even"])
def test_tma_kernels_are_bit_identical_to_the_plain_kernels(case, mode):
    df_t, mac_t, st_t = _run(case, mode)
    df_p, mac_p, st_p = _run(case, "0")
    if _tile_kernels_expected(case):
        assert st_t.tma_launches > 0, "the TMA kernels did not run (tile geometry?)"
    assert st_p.tma_launches == 0
    assert np.array_equal(df_t, df_p), f"{case.name}: {int((df_t != df_p).sum())} populations differ from the plain kernels"
    assert np.array_equal(mac_t, mac_p)


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_tma_kernels_match_the_oracle(case):
    df, mac, st = _run(case, "both")
    assert st.tma_launches > 0 or not _tile_kernels_expected(case)
    df_ref, mac_ref = gc.run_case(case, "port", nthreads=4)
    err = lc.rel_err_df(df, df_ref, case.desc)
    assert err <= TOL[case.desc.precision], f"{case.name}: rel err {err:.3e}"
    for lo, hi, label in lc.macro_groups(case.desc):
        # the ducts start from rest (|u| ~ 1e-4 after a few steps): velocities are compared on a scale of at least 1e-3 lattice units
        assert lc.rel_err(mac[lo:hi], mac_ref[lo:hi], floor=1e-3 if label == "velocity" else 0.0) <= TOL[case.desc.precision], label


@pytest.mark.parametrize("policy", [B.MACRO_EVERY_STEP, B.MACRO_LAST_STEP])
def test_tma_kernels_in_step_batches_and_with_every_step_macros(policy):
    case = CASES[0]
    a = _run(case, "both", macro_policy=policy)
    b = _run(case, "both", chunk=1, macro_policy=policy)
    c = _run(case, "0", chunk=2, macro_policy=policy)
    for x, y in ((a, b), (a, c)):
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1])


def test_tma_kernels_on_a_ghosted_slab_with_self_exchange():
    """One slab with ghost x-planes and periodic x (the configuration of the channel bench and of every rank of a multi-GPU run)."""
    case = gc.Case("tma_cum_f64_ghosted", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=6, Y=32, Z=8), P_BOX, lc.map_periodic, 8, "smooth")
    plain = _run(case, "0")
    ghosted = []
    for mode in ("0", "both"):
        old = os.environ.get("LBMX_TMA")
        os.environ["LBMX_TMA"] = mode
        try:
            d = case.desc
            port = O.Oracle(d, "port")
            df0 = gc.initial_df(case, port)
            with B.Engine(lattice=d.lattice, coll=d.coll, eq=d.eq, streaming=d.streaming, macro=d.macro, inflow=d.inflow, precision=d.precision, X=d.X, Y=d.Y, Z=d.Z,
                          ghost_x=1, periodic_x=1, macro_policy=B.MACRO_LAST_STEP) as e:
                e.map_upload(case.make_map(d))
                e.df_upload(df0, which=0)
                e.df_sync_ghosts()
                e.set_params(lbmViscosity=P_BOX.lbmViscosity, fx=P_BOX.fx, fy=P_BOX.fy, fz=P_BOX.fz)
                e.macro_init()
                e.step(case.nsteps)
                e.sync()
                ghosted.append((e.df_download(0), e.macro_download(), e.stats()))
        finally:
            if old is None:
                del os.environ["LBMX_TMA"]
            else:
                os.environ["LBMX_TMA"] = old
    assert ghosted[1][2].tma_launches > 0 and ghosted[0][2].tma_launches == 0
    for got in ghosted:
        assert np.array_equal(got[0], plain[0]) and np.array_equal(got[1], plain[1])
