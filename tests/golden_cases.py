"""Registry of the golden cases: small, seeded runs whose outputs were produced by the REFERENCE's own code
(oracle/_ref, see tests/golden/make_golden.py) and committed under tests/golden/*.npz."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np

import lbm_cases as lc
from oracle import oracle as O


@dataclass
class Case:
    name: str
    desc: O.Desc
    params: O.Params
    make_map: Callable
    nsteps: int
    init: str = "noisy"  # "noisy" | "smooth" | "uniform"
    seed: int = 11


def _p3(**kw):
    base = dict(lbmViscosity=0.013, fx=3e-5, fy=-2e-5, fz=1e-5, inflow_vx=0.04, inflow_vy=0.01, inflow_vz=-0.02)
    base.update(kw)
    return O.Params(**base)


def _p2(**kw):
    base = dict(lbmViscosity=0.02, fx=2e-5, fy=-1e-5, inflow_vx=0.05, inflow_vy=-0.01)
    base.update(kw)
    return O.Params(**base)


def zoo(d):
    return lc.map_random_ab(d) if d.streaming == O.AB else lc.map_random_aa(d)


CASES = [
    Case("cum_f64_ab_zoo", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cum_f64_aa_zoo", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cum_f32_ab_zoo", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cumstd_f64_aa_zoo", O.Desc(coll=O.CUM, eq=O.EQ_STD, streaming=O.AA, X=8, Y=8, Z=6), _p3(), zoo, 3),
    Case("cum_f64_aa_box", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=16, Y=16, Z=16), O.Params(lbmViscosity=1e-3, fx=1e-6), lc.map_periodic, 50, "smooth"),
    Case("cum_f64_ab_box", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, X=16, Y=16, Z=16), O.Params(lbmViscosity=1e-3, fx=1e-6), lc.map_periodic, 50, "smooth"),
    Case("cum_f64_ab_sim1", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, X=40, Y=16, Z=16), O.Params(lbmViscosity=1e-3, inflow_vx=0.05), lc.map_sim1_channel, 30, "uniform"),
    Case("cum_f64_ab_sim1_momentbc", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, X=32, Y=14, Z=14), O.Params(lbmViscosity=2e-3, inflow_vx=0.04),
         lambda d: lc.map_sim1_channel(d, "INFLOW_LEFT"), 30, "uniform"),
    Case("cum_f64_aa_duct", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=12, Y=12, Z=12), O.Params(lbmViscosity=5e-3, fx=1e-5),
         lambda d: _duct_all_periodic_faces(d), 40, "uniform"),
    Case("srt_f64_ab_zoo", O.Desc(coll=O.SRT, eq=O.EQ_STD, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("srtinv_f32_aa_zoo", O.Desc(coll=O.SRT, eq=O.EQ_INV_CUM, streaming=O.AA, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("bgk_f64_aa_zoo", O.Desc(coll=O.BGK, eq=O.EQ_STD, streaming=O.AA, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("bgk_f32_ab_zoo", O.Desc(coll=O.BGK, eq=O.EQ_STD, streaming=O.AB, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    # D3Q27_BGK built with USE_GALILEAN_CORRECTION (defs.h:253, col_bgk.h:20-45)
    Case("bgkgal_f64_ab_zoo", O.Desc(coll=O.BGK_GALILEAN, eq=O.EQ_STD, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("bgkgal_f32_aa_zoo", O.Desc(coll=O.BGK_GALILEAN, eq=O.EQ_STD, streaming=O.AA, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("bgkgal_f64_aa_box", O.Desc(coll=O.BGK_GALILEAN, eq=O.EQ_STD, streaming=O.AA, X=12, Y=12, Z=12), O.Params(lbmViscosity=1e-3, fx=1e-6), lc.map_periodic, 40, "smooth"),
    # D3Q27_CUM built with USE_HIGH_PRECISION_RHO (defs.h:252, d3q27/common.h:19-29: Kahan-summed density)
    Case("cumhp_f64_ab_zoo", O.Desc(coll=O.CUM_HP_RHO, eq=O.EQ_INV_CUM, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cumhp_f32_aa_zoo", O.Desc(coll=O.CUM_HP_RHO, eq=O.EQ_INV_CUM, streaming=O.AA, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cumhp_f32_aa_box", O.Desc(coll=O.CUM_HP_RHO, eq=O.EQ_INV_CUM, streaming=O.AA, precision=O.F32, X=12, Y=12, Z=12), O.Params(lbmViscosity=1e-3, fx=1e-6), lc.map_periodic, 40, "smooth"),
    Case("mrt_f64_ab_zoo", O.Desc(coll=O.MRT_LES, eq=O.EQ_STD, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("mrt_f32_aa_zoo", O.Desc(coll=O.MRT_LES, eq=O.EQ_STD, streaming=O.AA, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("d2q9_srt_f64_ab_cavity", O.Desc(lattice=O.D2Q9, coll=O.SRT, eq=O.EQ_STD, streaming=O.AB, X=24, Y=24, Z=1), O.Params(lbmViscosity=0.05, inflow_vx=0.1), lc.map_cavity_2d, 40, "uniform"),
    Case("d2q9_clbm_f64_ab_channel", O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AB, X=40, Y=16, Z=1), O.Params(lbmViscosity=0.01, inflow_vx=0.05), lc.map_sim2d1_channel, 40, "uniform"),
    Case("d2q9_srt_f32_aa_zoo", O.Desc(lattice=O.D2Q9, coll=O.SRT, eq=O.EQ_STD, streaming=O.AA, precision=O.F32, X=13, Y=11, Z=1), _p2(), zoo, 4),
    Case("d2q9_clbm_f64_aa_zoo", O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AA, X=13, Y=11, Z=1), _p2(), zoo, 4),
    Case("d2q9_clbm_f32_ab_zoo", O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AB, precision=O.F32, X=13, Y=11, Z=1), _p2(), zoo, 4),
    Case("cum_f64_ab_mean", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, macro=O.MACRO_MEAN, X=8, Y=7, Z=6), _p3(), lambda d: lc.map_random_ab(d, seed=3), 5),
    Case("d2q9_srt_f64_ab_bouzidi", O.Desc(lattice=O.D2Q9, coll=O.SRT, eq=O.EQ_STD, streaming=O.AB, X=13, Y=11, Z=1),
         _p2(bouzidi=lc.map_and_coeffs_bouzidi(O.Desc(lattice=O.D2Q9, coll=O.SRT, eq=O.EQ_STD, streaming=O.AB, X=13, Y=11, Z=1))[1]),
         lambda d: lc.map_and_coeffs_bouzidi(d)[0], 4, seed=5),
    Case("cum_f64_ab_profile", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, inflow=O.INFLOW_PROFILE_YZ, X=10, Y=9, Z=8),
         O.Params(lbmViscosity=0.004, fx=1e-5, vx_profile=(0.05 * np.random.RandomState(5).random_sample((8, 9)))), lambda d: _profile_map(d), 6),
    Case("clbm3d_f64_ab_zoo", O.Desc(coll=O.CLBM, eq=O.EQ_STD, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("clbm3d_f32_aa_zoo", O.Desc(coll=O.CLBM, eq=O.EQ_STD, streaming=O.AA, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("clbm3d_f64_aa_box", O.Desc(coll=O.CLBM, eq=O.EQ_STD, streaming=O.AA, X=12, Y=12, Z=12), O.Params(lbmViscosity=1e-3, fx=1e-6), lc.map_periodic, 40, "smooth"),
    Case("srtmf_f64_aa_zoo", O.Desc(coll=O.SRT_MODIF_FORCE, eq=O.EQ_STD, streaming=O.AA, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("srtmf_f32_ab_zoo", O.Desc(coll=O.SRT_MODIF_FORCE, eq=O.EQ_INV_CUM, streaming=O.AB, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cum2017_f64_ab_zoo", O.Desc(coll=O.CUM_2017, eq=O.EQ_INV_CUM, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cum2017_f32_aa_zoo", O.Desc(coll=O.CUM_2017, eq=O.EQ_INV_CUM, streaming=O.AA, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cumaa_f64_aa_zoo", O.Desc(coll=O.CUM_ANTIALIAS, eq=O.EQ_INV_CUM, streaming=O.AA, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cum2017aa_f64_ab_zoo", O.Desc(coll=O.CUM_2017_ANTIALIAS, eq=O.EQ_INV_CUM, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cum2017aa_f32_ab_zoo", O.Desc(coll=O.CUM_2017_ANTIALIAS, eq=O.EQ_INV_CUM, streaming=O.AB, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("cum2017aa_f64_aa_box", O.Desc(coll=O.CUM_2017_ANTIALIAS, eq=O.EQ_INV_CUM, streaming=O.AA, X=12, Y=12, Z=12), O.Params(lbmViscosity=1e-3, fx=1e-6), lc.map_periodic, 40, "smooth"),
    Case("kbcn1_f64_ab_zoo", O.Desc(coll=O.KBC_N1, eq=O.EQ_STD, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("kbcn2_f32_aa_zoo", O.Desc(coll=O.KBC_N2, eq=O.EQ_ENTROPIC, streaming=O.AA, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("kbcn3_f64_aa_zoo", O.Desc(coll=O.KBC_N3, eq=O.EQ_ENTROPIC, streaming=O.AA, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("kbcn4_f32_ab_zoo", O.Desc(coll=O.KBC_N4, eq=O.EQ_ENTROPIC, streaming=O.AB, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("kbcc1_f32_ab_zoo", O.Desc(coll=O.KBC_C1, eq=O.EQ_STD, streaming=O.AB, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("kbcc2_f64_ab_zoo", O.Desc(coll=O.KBC_C2, eq=O.EQ_ENTROPIC, streaming=O.AB, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("kbcc3_f32_aa_zoo", O.Desc(coll=O.KBC_C3, eq=O.EQ_ENTROPIC, streaming=O.AA, precision=O.F32, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("kbcc4_f64_aa_zoo", O.Desc(coll=O.KBC_C4, eq=O.EQ_ENTROPIC, streaming=O.AA, X=9, Y=8, Z=7), _p3(), zoo, 4),
    Case("kbcn4_f64_aa_box", O.Desc(coll=O.KBC_N4, eq=O.EQ_ENTROPIC, streaming=O.AA, X=12, Y=12, Z=12), O.Params(lbmViscosity=1e-3, fx=1e-6), lc.map_periodic, 40, "smooth"),
    Case("d2q9_clbm_f64_ab_parabolic", O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AB, inflow=O.INFLOW_PARABOLIC_Y, X=40, Y=16, Z=1),
         O.Params(lbmViscosity=0.01, inflow_vx=0.05, inflow_vy=1.0, inflow_vz=1.0 / 13), lc.map_sim2d1_channel, 40, "uniform"),
    Case("d2q9_srt_f32_ab_parabolic_zoo", O.Desc(lattice=O.D2Q9, coll=O.SRT, eq=O.EQ_STD, streaming=O.AB, inflow=O.INFLOW_PARABOLIC_Y, precision=O.F32, X=13, Y=11, Z=1),
         _p2(inflow_vx=0.07, inflow_vy=1.0, inflow_vz=0.125), zoo, 4),
    # sim_2D/sim2d_2.cu's macro class: velocity sums for the first half, host freeze, fluctuation sums for the second (run_case below)
    Case("d2q9_clbm_f64_ab_withmean", O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AB, macro=O.MACRO_WITH_MEAN_2D, inflow=O.INFLOW_PARABOLIC_Y, X=40, Y=16, Z=1),
         O.Params(lbmViscosity=0.01, inflow_vx=0.05, inflow_vy=1.0, inflow_vz=1.0 / 13), lc.map_sim2d1_channel, 40, "uniform"),
    Case("d2q9_srt_f32_aa_withmean_zoo", O.Desc(lattice=O.D2Q9, coll=O.SRT, eq=O.EQ_STD, streaming=O.AA, macro=O.MACRO_WITH_MEAN_2D, inflow=O.INFLOW_PARABOLIC_Y, precision=O.F32, X=13, Y=11, Z=1),
         _p2(inflow_vx=0.07, inflow_vy=1.0, inflow_vz=0.125), zoo, 6),
    Case("cum_f64_ab_void", O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, macro=O.MACRO_VOID, X=8, Y=7, Z=6), _p3(), lambda d: lc.map_random_ab(d, seed=3), 3),
]


def _profile_map(d):
    m = d.new_map(lc.G3["FLUID"])
    m[0] = lc.G3["INFLOW_LEFT"]
    m[1, 2:5, 2:6] = lc.G3["INFLOW"]
    m[d.X - 1] = lc.G3["OUTFLOW_RIGHT"]
    return m


def _duct_all_periodic_faces(d):
    """A-A-safe duct: periodic x planes, WALL ring at y,z = 1 / N-2 behind a NOTHING shell, but the
    x-face cells of the ring/shell are GEO_NOTHING too so that no non-periodic cell touches the unghosted x faces."""
    g = lc.G3
    m = lc.map_duct_periodic_x(d)
    for xf in (0, d.X - 1):
        plane = m[xf]
        plane[plane != g["PERIODIC"]] = g["NOTHING"]
    return m


BY_NAME = {c.name: c for c in CASES}


def initial_df(case: Case, orc) -> np.ndarray:
    d = case.desc
    if case.init == "noisy":
        return lc.noisy_df(d, orc, seed=case.seed)
    df = d.new_df()
    if case.init == "smooth":
        orc.set_equilibrium_field(df, *lc.smooth_fields(d))
    else:
        orc.set_equilibrium(df, 1.0, 0.0, 0.0, 0.0)
    return df


def freeze_means(mac: np.ndarray, samples: int) -> None:
    """What the solver does on the host between the two phases (sim_2D/sim2d_2.cu:471-505): frozen mean = sum / samples on the interior
    cells, fluctuation sums cleared; the array then goes back to the device."""
    inner = (slice(1, -1), slice(None), slice(1, -1))
    for sum_ch, mean_ch in ((3, 5), (4, 6)):
        mac[mean_ch][inner] = mac[sum_ch][inner] / mac.dtype.type(samples)
    for ch in (7, 8, 9):
        mac[ch][inner] = 0


def run_case(case: Case, kind: str, nthreads: int = 1, fast: bool = False, init_kind: str | None = None):
    """Run a case on a CPU checker; returns (df holding the current state, macro).  `init_kind`: take the initial state from that
    checker instead (the way the engine tests upload the port's initial arrays)."""
    d, p = case.desc, case.params
    orc = O.Oracle(d, kind, fast=fast)
    a = initial_df(case, O.Oracle(d, init_kind) if init_kind else orc)
    b = a.copy()
    mac = d.new_macro()
    m = case.make_map(d)
    p.stat_counter = 0
    if d.macro != O.MACRO_VOID:
        orc.initial_macro(p, a, mac)
    if d.macro == O.MACRO_MEAN:
        for it in range(case.nsteps):
            p.stat_counter = it
            orc.step(p, a, b, mac, m, it, 1, nthreads)
    elif d.macro == O.MACRO_WITH_MEAN_2D:
        half = case.nsteps // 2
        p.macro_gates = O.GATE_MEANS
        orc.step(p, a, b, mac, m, 0, half, nthreads)
        freeze_means(mac, half)
        p.macro_gates = O.GATE_FLUCS
        orc.step(p, a, b, mac, m, half, case.nsteps - half, nthreads)
        p.macro_gates = 0
    else:
        orc.step(p, a, b, mac, m, 0, case.nsteps, nthreads)
    cur = a if (d.streaming == O.AA or case.nsteps % 2 == 0) else b
    return cur, mac


def sample_stride(case: Case, max_cells: int = 1500) -> int:
    n = case.desc.XYZ
    s = max(1, -(-n // max_cells))
    while s > 1 and n % s == 0:  # avoid strides that alias with the row length
        s += 1
    return s


def sample(a: np.ndarray, stride: int) -> np.ndarray:
    return np.ascontiguousarray(a.reshape(a.shape[0], -1)[:, ::stride])
