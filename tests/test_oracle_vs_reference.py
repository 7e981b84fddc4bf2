"""Pin of the CPU restatement (oracle/lbm_oracle.cpp) against the reference's own per-cell code.

The reference build (oracle/_ref/libref_*.so) exists wherever oracle/Makefile could see /root/reference
(this container; the built .so travels to the GPU box).  Strict IEEE builds on both sides, so the
comparison is bit for bit: distributions, macroscopic fields, for every operator / equilibrium /
precision / streaming pattern, over random maps that exercise every cell type (d3q27/bc.h:51-241,
d2q9/bc.h:89-196).
"""
import numpy as np
import pytest

import lbm_cases as lc
from oracle import oracle as O

pytestmark = pytest.mark.skipif(
    not (O.available("reference", O.AB) and O.available("reference", O.AA) and O.available("port")),
    reason="oracle/_ref (reference build) or the port library is not built",
)

COMBOS_3D = [(O.CUM, O.EQ_INV_CUM), (O.CUM, O.EQ_STD), (O.CUM_HP_RHO, O.EQ_INV_CUM), (O.CUM_HP_RHO, O.EQ_STD), (O.SRT, O.EQ_STD), (O.SRT, O.EQ_INV_CUM), (O.BGK, O.EQ_STD), (O.BGK_GALILEAN, O.EQ_STD), (O.MRT_LES, O.EQ_STD), (O.MRT_LES, O.EQ_INV_CUM),
             (O.CLBM, O.EQ_STD), (O.CLBM, O.EQ_INV_CUM), (O.SRT_MODIF_FORCE, O.EQ_STD), (O.SRT_MODIF_FORCE, O.EQ_INV_CUM),
             (O.CUM_2017, O.EQ_INV_CUM), (O.CUM_ANTIALIAS, O.EQ_INV_CUM), (O.CUM_2017_ANTIALIAS, O.EQ_INV_CUM), (O.CUM_2017_ANTIALIAS, O.EQ_STD),
             (O.KBC_N1, O.EQ_STD), (O.KBC_N2, O.EQ_ENTROPIC), (O.KBC_N3, O.EQ_ENTROPIC), (O.KBC_N4, O.EQ_ENTROPIC), (O.KBC_N4, O.EQ_INV_CUM),
             (O.KBC_C1, O.EQ_STD), (O.KBC_C2, O.EQ_ENTROPIC), (O.KBC_C3, O.EQ_ENTROPIC), (O.KBC_C4, O.EQ_ENTROPIC), (O.KBC_C1, O.EQ_ENTROPIC)]


def run_pair(d: O.Desc, m: np.ndarray, p: O.Params, nsteps: int, seed=11, noise=0.05):
    out = []
    for kind in ("reference", "port"):
        orc = O.Oracle(d, kind)
        a = lc.noisy_df(d, orc, seed=seed, noise=noise)
        b = a.copy()
        mac = d.new_macro()
        p.stat_counter = 0
        if d.macro != O.MACRO_VOID:
            orc.initial_macro(p, a, mac)
        for it in range(nsteps):
            p.stat_counter = it
            orc.step(p, a, b, mac, m, it, 1, 1)
        out.append((a, b, mac))
    return out


def assert_same(ref, port, what):
    for name, r, q in zip(("df_a", "df_b", "macro"), ref, port):
        assert np.isfinite(r).all(), f"{what}: reference {name} not finite"
        if not np.array_equal(r, q):
            diff = np.abs(r.astype(np.float64) - q.astype(np.float64))
            i = np.unravel_index(np.argmax(diff), diff.shape)
            raise AssertionError(f"{what}: {name} differs, max |d|={diff.max():.3e} at {i}: ref={r[i]!r} port={q[i]!r}")


@pytest.mark.parametrize("prec", [O.F64, O.F32])
@pytest.mark.parametrize("streaming", [O.AB, O.AA])
@pytest.mark.parametrize("coll,eq", COMBOS_3D)
def test_d3q27_random_zoo(coll, eq, streaming, prec):
    d = O.Desc(lattice=O.D3Q27, coll=coll, eq=eq, streaming=streaming, precision=prec, X=9, Y=8, Z=7)
    m = lc.map_random_ab(d) if streaming == O.AB else lc.map_random_aa(d)
    p = O.Params(lbmViscosity=0.013, fx=3e-5, fy=-2e-5, fz=1e-5, inflow_vx=0.04, inflow_vy=0.01, inflow_vz=-0.02)
    ref, port = run_pair(d, m, p, nsteps=4)
    assert_same(ref, port, f"coll={coll} eq={eq} st={streaming} prec={prec}")


@pytest.mark.parametrize("prec", [O.F64, O.F32])
@pytest.mark.parametrize("streaming", [O.AB, O.AA])
@pytest.mark.parametrize("coll", [O.SRT, O.CLBM])
def test_d2q9_random_zoo(coll, streaming, prec):
    d = O.Desc(lattice=O.D2Q9, coll=coll, eq=O.EQ_STD, streaming=streaming, precision=prec, X=13, Y=11, Z=1)
    m = lc.map_random_ab(d) if streaming == O.AB else lc.map_random_aa(d)
    p = O.Params(lbmViscosity=0.02, fx=2e-5, fy=-1e-5, inflow_vx=0.05, inflow_vy=-0.01)
    ref, port = run_pair(d, m, p, nsteps=4)
    assert_same(ref, port, f"2d coll={coll} st={streaming} prec={prec}")


@pytest.mark.parametrize("macro,inflow", [(O.MACRO_VOID, O.INFLOW_CONST), (O.MACRO_MEAN, O.INFLOW_CONST), (O.MACRO_DEFAULT, O.INFLOW_PROFILE_YZ), (O.MACRO_DEFAULT, O.INFLOW_NONE)])
@pytest.mark.parametrize("prec", [O.F64, O.F32])
def test_d3q27_cum_macro_and_inflow_flavours(macro, inflow, prec):
    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, macro=macro, inflow=inflow, precision=prec, X=8, Y=7, Z=6)
    m = lc.map_random_ab(d, seed=3)
    rs = np.random.RandomState(5)
    prof = (0.05 * rs.random_sample((d.Z, d.Y))).astype(d.dtype)
    p = O.Params(lbmViscosity=0.004, fx=1e-5, inflow_vx=0.03, vx_profile=prof if inflow == O.INFLOW_PROFILE_YZ else None)
    ref, port = run_pair(d, m, p, nsteps=5)
    assert_same(ref, port, f"macro={macro} inflow={inflow} prec={prec}")


@pytest.mark.parametrize("prec", [O.F64, O.F32])
@pytest.mark.parametrize("streaming", [O.AB, O.AA])
@pytest.mark.parametrize("gates", [0, O.GATE_MEANS, O.GATE_FLUCS, O.GATE_MEANS | O.GATE_FLUCS])
def test_d2q9_with_mean_macro(gates, streaming, prec):
    """The solver-defined macro class of sim_2D/sim2d_2.cu:53-104 (gated velocity sums and fluctuation sums about a frozen mean),
    restated on the reference's D2Q9_MACRO_Base in oracle/ref_d2q9.cpp and driven by the reference's own kernel."""
    d = O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=streaming, macro=O.MACRO_WITH_MEAN_2D, inflow=O.INFLOW_PARABOLIC_Y, precision=prec, X=13, Y=11, Z=1)
    m = lc.map_random_ab(d) if streaming == O.AB else lc.map_random_aa(d)
    p = O.Params(lbmViscosity=0.02, fx=2e-5, fy=-1e-5, inflow_vx=0.07, inflow_vy=1.0, inflow_vz=0.125, macro_gates=gates)
    out = []
    for kind in ("reference", "port"):
        orc = O.Oracle(d, kind)
        a = lc.noisy_df(d, orc, seed=11)
        b = a.copy()
        mac = d.new_macro()
        mac[5:7] = (0.01 * np.random.RandomState(3).standard_normal(mac[5:7].shape)).astype(d.dtype)  # a frozen mean to fluctuate about
        orc.initial_macro(p, a, mac)
        orc.step(p, a, b, mac, m, 0, 5, 1)
        out.append((a, b, mac))
    assert_same(out[0], out[1], f"with-mean gates={gates} st={streaming} prec={prec}")
    mac = out[1][2]
    assert (np.abs(mac[3:5]).max() > 0) == bool(gates & O.GATE_MEANS) and (mac[7:].max() > 0) == bool(gates & O.GATE_FLUCS)


@pytest.mark.parametrize("macro", [O.MACRO_VOID, O.MACRO_MEAN])
def test_d2q9_macro_flavours(macro):
    d = O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=O.AB, macro=macro, X=10, Y=9, Z=1)
    m = lc.map_random_ab(d, seed=4)
    p = O.Params(lbmViscosity=0.01, fx=1e-5, inflow_vx=0.03)
    ref, port = run_pair(d, m, p, nsteps=5)
    assert_same(ref, port, f"2d macro={macro}")


@pytest.mark.parametrize("streaming", [O.AB, O.AA])
def test_ghost_plane_rule(streaming):
    """nproc>1 index rule with one ghost x-plane per side (kernels.h:21-29,39-48): no wrapping in x."""
    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=streaming, X=6, Y=8, Z=8, ox=1, nproc=2)
    m = lc.map_duct_periodic_x(d)
    m[0], m[-1] = m[-2], m[1]  # ghost map planes = periodic neighbours
    p = O.Params(lbmViscosity=0.01, fx=1e-5)
    ref, port = run_pair(d, m, p, nsteps=2)
    assert_same(ref, port, f"ghost st={streaming}")


def test_long_run_smooth_box_stays_identical():
    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=12, Y=12, Z=12)
    m = lc.map_periodic(d)
    p = O.Params(lbmViscosity=1e-3, fx=1e-6)
    out = []
    for kind in ("reference", "port"):
        orc = O.Oracle(d, kind)
        a = d.new_df()
        orc.set_equilibrium_field(a, *lc.smooth_fields(d))
        mac = d.new_macro()
        orc.step(p, a, a, mac, m, 0, 200, 2)
        out.append((a, a, mac))
    assert_same(out[0], out[1], "200 steps A-A")


@pytest.mark.parametrize("prec", [O.F64, O.F32])
@pytest.mark.parametrize("coll", [O.SRT, O.CLBM])
def test_d2q9_bouzidi_near_wall(coll, prec):
    """GEO_FLUID_NEAR_WALL with random interpolation coefficients in [-0.8, 1.2] (both Bouzidi branches and "no wall")."""
    d = O.Desc(lattice=O.D2Q9, coll=coll, eq=O.EQ_STD, streaming=O.AB, precision=prec, X=13, Y=11, Z=1)
    m, bz = lc.map_and_coeffs_bouzidi(d)
    p = O.Params(lbmViscosity=0.02, fx=2e-5, fy=-1e-5, inflow_vx=0.05, inflow_vy=-0.01, bouzidi=bz)
    ref, port = run_pair(d, m, p, nsteps=3, seed=5)
    assert_same(ref, port, f"bouzidi coll={coll} prec={prec}")


@pytest.mark.parametrize("prec", [O.F64, O.F32])
@pytest.mark.parametrize("streaming", [O.AB, O.AA])
def test_d2q9_parabolic_inflow(streaming, prec):
    """NSE2D_Data_ParabolicInflow (sim_2D/sim2d_3.cu:36-55): Poiseuille profile over y evaluated per inflow cell, with the double
    literals of the reference."""
    d = O.Desc(lattice=O.D2Q9, coll=O.CLBM, eq=O.EQ_STD, streaming=streaming, inflow=O.INFLOW_PARABOLIC_Y, precision=prec, X=13, Y=11, Z=1)
    m = lc.map_random_ab(d) if streaming == O.AB else lc.map_random_aa(d)
    p = O.Params(lbmViscosity=0.02, fx=2e-5, fy=-1e-5, inflow_vx=0.07, inflow_vy=1.0, inflow_vz=1.0 / (d.Y - 3))
    ref, port = run_pair(d, m, p, nsteps=4)
    assert_same(ref, port, f"2d parabolic st={streaming} prec={prec}")


def test_the_high_precision_rho_build_is_a_different_build():
    """USE_HIGH_PRECISION_RHO (defs.h:252) changes the summation of the density (d3q27/common.h:19-29): the reference compiled with the
    switch must differ from the default build in the last bits (otherwise oracle/ref_d3q27_cum_hprho.cpp did not get the switch into
    common.h), and only there."""
    outs = []
    for coll in (O.CUM, O.CUM_HP_RHO):
        d = O.Desc(lattice=O.D3Q27, coll=coll, eq=O.EQ_INV_CUM, streaming=O.AB, precision=O.F32, X=9, Y=8, Z=7)
        m = lc.map_random_ab(d)
        p = O.Params(lbmViscosity=0.013, fx=3e-5, fy=-2e-5, fz=1e-5, inflow_vx=0.04, inflow_vy=0.01, inflow_vz=-0.02)
        orc = O.Oracle(d, "reference")
        a = lc.noisy_df(d, orc, seed=11, noise=0.05)
        b = a.copy()
        mac = d.new_macro()
        orc.step(p, a, b, mac, m, 0, 1, 1)
        outs.append(b)
    assert not np.array_equal(outs[0], outs[1])
    np.testing.assert_allclose(outs[0], outs[1], rtol=0, atol=2e-6)
