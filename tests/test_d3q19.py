"""D3Q19 -- named by BASELINE.json (configs[4]) but absent from the reference: PARITY UNPINNED by construction.
What can be checked: conservation laws and the analytical duct profile (physics), and agreement between the engine's kernels
and the independent CPU implementation in oracle/lbm_oracle.cpp (L19), which applies the reference's SRT / MRT_LES formulas
to the 19-velocity set."""
import numpy as np
import pytest

import golden_cases as gc
import lbm_cases as lc
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.available("port"), reason="oracle port not built")


def case19(coll, streaming, prec, nsteps=4, mapper=None, init="noisy"):
    d = O.Desc(lattice=O.D3Q19, coll=coll, eq=O.EQ_STD, streaming=streaming, precision=prec, X=9, Y=8, Z=7)
    p = O.Params(lbmViscosity=0.013, fx=3e-5, fy=-2e-5, fz=1e-5, inflow_vx=0.04, inflow_vy=0.01, inflow_vz=-0.02)
    mk = mapper or (lambda dd: lc.map_random_ab(dd) if dd.streaming == O.AB else lc.map_random_aa(dd))
    return gc.Case("q19", d, p, mk, nsteps, init)


@pytest.mark.parametrize("coll", [O.SRT, O.MRT_LES])
@pytest.mark.parametrize("streaming", [O.AB, O.AA])
def test_cpu_d3q19_conserves_mass_and_momentum(coll, streaming):
    d = O.Desc(lattice=O.D3Q19, coll=coll, eq=O.EQ_STD, streaming=streaming, X=10, Y=9, Z=8)
    fx = 2e-6 if coll == O.SRT else 0.0  # MRT_LES ignores the body force (col_mrt.h: no force term)
    case = gc.Case("box19", d, O.Params(lbmViscosity=0.01, fx=fx), lc.map_periodic, 20, "noisy")
    orc = O.Oracle(d, "port")
    df0 = gc.initial_df(case, orc)
    df, mac = gc.run_case(case, "port")
    n = d.X * d.Y * d.Z
    c = lc.C27[:19]
    assert abs(df.sum() - df0.sum()) / n < 1e-13
    for a in range(3):
        j0 = float((df0 * c[:, a][:, None, None, None]).sum())
        j1 = float((df * c[:, a][:, None, None, None]).sum())
        expect = case.nsteps * (fx if a == 0 else 0.0)
        assert abs((j1 - j0) / n - expect) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("prec", [O.F64, O.F32])
@pytest.mark.parametrize("streaming", [O.AB, O.AA])
@pytest.mark.parametrize("coll", [O.SRT, O.MRT_LES])
def test_engine_d3q19_matches_cpu_implementation(coll, streaming, prec):
    from engine_runner import run_case_engine

    case = case19(coll, streaming, prec)
    df, mac, stats = run_case_engine(case)
    assert df.shape[0] == 19 and stats.kernel_launches > 0
    ref_df, ref_mac = gc.run_case(case, "port")
    tol = 1e-12 if prec == O.F64 else 1e-5
    assert lc.rel_err_df(df, ref_df, case.desc) <= tol
    for lo, hi, label in lc.macro_groups(case.desc):
        assert lc.rel_err(mac[lo:hi], ref_mac[lo:hi]) <= tol, label


@pytest.mark.gpu
def test_engine_d3q19_duct_profile():
    from test_gpu_physics_and_edges import analytical_ux
    from tnl_lbm_b200 import binding as B

    X, Y, Z, nu, fx = 8, 36, 36, 1.0 / 6.0, 1e-6
    d = O.Desc(lattice=O.D3Q19, coll=O.SRT, eq=O.EQ_STD, streaming=O.AA, X=X, Y=Y, Z=Z)
    with B.Engine(lattice=B.D3Q19, coll=B.SRT, eq=B.EQ_STD, streaming=B.AA, precision=B.F64, X=X, Y=Y, Z=Z) as e:
        e.map_upload(lc.map_duct_slab_safe(d))
        e.set_equilibrium(1.0, 0, 0, 0)
        e.set_params(lbmViscosity=nu, fx=fx)
        e.step(8000)
        mac = e.macro_download()
    ux, ref = mac[1, X // 2].astype(np.float64), analytical_ux(Y, Z, fx, nu)
    inner = (slice(2, Z - 2), slice(2, Y - 2))
    l2 = np.sqrt(np.sum((ux[inner] - ref[inner]) ** 2) / np.sum(ref[inner] ** 2))
    assert l2 < 0.03, l2


@pytest.mark.gpu
def test_engine_d3q19_two_slabs_self_consistency():
    """Ghost planes + self-exchange carry 5 populations per direction on D3Q19; must equal the run without ghost planes."""
    from engine_runner import engine_for, run_case_engine, set_params

    case = case19(O.SRT, O.AA, O.F64, nsteps=7, mapper=lc.map_periodic)
    plain_df, _, _ = run_case_engine(case)
    port = O.Oracle(case.desc, "port")
    df0 = gc.initial_df(case, port)
    with engine_for(case, ghost_x=1, periodic_x=1) as e:
        e.map_upload(case.make_map(case.desc))
        e.df_upload(df0, 0)
        e.df_sync_ghosts()
        set_params(e, case.params)
        e.step(case.nsteps)
        # A-A after an odd number of steps holds the swapped-slot form in both runs alike
        assert np.array_equal(e.df_download(0), plain_df)
        assert e.stats().halo_bytes_sent == case.nsteps * 2 * 5 * case.desc.Y * case.desc.Z * 8
