"""BASELINE.json configs[0] at its full size against the CPU oracle, and properties at BASELINE.json's full single-GPU size (512^3, D3Q27 cumulant fp64 A-A -- 29 GB of distributions), where the CPU
oracle cannot follow:

* periodic replication -- a 512^3 periodic box initialised with a field of period 64 must reproduce, bit for bit, the 64^3 box
  (which IS checked against the oracle) tiled 8 x 8 x 8: every cell sees the same neighbourhood;
* conservation -- total mass is constant, total x-momentum grows by F per cell per step.
Skipped when the device has less than 60 GB free."""
import numpy as np
import pytest

import golden_cases as gc
import lbm_cases as lc
from oracle import oracle as O
from tnl_lbm_b200 import binding as B

pytestmark = pytest.mark.gpu

S, T, STEPS = 512, 64, 30
NU, FX = 1e-3, 1e-6


def run(size, fields, steps):
    with B.Engine(coll=B.CUM, eq=B.EQ_INV_CUM, streaming=B.AA, precision=B.F64, inflow=B.INFLOW_NONE, X=size, Y=size, Z=size) as e:
        e.map_upload(np.full((size, size, size), 7, dtype=np.int16))
        e.set_params(lbmViscosity=NU, fx=FX)
        e.set_equilibrium_field(*fields)
        e.macro_init()
        m0 = e.macro_download()
        e.step(steps)
        return m0, e.macro_download()


def test_512_cube_replicates_the_64_cube_and_conserves():
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs 60 GB of free device memory")
    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA, X=T, Y=T, Z=T)
    small_fields = [np.ascontiguousarray(f) for f in lc.smooth_fields(d)]
    _, small = run(T, small_fields, STEPS)
    # the small run itself against the CPU oracle
    case = gc.Case("tile", d, O.Params(lbmViscosity=NU, fx=FX), lc.map_periodic, STEPS, "smooth")
    _, ref = gc.run_case(case, "port", nthreads=8)
    for lo, hi, label in lc.macro_groups(d):
        assert lc.rel_err(small[lo:hi], ref[lo:hi]) <= 1e-12, label
    # the full-size run
    r = S // T
    big_fields = [np.tile(f, (r, r, r)) for f in small_fields]
    m0, big = run(S, big_fields, STEPS)
    del big_fields
    blocks = big.reshape(4, r, T, r, T, r, T)
    assert np.array_equal(blocks, np.broadcast_to(small[:, None, :, None, :, None, :], blocks.shape)), "512^3 differs from the tiled 64^3 result"
    # conservation (macros are the pre-collision rho, u with the half-force shift: rho*u = j + F/2)
    n = float(S) ** 3
    mass0, mass1 = float(m0[0].sum(dtype=np.float64)), float(big[0].sum(dtype=np.float64))
    assert abs(mass1 - mass0) / n < 1e-13
    jx0 = float((m0[0] * m0[1]).sum(dtype=np.float64))   # macro_init zeroes the force: rho*u = j
    jx1 = float((big[0] * big[1]).sum(dtype=np.float64))  # last step's pre-collision state: j after STEPS-1 steps, + F/2
    assert abs((jx1 - jx0) / n - (STEPS - 0.5) * FX) < 1e-12


def test_sim1_resolution_4_against_the_cpu_oracle():
    """BASELINE.json configs[0] / SURVEY 8d cfg 1: the geometry of sim_NSE/sim_1.cu at resolution 4 (512 x 128 x 128, orifice channel,
    inflow / outflow / walls / GEO_NOTHING shell), D3Q27 cumulant fp64, A-B, 40 steps from rest -- the case the reference runs on
    its CPU build.  Default kernels within 1e-12 of the CPU oracle, parity-arithmetic kernels bit-identical, at full size."""
    import os

    from engine_runner import run_case_engine

    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AB, precision=O.F64, X=512, Y=128, Z=128)
    dl = 0.41 / 126
    dt = 1e-5 / 1.5e-5 * dl * dl
    case = gc.Case("sim1_res4", d, O.Params(lbmViscosity=1e-5, inflow_vx=1.0 * dt / dl), lc.map_sim1_channel, 40, "uniform")
    df_ref, mac_ref = gc.run_case(case, "port", nthreads=os.cpu_count() or 8)
    df, mac, _ = run_case_engine(case)
    assert np.isfinite(df).all()
    assert lc.rel_err_df(df, df_ref, d) <= 1e-12
    for lo, hi, label in lc.macro_groups(d):
        assert lc.rel_err(mac[lo:hi], mac_ref[lo:hi]) <= 1e-12, label
    del df, mac
    df, mac, _ = run_case_engine(case, flags=B.FLAG_STRICT_ARITH)
    assert np.array_equal(df, df_ref) and np.array_equal(mac, mac_ref), "parity-arithmetic kernels differ from the CPU oracle at full size"
