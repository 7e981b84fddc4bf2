"""Drive the CUDA engine (through the C ABI) on the golden-case registry, with exactly the inputs the CPU checkers get."""
from __future__ import annotations

import numpy as np

import golden_cases as gc
from oracle import oracle as O
from tnl_lbm_b200 import binding as B


def engine_for(case: gc.Case, **kw) -> B.Engine:
    d = case.desc
    # the selector values of include/lbmx.h and oracle/oracle_api.h are the same numbers by construction
    args = dict(lattice=d.lattice, coll=d.coll, eq=d.eq, streaming=d.streaming, macro=d.macro, inflow=d.inflow, precision=d.precision,
                X=d.X, Y=d.Y, Z=d.Z, macro_policy=B.MACRO_LAST_STEP)
    args.update(kw)
    return B.Engine(**args)


def set_params(e: B.Engine, p: O.Params):
    e.set_params(lbmViscosity=p.lbmViscosity, fx=p.fx, fy=p.fy, fz=p.fz, inflow_vx=p.inflow_vx, inflow_vy=p.inflow_vy, inflow_vz=p.inflow_vz,
                 stat_counter=0)
    if p.vx_profile is not None:
        e.set_inflow_profile(p.vx_profile)
    if p.bouzidi is not None:
        e.bouzidi_upload(p.bouzidi)


def run_case_engine(case: gc.Case, chunk: int | None = None, **kw):
    """Returns (df_cur, macro) after case.nsteps; the initial state is the bit-identical array the port starts from."""
    d = case.desc
    port = O.Oracle(d, "port")
    df0 = gc.initial_df(case, port)
    with engine_for(case, **kw) as e:
        e.map_upload(case.make_map(d))
        e.df_upload(df0, which=0)
        if d.streaming == O.AB:
            e.df_upload(df0, which=1)
        set_params(e, case.params)
        e.macro_init()
        n = case.nsteps
        if d.macro == O.MACRO_WITH_MEAN_2D:  # the two phases of gc.run_case, through the C ABI
            half = n // 2
            e.set_params(macro_gates=O.GATE_MEANS)
            for k in ([1] * half if chunk == 1 else [half]):
                e.step(k)
            mac = e.macro_download()
            gc.freeze_means(mac, half)
            e.macro_upload(mac)
            e.set_params(macro_gates=O.GATE_FLUCS)
            for k in ([1] * (n - half) if chunk == 1 else [n - half]):
                e.step(k)
        elif chunk:
            done = 0
            while done < n:
                k = min(chunk, n - done)
                e.step(k)
                done += k
        else:
            e.step(n)
        e.sync()
        assert e.iterations == n
        return e.df_download(0), e.macro_download(), e.stats()
