"""Obstacle-heavy maps (VERDICT r1 'next' #6): step time and DRAM bytes per lattice update as a function of what the map holds.

  python tools/solid_bench.py [--size 384] [--steps 20] [--maps periodic,duct,sim1,sphere,wall30,nothing30] [--streaming AA|AB]
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:'k_bulk|k_boundary' --csv ... \
      python tools/solid_bench.py --maps wall30 --steps 2 --warmup 1      (bytes per launch; the script prints the cell counts to divide by)

A cell that the reference streams (everything but GEO_NOTHING, d3q27/bc.h:53-60) costs Q x 2 x sizeof(real) algorithmic bytes; a
GEO_NOTHING cell costs none.  The table line is  <map>: cells by owner (bulk kernel / boundary list), ms per step, GB/s of algorithmic bytes.
Development tool: uses the product library only (no CPU checker)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tnl_lbm_b200 import binding as B  # noqa: E402

FLUID, WALL, INFLOW, OUTFLOW_RIGHT, PERIODIC, NOTHING = 0, 1, 2, 5, 7, 8


def make_map(kind, S, streaming):
    rng = np.random.default_rng(12345)
    m = np.full((S, S, S), PERIODIC, dtype=np.int16)  # (x, z, y)
    if kind == "periodic":
        return m
    if kind == "fluid":  # every cell GEO_FLUID: under A-B the face cells clamp (and are loaded twice by the bulk kernel)
        m[...] = FLUID
        return m
    if kind == "shell":  # only the GEO_NOTHING skin of the duct
        m[:, 0, :] = m[:, S - 1, :] = m[:, :, 0] = m[:, :, S - 1] = NOTHING
        return m
    if kind == "ring":  # only the wall ring of the duct
        m[:, 1, :] = m[:, S - 2, :] = m[:, :, 1] = m[:, :, S - 2] = WALL
        return m
    if kind == "plane":  # one wall plane: whole rows of obstacle cells, none inside the other rows
        m[:, 1, :] = WALL
        return m
    if kind == "columns":  # two obstacle cells in every row
        m[:, :, 1] = m[:, :, S - 2] = WALL
        return m
    if kind == "duct":  # sim_NSE/sim_2.cu:125-138
        m[:, 1, :] = m[:, S - 2, :] = m[:, :, 1] = m[:, :, S - 2] = WALL
        m[:, 0, :] = m[:, S - 1, :] = m[:, :, 0] = m[:, :, S - 1] = NOTHING
        return m
    if kind in ("sim1", "sphere"):
        # walls behind a GEO_NOTHING shell; A-B: inflow / outflow faces as sim_NSE/sim_1.cu:25-52, A-A: periodic in x (inflow cells on a bare face are ill-defined there)
        m[...] = FLUID
        if streaming == "AB":
            m[0] = INFLOW
            m[S - 1] = OUTFLOW_RIGHT
        else:
            m[0] = m[S - 1] = PERIODIC
        m[:, 1, :] = m[:, S - 2, :] = m[:, :, 1] = m[:, :, S - 2] = WALL
        m[:, 0, :] = m[:, S - 1, :] = m[:, :, 0] = m[:, :, S - 1] = NOTHING
        if kind == "sim1":  # wall slab with a centred hole (sim_1.cu:38-50), 12 cells thick at res 4 of a 126-cell height
            cx, w = S // 5, max(1, S * 12 // 128)
            blk = m[cx : cx + w + 1, 1 : S - 1, 1 : S - 1]
            z, y = np.meshgrid(np.arange(1, S - 1), np.arange(1, S - 1), indexing="ij")
            hole = (z >= S * 4 // 10) & (z <= S * 6 // 10) & (y >= S * 4 // 10) & (y <= S * 6 // 10)
            blk[:, ~hole] = WALL
        else:  # sim_NSE/sim_3.cu: a solid sphere (lbmDrawSphere paints every cell within the radius)
            x, z, y = np.meshgrid(np.arange(S), np.arange(S), np.arange(S), indexing="ij", sparse=True)
            r2 = (x - S // 3) ** 2 + (z - S // 2) ** 2 + (y - S // 2) ** 2
            m[(r2 <= (S // 5) ** 2) & (m == FLUID)] = WALL
        return m
    if kind == "nothinghalf":  # a large inert region (the outside of an immersed body, say): the upper half in z
        m[:, S // 2 :, :] = NOTHING
        m[:, S // 2 - 1, :] = WALL
        m[:, 1, :] = WALL
        m[:, 0, :] = NOTHING
        return m
    if kind in ("wall30", "nothing30", "wall30blocks"):
        if kind == "wall30blocks":  # 8^3 blocks, 30 % of them solid: a porous medium
            b = rng.random((S // 8, S // 8, S // 8)) < 0.30
            sel = np.repeat(np.repeat(np.repeat(b, 8, 0), 8, 1), 8, 2)
        else:
            sel = rng.random((S, S, S)) < 0.30
        m[sel] = WALL if kind.startswith("wall") else NOTHING
        return m
    raise SystemExit(f"unknown map {kind}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--maps", default="periodic,duct,sim1,sphere,wall30blocks,wall30,nothing30,nothinghalf")
    ap.add_argument("--streaming", default="AA")
    ap.add_argument("--precision", default="f64")
    a = ap.parse_args()
    S = a.size
    rs = 8 if a.precision == "f64" else 4
    for kind in a.maps.split(","):
        m = make_map(kind, S, a.streaming)
        with B.Engine(lattice=B.D3Q27, coll=B.CUM, eq=B.EQ_INV_CUM, streaming=B.AA if a.streaming == "AA" else B.AB, macro=B.MACRO_DEFAULT, inflow=B.INFLOW_CONST,
                      precision=B.F64 if rs == 8 else B.F32, X=S, Y=S, Z=S, periodic_x=1, macro_policy=B.MACRO_LAST_STEP) as e:
            e.map_upload(m)
            e.set_equilibrium(1.0, 0.0, 0.0, 0.0)
            e.set_params(lbmViscosity=1e-3, fx=1e-6, inflow_vx=0.02)
            e.step(a.warmup)
            e.sync()
            ms = e.step_timed(a.steps) / a.steps
            st = e.stats()
            assert not e.has_nan()
        streamed = int(np.sum(m != NOTHING))
        alg = streamed * 27 * 2 * rs
        print(json.dumps({"map": kind, "size": S, "streaming": a.streaming, "precision": a.precision, "cells": int(m.size), "streamed_cells": streamed, "solid_fraction": float(np.mean((m == WALL) | (m == NOTHING))),
                          "bulk_kernel_cells": int(st.bulk_cells), "boundary_list_cells": int(st.boundary_cells), "ms_per_step": ms, "MLUPS_all_cells": m.size / ms / 1e3,
                          "algorithmic_bytes_per_step": alg, "GBs_algorithmic": alg / ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
