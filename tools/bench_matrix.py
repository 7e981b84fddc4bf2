#!/usr/bin/env python
"""Throughput of every kernel family of liblbmx.so on one B200 (development tool; writes a markdown table).

    python tools/bench_matrix.py [--out profiles/bench_matrix_r1.md] [--steps 40]

Periodic boxes (all GEO_PERIODIC), uniform equilibrium + body force, device-resident, CUDA events on the engine's stream.
Algorithmic bytes per update B = Q * 2 * sizeof(real); GB/s = MLUPS * B / 1000."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tnl_lbm_b200 import binding as B  # noqa: E402

NAMES = {B.CUM: "CUM", B.SRT: "SRT", B.BGK: "BGK", B.MRT_LES: "MRT_LES", B.CLBM: "CLBM", B.SRT_MODIF_FORCE: "SRT_MODIF_FORCE", B.KBC_N1: "KBC_N1", B.KBC_N4: "KBC_N4", B.KBC_C4: "KBC_C4",
         B.CUM_2017: "CUM (USE_GEIER_CUM_2017)", B.CUM_ANTIALIAS: "CUM (ANTIALIAS)", B.CUM_2017_ANTIALIAS: "CUM (2017 + ANTIALIAS)"}


def run(lattice, coll, eq, prec, streaming, shape, steps, map_kind="periodic", macro=B.MACRO_DEFAULT, flags=0):
    X, Y, Z = shape
    e = B.Engine(lattice=lattice, coll=coll, eq=eq, streaming=streaming, macro=macro, flags=flags, inflow=B.INFLOW_CONST, precision=prec, X=X, Y=Y, Z=Z)
    per = 6 if lattice == B.D2Q9 else 7
    m = np.full((X, Z, Y), per, dtype=np.int16)
    if map_kind == "cavity":  # SURVEY §8d cfg 2: walls on x faces and y=0, lid row y=Y-1 = GEO_INFLOW
        m[...] = 0
        m[0], m[X - 1] = 1, 1
        m[:, :, 0] = 1
        m[:, :, Y - 1] = 2
    e.map_upload(m)
    e.set_equilibrium(1.0, 0.02, 0.01, 0.0 if lattice == B.D2Q9 else -0.01)
    e.set_params(lbmViscosity=0.05 if map_kind == "cavity" else 1e-3, fx=1e-6, inflow_vx=0.1)
    e.step(6)
    e.sync()
    ms = e.step_timed(steps)
    assert not e.has_nan()
    st = e.stats()
    e.close()
    cells = X * Y * Z
    mlups = cells * steps / (ms * 1e-3) / 1e6
    q = {B.D2Q9: 9, B.D3Q19: 19}.get(lattice, 27)
    bpu = q * 2 * (8 if prec == B.F64 else 4)
    return mlups, mlups * bpu / 1e3, st.bulk_regs, bpu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "bench_matrix_r2.md"))
    ap.add_argument("--round", default="2")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--only", default="", help="run only the rows whose name contains this string (e.g. 'ext' = the rows added after the first table)")
    ap.add_argument("--operator", default="", help="run only the rows whose operator name contains this string (e.g. SRT_MODIF)")
    a = ap.parse_args()
    peak = 6446.9
    pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pj):
        peak = float(json.load(open(pj))["hbm_gbs"])
    rows = []
    cfgs = []
    for coll, eq in ((B.CUM, B.EQ_INV_CUM), (B.SRT, B.EQ_STD), (B.BGK, B.EQ_STD), (B.MRT_LES, B.EQ_STD)):
        for prec in (B.F64, B.F32):
            for st in (B.AA, B.AB):
                shape = (384, 384, 384) if (prec == B.F64 or st == B.AB) else (512, 512, 512)
                cfgs.append(("D3Q27", B.D3Q27, coll, eq, prec, st, shape, "periodic"))
    for coll in (B.MRT_LES, B.SRT):  # BASELINE.json configs[4]: "D3Q19 MRT fp32 ... (A-A vs A-B)"; no reference implementation (parity unpinned)
        for prec in (B.F32, B.F64):
            for st in (B.AA, B.AB):
                cfgs.append(("D3Q19", B.D3Q19, coll, B.EQ_STD, prec, st, (512, 512, 512) if prec == B.F32 else (384, 384, 384), "periodic"))
    for coll in (B.SRT, B.CLBM):
        for prec in (B.F64, B.F32):
            for st in (B.AA, B.AB):
                cfgs.append(("D2Q9", B.D2Q9, coll, B.EQ_STD, prec, st, (8192, 8192, 1), "periodic"))
    cfgs.append(("D2Q9 cavity 1024^2 (cfg 2, 151 MB: L2-resident)", B.D2Q9, B.SRT, B.EQ_STD, B.F64, B.AB, (1024, 1024, 1), "cavity"))
    cfgs.append(("D2Q9 cavity 8192^2", B.D2Q9, B.SRT, B.EQ_STD, B.F64, B.AB, (8192, 8192, 1), "cavity"))
    cfgs = [c + (B.MACRO_DEFAULT, 0) for c in cfgs]
    # rows added after the first table: further D3Q27 operators, MACRO_Mean (13 read-modify-write fields per cell on top of the DFs:
    # its own bytes-per-update figure), and the parity-arithmetic build of the headline kernel
    for coll, eq in ((B.CLBM, B.EQ_STD), (B.SRT_MODIF_FORCE, B.EQ_STD)):
        for prec in (B.F64, B.F32):
            for st in (B.AA, B.AB):
                cfgs.append(("D3Q27 ext", B.D3Q27, coll, eq, prec, st, (384, 384, 384), "periodic", B.MACRO_DEFAULT, 0))
    for coll, eq in ((B.KBC_N1, B.EQ_STD), (B.KBC_N4, B.EQ_ENTROPIC), (B.KBC_C4, B.EQ_ENTROPIC), (B.CUM_2017, B.EQ_INV_CUM), (B.CUM_ANTIALIAS, B.EQ_INV_CUM), (B.CUM_2017_ANTIALIAS, B.EQ_INV_CUM)):
        for prec in (B.F64, B.F32):
            for st in (B.AA, B.AB):
                cfgs.append(("D3Q27 ext2", B.D3Q27, coll, eq, prec, st, (384, 384, 384), "periodic", B.MACRO_DEFAULT, 0))
    for st in (B.AA, B.AB):
        cfgs.append(("D3Q27 ext MACRO_Mean (+13 RMW fields: 640 B per update)", B.D3Q27, B.CUM, B.EQ_INV_CUM, B.F64, st, (320, 320, 320), "periodic", B.MACRO_MEAN, 0))
        cfgs.append(("D3Q27 ext parity arithmetic", B.D3Q27, B.CUM, B.EQ_INV_CUM, B.F64, st, (384, 384, 384), "periodic", B.MACRO_DEFAULT, B.FLAG_STRICT_ARITH))
        cfgs.append(("D3Q27 ext parity arithmetic", B.D3Q27, B.CUM, B.EQ_INV_CUM, B.F32, st, (384, 384, 384), "periodic", B.MACRO_DEFAULT, B.FLAG_STRICT_ARITH))
    if a.only:
        cfgs = [c for c in cfgs if a.only in c[0]]
    if a.operator:
        cfgs = [c for c in cfgs if a.operator in NAMES[c[2]]]
    for name, lat, coll, eq, prec, st, shape, mk, macro, flags in cfgs:
        mlups, gbs, regs, bpu = run(lat, coll, eq, prec, st, shape, a.steps, mk, macro, flags)
        row = f"| {name} | {NAMES[coll]} | {'fp64' if prec == B.F64 else 'fp32'} | {'A-A' if st == B.AA else 'A-B'} | {shape[0]}x{shape[1]}x{shape[2]} | {bpu} | {mlups:,.0f} | {gbs:,.0f} | {gbs / peak * 100:.0f} % | {regs} |"
        print(row, flush=True)
        rows.append(row)
    with open(a.out, "w") as f:
        f.write(f"# Kernel-family throughput on one B200 (round {a.round})\n\n`python tools/bench_matrix.py` -- periodic boxes, device-resident, "
                f"{a.steps} timed steps after 6 warm-up, CUDA events on the engine stream.\nB = algorithmic bytes per update (Q x 2 x sizeof real); "
                f"GB/s = MLUPS x B; % of the measured copy bandwidth {peak:.0f} GB/s (MEASURED_PEAKS.json).\n\n"
                "| lattice | operator | real | streaming | lattice size | B | MLUPS | GB/s | of measured peak | regs (A-A even / A-B) |\n|---|---|---|---|---|---|---|---|---|---|\n")
        f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
