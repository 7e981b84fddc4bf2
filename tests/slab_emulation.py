"""Multi-slab runs of the CPU oracle with hand-exchanged ghost planes (SURVEY.md App. B "ghost-plane oracle mode").

Every slab is an oracle instance with one ghost x-plane per side (ox = 1) and the reference's nproc > 1 index rule
(kernels.h:21-29,39-48).  After every step the planes named by lbmx_halo_plan() -- the host-side plan the CUDA engine itself
executes with NCCL -- are copied between neighbouring slabs.  This is how the reference's DistributedNDArraySynchronizer
semantics (un-vendored TNL; call sites lbm_block.hpp:410-451, lbm.hpp:196-280) are pinned: N slabs == 1 slab."""
from __future__ import annotations

import copy

import numpy as np

import golden_cases as gc
from oracle import oracle as O
from tnl_lbm_b200 import binding as B


def split(a_global: np.ndarray, nslabs: int, axis: int, periodic: bool = True):
    """Slabs with one ghost plane per side taken from the periodic neighbours."""
    X = a_global.shape[axis]
    out = []
    for r in range(nslabs):
        x0, xl = B.decompose_x(X, nslabs, r)
        idx = np.arange(x0 - 1, x0 + xl + 1)
        if periodic:
            idx %= X
        else:
            idx = np.clip(idx, 0, X - 1)
        out.append(np.ascontiguousarray(np.take(a_global, idx, axis=axis)))
    return out


def exchange(slabs, plan, periodic=True):
    """Apply one step's halo plan to a list of [Q, X+2, Z, Y] arrays (all reads before all writes)."""
    n = len(slabs)
    staged = []
    for r, a in enumerate(slabs):
        for m in plan:
            dst = r + 1 if m["to_right"] else r - 1
            if periodic:
                dst %= n
            elif dst < 0 or dst >= n:
                continue
            staged.append((dst, m["dirs"], m["dst_plane"], a[m["dirs"], m["src_plane"]].copy()))
    for dst, dirs, plane, data in staged:
        slabs[dst][dirs, plane] = data


def run_slabs_oracle(case: gc.Case, nslabs: int, kind: str = "port", nthreads: int = 1):
    """Run `case` (global desc, ox = 0) as `nslabs` ghosted slabs; returns the assembled global (df_cur, macro)."""
    dg = case.desc
    glob = O.Oracle(dg, kind)
    df0 = gc.initial_df(case, glob)
    mapg = case.make_map(dg)
    dfs_a = split(df0, nslabs, 1)
    dfs_b = [a.copy() for a in dfs_a]
    maps = split(mapg, nslabs, 0)
    descs, orcs, macs = [], [], []
    for r in range(nslabs):
        x0, xl = B.decompose_x(dg.X, nslabs, r)
        d = copy.copy(dg)
        d.X, d.ox, d.nproc = xl, 1, 2
        descs.append(d)
        orcs.append(O.Oracle(d, kind))
        macs.append(d.new_macro())
    p = case.params
    p.stat_counter = 0
    lengths = [d.X for d in descs]
    assert len(set(lengths)) == 1, "the emulation keeps equal slabs so that one halo plan fits all"
    aa = dg.streaming == O.AA
    for it in range(case.nsteps):
        for r in range(nslabs):
            orcs[r].step(p, dfs_a[r], dfs_b[r], macs[r], maps[r], it, 1, nthreads)
        plan = B.halo_plan(dg.lattice, dg.streaming, it, lengths[0])
        written = dfs_a if (aa or it % 2 == 1) else dfs_b  # A-B: even iterations write df_b (lbm.hpp:320-327)
        exchange(written, plan)
    cur = dfs_a if (aa or case.nsteps % 2 == 0) else dfs_b
    df = np.concatenate([a[:, 1:-1] for a in cur], axis=1)
    mac = np.concatenate([m[:, 1:-1] for m in macs], axis=1)
    return df, mac
