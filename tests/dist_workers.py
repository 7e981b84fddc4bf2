"""Worker functions for the multi-process tests (spawned with torch.multiprocessing; must live in an importable module)."""
from __future__ import annotations

import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


def _init(rank, world, port, backend):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    return dist


def gloo_slab_worker(rank, world, port, case_name, outdir, kind="port"):
    """CPU: one slab per process, stepped by the CPU checker (`kind` = "port") or by the engine's own kernels compiled for the host
    (`kind` = "engine_host", tests/host_harness/engine_host.cpp); halo planes travel through torch.distributed (gloo) exactly as lbmx_halo_plan says."""
    import torch

    import golden_cases as gc
    import lbm_cases as lc
    from oracle import oracle as O
    from slab_emulation import split
    from tnl_lbm_b200 import binding as B

    dist = _init(rank, world, port, "gloo")
    case = DIST_CASES[case_name]()
    dg = case.desc
    glob = O.Oracle(dg, "port")
    df0 = gc.initial_df(case, glob)
    a = split(df0, world, 1)[rank]
    b = a.copy()
    m = split(case.make_map(dg), world, 0)[rank]
    x0, xl = B.decompose_x(dg.X, world, rank)
    d = copy.copy(dg)
    d.X, d.ox, d.nproc = xl, 1, 2
    orc = O.Oracle(d, kind)
    mac = d.new_macro()
    aa = dg.streaming == O.AA
    left, right = (rank - 1) % world, (rank + 1) % world
    for it in range(case.nsteps):
        orc.step(case.params, a, b, mac, m, it, 1, 1)
        arr = a if (aa or it % 2 == 1) else b
        for msg in B.halo_plan(dg.lattice, dg.streaming, it, xl):
            send_to, recv_from = (right, left) if msg["to_right"] else (left, right)
            out = torch.from_numpy(np.ascontiguousarray(arr[msg["dirs"], msg["src_plane"]]))
            inc = torch.empty_like(out)
            reqs = [dist.isend(out, send_to), dist.irecv(inc, recv_from)]
            for r in reqs:
                r.wait()
            arr[msg["dirs"], msg["dst_plane"]] = inc.numpy()
    cur = a if (aa or case.nsteps % 2 == 0) else b
    np.save(os.path.join(outdir, f"df_{rank}.npy"), cur[:, 1:-1])
    np.save(os.path.join(outdir, f"mac_{rank}.npy"), mac[:, 1:-1])
    dist.barrier()
    dist.destroy_process_group()


def nccl_engine_worker(rank, world, port, case_name, outdir, transport="auto", delay=0.0):
    """GPU: one engine slab per process / device; ghost planes exchanged by the engine itself (peer memory or NCCL send/recv).
    delay > 0: the state arrives as arrays WITH ghost planes (the restore path of LBM_BLOCK checkpoints: no collective call after the
    upload) and the last rank uploads `delay` seconds late, while its neighbours are already stepping -- their one-sided halo stores
    must wait for it (receiver-ready handshake of the peer-memory exchange)."""
    if transport == "nccl":
        os.environ["LBMX_HALO"] = "nccl"
    else:
        os.environ.pop("LBMX_HALO", None)
    import torch

    import golden_cases as gc
    from engine_runner import set_params
    from oracle import oracle as O
    from tnl_lbm_b200 import binding as B

    torch.cuda.set_device(rank)
    dist = _init(rank, world, port, "gloo")  # the process group only carries the NCCL id; the data path is the engine's own communicator
    case = DIST_CASES[case_name]()
    dg = case.desc
    glob = O.Oracle(dg, "port")
    df0 = gc.initial_df(case, glob)
    mapg = case.make_map(dg)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(B.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, 0)
    e = B.Engine(lattice=dg.lattice, coll=dg.coll, eq=dg.eq, streaming=dg.streaming, macro=dg.macro, inflow=dg.inflow, precision=dg.precision,
                 X=dg.X, Y=dg.Y, Z=dg.Z, rank=rank, nranks=world, device=rank, ghost_x=1, periodic_x=1)
    e.comm_init(idt.numpy().tobytes())
    x0, xl = e.layout.x_offset, e.layout.X_local
    e.map_upload(np.ascontiguousarray(mapg[x0 : x0 + xl]))
    if delay > 0:
        import time

        from slab_emulation import split

        ghosted = split(df0, world, 1)[rank]
        if rank == world - 1:
            time.sleep(delay)
        e.df_upload(ghosted, 0, with_ghosts=True)
        if dg.streaming == O.AB:
            e.df_upload(ghosted, 1, with_ghosts=True)
    else:
        mine = np.ascontiguousarray(df0[:, x0 : x0 + xl])
        e.df_upload(mine, 0)
        e.df_sync_ghosts()
        if dg.streaming == O.AB:
            e.df_upload(mine, 1)
    set_params(e, case.params)
    e.macro_init()
    e.step(case.nsteps)
    e.sync()
    np.save(os.path.join(outdir, f"df_{rank}.npy"), e.df_download(0))
    np.save(os.path.join(outdir, f"mac_{rank}.npy"), e.macro_download())
    st = e.stats()
    np.save(os.path.join(outdir, f"halo_{rank}.npy"), np.array([st.halo_bytes_sent, st.kernel_launches, st.halo_peer_memory]))
    dist.barrier()
    e.close()
    dist.destroy_process_group()


def shared_device_worker(rank, nprocs, names, reps, outdir, runner="engine"):
    """Several processes drive small engines on ONE GPU at the same time (time-slicing stretches the gap between a host call returning
    and its device work landing: the situation that exposed the copies on the legacy default stream, profiles/gpu_suite_r1_shared_gpu.md).
    Every repetition of a case must reproduce the first bit for bit; the first result goes to `outdir` for the comparison with the
    CPU checker.  runner = "engine_host": the same loop over the kernels' host build (checks this worker without a GPU)."""
    import golden_cases as gc

    if runner == "engine":
        from engine_runner import run_case_engine

        def run(case):
            df, mac, _ = run_case_engine(case, chunk=9)  # batches of 9: graph replay on even starts, plain launches on odd ones
            return df, mac
    else:
        def run(case):
            return gc.run_case(case, "engine_host", fast=True, init_kind="port")

    first, changed = {}, []
    for r in range(reps):
        for n in names:
            df, mac = run(gc.BY_NAME[n])
            if n not in first:
                first[n] = (df, mac)
                np.save(os.path.join(outdir, f"df_{n}_{rank}.npy"), df)
                np.save(os.path.join(outdir, f"mac_{n}_{rank}.npy"), mac)
            elif not (np.array_equal(df, first[n][0]) and np.array_equal(mac, first[n][1])):
                changed.append((n, r))
    with open(os.path.join(outdir, f"changed_{rank}.txt"), "w") as f:
        f.write(repr(changed))


def _duct(streaming, X=16, nsteps=9):
    import golden_cases as gc
    import lbm_cases as lc
    from oracle import oracle as O

    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=streaming, X=X, Y=10, Z=9)
    return gc.Case("duct", d, O.Params(lbmViscosity=5e-3, fx=2e-5, fy=1e-6), lc.map_duct_slab_safe, nsteps, "noisy")


def _box(streaming, X=16, nsteps=10):
    import golden_cases as gc
    import lbm_cases as lc
    from oracle import oracle as O

    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=streaming, X=X, Y=12, Z=10)
    return gc.Case("box", d, O.Params(lbmViscosity=1e-3, fx=1e-6), lc.map_periodic, nsteps, "noisy")


DIST_CASES = {
    "duct_ab": lambda: _duct(0),
    "duct_aa": lambda: _duct(1),
    "box_ab": lambda: _box(0),
    "box_aa": lambda: _box(1),
}


def gather(outdir, world):
    df = np.concatenate([np.load(os.path.join(outdir, f"df_{r}.npy")) for r in range(world)], axis=1)
    mac = np.concatenate([np.load(os.path.join(outdir, f"mac_{r}.npy")) for r in range(world)], axis=1)
    return df, mac
