#!/usr/bin/env python
"""BASELINE.json configs[4]: "D3Q19 MRT fp32 1024^3 weak-scaling on 8xB200 (A-A vs A-B streaming)".

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_cfg5.py --gpus N

Per GPU a 128 x 1024 x 1024 slab (global 128*N x 1024 x 1024 = 1024^3 at N = 8), periodic box, fp32, the engine's D3Q19 MRT_LES
kernels (the reference has no D3Q19: parity unpinned), both streaming patterns.  One JSON line per pattern on rank 0.
Development tool: the contract benchmark is /bench.py."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--slab", type=int, default=128)
    ap.add_argument("--yz", type=int, default=1024)
    a = ap.parse_args()
    import torch

    from tnl_lbm_b200 import binding as B

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    N = a.gpus
    assert world == N
    torch.cuda.set_device(local)
    if N > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for streaming, name in ((B.AA, "A-A"), (B.AB, "A-B")):
        Xg, S = a.slab * N, a.yz
        e = B.Engine(lattice=B.D3Q19, coll=B.MRT_LES, eq=B.EQ_STD, streaming=streaming, precision=B.F32, inflow=B.INFLOW_NONE, X=Xg, Y=S, Z=S,
                     rank=rank, nranks=N, device=local, ghost_x=1 if N > 1 else 0, periodic_x=1)
        if N > 1:
            idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idbuf.copy_(torch.frombuffer(bytearray(B.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idbuf, 0)
            e.comm_init(bytes(idbuf.cpu().numpy().tobytes()))
        xl = e.layout.X_local
        e.map_upload(np.full((xl, S, S), 7, dtype=np.int16))
        e.set_equilibrium(1.0, 0.03, 0.01, -0.02)
        e.set_params(lbmViscosity=1e-3)
        e.step(a.warmup)
        e.sync()
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e.step_timed(a.steps)
        if N > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        assert not e.has_nan()
        st = e.stats()
        e.close()
        if rank == 0:
            cells = Xg * S * S
            mlups = cells * a.steps / (ms * 1e-3) / 1e6
            bpu = 19 * 2 * 4
            print(json.dumps({"config": "BASELINE.json configs[4]", "lattice": "D3Q19", "operator": "MRT_LES", "dtype": "f32", "streaming": name, "n_gpus": N,
                              "global_lattice": [Xg, S, S], "steps": a.steps, "ms_per_step": ms / a.steps, "MLUPS": mlups, "MLUPS_per_gpu": mlups / N,
                              "GBs_per_gpu_algorithmic": mlups / N * bpu / 1e3, "halo_bytes_per_step_per_gpu": st.halo_bytes_sent / max(a.steps + a.warmup, 1),
                              "parity": "unpinned (no D3Q19 in the reference)"}), flush=True)
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
