#!/usr/bin/env python
"""bench.py -- MLUPS of the fused collide-and-stream path (D3Q27 cumulant, fp64, A-A streaming) on 1..8 B200.

    python bench.py --gpus 1 --steps 200 --warmup 5                      # this engine, through the C ABI (liblbmx.so)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps K --warmup W        # the reference's CPU implementation (oracle/_ref)

Workload (BASELINE.json configs[2], SURVEY.md §8d cfg 3): periodic box of GEO_PERIODIC cells, 512^3 cells per GPU,
D3Q27_CUM + EQ_INV_CUM, fp64, A-A streaming, nu_lbm = 1e-3, body force fx = 1e-6, smooth sinusoidal initial field.
At N > 1 the box is N*512 x 512 x 512, split into x-slabs of 512 planes (weak scaling): one process per GPU, ghost planes
exchanged every step with NCCL send/recv (9 populations per direction) overlapped with the interior update.

One JSON line on stdout (rank 0).  `value` = whole-job MLUPS with the state resident in HBM (device time, CUDA events on the
engine's stream, max over ranks); `e2e` = the same job through the public C ABI with HOST buffers: upload of the cell map and
of the initial macroscopic fields from pinned host memory, all time steps, download of the macroscopic result.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q, SIZEOF, B_PER_UPDATE = 27, 8, 27 * 2 * 8  # algorithmic bytes per lattice update: 432 B


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------------------ inputs
def initial_fields(X_global, x0, xl, Y, Z):
    """SURVEY §8d cfg 3: rho = 1 + 0.01 sin(2 pi x/X), vx = 0.05 sin(2 pi y/Y), vy = 0.02 cos(2 pi z/Z), vz = 0.01 as
    float64 [x][z][y] arrays of the local slab (closed form: no RNG, identical on every platform)."""
    x = np.arange(x0, x0 + xl, dtype=np.float64)[:, None, None]
    z = np.arange(Z, dtype=np.float64)[None, :, None]
    y = np.arange(Y, dtype=np.float64)[None, None, :]
    shape = (xl, Z, Y)
    rho = np.broadcast_to(1.0 + 0.01 * np.sin(2 * np.pi * x / X_global), shape)
    vx = np.broadcast_to(0.05 * np.sin(2 * np.pi * y / Y), shape)
    vy = np.broadcast_to(0.02 * np.cos(2 * np.pi * z / Z), shape)
    return rho, vx, vy, 0.01


def pinned(shape, dtype):
    """Page-locked host buffer as a numpy array (torch is only the allocator here)."""
    import torch

    t = torch.empty(tuple(shape), dtype={np.float64: torch.float64, np.int16: torch.int16, np.float32: torch.float32}[dtype], pin_memory=True)
    return t, t.numpy()


# ------------------------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region.  Default: NVML in this process (one light query set every 50 ms from a
    thread); --clock-sampler smi: an `nvidia-smi -lms` child as in /opt/skills/guides/B200_PROFILING.md; none: no sampling."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, how: str = "nvml"):
        self.index, self.rows, self.proc, self.how, self.stop_flag, self.thread, self.smax = index, [], None, how, False, None, None

    def start(self):
        if self.how == "none":
            return
        if self.how == "nvml":
            try:
                import pynvml

                pynvml.nvmlInit()
                # CUDA_VISIBLE_DEVICES may renumber: find the handle by the PCI bus id of the CUDA device
                import torch

                bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
                h = None
                if bus is not None:
                    for i in range(pynvml.nvmlDeviceGetCount()):
                        hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                        if pynvml.nvmlDeviceGetPciInfo(hi).bus == bus:
                            h = hi
                            break
                if h is None:
                    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                    phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
                    h = pynvml.nvmlDeviceGetHandleByIndex(phys)
                self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
                bits = [(pynvml.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"), (pynvml.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                        (pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"), (pynvml.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap")]

                def pump():
                    while not self.stop_flag:
                        try:
                            sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                            mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                            self.rows.append((time.perf_counter(), sm, [n for b, n in bits if mask & b]))
                        except Exception:
                            pass
                        time.sleep(0.05)

                self.thread = threading.Thread(target=pump, daemon=True)
                self.thread.start()
                return
            except Exception:
                self.how = "smi"  # fall through to the child process
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.how == "none":
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled (--clock-sampler none)"]}
        if self.how == "nvml":
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=1.0)
            sm, reasons = [], set()
            for t, mhz, why in self.rows:
                if t0 <= t <= t1 + 0.08:
                    sm.append(mhz)
                    reasons.update(why)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(reasons), "samples": len(sm), "how": "NVML, 50 ms"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            if not (t0 <= t <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except Exception:
                continue
            for n, v in zip(self.NAMES, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm), "how": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(size: int, steps: int, warmup: int, threads: int, streaming_aa: bool = True):
    """The reference's own per-cell code on the host cores (oracle/_ref, `fast` build), else the C++ restatement.
    Returns (mlups, kind, ms_per_step)."""
    from oracle import oracle as O

    st = O.AA if streaming_aa else O.AB
    kind = "reference" if O.available("reference", st, fast=True) else "port"
    if kind == "port" and not O.available("port", fast=True):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, stdout=subprocess.DEVNULL)
    d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=st, precision=O.F64, X=size, Y=size, Z=size)
    orc = O.Oracle(d, kind, fast=True)
    df = d.new_df()
    rho, vx, vy, vz = initial_fields(size, 0, size, size, size)
    orc.set_equilibrium_field(df, np.ascontiguousarray(rho), np.ascontiguousarray(vx), np.ascontiguousarray(vy), np.full(rho.shape, vz))
    other = df if streaming_aa else df.copy()
    mac = d.new_macro()
    m = d.new_map(7)  # GEO_PERIODIC
    p = O.Params(lbmViscosity=1e-3, fx=1e-6)
    if warmup:
        orc.step(p, df, other, mac, m, 0, warmup, threads)
    t0 = time.perf_counter()
    orc.step(p, df, other, mac, m, warmup, steps, threads)
    dt = time.perf_counter() - t0
    assert np.isfinite(mac).all()
    return size ** 3 * steps / dt / 1e6, kind, dt / steps * 1e3


def macro_every_step_run(a):
    """Re-run the device-resident measurement with LBMX_MACRO_EVERY_STEP in a child process; returns a small dict for the JSON line."""
    steps = int(min(a.steps, 50))
    cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1", "--steps", str(steps), "--warmup", "3", "--size", str(a.size), "--streaming", a.streaming,
           "--macro-policy", "every", "--no-cpu-baseline", "--no-extras"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
        rows = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not rows:
            return {"error": f"child exited {r.returncode}: {r.stderr.strip()[-300:]}"}
        child = json.loads(rows[-1])
        bytes_per_update = B_PER_UPDATE + 4 * SIZEOF + 2  # populations + rho,u + the cell type
        return {"value": child["value"], "unit": "MLUPS", "steps": steps, "ms_per_step": child["ms_per_step"], "clocks": child.get("clocks"),
                "bytes_per_update_incl_macro_and_map": bytes_per_update, "GBs_incl_macro_and_map": child["value"] * 1e6 * bytes_per_update / 1e9,
                "what": "same workload, rho and u written by every step as the reference kernel does (lbmx_desc.macro_policy = LBMX_MACRO_EVERY_STEP)"}
    except Exception as ex:
        return {"error": repr(ex)}


def dropin_mirror_run(a):
    """The same workload as an ordinary solver written against the host mirror of the reference's interface (examples/box3d.cpp:
    LBM_CONFIG + State<NSE> + execute(), nothing batch-aware in it), in a child process; the figure is the mirror's own GLUPS= line
    (State::AfterSimUpdate, state.hpp:1244-1262).  VERDICT r1 'next' #2: an unmodified solver has to reach the headline rate."""
    exe = os.path.join(ROOT, "examples", "bin", "box3d_aa" if a.streaming == "AA" else "box3d")
    if not os.path.exists(exe):
        return {"error": "examples/bin/box3d[_aa] not built (python -m tnl_lbm_b200.build_examples)"}
    import re
    import tempfile

    steps, period = 330, 110
    out = {}
    try:
        for mode in ("1", "0"):
            with tempfile.TemporaryDirectory() as tmp:
                n = steps if mode == "1" else 120
                r = subprocess.run([exe, str(a.size), str(a.size), str(a.size), str(n), str(period if mode == "1" else 60)], capture_output=True, text=True, timeout=420, cwd=tmp,
                                   env=dict(os.environ, LBMX_HOST_BATCH=mode))
            g = [float(x) for x in re.findall(r"GLUPS=([0-9.]+)", r.stdout)]
            b = re.search(r"lbmx: (\d+) steps enqueued in (\d+) batches", r.stdout)
            if r.returncode != 0 or not g or not b:
                return {"error": f"child exited {r.returncode}: {(r.stdout + r.stderr).strip()[-300:]}"}
            out[mode] = {"MLUPS": g[-1] * 1e3, "all_GLUPS_lines": g, "steps": int(b.group(1)), "batches": int(b.group(2))}
        return {"value": out["1"]["MLUPS"], "unit": "MLUPS", "steps": out["1"]["steps"], "batches": out["1"]["batches"], "glups_lines": out["1"]["all_GLUPS_lines"],
                "one_launch_per_SimUpdate": {"value": out["0"]["MLUPS"], "steps": out["0"]["steps"], "batches": out["0"]["batches"], "how": "LBMX_HOST_BATCH=0"},
                "what": f"examples/box3d.cpp ({a.size}^3 periodic, D3Q27 CUM fp64, {'A-A' if a.streaming == 'AA' else 'A-B'}) through State<NSE>/execute() of the host mirror; "
                        f"last GLUPS= line of the run (interval of {period} steps, includes the NaN scan of AfterSimUpdate)"}
    except Exception as ex:
        return {"error": repr(ex)}


# ------------------------------------------------------------------------------------------------------------------ multi-GPU parity leg
def duct_map(Xg, x0, xl, S, out=None):
    """sim_NSE/sim_2.cu:125-138, painted in that order: global x planes 0 and X-1 GEO_PERIODIC, GEO_WALL at y,z = 1 / N-2, GEO_NOTHING
    at y,z = 0 / N-1 (cell types d3q27/bc.h:17-34: 0 fluid, 1 wall, 7 periodic, 8 nothing); local planes [x0, x0 + xl)."""
    m = out if out is not None else np.empty((xl, S, S), dtype=np.int16)
    m[...] = 0
    if x0 == 0:
        m[0] = 7
    if x0 + xl == Xg:
        m[xl - 1] = 7
    m[:, 1, :] = 1
    m[:, S - 2, :] = 1
    m[:, :, 1] = 1
    m[:, :, S - 2] = 1
    m[:, 0, :] = 8
    m[:, S - 1, :] = 8
    m[:, :, 0] = 8
    m[:, :, S - 1] = 8
    return m


def halo_parity(rank, N, local_rank, dist, torch, B, threads):
    """CHECKER LEG, before any timing (SURVEY 8d cfg 4): the 256 x 64 x 64 miniature of the duct, 12 steps, A-A and A-B,
    over the N ranks with each halo transport (peer-memory stores over NVLink, NCCL send/recv).  The slabs are gathered on rank 0 and must
    be BIT-IDENTICAL to one slab with ghost planes and a self-exchange on rank 0's GPU, which in turn must agree to 1e-12 with the CPU
    checker (oracle port under the reference's ghost-plane index rule, kernels.h:39-48, planes exchanged by hand as lbmx_halo_plan says).
    Replaces nothing in the reference -- it pins what lbm.hpp:196-280 / lbm_block.hpp:428-442 do through TNL's synchroniser."""
    X, S, STEPS = (256 if 256 % N == 0 else 32 * N), 64, 12
    prm = dict(lbmViscosity=1e-3, fx=1e-5, fy=2e-6, fz=-1e-6)

    def fields(x0, xl):
        rho, vx, vy, vz = initial_fields(X, x0, xl, S, S)
        return [np.ascontiguousarray(f) for f in (rho, vx, vy)] + [np.full((xl, S, S), vz)]

    def run_engine(streaming, nranks, r, comm):
        e = B.Engine(lattice=B.D3Q27, coll=B.CUM, eq=B.EQ_INV_CUM, streaming=streaming, macro=B.MACRO_DEFAULT, inflow=B.INFLOW_NONE, precision=B.F64, X=X, Y=S, Z=S,
                     rank=r, nranks=nranks, device=local_rank, ghost_x=1, periodic_x=1, macro_policy=B.MACRO_LAST_STEP)
        try:
            if nranks > 1:
                idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
                if r == 0:
                    idbuf.copy_(torch.frombuffer(bytearray(B.comm_unique_id()), dtype=torch.uint8))
                dist.broadcast(idbuf, 0)
                e.comm_init(bytes(idbuf.cpu().numpy().tobytes()))
            x0, xl = e.layout.x_offset, e.layout.X_local
            e.map_upload(duct_map(X, x0, xl, S))
            e.set_equilibrium_field(*fields(x0, xl))
            e.set_params(**prm)
            e.step(STEPS)
            e.sync()
            return e.df_download(0), e.macro_download(), e.stats().halo_peer_memory
        finally:
            e.close()

    cases, worst = [], 0.0
    for streaming, sname in ((B.AA, "AA"), (B.AB, "AB")):
        one = None
        for transport in ("peer_memory", "nccl"):
            if transport == "nccl":
                os.environ["LBMX_HALO"] = "nccl"
            else:
                os.environ.pop("LBMX_HALO", None)
            df, mac, peer = run_engine(streaming, N, rank, True)
            os.environ.pop("LBMX_HALO", None)
            parts = [torch.empty_like(torch.from_numpy(df).cuda()) for _ in range(N)] if rank == 0 else None
            dist.gather(torch.from_numpy(df).cuda(), parts, dst=0)
            mparts = [torch.empty_like(torch.from_numpy(mac).cuda()) for _ in range(N)] if rank == 0 else None
            dist.gather(torch.from_numpy(mac).cuda(), mparts, dst=0)
            if rank == 0:
                df_all = np.concatenate([t.cpu().numpy() for t in parts], axis=1)
                mac_all = np.concatenate([t.cpu().numpy() for t in mparts], axis=1)
                if one is None:
                    one = run_engine(streaming, 1, 0, False)
                same = bool(np.array_equal(df_all, one[0]) and np.array_equal(mac_all, one[1]))
                cases.append({"streaming": sname, "transport": transport if peer or transport == "nccl" else "nccl (peer memory unavailable)", "bit_identical_to_one_slab": same})
                del parts, mparts, df_all, mac_all
        if rank == 0:
            # the CPU checker on the same inputs: one ghosted slab under the nproc > 1 index rule, ghost planes exchanged by hand
            from oracle import oracle as O

            if not O.available("port"):
                subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, stdout=subprocess.DEVNULL)
            d = O.Desc(coll=O.CUM, eq=O.EQ_INV_CUM, streaming=O.AA if streaming == B.AA else O.AB, precision=O.F64, X=X, Y=S, Z=S, ox=1, nproc=2)
            orc = O.Oracle(d, "port")
            idx = np.arange(-1, X + 1) % X
            f = fields(0, X)
            a = d.new_df()
            orc.set_equilibrium_field(a, *[np.ascontiguousarray(v[idx]) for v in f])
            b = a.copy()
            m = np.ascontiguousarray(duct_map(X, 0, X, S)[idx])
            mac = d.new_macro()
            p = O.Params(**prm)
            aa = streaming == B.AA
            for it in range(STEPS):
                orc.step(p, a, b, mac, m, it, 1, threads)
                arr = a if (aa or it % 2 == 1) else b  # A-B: even iterations write df[1] (lbm.hpp:320-327)
                for msg in B.halo_plan(B.D3Q27, streaming, it, X):
                    arr[msg["dirs"], msg["dst_plane"]] = arr[msg["dirs"], msg["src_plane"]]
            cur = a if (aa or STEPS % 2 == 0) else b
            ref = cur[:, 1:-1]
            w = np.array([8 / 27] + [2 / 27] * 6 + [1 / 54] * 12 + [1 / 216] * 8).reshape(27, 1, 1, 1)
            err = float((np.abs(one[0] - ref) / np.maximum(np.abs(ref), w / 2)).max())
            worst = max(worst, err)
            for c in cases:
                if c["streaming"] == sname:
                    c["vs_oracle"] = err
    out = None
    if rank == 0:
        ok = all(c["bit_identical_to_one_slab"] for c in cases)
        out = {"slabs": N, "lattice": [X, S, S], "steps": STEPS, "bit_identical": ok, "vs_oracle": worst, "tolerance": 1e-12, "cases": cases,
               "what": "N ghosted slabs (one process per GPU, both halo transports) == one slab with self-exchange, bit for bit; that run vs the CPU checker under the ghost-plane rule"}
        assert ok, f"multi-GPU halo exchange is NOT bit-identical to the single-slab run: {cases}"
        assert worst <= 1e-12, f"single-slab run differs from the CPU checker by {worst:.3e}"
    dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------------------------ strong-scaling leg
def channel_strong(rank, N, local_rank, dist, torch, B, S, steps=10, warmup=3):
    """BASELINE.json configs[3] inside every default run (VERDICT r1 'next' #5): the 4S x S x S body-force duct of sim_NSE/sim_2.cu
    (2048 x 512 x 512 at S = 512; 116 GB of distributions + 17 GB of rho,u: one B200 holds it) stepped `steps` times
      (1) on rank 0's GPU alone, as ONE slab with ghost planes and a self-exchange, and
      (2) split into N x-slabs over the N ranks (default halo transport),
    both timed with CUDA events on the engines' compute streams, (2) as the max over ranks.  Strong scaling: efficiency = t1 / (N tN)."""
    Xg = 4 * S

    def run(nranks, r):
        e = B.Engine(lattice=B.D3Q27, coll=B.CUM, eq=B.EQ_INV_CUM, streaming=B.AA, macro=B.MACRO_DEFAULT, inflow=B.INFLOW_NONE, precision=B.F64, X=Xg, Y=S, Z=S,
                     rank=r, nranks=nranks, device=local_rank, ghost_x=1, periodic_x=1, macro_policy=B.MACRO_LAST_STEP)
        try:
            if nranks > 1:
                idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
                if r == 0:
                    idbuf.copy_(torch.frombuffer(bytearray(B.comm_unique_id()), dtype=torch.uint8))
                dist.broadcast(idbuf, 0)
                e.comm_init(bytes(idbuf.cpu().numpy().tobytes()))
            x0, xl = e.layout.x_offset, e.layout.X_local
            e.map_upload(duct_map(Xg, x0, xl, S))
            e.set_equilibrium(1.0, 0.0, 0.0, 0.0)
            e.set_params(lbmViscosity=1e-3, fx=1e-7, fy=0.0, fz=0.0)
            e.step(warmup)
            e.sync()
            if nranks > 1:
                dist.barrier()
            ms = e.step_timed(steps)
            if nranks > 1:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            assert not e.has_nan()
            return ms / steps
        finally:
            e.close()

    cells = Xg * S * S
    out = {"lattice": [Xg, S, S], "steps": steps, "warmup": warmup, "unit": "MLUPS",
           "what": "D3Q27 cumulant fp64 A-A duct of sim_NSE/sim_2.cu:125-138 (periodic x, GEO_WALL ring behind a GEO_NOTHING shell), fixed global lattice"}
    one = None
    if rank == 0:
        try:
            one = run(1, 0)
            out["n1"] = {"ms_per_step": one, "value": cells / (one * 1e-3) / 1e6, "how": "one slab with ghost planes and a self-exchange on rank 0's GPU"}
        except Exception as ex:  # e.g. a GPU with less memory
            out["n1"] = {"error": repr(ex)[:300]}
    if N > 1:
        dist.barrier()
        msN = run(N, rank)
        out["value"] = cells / (msN * 1e-3) / 1e6
        out["ms_per_step"] = msN
        out["n_gpus"] = N
        out["efficiency_vs_n1"] = (one / (N * msN)) if one else None
    elif one:
        out["value"], out["ms_per_step"], out["n_gpus"], out["efficiency_vs_n1"] = out["n1"]["value"], one, 1, 1.0
    return out


# ------------------------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="lbmx", choices=["lbmx", "reference"])
    ap.add_argument("--size", type=int, default=512, help="cells per GPU along each axis (default: the 512^3 configuration)")
    ap.add_argument("--cpu-sample", type=int, default=128, help="edge of the periodic sub-box the CPU arm times")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streaming", default="AA", choices=["AA", "AB"])
    ap.add_argument("--macro-policy", default="last", choices=["last", "every"],
                    help="last: rho,u written by the last step of a batch (default; identical values wherever the host can observe them); "
                         "every: written by every step, as the reference kernel does (+32 B per update)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements that start child processes")
    ap.add_argument("--no-channel-strong", action="store_true", help="skip the strong-scaling leg (BASELINE.json configs[3]) that follows the headline measurement")
    ap.add_argument("--no-halo-parity", action="store_true", help="skip the multi-GPU bit-identity leg that precedes the timing at N > 1")
    ap.add_argument("--clock-sampler", default="nvml", choices=["nvml", "smi", "none"], help="how SM clocks / throttle reasons are sampled during the timed region")
    ap.add_argument("--workload", default="box", choices=["box", "channel"],
                    help="box: periodic 512^3 per GPU, weak scaling (the headline metric); channel: BASELINE.json configs[3], the 2048x512x512 "
                         "body-force duct of sim_NSE/sim_2.cu split into N x-slabs, strong scaling (N >= 2)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = a.gpus
    host_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    channel = a.workload == "channel"
    Xg = 4 * a.size if channel else a.size * N
    if channel:
        workload = (f"D3Q27 cumulant (EQ_INV_CUM) fp64 A-{'A' if a.streaming == 'AA' else 'B'} channel {Xg}x{a.size}x{a.size} (sim_2 duct: periodic x, "
                    f"GEO_WALL ring behind a GEO_NOTHING shell), nu=1e-3, fx=1e-7, split into {N} x-slabs")
    else:
        workload = f"D3Q27 cumulant (EQ_INV_CUM) fp64 A-{'A' if a.streaming == 'AA' else 'B'} periodic box, {a.size}^3 cells per GPU, nu=1e-3, fx=1e-6"
    config = {"workload": workload, "global_lattice": [Xg, a.size, a.size], "decomposition": f"x-slabs x{N}", "l2_policy": "working set (29 GB/GPU) larger than L2; no flush needed",
              "macro_policy": ("written by the last step of the batch (values identical at every host-observable point)" if a.macro_policy == "last"
                               else "written by every step, as the reference kernel does (d3q27/macro.h:64-71)")}

    # ---------------------------------------------------------------- reference arm: the CPU implementation, rank 0 only
    if a.impl == "reference":
        if rank != 0:
            return
        mlups, kind, ms = cpu_reference_run(a.cpu_sample, a.steps, a.warmup, host_threads, a.streaming == "AA")
        sample = f"{a.cpu_sample}^3 periodic sub-box of the same field and operator, {a.steps} steps after {a.warmup} warm-up, OpenMP collapse(2) over (x,z) as state.hpp:1116"
        line = {"impl": "reference", "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": N, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": host_threads, "kind": kind, "sample": sample},
                "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------- this engine
    import torch  # plumbing only: pinned host memory, process group

    from tnl_lbm_b200 import binding as B

    assert world == N, f"--gpus {N} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {N}"
    torch.cuda.set_device(local_rank)
    if N > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v: float) -> float:
        if N == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    S = a.size
    parity = None
    if N > 1 and not a.no_halo_parity:
        t_par = time.perf_counter()
        parity = halo_parity(rank, N, local_rank, dist, torch, B, host_threads)
        log(f"[rank {rank}] halo parity leg: {time.perf_counter() - t_par:.1f}s {'' if parity is None else json.dumps(parity['cases'])}")
    streaming = B.AA if a.streaming == "AA" else B.AB
    eng = B.Engine(lattice=B.D3Q27, coll=B.CUM, eq=B.EQ_INV_CUM, streaming=streaming, macro=B.MACRO_DEFAULT, inflow=B.INFLOW_NONE, precision=B.F64,
                   X=Xg, Y=S, Z=S, rank=rank, nranks=N, device=local_rank, ghost_x=1 if (N > 1 or channel) else 0, periodic_x=1,
                   macro_policy=B.MACRO_LAST_STEP if a.macro_policy == "last" else B.MACRO_EVERY_STEP)
    if N > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idbuf.copy_(torch.frombuffer(bytearray(B.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        eng.comm_init(bytes(idbuf.cpu().numpy().tobytes()))
    lay = eng.layout
    xl, x0 = lay.X_local, lay.x_offset
    cells_local = xl * S * S
    cells_global = Xg * S * S

    # host-side inputs in pinned memory (created on the host, as a solver's setupBoundaries()/initial condition would)
    t_in = time.perf_counter()
    keep = []
    tm, h_map = pinned((xl, S, S), np.int16)
    keep.append(tm)
    fields = []
    if channel:
        duct_map(Xg, x0, xl, S, out=h_map)
    else:
        h_map[...] = 7  # GEO_PERIODIC (d3q27/bc.h:25)
        rho, vx, vy, vz = initial_fields(Xg, x0, xl, S, S)
        for src in (rho, vx, vy, vz):
            t, arr = pinned((xl, S, S), np.float64)
            arr[...] = src
            keep.append(t)
            fields.append(arr)
    tmac, h_mac = pinned(eng.macro_shape(), np.float64)
    keep.append(tmac)
    log(f"[rank {rank}] host inputs ready in {time.perf_counter() - t_in:.1f}s; slab x0={x0} xl={xl}")

    upload_log = []

    def upload_state():
        t0 = time.perf_counter()
        eng.map_upload(h_map)
        t1 = time.perf_counter()
        if channel:
            eng.set_equilibrium(1.0, 0.0, 0.0, 0.0)  # State::resetDFs (state.hpp:880-896)
        else:
            eng.set_equilibrium_field(*fields)
        eng.iterations = 0
        upload_log.append((t1 - t0, time.perf_counter() - t1))

    eng.set_params(lbmViscosity=1e-3, fx=1e-7 if channel else 1e-6, fy=0.0, fz=0.0)

    # ---- device-resident measurement: W warm-up steps, then exactly K timed steps
    upload_state()
    eng.step(a.warmup)
    eng.sync()
    barrier()
    sampler = ClockSampler(local_rank, a.clock_sampler)
    sampler.start()
    time.sleep(0.25)
    s0 = eng.stats()
    l0, hb0 = s0.kernel_launches, s0.halo_bytes_sent
    barrier()
    t0 = time.perf_counter()
    ms = eng.step_timed(a.steps)  # CUDA events on the engine's compute stream, which joins the edge and comm streams
    barrier()
    t1 = time.perf_counter()
    s1 = eng.stats()
    launches, halo_bytes = s1.kernel_launches - l0, s1.halo_bytes_sent - hb0
    clocks = sampler.stop(t0, t1)
    ms_max = allmax(ms)
    value = cells_global * a.steps / (ms_max * 1e-3) / 1e6
    assert not eng.has_nan(), "NaN in the density field after the timed steps"
    # the halo exchange of one step, alone on the communication stream (inside a step it runs concurrently with the interior update)
    halo_ms = None
    if halo_bytes > 0:
        barrier()
        halo_ms = allmax(eng.halo_time(20))

    # ---- end-to-end through the C ABI with host buffers: upload map + initial fields, K steps, download macros
    barrier()
    te0 = time.perf_counter()
    upload_state()
    te_up = time.perf_counter()
    eng.step(a.steps)
    eng.sync()
    te_st = time.perf_counter()
    eng.macro_download(out=h_mac)
    eng.sync()
    torch.cuda.synchronize()
    te1 = time.perf_counter()
    log(f"[rank {rank}] e2e phases: upload {te_up - te0:.3f}s (map + boundary list {upload_log[-1][0]:.3f}s, initial fields {upload_log[-1][1]:.3f}s), "
        f"{a.steps} steps {te_st - te_up:.3f}s, macro download {te1 - te_st:.3f}s")
    te = allmax(te1 - te0)
    e2e_value = cells_global * a.steps / te / 1e6
    # the three phases as the slowest rank saw them, and the host-link rates they imply (all ranks copy at the same time: GPUs that share a
    # PCIe switch uplink or a memory controller share its bandwidth, which is what bounds e2e at N = 8)
    ph_up, ph_st, ph_dn = allmax(te_up - te0), allmax(te_st - te_up), allmax(te1 - te_st)
    h2d = (h_map.nbytes + sum(f.nbytes for f in fields)) * N
    d2h = h_mac.nbytes * N
    rho_sum = float(h_mac[0].sum())
    if N > 1:
        t = torch.tensor([rho_sum], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        rho_sum = float(t.item())
    assert abs(rho_sum / cells_global - 1.0) < 1e-6, rho_sum / cells_global  # mass is conserved (wall / NOTHING cells report rho = 1)

    st = eng.stats()
    line = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        # the bulk kernel is the only kernel of a step at N=1 (at N>1: + 2 edge launches and the exchange), so its average
        # duration is the event time per step; per launch it processes cells_local updates of 432 algorithmic bytes each
        achieved = cells_local * B_PER_UPDATE / (ms_max * 1e-3 / a.steps) / 1e9
        # DRAM bytes per launch from the committed ncu capture (dram__bytes_read.sum + dram__bytes_write.sum of k_bulk), which was
        # taken at this workload's size (512^3 per launch); for other sizes it is scaled by the cells per launch
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("streaming") == a.streaming and not channel:
                    traffic = tj["dram_bytes_per_update"] * cells_local
                    traffic_note = f"{tj['dram_bytes_per_update']:.1f} B per update ({tj['source']}) x {cells_local} updates per launch"
            except Exception:
                pass
        line = {"metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": N, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_max / a.steps,
                "higher_is_better": True, "scaling": "strong" if channel else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "MLUPS", "h2d_bytes_per_step": h2d / a.steps, "d2h_bytes_per_step": d2h / a.steps,
                        "what": "lbmx_map_upload + lbmx_df_set_equilibrium_field from pinned host buffers, lbmx_step(K), lbmx_macro_download",
                        "phases_s_slowest_rank": {"upload": ph_up, "steps": ph_st, "download": ph_dn},
                        "host_link_GBs_per_gpu": {"h2d": h2d / N / ph_up / 1e9, "d2h": d2h / N / ph_dn / 1e9},
                        "host_link_GBs_all_gpus": {"h2d": h2d / ph_up / 1e9, "d2h": d2h / ph_dn / 1e9}},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                             "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0, "algorithmic_bytes_per_update": B_PER_UPDATE,
                             "kernel": f"k_bulk<D3Q27,CUM,double,{'A-A even/odd' if a.streaming == 'AA' else 'A-B'}>", "registers": st.bulk_regs, "block": st.bulk_block},
                "halo_bytes_per_step_per_gpu": halo_bytes / a.steps, "halo_parity": parity}
        if halo_ms is not None:
            hb = halo_bytes / a.steps
            line["halo"] = {"bytes_per_step_per_gpu": hb, "exchange_ms_alone": halo_ms, "nvlink_bound_ms": hb / 900e9 * 1e3,
                            "nvlink_peak": "900 GB/s per direction (NVLink 5); each GPU sends and receives bytes_per_step_per_gpu per step",
                            "step_ms": ms_max / a.steps, "hidden_behind_interior": bool(halo_ms < ms_max / a.steps),
                            "transport": ("stores into the neighbour's planes over NVLink (CUDA IPC peer mapping) + arrival counter" if st.halo_peer_memory
                                          else "NCCL send/recv, 9 plane messages per direction in one group") if N > 1 else "device copy kernel (single slab with ghost planes)"}
    eng.close()
    del keep, h_mac, h_map, fields
    # ---- strong scaling of the channel (configs[3]) at this N against one GPU, same run
    if not channel and not a.no_channel_strong and not a.no_extras and a.streaming == "AA" and a.macro_policy == "last":
        try:
            cs = channel_strong(rank, N, local_rank, dist if N > 1 else None, torch, B, S)
        except Exception as ex:
            cs = {"error": repr(ex)[:300]}
            if N > 1:
                raise
        if rank == 0:
            line["channel_strong"] = cs
            log(f"[rank 0] channel_strong: {json.dumps(cs)}")

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's CPU code on a bounded sample of the same workload
    if rank == 0 and N == 1 and not a.no_cpu_baseline:
        try:
            probe, kind, _ = cpu_reference_run(a.cpu_sample, 2, 1, host_threads, a.streaming == "AA")
            nsteps = int(min(max(15.0 * probe * 1e6 / a.cpu_sample ** 3, 4), 400))  # about 15 s of CPU work
            mlups, kind, _ = cpu_reference_run(a.cpu_sample, nsteps, 1, host_threads, a.streaming == "AA")
            line["cpu_baseline"] = {"value": mlups, "unit": "MLUPS", "cores": host_threads, "kind": kind,
                                    "sample": f"{a.cpu_sample}^3 periodic sub-box of the same field and operator, {nsteps} steps, OpenMP over (x,z)"}
        except Exception as ex:  # the checker being unavailable must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "MLUPS", "cores": host_threads, "kind": "unavailable", "sample": repr(ex)}
    # ---- GPU baseline beside it (rank 0, N = 1): the reference's own cudaLBMKernel recompiled for sm_100a through the TNL
    #      stand-in (oracle/ref_gpu_bench.cu -> oracle/_ref/, built where the reference tree exists), same box, same lattice
    if rank == 0 and N == 1 and not a.no_cpu_baseline and not channel:
        exe = os.path.join(ROOT, "oracle", "_ref", f"ref_gpu_bench_{'aa' if a.streaming == 'AA' else 'ab'}")
        if os.path.exists(exe):
            try:
                r = subprocess.run([exe, str(S), "20"], capture_output=True, text=True, timeout=300)
                line["reference_gpu_kernel"] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as ex:
                line["reference_gpu_kernel"] = {"error": repr(ex)}
    # ---- the same workload with the macroscopic fields written by EVERY step (the reference's behaviour), in a child process after this
    #      engine has released its memory: a failure or time-out there cannot take the line above with it
    if rank == 0 and N == 1 and not channel and a.macro_policy == "last" and not a.no_extras and not a.no_cpu_baseline:
        line["macro_every_step"] = macro_every_step_run(a)
        line["dropin_mirror"] = dropin_mirror_run(a)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
