"""Slab decomposition on the GPU: ghost planes, boundary-planes-first ordering and the halo exchange of the engine.

* 1 GPU: a single slab with ghost planes and a periodic self-exchange must reproduce the run without ghost planes.
* >= 2 GPUs (gpurun --gpus 2): one process per GPU, the engine's own NCCL send/recv; the assembled result must equal the
  undivided oracle run within the fp64 tolerance, and be identical to the 1-slab self-exchange run of the engine."""
import socket
import tempfile

import numpy as np
import pytest

import golden_cases as gc
import lbm_cases as lc
from engine_runner import run_case_engine
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_ghost_single(case):
    """Engine, one slab, ghost planes + periodic self-exchange."""
    from engine_runner import engine_for, set_params

    d = case.desc
    port = O.Oracle(d, "port")
    df0 = gc.initial_df(case, port)
    with engine_for(case, ghost_x=1, periodic_x=1) as e:
        e.map_upload(case.make_map(d))
        e.df_upload(df0, 0)
        e.df_sync_ghosts()
        if d.streaming == O.AB:
            e.df_upload(df0, 1)
        set_params(e, case.params)
        e.macro_init()
        e.step(case.nsteps)
        e.sync()
        st = e.stats()
        assert st.halo_bytes_sent > 0
        # lbmx_halo_time repeats the last exchange alone on the communication stream: a positive device time, and the state
        # (ghost planes included) is untouched
        before = e.df_download(0, with_ghosts=True)
        assert e.halo_time(5) > 0.0
        assert np.array_equal(before, e.df_download(0, with_ghosts=True))
        return e.df_download(0), e.macro_download()


@pytest.mark.parametrize("case_name", ["duct_ab", "duct_aa", "box_ab", "box_aa"])
def test_ghost_planes_with_self_exchange_equal_plain_run(case_name):
    import dist_workers as W

    case = W.DIST_CASES[case_name]()
    plain_df, plain_mac, _ = run_case_engine(case)
    df, mac = run_ghost_single(case)
    assert np.array_equal(df, plain_df), f"max diff {np.abs(df - plain_df).max():.3e}"
    assert np.array_equal(mac, plain_mac)
    ref_df, ref_mac = gc.run_case(case, "port", nthreads=4)
    assert lc.rel_err_df(df, ref_df, case.desc) <= 1e-12


def _run_ranks(world, case_name, transport, delay=0.0):
    import torch.multiprocessing as mp

    import dist_workers as W

    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(W.nccl_engine_worker, args=(world, free_port(), case_name, tmp, transport, delay), nprocs=world, join=True)
        df, mac = W.gather(tmp, world)
        halo = np.load(f"{tmp}/halo_0.npy")
    return df, mac, halo


def _need_gpus(n):
    import torch

    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs (gpurun --gpus {n})")


@pytest.mark.parametrize("transport", ["peer_memory", "nccl"])
@pytest.mark.parametrize("case_name", ["duct_ab", "duct_aa", "box_ab", "box_aa"])
def test_two_gpus_halo_exchange(case_name, transport):
    """Two slabs, two processes, two GPUs.  Default transport: stores into the neighbour's array over NVLink (CUDA IPC peer
    mappings + arrival counters); LBMX_HALO=nccl: NCCL send/recv.  Either way identical to one slab with a self-exchange."""
    import dist_workers as W

    _need_gpus(2)
    df, mac, halo = _run_ranks(2, case_name, transport)
    case = W.DIST_CASES[case_name]()
    assert halo[0] == case.nsteps * 2 * 9 * case.desc.Y * case.desc.Z * 8, "9 populations per direction per step"
    if transport == "nccl":
        assert halo[2] == 0
    else:
        assert halo[2] == 1, "peer-memory halo exchange not available on this box (CUDA IPC between the two processes failed)"
    one_df, one_mac = run_ghost_single(case)
    assert np.array_equal(df, one_df), "2 slabs must be identical to 1 slab with self-exchange"
    assert np.array_equal(mac, one_mac)
    ref_df, ref_mac = gc.run_case(case, "port", nthreads=4)
    assert lc.rel_err_df(df, ref_df, case.desc) <= 1e-12


@pytest.mark.parametrize("transport", ["peer_memory", "nccl"])
@pytest.mark.parametrize("case_name", ["duct_aa", "duct_ab", "box_aa"])
@pytest.mark.parametrize("world", [3, 4, 8])
def test_n_gpus_halo_exchange(world, case_name, transport):
    """3 (uneven: 6 + 5 + 5 planes), 4 and 8 slabs (2 planes each: every plane is an edge plane) in a periodic ring: every rank has two
    different neighbours.  SURVEY 8d cfg 4: 8 == 4 == 2 == 1 slab with self-exchange, bit for bit."""
    import dist_workers as W

    _need_gpus(world)
    df, mac, halo = _run_ranks(world, case_name, transport)
    case = W.DIST_CASES[case_name]()
    assert halo[2] == (0 if transport == "nccl" else 1)
    one_df, one_mac = run_ghost_single(case)
    assert np.array_equal(df, one_df), f"{world} slabs must be identical to 1 slab with self-exchange"
    assert np.array_equal(mac, one_mac)


@pytest.mark.parametrize("case_name", ["duct_aa", "duct_ab"])
@pytest.mark.parametrize("world", [2, 4])
def test_late_rank_does_not_lose_halo_planes(world, case_name):
    """The last rank uploads its (ghosted) state two seconds after the others have started stepping.  The peer-memory exchange is
    one-sided: without the receiver-ready handshake the early neighbours' first planes would land in arrays the late rank then
    overwrites from the host, and step 1 would read stale ghost populations (engine.cu exchange())."""
    import dist_workers as W

    _need_gpus(world)
    df, mac, halo = _run_ranks(world, case_name, "peer_memory", delay=2.0)
    assert halo[2] == 1
    case = W.DIST_CASES[case_name]()
    one_df, one_mac = run_ghost_single(case)
    assert np.array_equal(df, one_df) and np.array_equal(mac, one_mac)
