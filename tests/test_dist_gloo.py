"""world_size-2 run on the CPU (gloo): two processes, one oracle slab each, ghost planes exchanged through
torch.distributed following lbmx_halo_plan -- the same host logic the engine drives NCCL with.  Must equal the undivided run."""
import socket
import tempfile

import numpy as np
import pytest

import golden_cases as gc
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.available("port"), reason="oracle port not built")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kind", ["port", "engine_host"])
@pytest.mark.parametrize("case_name", ["duct_ab", "duct_aa"])
def test_two_process_gloo_slabs(case_name, kind):
    """kind = "engine_host": each process steps its slab with the engine's own CUDA kernels compiled for the host (parity arithmetic,
    tests/host_harness/engine_host.cpp) -- the kernels' ghost-plane rule and the halo plan across two processes, without a GPU."""
    import torch.multiprocessing as mp

    import dist_workers as W

    if kind == "engine_host":
        from test_kernels_on_host import _build

        _build(strict=True)
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(W.gloo_slab_worker, args=(world, free_port(), case_name, tmp, kind), nprocs=world, join=True)
        df, mac = W.gather(tmp, world)
    case = W.DIST_CASES[case_name]()
    ref_df, ref_mac = gc.run_case(case, "port")
    assert np.array_equal(df, ref_df)
    assert np.array_equal(mac, ref_mac)
