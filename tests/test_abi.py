"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/lbmx.h declares, the host-only
helpers (slab decomposition, halo plan) behave, and the engine refuses to run without a GPU instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tnl_lbm_b200 import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "lbmx.h")).read()


def declared_symbols():
    return sorted(set(re.findall(r"^(?:const char\*|int)\s+(lbmx_\w+)\s*\(", HEADER, flags=re.M)))


def test_header_symbols_all_exported_and_bound():
    lib = B.lib()
    decl = declared_symbols()
    assert len(decl) >= 30
    for s in decl:
        assert hasattr(lib, s), f"liblbmx.so does not export {s}"
    assert sorted(B.SYMBOLS) == decl, "binding.SYMBOLS out of sync with include/lbmx.h"
    assert lib.lbmx_version() == int(re.search(r"#define LBMX_VERSION (\d+)", HEADER).group(1))


def test_no_torch_types_in_abi():
    assert "torch" not in HEADER and "at::" not in HEADER and "Tensor" not in HEADER


def test_struct_sizes_match_header_layout():
    assert C.sizeof(B.Desc) == 8 * 4 + 3 * 8 + 5 * 4 + 3 * 4
    assert C.sizeof(B.Params) == 7 * 8 + 8
    assert C.sizeof(B.Layout) == 6 * 8 + 4 * 4
    assert C.sizeof(B.HaloMsg) == 4 + 4 + 9 * 4 + 4 + 8 + 8  # 4 bytes padding before the int64 members


@pytest.mark.parametrize("X,n", [(512, 8), (2048, 8), (10, 3), (7, 7), (100, 1)])
def test_decompose_x_is_a_partition(X, n):
    parts = [B.decompose_x(X, n, r) for r in range(n)]
    assert parts[0][0] == 0
    for (o0, l0), (o1, _) in zip(parts, parts[1:]):
        assert o0 + l0 == o1
    assert parts[-1][0] + parts[-1][1] == X
    sizes = [l for _, l in parts]
    assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


def test_decompose_x_rejects_bad_arguments():
    with pytest.raises(B.LbmxError):
        B.decompose_x(4, 8, 0)
    with pytest.raises(B.LbmxError):
        B.decompose_x(16, 4, 4)


def test_halo_directions_are_the_x_movers():
    import lbm_cases as lc
    r, l = B.halo_directions(B.D3Q27)
    assert r == [q for q in range(27) if lc.C27[q][0] > 0] and len(r) == 9
    assert l == [q for q in range(27) if lc.C27[q][0] < 0] and len(l) == 9
    r2, l2 = B.halo_directions(B.D2Q9)
    assert r2 == [1, 5, 7] and l2 == [2, 6, 8]
    r3, l3 = B.halo_directions(B.D3Q19)
    assert r3 == [1, 7, 9, 11, 13] and l3 == [2, 8, 10, 12, 14]


def test_halo_plan_planes_and_slots():
    X = 6
    r, l = B.halo_directions(B.D3Q27)
    ab = B.halo_plan(B.D3Q27, B.AB, 0, X)
    assert ab[0] == dict(to_right=True, dirs=r, src_plane=X, dst_plane=0)
    assert ab[1] == dict(to_right=False, dirs=l, src_plane=1, dst_plane=X + 1)
    ev = B.halo_plan(B.D3Q27, B.AA, 4, X)  # even: opposite slots from the boundary planes into the ghost planes
    assert ev[0] == dict(to_right=True, dirs=l, src_plane=X, dst_plane=0)
    assert ev[1] == dict(to_right=False, dirs=r, src_plane=1, dst_plane=X + 1)
    od = B.halo_plan(B.D3Q27, B.AA, 5, X)  # odd: canonical slots from my ghost planes into the neighbour's boundary planes
    assert od[0] == dict(to_right=True, dirs=r, src_plane=X + 1, dst_plane=1)
    assert od[1] == dict(to_right=False, dirs=l, src_plane=0, dst_plane=X)


def test_create_rejects_unsupported_combinations():
    with pytest.raises(B.LbmxError, match="D3Q19"):
        B.Engine(lattice=B.D3Q19, coll=B.SRT, eq=B.EQ_INV_CUM)  # the product-form equilibrium needs 27 velocities
    with pytest.raises(B.LbmxError, match="no kernel family"):
        B.Engine(lattice=B.D3Q19, coll=B.CUM, eq=B.EQ_STD)
    with pytest.raises(B.LbmxError, match="Z == 1"):
        B.Engine(lattice=B.D2Q9, coll=B.SRT, eq=B.EQ_STD, Z=4)
    with pytest.raises(B.LbmxError):
        B.Engine(lattice=B.D2Q9, coll=B.CUM, eq=B.EQ_STD, Z=1)
    with pytest.raises(B.LbmxError):
        B.Engine(X=0)


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device the engine must fail loudly (LBMX_ERR_CUDA), never compute on the host."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(B.LbmxError, match="status 3"):
        B.Engine(X=4, Y=4, Z=4)


def test_product_does_not_import_the_oracle():
    """The checker is test infrastructure: the package, the ABI header, the examples and the measurement tools never name it (only
    tests/, __graft_entry__.py and bench.py's CPU-baseline legs do)."""
    for top in ("tnl_lbm_b200", "include", "examples", "tools"):
        for dirpath, dirs, files in os.walk(os.path.join(ROOT, top)):
            dirs[:] = [d for d in dirs if d not in ("bin", "build", "__pycache__")]
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c", ".sh")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "oracle" not in text.replace("no CPU fallback", ""), f"{top}/{f} mentions the oracle"


def test_engine_copies_are_ordered_on_its_own_streams():
    """The engine's streams are non-blocking: work on the legacy default stream is not ordered with them, and a plain cudaMemcpy from
    pageable memory may return before the DMA has landed.  With several processes sharing one GPU that showed up as a boundary list
    counted on a half-uploaded map.  Every copy / memset in the engine therefore names a stream."""
    text = open(os.path.join(ROOT, "tnl_lbm_b200", "csrc", "engine.cu")).read()
    code = "\n".join(line.split("//")[0] for line in text.splitlines())
    assert not re.findall(r"\bcudaMem(?:cpy|set)(?:2D|3D)?\s*\(", code), "synchronous cudaMemcpy/cudaMemset on the default stream in engine.cu"
    assert "cudaStreamNonBlocking" in code
