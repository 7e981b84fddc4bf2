"""tnl_lbm_b200 -- B200-native lattice-Boltzmann time-stepping engine (the fused collide-and-stream path of TNL-LBM).

The product is tnl_lbm_b200/liblbmx.so: hand-written CUDA for sm_100a behind the C ABI of include/lbmx.h.
`binding` is a ctypes caller used by the tests and bench.py; `build` compiles the library in-tree with nvcc.
"""
from . import binding  # noqa: F401
from .binding import Engine, LbmxError  # noqa: F401

__all__ = ["binding", "Engine", "LbmxError"]
