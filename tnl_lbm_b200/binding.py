"""ctypes binding of liblbmx.so (include/lbmx.h) for the tests and bench.py.

The product is the shared library; this module is only the thinnest possible caller: plain pointers and sizes in, status
codes out.  There is no fallback of any kind: if the library is missing, or there is no CUDA device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblbmx.so")

# selectors (include/lbmx.h)
D3Q27, D2Q9, D3Q19 = 0, 1, 2
CUM, SRT, BGK, MRT_LES, CLBM, SRT_MODIF_FORCE = 0, 1, 2, 3, 4, 5
CUM_2017, CUM_ANTIALIAS, CUM_2017_ANTIALIAS = 10, 11, 12  # D3Q27_CUM built with the switches of defs.h:254-255
KBC_N1, KBC_N2, KBC_N3, KBC_N4, KBC_C1, KBC_C2, KBC_C3, KBC_C4 = range(13, 21)
BGK_GALILEAN = 21  # D3Q27_BGK built with USE_GALILEAN_CORRECTION (defs.h:253)
CUM_HP_RHO = 22  # D3Q27_CUM built with USE_HIGH_PRECISION_RHO (defs.h:252)
EQ_STD, EQ_INV_CUM, EQ_ENTROPIC = 0, 1, 3
AB, AA = 0, 1
MACRO_VOID, MACRO_DEFAULT, MACRO_MEAN, MACRO_WITH_MEAN_2D = 0, 1, 2, 3
GATE_MEANS, GATE_FLUCS = 1, 2
INFLOW_NONE, INFLOW_CONST, INFLOW_PROFILE_YZ, INFLOW_PARABOLIC_Y = 0, 1, 2, 3
F32, F64 = 0, 1
MACRO_EVERY_STEP, MACRO_LAST_STEP, MACRO_NEVER = 0, 1, 2
FLAG_STRICT_ARITH = 1


class LbmxError(RuntimeError):
    pass


class Desc(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("lattice", "coll", "eq", "streaming", "macro", "inflow", "precision", "macro_policy")]
        + [(n, C.c_int64) for n in ("X", "Y", "Z")]
        + [(n, C.c_int32) for n in ("rank", "nranks", "device", "ghost_x", "periodic_x", "flags")]
        + [("reserved", C.c_int32 * 2)]
    )


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("lbmViscosity", "fx", "fy", "fz", "inflow_vx", "inflow_vy", "inflow_vz")] + [("stat_counter", C.c_int32), ("macro_gates", C.c_int32)]


class Layout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("X_local", "Y", "Z", "x_offset", "ghost_x", "XYZ")] + [(n, C.c_int32) for n in ("Q", "n_macro", "sizeof_real", "dfmax")]


class Ptrs(C.Structure):
    _fields_ = [("dfs", C.c_void_p * 2), ("dmacro", C.c_void_p), ("dmap", C.c_void_p), ("even_iter", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("kernel_launches", "halo_bytes_sent", "boundary_cells", "bulk_cells")] + [(n, C.c_int32) for n in ("bulk_regs", "boundary_regs", "bulk_block", "halo_peer_memory")] + [
        ("aa_cells_reaching_outside", C.c_int64), ("tma_launches", C.c_int64)]


class HaloMsg(C.Structure):
    _fields_ = [("to_right", C.c_int32), ("n_dirs", C.c_int32), ("dirs", C.c_int32 * 9), ("src_plane", C.c_int64), ("dst_plane", C.c_int64)]


# every symbol include/lbmx.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = [
    "lbmx_last_error", "lbmx_version", "lbmx_decompose_x", "lbmx_halo_directions", "lbmx_halo_plan", "lbmx_create", "lbmx_destroy", "lbmx_get_layout",
    "lbmx_device_count", "lbmx_comm_unique_id", "lbmx_comm_init", "lbmx_map_upload", "lbmx_map_download", "lbmx_df_set_equilibrium", "lbmx_df_set_equilibrium_field",
    "lbmx_df_upload", "lbmx_df_download", "lbmx_df_sync_ghosts", "lbmx_macro_init", "lbmx_macro_download", "lbmx_macro_upload", "lbmx_set_params",
    "lbmx_set_inflow_profile", "lbmx_bouzidi_upload", "lbmx_step", "lbmx_sync", "lbmx_step_timed", "lbmx_halo_time", "lbmx_get_iterations", "lbmx_set_iterations", "lbmx_has_nan",
    "lbmx_get_device_ptrs", "lbmx_get_stats",
]

_lib = None


def lib():
    """Load liblbmx.so; raises if it has not been built (python -m tnl_lbm_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LbmxError(f"{LIB_PATH} is missing: build it with `python -m tnl_lbm_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.lbmx_last_error.restype = C.c_char_p
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.lbmx_decompose_x.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
        L.lbmx_halo_directions.argtypes = [i32, C.POINTER(i32), C.POINTER(i32)]
        L.lbmx_halo_plan.argtypes = [i32, i32, i64, i64, C.POINTER(HaloMsg)]
        L.lbmx_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
        L.lbmx_destroy.argtypes = [vp]
        L.lbmx_get_layout.argtypes = [vp, C.POINTER(Layout)]
        L.lbmx_comm_unique_id.argtypes = [vp]
        L.lbmx_comm_init.argtypes = [vp, vp]
        L.lbmx_map_upload.argtypes = [vp, vp, C.c_int]
        L.lbmx_map_download.argtypes = [vp, vp, C.c_int]
        L.lbmx_df_set_equilibrium.argtypes = [vp] + [C.c_double] * 4
        L.lbmx_df_set_equilibrium_field.argtypes = [vp] + [vp] * 4
        L.lbmx_df_upload.argtypes = [vp, C.c_int, vp, C.c_int]
        L.lbmx_df_download.argtypes = [vp, C.c_int, vp, C.c_int]
        L.lbmx_df_sync_ghosts.argtypes = [vp]
        L.lbmx_macro_init.argtypes = [vp]
        L.lbmx_macro_download.argtypes = [vp, vp, C.c_int]
        L.lbmx_macro_upload.argtypes = [vp, vp, C.c_int]
        L.lbmx_set_params.argtypes = [vp, C.POINTER(Params)]
        L.lbmx_set_inflow_profile.argtypes = [vp, vp, i64, i64]
        L.lbmx_bouzidi_upload.argtypes = [vp, vp]
        L.lbmx_step.argtypes = [vp, i64]
        L.lbmx_sync.argtypes = [vp]
        L.lbmx_step_timed.argtypes = [vp, i64, C.POINTER(C.c_float)]
        L.lbmx_halo_time.argtypes = [vp, i32, C.POINTER(C.c_float)]
        L.lbmx_get_iterations.argtypes = [vp, C.POINTER(i64)]
        L.lbmx_set_iterations.argtypes = [vp, i64]
        L.lbmx_has_nan.argtypes = [vp, C.POINTER(i32)]
        L.lbmx_get_device_ptrs.argtypes = [vp, C.POINTER(Ptrs)]
        L.lbmx_get_stats.argtypes = [vp, C.POINTER(Stats)]
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise LbmxError(f"{what} failed with status {rc}: {lib().lbmx_last_error().decode()}")


def decompose_x(X: int, nranks: int, rank: int):
    off, loc = C.c_int64(), C.c_int64()
    _check(lib().lbmx_decompose_x(X, nranks, rank, C.byref(off), C.byref(loc)), "lbmx_decompose_x")
    return off.value, loc.value


def halo_directions(lattice: int):
    r, l = (C.c_int32 * 9)(), (C.c_int32 * 9)()
    n = lib().lbmx_halo_directions(lattice, r, l)
    return list(r[:n]), list(l[:n])


def halo_plan(lattice: int, streaming: int, iteration: int, X_local: int):
    msgs = (HaloMsg * 2)()
    _check(lib().lbmx_halo_plan(lattice, streaming, iteration, X_local, msgs), "lbmx_halo_plan")
    return [dict(to_right=bool(m.to_right), dirs=list(m.dirs[: m.n_dirs]), src_plane=m.src_plane, dst_plane=m.dst_plane) for m in msgs]


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().lbmx_comm_unique_id(buf), "lbmx_comm_unique_id")
    return buf.raw


class Engine:
    """One x-slab of a lattice on one GPU.  Host arrays use the reference layout: [q | component][x][z][y], y fastest."""

    def __init__(self, lattice=D3Q27, coll=CUM, eq=EQ_INV_CUM, streaming=AB, macro=MACRO_DEFAULT, inflow=INFLOW_CONST, precision=F64,
                 X=8, Y=8, Z=8, rank=0, nranks=1, device=-1, ghost_x=0, periodic_x=0, macro_policy=MACRO_LAST_STEP, flags=0):
        self._h = C.c_void_p()
        self.desc = Desc(lattice, coll, eq, streaming, macro, inflow, precision, macro_policy, X, Y, Z, rank, nranks, device, ghost_x, periodic_x, flags)
        _check(lib().lbmx_create(C.byref(self.desc), C.byref(self._h)), "lbmx_create")
        self.layout = Layout()
        _check(lib().lbmx_get_layout(self._h, C.byref(self.layout)), "lbmx_get_layout")
        self.dtype = np.float64 if precision == F64 else np.float32
        self.params = Params(lbmViscosity=0.01)

    # -- life cycle
    def close(self):
        if self._h:
            lib().lbmx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def comm_init(self, unique_id: bytes):
        _check(lib().lbmx_comm_init(self._h, unique_id), "lbmx_comm_init")

    # -- shapes
    def _xs(self, with_ghosts):
        return self.layout.X_local + (2 * self.layout.ghost_x if with_ghosts else 0)

    def df_shape(self, with_ghosts=False):
        return (self.layout.Q, self._xs(with_ghosts), self.layout.Z, self.layout.Y)

    def macro_shape(self, with_ghosts=False):
        return (max(self.layout.n_macro, 1), self._xs(with_ghosts), self.layout.Z, self.layout.Y)

    def map_shape(self, with_ghosts=False):
        return (self._xs(with_ghosts), self.layout.Z, self.layout.Y)

    @staticmethod
    def _ptr(a, dtype, shape=None):
        assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"] and a.dtype == dtype, (a.dtype, dtype)
        if shape is not None:
            assert tuple(a.shape) == tuple(shape), (a.shape, shape)
        return a.ctypes.data

    # -- state
    def map_upload(self, m, with_ghosts=False):
        _check(lib().lbmx_map_upload(self._h, self._ptr(m, np.int16, self.map_shape(with_ghosts)), int(with_ghosts)), "lbmx_map_upload")

    def map_download(self, with_ghosts=False):
        m = np.empty(self.map_shape(with_ghosts), dtype=np.int16)
        _check(lib().lbmx_map_download(self._h, m.ctypes.data, int(with_ghosts)), "lbmx_map_download")
        return m

    def set_equilibrium(self, rho=1.0, vx=0.0, vy=0.0, vz=0.0):
        _check(lib().lbmx_df_set_equilibrium(self._h, rho, vx, vy, vz), "lbmx_df_set_equilibrium")

    def set_equilibrium_field(self, rho, vx, vy, vz=None):
        shp = self.map_shape(False)
        f64 = np.float64
        _check(lib().lbmx_df_set_equilibrium_field(self._h, self._ptr(rho, f64, shp), self._ptr(vx, f64, shp), self._ptr(vy, f64, shp),
                                                   None if vz is None else self._ptr(vz, f64, shp)), "lbmx_df_set_equilibrium_field")

    def df_upload(self, df, which=0, with_ghosts=False):
        _check(lib().lbmx_df_upload(self._h, which, self._ptr(df, self.dtype, self.df_shape(with_ghosts)), int(with_ghosts)), "lbmx_df_upload")

    def df_download(self, which=0, with_ghosts=False, out=None):
        df = out if out is not None else np.empty(self.df_shape(with_ghosts), dtype=self.dtype)
        _check(lib().lbmx_df_download(self._h, which, self._ptr(df, self.dtype, self.df_shape(with_ghosts)), int(with_ghosts)), "lbmx_df_download")
        return df

    def df_sync_ghosts(self):
        _check(lib().lbmx_df_sync_ghosts(self._h), "lbmx_df_sync_ghosts")

    def macro_init(self):
        _check(lib().lbmx_macro_init(self._h), "lbmx_macro_init")

    def macro_download(self, with_ghosts=False, out=None):
        m = out if out is not None else np.zeros(self.macro_shape(with_ghosts), dtype=self.dtype)
        _check(lib().lbmx_macro_download(self._h, self._ptr(m, self.dtype, self.macro_shape(with_ghosts)), int(with_ghosts)), "lbmx_macro_download")
        return m

    def macro_upload(self, m, with_ghosts=False):
        _check(lib().lbmx_macro_upload(self._h, self._ptr(m, self.dtype, self.macro_shape(with_ghosts)), int(with_ghosts)), "lbmx_macro_upload")

    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise AttributeError(k)
            setattr(self.params, k, v)
        _check(lib().lbmx_set_params(self._h, C.byref(self.params)), "lbmx_set_params")

    def set_inflow_profile(self, prof):
        assert prof.ndim == 2
        _check(lib().lbmx_set_inflow_profile(self._h, self._ptr(prof, self.dtype), prof.shape[1], prof.shape[0]), "lbmx_set_inflow_profile")

    def bouzidi_upload(self, coeff):
        shp = (8,) + self.map_shape(False)
        _check(lib().lbmx_bouzidi_upload(self._h, self._ptr(coeff, self.dtype, shp)), "lbmx_bouzidi_upload")

    # -- stepping
    def step(self, n=1):
        _check(lib().lbmx_step(self._h, n), "lbmx_step")

    def sync(self):
        _check(lib().lbmx_sync(self._h), "lbmx_sync")

    def step_timed(self, n) -> float:
        ms = C.c_float()
        _check(lib().lbmx_step_timed(self._h, n, C.byref(ms)), "lbmx_step_timed")
        return ms.value

    def halo_time(self, reps: int = 20) -> float:
        """Device time (ms) of one halo exchange run alone on the communication stream; collective over the ranks."""
        ms = C.c_float()
        _check(lib().lbmx_halo_time(self._h, reps, C.byref(ms)), "lbmx_halo_time")
        return ms.value

    @property
    def iterations(self) -> int:
        it = C.c_int64()
        _check(lib().lbmx_get_iterations(self._h, C.byref(it)), "lbmx_get_iterations")
        return it.value

    @iterations.setter
    def iterations(self, v):
        _check(lib().lbmx_set_iterations(self._h, v), "lbmx_set_iterations")

    def has_nan(self) -> bool:
        f = C.c_int32()
        _check(lib().lbmx_has_nan(self._h, C.byref(f)), "lbmx_has_nan")
        return bool(f.value)

    def device_ptrs(self) -> Ptrs:
        p = Ptrs()
        _check(lib().lbmx_get_device_ptrs(self._h, C.byref(p)), "lbmx_get_device_ptrs")
        return p

    def stats(self) -> Stats:
        s = Stats()
        _check(lib().lbmx_get_stats(self._h, C.byref(s)), "lbmx_get_stats")
        return s
