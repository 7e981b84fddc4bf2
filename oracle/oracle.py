"""ctypes front-end to the CPU checkers declared in oracle/oracle_api.h.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product package.

Two kinds of library implement the same interface:
  kind="reference"  oracle/_ref/libref_{ab,aa}[_fast].so : the reference's own per-cell code
                    (include/lbm3d/kernels.h:60-100 + trait headers) compiled by oracle/Makefile
  kind="port"       oracle/liboracle_port[_fast].so      : the C++ restatement oracle/lbm_oracle.cpp
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

D3Q27, D2Q9, D3Q19 = 0, 1, 2
CUM, SRT, BGK, MRT_LES, CLBM, SRT_MODIF_FORCE, SRT_WELL, BGK_WELL, CLBM_WELL, CUM_WELL, CUM_2017, CUM_ANTIALIAS, CUM_2017_ANTIALIAS = range(13)
KBC_N1, KBC_N2, KBC_N3, KBC_N4, KBC_C1, KBC_C2, KBC_C3, KBC_C4 = range(13, 21)
BGK_GALILEAN = 21  # D3Q27_BGK built with -DUSE_GALILEAN_CORRECTION (defs.h:253)
CUM_HP_RHO = 22  # D3Q27_CUM built with -DUSE_HIGH_PRECISION_RHO (defs.h:252)
EQ_STD, EQ_INV_CUM, EQ_WELL, EQ_ENTROPIC = 0, 1, 2, 3
AB, AA = 0, 1
MACRO_VOID, MACRO_DEFAULT, MACRO_MEAN, MACRO_WITH_MEAN_2D = 0, 1, 2, 3
GATE_MEANS, GATE_FLUCS = 1, 2
INFLOW_NONE, INFLOW_CONST, INFLOW_PROFILE_YZ, INFLOW_PARABOLIC_Y = 0, 1, 2, 3
F32, F64 = 0, 1

Q_OF = {D3Q27: 27, D2Q9: 9, D3Q19: 19}


def n_macro(lattice: int, macro: int) -> int:
    """Number of macro components (d3q27/macro.h:56-63,89-105; d2q9/macro.h)."""
    if macro == MACRO_VOID:
        return 0
    if lattice == D2Q9:
        return {MACRO_DEFAULT: 3, MACRO_MEAN: 8, MACRO_WITH_MEAN_2D: 10}[macro]
    return 4 if macro == MACRO_DEFAULT else 13


class _Desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("lattice", "coll", "eq", "streaming", "macro", "inflow", "precision", "nproc")] + [
        (n, C.c_int64) for n in ("X", "Y", "Z", "ox")
    ]


class _Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("lbmViscosity", "fx", "fy", "fz", "inflow_vx", "inflow_vy", "inflow_vz")] + [
        ("vx_profile", C.c_void_p),
        ("profile_size_y", C.c_int64),
        ("stat_counter", C.c_int32),
        ("macro_gates", C.c_int32),
        ("bouzidi_coeff", C.c_void_p),
    ]


@dataclass
class Desc:
    lattice: int = D3Q27
    coll: int = CUM
    eq: int = EQ_INV_CUM
    streaming: int = AB
    macro: int = MACRO_DEFAULT
    inflow: int = INFLOW_CONST
    precision: int = F64
    nproc: int = 1
    X: int = 8
    Y: int = 8
    Z: int = 8
    ox: int = 0

    @property
    def Q(self) -> int:
        return Q_OF[self.lattice]

    @property
    def dtype(self):
        return np.float64 if self.precision == F64 else np.float32

    @property
    def XYZ(self) -> int:
        return (self.X + 2 * self.ox) * self.Y * self.Z

    @property
    def n_macro(self) -> int:
        return n_macro(self.lattice, self.macro)

    def c(self) -> _Desc:
        return _Desc(self.lattice, self.coll, self.eq, self.streaming, self.macro, self.inflow, self.precision, self.nproc, self.X, self.Y, self.Z, self.ox)

    # ---- array helpers in the reference layout (q, x+ox, z, y) ----
    def new_df(self) -> np.ndarray:
        return np.zeros((self.Q, self.X + 2 * self.ox, self.Z, self.Y), dtype=self.dtype)

    def new_macro(self) -> np.ndarray:
        return np.zeros((max(self.n_macro, 1), self.X + 2 * self.ox, self.Z, self.Y), dtype=self.dtype)

    def new_map(self, value: int = 0) -> np.ndarray:
        return np.full((self.X + 2 * self.ox, self.Z, self.Y), value, dtype=np.int16)


@dataclass
class Params:
    lbmViscosity: float = 0.01
    fx: float = 0.0
    fy: float = 0.0
    fz: float = 0.0
    inflow_vx: float = 0.0
    inflow_vy: float = 0.0
    inflow_vz: float = 0.0
    vx_profile: np.ndarray | None = None  # dreal[z, y]
    stat_counter: int = 0
    macro_gates: int = 0
    bouzidi: np.ndarray | None = None  # dreal[8, x, z, y] (D2Q9 near-wall interpolation coefficients)
    _keep: list = field(default_factory=list, repr=False)

    def c(self) -> _Params:
        ptr, sy = None, 0
        if self.vx_profile is not None:
            assert self.vx_profile.flags["C_CONTIGUOUS"]
            ptr = self.vx_profile.ctypes.data
            sy = self.vx_profile.shape[-1]
        bz = None
        if self.bouzidi is not None:
            assert self.bouzidi.flags["C_CONTIGUOUS"] and self.bouzidi.shape[0] == 8
            bz = self.bouzidi.ctypes.data
        return _Params(self.lbmViscosity, self.fx, self.fy, self.fz, self.inflow_vx, self.inflow_vy, self.inflow_vz, ptr, sy, self.stat_counter, self.macro_gates, bz)


def _path(kind: str, streaming: int, fast: bool) -> str:
    suf = "_fast" if fast else ""
    if kind == "reference":
        return os.path.join(HERE, "_ref", f"libref_{'aa' if streaming == AA else 'ab'}{suf}.so")
    if kind == "port":
        return os.path.join(HERE, f"liboracle_port{suf}.so")
    if kind == "engine_host":
        # not a checker: the engine's CUDA kernels compiled for the host (tests/host_harness/engine_host.cpp) behind this same C interface, so that the
        # tests can run them against the checkers without a GPU.  fast = the default arithmetic, otherwise the parity arithmetic.
        return os.path.join(os.path.dirname(HERE), "tests", "host_harness", "bin", f"libengine_host{suf}.so")
    raise ValueError(kind)


def available(kind: str, streaming: int = AB, fast: bool = False) -> bool:
    return os.path.exists(_path(kind, streaming, fast))


_libs: dict = {}


def _load(kind: str, streaming: int, fast: bool):
    key = _path(kind, streaming, fast)
    lib = _libs.get(key)
    if lib is None:
        lib = C.CDLL(key)
        lib.oracle_kind.restype = C.c_char_p
        lib.oracle_supported.argtypes = [C.POINTER(_Desc)]
        lib.oracle_step.argtypes = [C.POINTER(_Desc), C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
        lib.oracle_set_equilibrium.argtypes = [C.POINTER(_Desc), C.c_void_p] + [C.c_double] * 4
        lib.oracle_set_equilibrium_field.argtypes = [C.POINTER(_Desc), C.c_void_p] + [C.c_void_p] * 4
        lib.oracle_initial_macro.argtypes = [C.POINTER(_Desc), C.POINTER(_Params), C.c_void_p, C.c_void_p]
        _libs[key] = lib
    return lib


class Oracle:
    """One CPU checker bound to a descriptor.  `kind` is "reference" or "port" ("engine_host": see _path)."""

    def __init__(self, desc: Desc, kind: str = "port", fast: bool = False):
        self.desc = desc
        self.kind = kind
        self.lib = _load(kind, desc.streaming, fast)
        assert self.lib.oracle_kind().decode() == kind
        d = desc.c()
        if self.lib.oracle_supported(C.byref(d)) != 0:
            raise NotImplementedError(f"{kind} oracle has no instantiation for {desc}")

    def _chk(self, a: np.ndarray | None, dtype=None):
        if a is None:
            return None
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        if dtype is not None:
            assert a.dtype == dtype, (a.dtype, dtype)
        return a.ctypes.data

    def step(self, params: Params, df_a, df_b, macro, map_, iteration: int = 0, nsteps: int = 1, nthreads: int = 1):
        d, p = self.desc.c(), params.c()
        dt = self.desc.dtype
        r = self.lib.oracle_step(C.byref(d), C.byref(p), self._chk(df_a, dt), self._chk(df_b, dt), self._chk(macro, dt), self._chk(map_, np.int16), iteration, nsteps, nthreads)
        if r != 0:
            raise RuntimeError(f"oracle_step -> {r}")

    def set_equilibrium(self, df, rho=1.0, vx=0.0, vy=0.0, vz=0.0):
        d = self.desc.c()
        r = self.lib.oracle_set_equilibrium(C.byref(d), self._chk(df, self.desc.dtype), rho, vx, vy, vz)
        if r != 0:
            raise RuntimeError(f"oracle_set_equilibrium -> {r}")

    def set_equilibrium_field(self, df, rho, vx, vy, vz=None):
        d = self.desc.c()
        f64 = np.float64
        r = self.lib.oracle_set_equilibrium_field(C.byref(d), self._chk(df, self.desc.dtype), self._chk(rho, f64), self._chk(vx, f64), self._chk(vy, f64), self._chk(vz, f64))
        if r != 0:
            raise RuntimeError(f"oracle_set_equilibrium_field -> {r}")

    def initial_macro(self, params: Params, df, macro):
        d, p = self.desc.c(), params.c()
        r = self.lib.oracle_initial_macro(C.byref(d), C.byref(p), self._chk(df, self.desc.dtype), self._chk(macro, self.desc.dtype))
        if r != 0:
            raise RuntimeError(f"oracle_initial_macro -> {r}")
