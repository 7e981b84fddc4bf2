"""Shared synthetic inputs for the parity tests (maps, initial fields, error norms).

Layouts follow the reference: maps/macros are [x, z, y] and distributions
[q, x, z, y] C-contiguous arrays, y fastest.  Everything is seeded and platform independent
(numpy RandomState / closed-form fields), so CPU oracle and GPU engine see identical bits.
"""
from __future__ import annotations

import numpy as np

from oracle import oracle as O

# cell types: d3q27/bc.h:17-34 and d2q9/bc.h:16-34 (numeric values differ between the two!)
G3 = dict(FLUID=0, WALL=1, INFLOW=2, INFLOW_LEFT=3, OUTFLOW_EQ=4, OUTFLOW_RIGHT=5, OUTFLOW_RIGHT_INTERP=6, PERIODIC=7, NOTHING=8,
          SYM_TOP=9, SYM_BOTTOM=10, SYM_LEFT=11, SYM_RIGHT=12, SYM_BACK=13, SYM_FRONT=14)
G2 = dict(FLUID=0, WALL=1, INFLOW=2, OUTFLOW_EQ=3, OUTFLOW_RIGHT=4, OUTFLOW_RIGHT_INTERP=5, PERIODIC=6, NOTHING=7,
          SYM_TOP=8, SYM_BOTTOM=9, SYM_LEFT=10, SYM_RIGHT=11)

C27 = np.array([(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 0), (-1, -1, 0), (1, -1, 0), (-1, 1, 0),
                (1, 0, 1), (-1, 0, -1), (1, 0, -1), (-1, 0, 1), (0, 1, 1), (0, -1, -1), (0, 1, -1), (0, -1, 1), (1, 1, 1), (-1, -1, -1),
                (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, -1), (1, -1, -1), (-1, 1, 1)])  # defs.h:273-305
C9 = np.array([(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (1, 1, 0), (-1, -1, 0), (1, -1, 0), (-1, 1, 0)])  # defs.h:257-270


def geo(d: O.Desc) -> dict:
    return G2 if d.lattice == O.D2Q9 else G3


def smooth_fields(d: O.Desc, amp: float = 1.0):
    """rho, vx, vy, vz of SURVEY §8d cfg 3 (smooth, periodic) as float64 [x+2ox, z, y] arrays."""
    X, Y, Z = d.X, d.Y, d.Z
    x = (np.arange(-d.ox, X + d.ox) % X)[:, None, None]
    z = np.arange(Z)[None, :, None]
    y = np.arange(Y)[None, None, :]
    shape = (X + 2 * d.ox, Z, Y)
    rho = np.broadcast_to(1.0 + 0.01 * amp * np.sin(2 * np.pi * x / X), shape).copy()
    vx = np.broadcast_to(0.05 * amp * np.sin(2 * np.pi * y / Y), shape).copy()
    vy = np.broadcast_to(0.02 * amp * np.cos(2 * np.pi * z / max(Z, 1)) * (1.0 if Z > 1 else np.cos(2 * np.pi * x / X)), shape).copy()
    vz = np.full(shape, 0.01 * amp if Z > 1 else 0.0)
    return rho, vx, vy, vz


def noisy_df(d: O.Desc, orc: O.Oracle, seed: int = 1234, noise: float = 0.05, amp: float = 1.0) -> np.ndarray:
    """Equilibrium of the smooth field times (1 + noise*U[-1,1]) per population: positive, far from equilibrium."""
    df = d.new_df()
    rho, vx, vy, vz = smooth_fields(d, amp)
    orc.set_equilibrium_field(df, rho, vx, vy, vz)
    rs = np.random.RandomState(seed)
    df *= (1.0 + noise * (2.0 * rs.random_sample(df.shape) - 1.0)).astype(df.dtype)
    return df


def map_periodic(d: O.Desc) -> np.ndarray:
    return d.new_map(geo(d)["PERIODIC"])


def map_random_ab(d: O.Desc, seed: int = 7, frac_special: float = 0.5) -> np.ndarray:
    """Every cell type at random places: well defined under A-B because neighbour indices clamp (kernels.h:49-56)."""
    g = geo(d)
    rs = np.random.RandomState(seed)
    m = d.new_map(g["FLUID"])
    kinds = np.array(sorted(g.values()), dtype=np.int16)
    pick = rs.random_sample(m.shape) < frac_special
    m[pick] = kinds[rs.randint(0, len(kinds), size=int(pick.sum()))]
    return m


def map_random_aa(d: O.Desc, seed: int = 7, frac_special: float = 0.4) -> np.ndarray:
    """A-A zoo: GEO_PERIODIC shell (so every access wraps in bounds, kernels.h:21-29) and random interior types.
    OUTFLOW_RIGHT / OUTFLOW_RIGHT_INTERP are left out: under A-A they read slots owned by another cell, so the
    result depends on the visiting order (and the reference's A-A streaming lacks streamingInterpRight)."""
    g = geo(d)
    rs = np.random.RandomState(seed)
    m = d.new_map(g["PERIODIC"])
    skip = ("OUTFLOW_RIGHT", "OUTFLOW_RIGHT_INTERP")
    if d.lattice == O.D2Q9:
        # the reference's D2Q9 wall / SYM_TOP / SYM_BOTTOM rules index f[z-1], f[z+1] under A-A (shadowed enumerators,
        # d2q9/bc.h:90,135,184,190): undefined behaviour in the reference, so there is nothing to compare against
        skip += ("WALL", "SYM_TOP", "SYM_BOTTOM")
    kinds = np.array(sorted(v for k, v in g.items() if k not in skip), dtype=np.int16)
    zs = slice(1, -1) if d.Z > 2 else slice(None)
    inner = m[1:-1, zs, 1:-1]
    vals = np.full(inner.shape, g["FLUID"], dtype=np.int16)
    pick = rs.random_sample(inner.shape) < frac_special
    vals[pick] = kinds[rs.randint(0, len(kinds), size=int(pick.sum()))]
    m[1:-1, zs, 1:-1] = vals
    return m


def map_sim1_channel(d: O.Desc, inflow_kind: str = "INFLOW") -> np.ndarray:
    """Orifice channel painted in the order of sim_NSE/sim_1.cu:25-52 (x-planes, wall slab with a centred hole,
    walls at y,z = 1 / N-2, NOTHING shell at y,z = 0 / N-1), scaled to the lattice size."""
    g = G3
    X, Y, Z = d.X, d.Y, d.Z
    m = d.new_map(g["FLUID"])
    o = d.ox
    m[o + 0, :, :] = g[inflow_kind]
    m[o + X - 1, :, :] = g["OUTFLOW_RIGHT"]
    cx = int(np.floor(0.20 / (0.41 / (Y - 2))))
    width = max(X // 40, 1)
    for px in range(cx, min(cx + width, X - 2) + 1):
        for pz in range(1, Z - 1):
            for py in range(1, Y - 1):
                if not (0.4 * Y <= py < 0.6 * Y and 0.4 * Z <= pz < 0.6 * Z):
                    m[o + px, pz, py] = g["WALL"]
    m[:, 1, :] = g["WALL"]
    m[:, Z - 2, :] = g["WALL"]
    m[:, :, 1] = g["WALL"]
    m[:, :, Y - 2] = g["WALL"]
    m[:, 0, :] = g["NOTHING"]
    m[:, Z - 1, :] = g["NOTHING"]
    m[:, :, 0] = g["NOTHING"]
    m[:, :, Y - 1] = g["NOTHING"]
    return m


def map_duct_periodic_x(d: O.Desc) -> np.ndarray:
    """Body-force duct of sim_NSE/sim_2.cu:115-139: x-planes 0 and X-1 GEO_PERIODIC, walls at y,z = 1 / N-2, NOTHING outside."""
    g = G3
    X, Y, Z = d.X, d.Y, d.Z
    o = d.ox
    m = d.new_map(g["FLUID"])
    m[o + 0, :, :] = g["PERIODIC"]
    m[o + X - 1, :, :] = g["PERIODIC"]
    m[:, 1, :] = g["WALL"]
    m[:, Z - 2, :] = g["WALL"]
    m[:, :, 1] = g["WALL"]
    m[:, :, Y - 2] = g["WALL"]
    m[:, 0, :] = g["NOTHING"]
    m[:, Z - 1, :] = g["NOTHING"]
    m[:, :, 0] = g["NOTHING"]
    m[:, :, Y - 1] = g["NOTHING"]
    return m


def map_cavity_2d(d: O.Desc) -> np.ndarray:
    """Lid-driven cavity synthesised from reference cell types (SURVEY §8d cfg 2): walls on x=0, x=X-1, y=0; lid row y=Y-1 = GEO_INFLOW."""
    g = G2
    m = d.new_map(g["FLUID"])
    m[0, :, :] = g["WALL"]
    m[d.X - 1, :, :] = g["WALL"]
    m[:, :, 0] = g["WALL"]
    m[:, :, d.Y - 1] = g["INFLOW"]
    return m


def map_sim2d1_channel(d: O.Desc) -> np.ndarray:
    """2-D orifice channel in the painting order of sim_2D/sim2d_1.cu:56-76."""
    g = G2
    X, Y = d.X, d.Y
    m = d.new_map(g["FLUID"])
    m[0, :, :] = g["INFLOW"]
    m[X - 1, :, :] = g["OUTFLOW_RIGHT"]
    m[:, :, 0] = g["WALL"]
    m[:, :, Y - 1] = g["WALL"]
    cx = X // 5
    for py in range(Y):
        if not (0.4 * Y <= py < 0.6 * Y):
            m[cx : cx + max(X // 40, 1), 0, py] = g["WALL"]
    return m


def rel_err(a: np.ndarray, b: np.ndarray, floor: float = 0.0) -> float:
    """max |a-b| / max|b| -- the norm used for fields that cross zero (velocities).  floor: lower bound of the scale, for flows that
    have barely left rest (|u| ~ 1e-4: one ulp of the O(1) population sums is then 1e-12 of it)."""
    scale = max(float(np.max(np.abs(b))), floor)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) / (scale if scale > 0 else 1.0)


def rel_err_elementwise(a: np.ndarray, b: np.ndarray, floor: float = 1e-300) -> float:
    """max_i |a_i-b_i| / |b_i| -- used for distributions (strictly positive in all test cases)."""
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    return float(np.max(np.abs(a64 - b64) / np.maximum(np.abs(b64), floor)))


def macro_groups(d: O.Desc):
    """Component groups that share a physical scale: density | velocity vector | running means | co-moments.  A velocity
    component that is identically zero by symmetry (e.g. v_y in a duct) is compared relative to |u|_max, not to itself."""
    nd = 2 if d.lattice == O.D2Q9 else 3
    groups = [(0, 1, "rho"), (1, 1 + nd, "velocity")]
    if d.macro == O.MACRO_MEAN:
        groups += [(1 + nd, 1 + 2 * nd, "mean velocity"), (1 + 2 * nd, d.n_macro, "co-moments")]
    if d.macro == O.MACRO_WITH_MEAN_2D:
        groups += [(3, 5, "velocity sums"), (5, 7, "frozen mean"), (7, 8, "sum |u'|"), (8, 10, "sums of squares")]
    return groups


def df_floor(d: O.Desc) -> np.ndarray:
    """Per-population floor for the element-wise relative error: half the lattice weight w_q (the population's own scale at
    rho = 1).  Boundary rules (moment inflow, equilibrium decomposition) can drive single populations through zero; their
    error is then measured against w_q / 2 instead of against a value that happens to be ~0."""
    c = C9 if d.lattice == O.D2Q9 else C27[: d.Q]
    n = np.abs(c).sum(axis=1)
    table = {O.D2Q9: [4 / 9, 1 / 9, 1 / 36], O.D3Q27: [8 / 27, 2 / 27, 1 / 54, 1 / 216], O.D3Q19: [1 / 3, 1 / 18, 1 / 36]}[d.lattice]
    return 0.5 * np.array(table)[n]


def rel_err_df(a: np.ndarray, b: np.ndarray, d: O.Desc) -> float:
    fl = df_floor(d).reshape((-1,) + (1,) * (a.ndim - 1))
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    return float(np.max(np.abs(a64 - b64) / np.maximum(np.abs(b64), fl)))


def map_duct_slab_safe(d: O.Desc) -> np.ndarray:
    """Duct with periodic x whose only non-inert cells on the x faces are GEO_PERIODIC (the rest of the wall ring / shell is
    GEO_NOTHING there), and whose periodic cells never touch a y/z face.  On such a map the reference's 1-process wrap rule
    and its ghost-plane rule (kernels.h:21-57) address exactly the same neighbours, under A-B and A-A alike -- so an
    undivided run and an N-slab run must agree bit for bit."""
    g = geo(d)
    if d.lattice == O.D2Q9:
        m = d.new_map(g["FLUID"])
        m[0, :, :] = g["PERIODIC"]
        m[d.X - 1, :, :] = g["PERIODIC"]
        m[:, :, 0] = g["WALL"]
        m[:, :, d.Y - 1] = g["WALL"]
    else:
        m = map_duct_periodic_x(d)
    for xf in (0, d.X - 1):
        plane = m[xf]
        plane[plane != g["PERIODIC"]] = g["NOTHING"]
    return m


def map_and_coeffs_bouzidi(d: O.Desc, seed: int = 9):
    """Random D2Q9 map with ~30 % GEO_FLUID_NEAR_WALL (=12) cells and coefficients theta in [-0.8, 1.2] for all 8 links."""
    rs = np.random.RandomState(seed)
    m = map_random_ab(d, seed=seed)
    m[rs.random_sample(m.shape) < 0.3] = 12
    bz = (2.0 * rs.random_sample((8,) + m.shape) - 0.8).astype(d.dtype)
    return m, bz
