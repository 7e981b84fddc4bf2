"""Physics known-answer test and edge cases of the engine (through the C ABI).

* Square-duct flow driven by a body force against the analytical Fourier-series profile the reference's validation solver
  uses (sim_NSE/sim_2.cu:63-88; geometry :115-139) -- the only "known answer" the reference has.
* Degenerate and ragged lattices, argument validation, state errors."""
import math

import numpy as np
import pytest

import lbm_cases as lc
from oracle import oracle as O
from tnl_lbm_b200 import binding as B


def analytical_ux(Y, Z, fx, nu, n=60):
    """Steady body-force flow in a rectangular duct (the Fourier series the reference's validation solver evaluates,
    sim_NSE/sim_2.cu:63-88), as a [z, y] array in lattice units.  Full-way bounce-back puts the no-slip plane half-way between
    the GEO_WALL node (y = 1, Y-2) and the first fluid node, so the half-widths are a = (Y-4)/2, b = (Z-4)/2 around the centre
    (Y-1)/2.  (sim_2.cu uses Y/2-1 with a half-cell shifted origin; the series is the same.)"""
    a = (Y - 4) / 2.0
    b = (Z - 4) / 2.0
    y = (np.arange(Y) - (Y - 1) / 2.0)[None, :]
    z = (np.arange(Z) - (Z - 1) / 2.0)[:, None]
    s = np.zeros((Z, Y))
    for i in range(n + 1):
        k = 2 * i + 1
        s += (-1.0) ** i * (1.0 - np.cosh(k * math.pi * z / (2 * a)) / np.cosh(k * math.pi * b / (2 * a))) * np.cos(k * math.pi * y / (2 * a)) / k ** 3
    return 16.0 * a * a * fx / (nu * math.pi ** 3) * s


@pytest.mark.gpu
@pytest.mark.parametrize("coll,eq,streaming", [(B.CUM, B.EQ_INV_CUM, B.AA), (B.CUM, B.EQ_INV_CUM, B.AB), (B.SRT, B.EQ_STD, B.AB), (B.BGK, B.EQ_STD, B.AA)])
def test_duct_flow_reaches_the_analytical_profile(coll, eq, streaming):
    X, Y, Z = 16, 36, 36
    nu, fx = 1.0 / 6.0, 1e-6
    d = O.Desc(coll=coll, eq=eq, streaming=streaming, X=X, Y=Y, Z=Z)
    with B.Engine(lattice=B.D3Q27, coll=coll, eq=eq, streaming=streaming, precision=B.F64, X=X, Y=Y, Z=Z) as e:
        # A-B: the reference's own duct map (walls reach the periodic x faces and clamp there: x-invariant flow);
        # A-A needs the variant whose non-periodic face cells are GEO_NOTHING, which disturbs the flow next to the x faces
        e.map_upload(lc.map_duct_periodic_x(d) if streaming == B.AB else lc.map_duct_slab_safe(d))
        e.set_equilibrium(1.0, 0, 0, 0)
        e.set_params(lbmViscosity=nu, fx=fx)
        e.step(8000)
        mac = e.macro_download()
        assert not e.has_nan()
    ux = mac[1, X // 2].astype(np.float64)  # [z, y]
    ref = analytical_ux(Y, Z, fx, nu)
    inner = (slice(2, Z - 2), slice(2, Y - 2))
    l2 = np.sqrt(np.sum((ux[inner] - ref[inner]) ** 2) / np.sum(ref[inner] ** 2))
    assert l2 < 0.03, f"relative L2 error vs analytical duct profile {l2:.4f}"  # bounce-back wall-position error at tau = 1, 32 cells across
    if streaming == B.AB:
        assert np.allclose(mac[1, 1], mac[1, X - 2], rtol=0, atol=1e-12 * ref[inner].max())  # x-invariant
    assert abs(mac[2]).max() < 1e-9 and abs(mac[3]).max() < 1e-9  # no cross flow


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 5, 1), (3, 1, 2), (2, 33, 3), (5, 129, 2), (4, 7, 131)])
@pytest.mark.parametrize("streaming", [B.AB, B.AA])
def test_ragged_and_degenerate_lattices(shape, streaming):
    """Uniform equilibrium at rest without force is a fixed point on any periodic lattice, including 1-cell axes and sizes that
    are not multiples of the warp / CTA width."""
    X, Y, Z = shape
    with B.Engine(coll=B.CUM, eq=B.EQ_INV_CUM, streaming=streaming, precision=B.F64, X=X, Y=Y, Z=Z) as e:
        e.map_upload(np.full((X, Z, Y), 7, dtype=np.int16))
        e.set_equilibrium(1.0, 0.03, -0.02, 0.01)
        e.set_params(lbmViscosity=0.01)
        f0 = e.df_download(0)
        e.step(6)  # even count: the A-A array is back in its canonical (streamed) form
        f1 = e.df_download(0)
        mac = e.macro_download()
    assert np.abs(f1 - f0).max() < 1e-15
    assert np.allclose(mac[0], 1.0, atol=1e-14) and np.allclose(mac[1], 0.03, atol=1e-15)


@pytest.mark.gpu
def test_empty_lattice_of_inert_cells():
    """A map of GEO_NOTHING only: nothing is read or written, macros report rho=1, u=0 (d3q27/bc.h:53-60)."""
    with B.Engine(X=4, Y=6, Z=5) as e:
        e.map_upload(np.full((4, 5, 6), 8, dtype=np.int16))
        marker = np.full(e.df_shape(), 0.125)
        e.df_upload(marker, 0)
        e.df_upload(marker, 1)
        e.set_params(lbmViscosity=0.01)
        e.step(3)
        assert np.array_equal(e.df_download(0), marker) and np.array_equal(e.df_download(1), marker)
        mac = e.macro_download()
        assert np.all(mac[0] == 1.0) and np.all(mac[1:] == 0.0)
        st = e.stats()
        assert st.bulk_cells == 120 and st.boundary_cells == 0  # inert cells cost the bulk kernel one macro write, no list entry


@pytest.mark.gpu
def test_state_and_argument_errors():
    with B.Engine(X=4, Y=4, Z=4) as e:
        with pytest.raises(B.LbmxError, match="upload a map first"):
            e.step(1)
        e.map_upload(np.zeros((4, 4, 4), dtype=np.int16))
        with pytest.raises(B.LbmxError, match="lbmViscosity must not be 0"):
            e.set_params(lbmViscosity=0.0)
        e.set_params(lbmViscosity=0.01)
        with pytest.raises(B.LbmxError):
            e.step(-1)
        e.step(0)
        assert e.iterations == 0
        with pytest.raises(AssertionError):
            e.map_upload(np.zeros((4, 4, 5), dtype=np.int16))  # wrong shape is caught by the binding before the call
    with B.Engine(X=4, Y=4, Z=4, inflow=B.INFLOW_PROFILE_YZ) as e:
        e.map_upload(np.zeros((4, 4, 4), dtype=np.int16))
        with pytest.raises(B.LbmxError, match="inflow profile"):
            e.step(1)
        with pytest.raises(B.LbmxError, match="cross-section"):
            e.set_inflow_profile(np.zeros((3, 4), dtype=e.dtype))  # 4 x 3 (y, z): smaller than the 4 x 4 cross-section
        e.set_inflow_profile(np.zeros((4, 4), dtype=e.dtype))
        e.step(1)
    with B.Engine(X=4, Y=4, Z=4, macro=B.MACRO_VOID) as e:
        e.map_upload(np.zeros((4, 4, 4), dtype=np.int16))
        e.step(2)
        assert e.layout.n_macro == 0 and not e.has_nan()
    with B.Engine(X=4, Y=4, Z=4, streaming=B.AA) as e:
        with pytest.raises(B.LbmxError, match="single array"):
            e.df_download(1)


def test_slab_larger_than_32bit_cell_index_is_refused():
    """Maximum size: cell indices are 32-bit per slab; a larger slab must be refused with a clear message, before any allocation."""
    with pytest.raises(B.LbmxError, match="2\\^31"):
        B.Engine(X=4096, Y=1024, Z=1024)
