// lattice.cuh -- velocity sets in the reference's numbering (include/lbm3d/defs.h:257-305) as compile-time tables.
#pragma once
#include <cstdint>
#include <utility>

#define LBMX_HD __host__ __device__ __forceinline__
#define LBMX_D __device__ __forceinline__
#ifdef __CUDACC__
	#define LBMX_GRID_CONSTANT __grid_constant__  // kernel parameter whose address may be handed to an out-of-line device function
#else
	#define LBMX_GRID_CONSTANT
#endif

namespace lbmx {

// compile-time loop: body(std::integral_constant<int, i>) for i in [0, N)
template <typename F, int... I>
LBMX_HD void static_for_impl(F&& f, std::integer_sequence<int, I...>)
{
	(f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F>
LBMX_HD void static_for(F&& f)
{
	static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

struct D3Q27
{
	static constexpr int Q = 27;
	static constexpr int NDIM = 3;
	// 0 zzz 1 pzz 2 mzz 3 zpz 4 zmz 5 zzp 6 zzm 7 ppz 8 mmz 9 pmz 10 mpz 11 pzp 12 mzm 13 pzm 14 mzp
	// 15 zpp 16 zmm 17 zpm 18 zmp 19 ppp 20 mmm 21 ppm 22 mmp 23 pmp 24 mpm 25 pmm 26 mpp
	LBMX_HD static constexpr int cx(int q)
	{
		constexpr int t[27] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1};
		return t[q];
	}
	LBMX_HD static constexpr int cy(int q)
	{
		constexpr int t[27] = {0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, -1, 1, -1, 1};
		return t[q];
	}
	LBMX_HD static constexpr int cz(int q)
	{
		constexpr int t[27] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1, 1, -1, -1, 1, 1, -1, -1, 1};
		return t[q];
	}
	LBMX_HD static constexpr int find(int x, int y, int z)
	{
		for (int q = 0; q < 27; q++)
			if (cx(q) == x && cy(q) == y && cz(q) == z)
				return q;
		return -1;
	}
	LBMX_HD static constexpr int opp(int q) { return find(-cx(q), -cy(q), -cz(q)); }
	// cell types, d3q27/bc.h:17-34
	enum : int { FLUID = 0, WALL, INFLOW, INFLOW_LEFT, OUTFLOW_EQ, OUTFLOW_RIGHT, OUTFLOW_RIGHT_INTERP, PERIODIC, NOTHING, SYM_TOP, SYM_BOTTOM, SYM_LEFT, SYM_RIGHT, SYM_BACK, SYM_FRONT };
	LBMX_HD static constexpr bool collides(int m) { return m == FLUID || m == PERIODIC || m == OUTFLOW_RIGHT || m == OUTFLOW_RIGHT_INTERP || m == INFLOW_LEFT; }  // bc.h:243-248
	LBMX_HD static constexpr bool bulk(int m) { return m == FLUID || m == PERIODIC; }
};

// D3Q19: the first 19 directions of the D3Q27 numbering (the "+Q19" block of defs.h:273-295).  The reference has no such
// lattice: BASELINE.json names it, so it exists here, with PARITY UNPINNED by construction.
struct D3Q19
{
	static constexpr int Q = 19;
	static constexpr int NDIM = 3;
	LBMX_HD static constexpr int cx(int q) { return D3Q27::cx(q); }
	LBMX_HD static constexpr int cy(int q) { return D3Q27::cy(q); }
	LBMX_HD static constexpr int cz(int q) { return D3Q27::cz(q); }
	LBMX_HD static constexpr int find(int x, int y, int z)
	{
		const int q = D3Q27::find(x, y, z);
		return q < 19 ? q : -1;
	}
	LBMX_HD static constexpr int opp(int q) { return find(-cx(q), -cy(q), -cz(q)); }
	enum : int { FLUID = 0, WALL, INFLOW, INFLOW_LEFT, OUTFLOW_EQ, OUTFLOW_RIGHT, OUTFLOW_RIGHT_INTERP, PERIODIC, NOTHING, SYM_TOP, SYM_BOTTOM, SYM_LEFT, SYM_RIGHT, SYM_BACK, SYM_FRONT };
	LBMX_HD static constexpr bool collides(int m) { return D3Q27::collides(m); }
	LBMX_HD static constexpr bool bulk(int m) { return m == FLUID || m == PERIODIC; }
};

struct D2Q9
{
	static constexpr int Q = 9;
	static constexpr int NDIM = 2;
	// 0 zz 1 pz 2 mz 3 zp 4 zm 5 pp 6 mm 7 pm 8 mp
	LBMX_HD static constexpr int cx(int q)
	{
		constexpr int t[9] = {0, 1, -1, 0, 0, 1, -1, 1, -1};
		return t[q];
	}
	LBMX_HD static constexpr int cy(int q)
	{
		constexpr int t[9] = {0, 0, 0, 1, -1, 1, -1, -1, 1};
		return t[q];
	}
	LBMX_HD static constexpr int cz(int) { return 0; }
	LBMX_HD static constexpr int find(int x, int y, int = 0)
	{
		for (int q = 0; q < 9; q++)
			if (cx(q) == x && cy(q) == y)
				return q;
		return -1;
	}
	LBMX_HD static constexpr int opp(int q) { return find(-cx(q), -cy(q)); }
	// cell types, d2q9/bc.h:16-34 (numeric values differ from D3Q27!)
	enum : int { FLUID = 0, WALL, INFLOW, OUTFLOW_EQ, OUTFLOW_RIGHT, OUTFLOW_RIGHT_INTERP, PERIODIC, NOTHING, SYM_TOP, SYM_BOTTOM, SYM_LEFT, SYM_RIGHT, FLUID_NEAR_WALL, INFLOW_LEFT = -100, SYM_BACK = -101, SYM_FRONT = -102 };
	LBMX_HD static constexpr bool collides(int m) { return m == FLUID || m == FLUID_NEAR_WALL || m == PERIODIC || m == OUTFLOW_RIGHT || m == OUTFLOW_RIGHT_INTERP; }  // d2q9/bc.h:198-203
	LBMX_HD static constexpr bool bulk(int m) { return m == FLUID || m == PERIODIC; }
};

}  // namespace lbmx
