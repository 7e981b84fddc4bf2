// kernels.cuh -- the fused collide-and-stream kernels (replaces cudaLBMKernel<NSE>, include/lbm3d/kernels.h:60-100).
//
// Launch plan per time step and slab:
//   k_bulk      one thread per lattice cell over an x-range of the slab; handles GEO_FLUID and GEO_PERIODIC cells
//               (the overwhelming majority) with compile-time streaming pattern and operator, everything else returns;
//   k_boundary  one thread per entry of the compact boundary list built at map upload; the full cell-type dispatch of
//               D3Q27_BC_All / D2Q9_BC_All (d3q27/bc.h:51-258, d2q9/bc.h:89-213).
// Under both streaming patterns every population slot is read and written by exactly one cell per step, so the two
// kernels (and the slab's x-ranges) are independent of each other and may run in any order or concurrently.
//
// Data layout (HBM): structure of arrays, the reference's storage order kept so that raw pointers stay interchangeable
// with block.data (lbm_data.h:49-67): element (q, x, z, y) at q*XYZ + ((x+ox)*Z + z)*Y + y, y fastest.  A warp therefore
// touches 32 consecutive reals of one population: fully coalesced for the x/z-shifted and unshifted accesses, one extra
// 32-byte sector for the +-1 shifts in y.  Cell indices are 32-bit; only the per-population base q*XYZ is 64-bit.
#pragma once
#include "collide.cuh"

// ---- tuning knobs of the bulk kernel (defaults = the shipped configuration; tools/kbench.cu sweeps them) ----
// Measured on B200, D3Q27 cumulant fp64, 512^3 (profiles/kbench_r1.txt): every population is touched exactly once per step,
// so L1 allocation (plain ld) and normal-priority L2 lines (plain st) only cost: ld.cg + st.cs is worth +9 % over plain.
#ifndef LBMX_BULK_BLOCK
	#define LBMX_BULK_BLOCK 128	 // threads per CTA, along the (y,z) plane
#endif
#ifndef LBMX_BULK_MINBLOCKS
	#define LBMX_BULK_MINBLOCKS 4  // __launch_bounds__ second argument (register cap = 65536 / (BLOCK * MINBLOCKS) = 128: 16 warps per SM)
#endif
// KBC kernels (collide_kbc_fast: delta-h in place of f, 13 moments, 9 equilibrium factors and their reciprocals live).  fp64 takes 168
// registers = 3 CTAs per SM in all three kernels (the odd one spills 40 bytes there and still gains 17 % over 2 CTAs; 128 registers spill
// ~450 bytes and lose 16-30 %); fp32: 128 registers / 4 CTAs for the A-A kernels, 96 / 5 CTAs for A-B (+14 %)
// -- tools/kbench, profiles/kbench_r2_kbc_clbm_reorganised.txt (profiles/kbench_r2_kbc.txt: the reference statement order)
#ifndef LBMX_KBC_MINBLOCKS_F64
	#define LBMX_KBC_MINBLOCKS_F64 3
#endif
#ifndef LBMX_KBC_MINBLOCKS_F64_ODD
	#define LBMX_KBC_MINBLOCKS_F64_ODD 3
#endif
#ifndef LBMX_KBC_MINBLOCKS_F32
	#define LBMX_KBC_MINBLOCKS_F32 4
#endif
#ifndef LBMX_KBC_MINBLOCKS_F32_AB
	#define LBMX_KBC_MINBLOCKS_F32_AB 5
#endif
#ifndef LBMX_BULK_MINBLOCKS_AB
	#define LBMX_BULK_MINBLOCKS_AB 5  // the A-B kernel fits 96 registers without spilling and likes the extra occupancy (kbench: 6.68 vs 6.33 TB/s)
#endif
#ifndef LBMX_AB_WHOLE_SECTORS
	#define LBMX_AB_WHOLE_SECTORS 1	 // A-B bulk kernel: obstacle / inert lanes store with their warp (k_bulk phase 2)
#endif
#ifndef LBMX_BOUNDARY_MINBLOCKS
	#define LBMX_BOUNDARY_MINBLOCKS 1  // __launch_bounds__ second argument of the boundary-list kernel (4 CTAs per SM, i.e. 128 registers with ~100 spilled doubles in fp64, changed nothing measurable: profiles/solid_maps_r2.md)
#endif
#ifndef LBMX_LD_HINT
	#define LBMX_LD_HINT 2	// 0 plain, 1 ld.global.cs (evict-first), 2 ld.global.cg (L2 only), 3 ld.global.lu
#endif
#ifndef LBMX_ST_HINT
	#define LBMX_ST_HINT 1	// 0 plain, 1 st.global.cs, 2 st.global.cg, 3 st.global.wt
#endif
// macroscopic fields: written once per step and never read back by a step kernel (MACRO_Mean: read and written once), so they
// should not take L2 lines from the populations either -- 0 plain, 1 streaming (ld.global.cs / st.global.cs)
#ifndef LBMX_MACRO_HINT
	#define LBMX_MACRO_HINT 1
#endif

namespace lbmx {

enum StreamMode : int { S_AB = 0, S_AA_EVEN = 1, S_AA_ODD = 2 };
enum OutMode : int { OUT_NONE = 0, OUT_DEFAULT = 1, OUT_MEAN = 2, OUT_WITH_MEAN_2D = 4 };  // OUT_WITH_MEAN_2D + LBMX_GATE_* bits (4..7)

template <typename R>
struct KParams
{
	R* cur;			   // df_cur (A-A: the only array)
	R* out;			   // df_out (A-B)
	R* rd[27];		   // per-population read bases  cur + q*XYZ (kernel-parameter constants: one IMAD.WIDE per access)
	R* wr[27];		   // per-population write bases (A-B: out + q*XYZ, A-A: cur + q*XYZ)
	R* macro;		   // [n_macro][XYZ]
	const int16_t* map; // [XYZ]
	const R* profile;  // inflow vx profile [z*profile_sy + y] or nullptr
	const R* bouzidi;  // D2Q9 near-wall interpolation coefficients [8][XYZ] or nullptr (lbm_data.h:69-83)
	const uint32_t* blist;	// boundary list: storage cell indices
	const uint8_t* inert;	// nullptr, or [storage plane][inert_stride] flags, one per LBMX_BULK_BLOCK consecutive cells of a plane: 1 = all of them GEO_NOTHING
	int inert_stride;		// = ceil(YZ / LBMX_BULK_BLOCK)
	long long XYZ;	   // storage cells per component
	int X, Y, Z, ox;   // local slab size (no ghosts), ghost planes per side
	int YZ;
	unsigned ydiv_mul, ydiv_shift;	// floor(n / Y) = umulhi(n, ydiv_mul) >> ydiv_shift for 0 <= n < 2^31 (ydiv_mul == 0: plain division)
	int x_begin;	   // first plane handled by this launch (local coordinate in [0, X)); the grid's y extent is the plane count
	int nb_begin, nb_end; // boundary-list range handled by this launch
	int wrap;		   // 1: the reference's nproc==1 rule (GEO_PERIODIC cells wrap), 0: ghost-plane rule
	int profile_sy;
	int tile_y, tile_y_shift;  // k_bulk_tma: cells of one row per CTA (power of two dividing Y) and its log2
	int eq, inflow, stream, out_mode, stat_counter;
	int kahan_rho;	   // 1: the operator is a build with USE_HIGH_PRECISION_RHO (k_initial_macro has no operator template argument)
	int pdl;		   // k_boundary under programmatic dependent launch: 1 = runs beside the bulk kernel launched just before it (see pdl_wait)
	Phys<R> phys;
	R in_vx, in_vy, in_vz;
};

template <typename R>
LBMX_D int div_by_Y(const KParams<R>& p, int n)
{
	return p.ydiv_mul ? (int) (__umulhi((unsigned) n, p.ydiv_mul) >> p.ydiv_shift) : n / p.Y;
}

// neighbour deltas in storage cells for one cell: kernelInitIndices (kernels.h:6-58)
struct Deltas
{
	int xm, xp, ym, yp, zm, zp;
};

template <bool AA, typename R>
LBMX_D Deltas neighbour_deltas(const KParams<R>& p, bool periodic_cell, int x, int y, int z)
{
	Deltas d;
	const int YZ = p.YZ, Y = p.Y;
	if (periodic_cell) {
		// wrap only in the 1-process rule; with ghost planes x+-1 always exists, and y/z are never cut
		d.xp = (p.wrap && x == p.X - 1) ? -(p.X - 1) * YZ : YZ;
		d.xm = (p.wrap && x == 0) ? (p.X - 1) * YZ : -YZ;
		// y and z are never cut by the slab decomposition: periodic cells always wrap there.  (With nproc > 1 the reference
		// would step out of the array instead, kernels.h:24-28 "TODO: use nproc_y and nproc_z"; identical wherever the
		// reference is well defined, i.e. for periodic cells away from the y/z faces.)
		d.yp = (y == p.Y - 1) ? -(p.Y - 1) : 1;
		d.ym = (y == 0) ? (p.Y - 1) : -1;
		d.zp = (z == p.Z - 1) ? -(p.Z - 1) * Y : Y;
		d.zm = (z == 0) ? (p.Z - 1) * Y : -Y;
#ifdef LBMX_EXP_YSHIFT4	 // development experiment only (wrong physics): y-neighbours 4 cells away = sector-aligned but line-misaligned accesses
		d.yp = (y >= p.Y - 4) ? -(p.Y - 4) : 4;
		d.ym = (y < 4) ? (p.Y - 4) : -4;
#endif
	}
	else if (AA) {
		d.xp = YZ;
		d.xm = -YZ;
		d.yp = 1;
		d.ym = -1;
		d.zp = Y;
		d.zm = -Y;
	}
	else {
		d.xp = (x == p.X - 1 + p.ox) ? 0 : YZ;
		d.xm = (x == -p.ox) ? 0 : -YZ;
		d.yp = (y == p.Y - 1) ? 0 : 1;
		d.ym = (y == 0) ? 0 : -1;
		d.zp = (z == p.Z - 1) ? 0 : Y;
		d.zm = (z == 0) ? 0 : -Y;
	}
	return d;
}

template <typename L, int WHICH = 0>
LBMX_D int dir_offset(const Deltas& d, int q, int sign)
{
	// storage offset of the neighbour in direction sign*c_q
	const int cx = sign * L::cx(q), cz = sign * L::cz(q);
	// development experiments only (tools/kbench.cu): wrong physics, isolate the cost of the +-1 shifts in y (WHICH: 1 = load, 2 = store)
#if defined(LBMX_EXP_NOYSHIFT)
	const int cy = 0;
#elif defined(LBMX_EXP_NOYSHIFT_LD)
	const int cy = WHICH == 1 ? 0 : sign * L::cy(q);
#elif defined(LBMX_EXP_NOYSHIFT_ST)
	const int cy = WHICH == 2 ? 0 : sign * L::cy(q);
#else
	const int cy = sign * L::cy(q);
#endif
	return (cx > 0 ? d.xp : cx < 0 ? d.xm : 0) + (cy > 0 ? d.yp : cy < 0 ? d.ym : 0) + (cz > 0 ? d.zp : cz < 0 ? d.zm : 0);
}

template <int HINT = LBMX_LD_HINT, typename R>
LBMX_D R ld_df(const R* ptr)
{
	if constexpr (HINT == 1)
		return __ldcs(ptr);
	else if constexpr (HINT == 2)
		return __ldcg(ptr);
	else if constexpr (HINT == 3)
		return __ldlu(ptr);
	else
		return *ptr;
}
template <int HINT = LBMX_ST_HINT, typename R>
LBMX_D void st_df(R* ptr, R v)
{
	if constexpr (HINT == 1)
		__stcs(ptr, v);
	else if constexpr (HINT == 2)
		__stcg(ptr, v);
	else if constexpr (HINT == 3)
		__stwt(ptr, v);
	else
		*ptr = v;
}
// A-A odd steps: the accesses of the populations that move in y are shifted by one element, so the last 32-byte sector of a warp's
// request is the first sector of the next warp's (same CTA, same row): there the plain, L1-allocating forms win over the streaming
// hints of the even / A-B kernels (tools/kbench sweeps of round 2, profiles/kbench_r2_odd_hints.txt: D3Q27 fp64 +0.4 %, D3Q19 fp32 +2 %,
// D2Q9 fp32 +1 %, the rest neutral).  Same encoding as LBMX_LD_HINT / LBMX_ST_HINT; *_YSHIFT for the populations that move in y.
#ifndef LBMX_LD_HINT_YSHIFT
	#define LBMX_LD_HINT_YSHIFT 0
#endif
#ifndef LBMX_ST_HINT_YSHIFT
	#define LBMX_ST_HINT_YSHIFT 0
#endif
#ifndef LBMX_LD_HINT_ODD
	#define LBMX_LD_HINT_ODD 0
#endif
#ifndef LBMX_ST_HINT_ODD
	#define LBMX_ST_HINT_ODD 0
#endif

// ---- programmatic dependent launch (small lattices: a step is tens of microseconds, the launch gap between two steps 2-3 of them) ----
// Kernels of a step chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel may start while its
// predecessor in the stream drains, and pdl_wait() holds it until that predecessor has completed and its stores are visible.
// Launched without the attribute both instructions do nothing.
LBMX_D void pdl_wait()
{
#ifdef __CUDA_ARCH__
	asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
LBMX_D void pdl_trigger()
{
#ifdef __CUDA_ARCH__
	asm volatile("griddepcontrol.launch_dependents;");
#endif
}

// ---- streaming (d3q27/streaming_AB.h:12-58, streaming_AA.h:12-116 and the D2Q9 twins) ----
// NONNEG: the caller guarantees cell + offset >= 0 (bulk kernel: offsets wrap), so the index is zero-extended for free.
// The boundary kernel keeps signed indices: under A-A a non-periodic cell on an unghosted face addresses x-1 / x+1 outside
// its slice exactly as the reference does (kernels.h:31-38) -- that lands in the neighbouring population's slice, never outside
// the allocation (slot 0 has no shift, the last slot only shifts towards lower x), so it is memory-safe though ill-defined.
template <bool NONNEG>
LBMX_D long long cell_index(int i)
{
	if constexpr (NONNEG)
		return (long long) (unsigned) i;
	else
		return (long long) i;
}

template <typename L, int MODE, bool NONNEG = false, typename R>
LBMX_D void stream_in(const KParams<R>& p, R (&f)[L::Q], int c, const Deltas& d)
{
	static_for<L::Q>([&](auto qc) {
		constexpr int q = qc;
		if constexpr (MODE == S_AB)
			f[q] = __ldg(p.rd[q] + cell_index<NONNEG>(c + dir_offset<L, 1>(d, q, -1)));
		else if constexpr (MODE == S_AA_EVEN)
			f[q] = ld_df(p.rd[q] + cell_index<NONNEG>(c));
		else
			f[L::opp(q)] = ld_df<(L::cy(q) != 0 ? LBMX_LD_HINT_YSHIFT : LBMX_LD_HINT_ODD)>(p.rd[q] + cell_index<NONNEG>(c + dir_offset<L, 1>(d, q, +1)));
	});
}

template <typename L, int MODE, bool NONNEG = false, typename R>
LBMX_D void stream_out(const KParams<R>& p, const R (&f)[L::Q], int c, const Deltas& d)
{
	static_for<L::Q>([&](auto qc) {
		constexpr int q = qc;
		if constexpr (MODE == S_AB)
			st_df(p.wr[q] + cell_index<NONNEG>(c), f[q]);
		else if constexpr (MODE == S_AA_EVEN)
			st_df(p.wr[L::opp(q)] + cell_index<NONNEG>(c), f[q]);
		else
			st_df<(L::cy(q) != 0 ? LBMX_ST_HINT_YSHIFT : LBMX_ST_HINT_ODD)>(p.wr[q] + cell_index<NONNEG>(c + dir_offset<L, 2>(d, q, +1)), f[q]);
	});
}

// ---- macroscopic output (d3q27/macro.h:50-171, d2q9/macro.h:49-140) ----
template <typename R>
LBMX_D R ld_macro(const R* ptr)
{
#if LBMX_MACRO_HINT == 1
	return __ldcs(ptr);
#else
	return *ptr;
#endif
}
template <typename R>
LBMX_D void st_macro(R* ptr, R v)
{
#if LBMX_MACRO_HINT == 1
	__stcs(ptr, v);
#else
	*ptr = v;
#endif
}

// The read-modify-write fields of MACRO_Mean / D2Q9_MACRO_WithMean are needed only after the collision, at the end of a thread's life:
// asking L2 for them together with the populations takes one DRAM latency off that tail without holding registers for them.
template <typename R>
LBMX_D void prefetch_l2(const R* ptr)
{
#ifdef __CUDA_ARCH__
	asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#endif
}
template <typename L, typename R>
LBMX_D void prefetch_macro_sums(const KParams<R>& p, int c)
{
	constexpr int nd = L::NDIM;
	const R* M = p.macro;
	const long long S = p.XYZ;
	// unrolled on purpose: rolled loops (or an out-of-line call) here cost the fp32 A-A odd kernel 10 % (profiles/kbench_r2_prefetch.txt)
	if (p.out_mode == OUT_MEAN) {
#pragma unroll
		for (int i = 1 + nd; i < 1 + 2 * nd + (nd == 3 ? 6 : 3); i++)
			prefetch_l2(M + (i * S + c));
	}
	else if (nd == 2 && p.out_mode >= OUT_WITH_MEAN_2D) {
		if (p.out_mode & 1) {
			prefetch_l2(M + (3 * S + c));
			prefetch_l2(M + (4 * S + c));
		}
		if (p.out_mode & 2) {
#pragma unroll
			for (int i = 5; i < 10; i++)
				prefetch_l2(M + (i * S + c));
		}
	}
}

template <typename L, typename R>
LBMX_D void output_macro_impl(R* M, const long long S, const int out_mode, const int stat_counter, int c, R rho, R vx, R vy, R vz)
{
	if (out_mode == OUT_NONE)
		return;
	constexpr int nd = L::NDIM;
	const R v[3] = {vx, vy, vz};
	st_macro(M + c, rho);
#pragma unroll
	for (int a = 0; a < nd; a++)
		st_macro(M + ((1 + a) * S + c), v[a]);
	if constexpr (nd == 2) {
		// D2Q9_MACRO_WithMean (sim_2D/sim2d_2.cu:75-95): gated velocity sums, and fluctuation sums about a mean the host froze
		if (out_mode >= OUT_WITH_MEAN_2D) {
			if (out_mode & 1) {
				st_macro(M + (3 * S + c), ld_macro(M + (3 * S + c)) + vx);
				st_macro(M + (4 * S + c), ld_macro(M + (4 * S + c)) + vy);
			}
			if (out_mode & 2) {
				const R dux = vx - ld_macro(M + (5 * S + c));
				const R duy = vy - ld_macro(M + (6 * S + c));
				st_macro(M + (7 * S + c), ld_macro(M + (7 * S + c)) + sqrt(dux * dux + duy * duy));  // IEEE sqrt in R = the reference's sqrt(double) rounded to dreal
				st_macro(M + (8 * S + c), ld_macro(M + (8 * S + c)) + dux * dux);
				st_macro(M + (9 * S + c), ld_macro(M + (9 * S + c)) + duy * duy);
			}
			return;
		}
	}
	if (out_mode != OUT_MEAN)
		return;
	// running mean + Welford co-moments, components: means then xx,yy,zz,xy,xz,yz (3-D) / xx,yy,xy (2-D)
	const R denom = R(1) / R(stat_counter + 1);
	R delta[3], delta_new[3];
#pragma unroll
	for (int a = 0; a < nd; a++) {
		const R old = ld_macro(M + ((1 + nd + a) * S + c));
		delta[a] = v[a] - old;
		const R now = old + delta[a] * denom;
		delta_new[a] = v[a] - now;
		st_macro(M + ((1 + nd + a) * S + c), now);
	}
	constexpr int np = nd == 3 ? 6 : 3;
	constexpr int pa[6] = {0, 1, nd == 3 ? 2 : 0, 0, 0, 1};
	constexpr int pb[6] = {0, 1, nd == 3 ? 2 : 1, 1, 2, 2};
#pragma unroll
	for (int i = 0; i < np; i++) {
		const long long o = (long long) (1 + 2 * nd + i) * S + c;
		st_macro(M + o, ld_macro(M + o) + delta_new[pa[i]] * delta[pb[i]]);
	}
}

template <typename L, typename R>
LBMX_D void output_macro(const KParams<R>& p, int c, R rho, R vx, R vy, R vz)
{
	output_macro_impl<L, R>(p.macro, p.XYZ, p.out_mode, p.stat_counter, c, rho, vx, vy, vz);
}
// full-way bounce-back: swap opposite populations, no collision (d3q27/bc.h:147-165).
// REFERENCE QUIRK kept for parity: in D2Q9_BC_All::preCollision the coordinate parameters zm/zp shadow the direction enumerators
// (d2q9/bc.h:90 vs defs.h:262-263), so its swap(f[zm], f[zp]) degenerates to a no-op on the X x Y x 1 lattice: the straight +-y
// populations are NOT bounced (nor mirrored by SYM_TOP/SYM_BOTTOM).
template <typename L, typename R>
LBMX_D void bounce_back(R (&f)[L::Q])
{
	static_for<L::Q>([&](auto qc) {
		constexpr int q = qc;
		constexpr int o = L::opp(q);
		if constexpr (o > q && ! (L::NDIM == 2 && L::cx(q) == 0)) {
			const R t = f[q];
			f[q] = f[o];
			f[o] = t;
		}
	});
}

// Which kernel owns a cell.  The bulk kernel takes GEO_FLUID and GEO_PERIODIC cells and -- so that obstacle-heavy maps (spheres,
// cylinders, porous blocks of GEO_WALL) stream every population exactly once -- GEO_WALL cells away from the lattice faces, where the
// wrapping neighbour rule it loads with coincides with the rule of a non-periodic cell (kernels.h:30-56), and GEO_NOTHING cells, which
// only report rho = 1, u = 0.  Everything else, walls on a face, and under A-B the GEO_FLUID cells on a face (they clamp instead of
// wrapping) go to the boundary list.  Shared by lbmx_map_upload (list construction) and the kernels.
LBMX_HD bool cell_on_face(int ndim, int ox, int X, int Y, int Z, int x, int y, int z)
{
	return (ox == 0 && (x == 0 || x == X - 1)) || y == 0 || y == Y - 1 || (ndim == 3 && (z == 0 || z == Z - 1));
}
LBMX_HD bool cell_in_boundary_list(int m, int fluid, int periodic, int wall, int nothing, bool face, bool ab)
{
	if (m == fluid)
		return ab && face;	// A-B: a GEO_FLUID cell on a lattice face clamps its neighbour indices (kernels.h:49-56): not the rule the bulk kernel loads with
	return ! (m == periodic || m == nothing || (m == wall && ! face));
}

// =====================================================================================================================
// bulk kernel: GEO_FLUID / GEO_PERIODIC cells, GEO_WALL cells away from the faces
// =====================================================================================================================
// what GEO_WALL and GEO_NOTHING cells report: rho = 1, u = 0 (d3q27/bc.h:53-60, 147-165), out of line
template <typename L, typename R>
__device__ __noinline__ void output_macro_at_rest(R* M, const long long S, const int out_mode, const int stat_counter, int c)
{
	output_macro_impl<L, R>(M, S, out_mode, stat_counter, c, R(1), R(0), R(0), R(0));
}
// How the bulk kernel carries its cold path (phase 3), from the tools/kbench comparison of round 2 (profiles/kbench_r2_cold_path.txt): as a
// call to bulk_cold_cell everywhere except the A-A odd kernels of D3Q27 and D2Q9, which lose 4-5 % with the call in them and nothing with
// the same code in line (the fp32 kernels with two cells per thread and the D3Q19 A-B kernel lose 6-10 % with it in line).
#ifndef LBMX_COLD_INLINE_ODD_Q27
	#define LBMX_COLD_INLINE_ODD_Q27 1
#endif
// ... in fp64 only: by the end of round 2 the fp32 A-A odd kernel of the D3Q27 cumulant ran at 5.02 TB/s with the cold path in line and at
// 5.71 TB/s with the call (profiles/kbench_r2_f32_odd_cold_path.txt)
#ifndef LBMX_COLD_INLINE_ODD_F32
	#define LBMX_COLD_INLINE_ODD_F32 0
#endif
template <typename L, typename R, int MODE>
constexpr bool bulk_cold_inline()
{
#ifdef LBMX_COLD_INLINE_ALL
	return LBMX_COLD_INLINE_ALL;
#else
	return LBMX_COLD_INLINE_ODD_Q27 && MODE == S_AA_ODD && L::Q != 19 && (sizeof(R) == 8 || LBMX_COLD_INLINE_ODD_F32);
#endif
}

// Which bulk kernels ask L2 for the MACRO_Mean / WithMean sums while the populations load (prefetch_macro_sums: MACRO_Mean A-A 8.2 -> 10.2
// GLUPS, i.e. 78 % -> 100 % of the measured HBM rate on 640 B per update).  The few instructions cost kernels that never use them nothing
// on D3Q27 cumulant fp64 / fp32, but 5-9 % on the D3Q19 A-B and the KBC A-A odd kernels (code placement; profiles/kbench_r2_prefetch.txt),
// and neither has a mean-collecting macro class in the reference: left out there.
template <typename L, int KIND>
constexpr bool bulk_prefetches_macro()
{
	return L::Q != 19 && ! (KIND >= K_KBC_N1 && KIND <= K_KBC_C4);
}

// Cold path of the bulk kernel, out of line so that the fluid path keeps its size (instruction cache: the fp32 kernels with two cells per
// thread sit right below the 32 KB where fetch starts to cost) and its registers: obstacle cells away from the faces bounce back without
// colliding, inert cells only report rho = 1, u = 0 (d3q27/bc.h:53-60, 147-165).  The populations are loaded again here rather than handed
// over in registers (they are in L2: the calling warp has just fetched these lines).
template <typename L, int MODE, typename R>
__device__ __noinline__ void bulk_cold_cell(const KParams<R>& p, int x, int yz, int m)
{
	const int c = (x + p.ox) * p.YZ + yz;
	if (m == L::WALL) {
		const int z = div_by_Y(p, yz);
		const int y = yz - z * p.Y;
		if (cell_on_face(L::NDIM, p.ox, p.X, p.Y, p.Z, x, y, z))
			return;	 // a boundary-list cell
		const Deltas d = neighbour_deltas<true>(p, true, x, y, z);
		R f[L::Q];
		stream_in<L, MODE, true>(p, f, c, d);
		bounce_back<L>(f);
		stream_out<L, MODE, true>(p, f, c, d);
	}
	output_macro_impl<L, R>(p.macro, p.XYZ, p.out_mode, p.stat_counter, c, R(1), R(0), R(0), R(0));
}

// A-B bulk kernel: do obstacle / inert lanes store together with their warp (phase 2)?  Measured per kernel on a map without obstacles
// (profiles/kbench_r2_ab_whole_sectors_per_kernel.txt): free or slightly positive on the D3Q27 kernels, D3Q19 MRT_LES and D2Q9 CLBM fp64,
// but the SRT kernels of D3Q19 and D2Q9 lose 6-26 % to the extra path (D3Q19 fp64 5.13 -> 6.09 TB/s without it, fp32 4.67 -> 5.90; D2Q9 fp64
// 5.43 -> 6.16, fp32 5.24 -> 5.55): there those lanes take the cold path of phase 3 like under A-A.
template <typename L, int KIND, typename R>
constexpr bool ab_whole_sectors()
{
	return LBMX_AB_WHOLE_SECTORS && ! (L::Q != 27 && KIND == K_SRT);
}

// resident CTAs per SM the register allocation is sized for: the cumulant / MRT_LES kernels fit 128 (A-A) and 96 (A-B)
// registers without spilling; fp64 SRT and BGK keep f[27], feq[27] and the source terms live and get 170
template <int KIND, typename R, int MODE, int L_Q = 27>
constexpr int bulk_minblocks()
{
	if (KIND >= K_KBC_N1 && KIND <= K_KBC_C4)
	{
		// fp32 A-B, measured model by model (profiles/kbench_r2_kbc_f32_ab_per_model.txt): the models whose shear part includes the trace
		// (N2, N4, C2, C4) run 11-13 % faster at 96 registers / 5 CTAs, the others (N1, N3, C1, C3) 14 % faster at 128 / 4
		constexpr bool with_trace = (KIND - K_KBC_N1) % 4 == 1 || (KIND - K_KBC_N1) % 4 == 3;
		return sizeof(R) == 8 ? (MODE == S_AA_ODD ? LBMX_KBC_MINBLOCKS_F64_ODD : LBMX_KBC_MINBLOCKS_F64)
							  : (MODE == S_AB && with_trace ? LBMX_KBC_MINBLOCKS_F32_AB : LBMX_KBC_MINBLOCKS_F32);
	}
	// (the default-arithmetic cascaded operator is as light as the cumulant one; its parity-arithmetic form keeps 27 moments of each kind live)
	// (D3Q19 SRT in its default-arithmetic form fits 128 / 96 registers: 4 / 5 CTAs run the chained A-A steps at 6.39 instead of 5.69 TB/s,
	// profiles/kbench_r2_srt_reorganised.txt; D3Q27 SRT is best at 3)
	if (sizeof(R) == 8 && ((KIND == K_SRT && (kStrict || L_Q != 19)) || KIND == K_BGK || KIND == K_BGK_GAL || KIND == K_SRT_MF || (KIND == K_CLBM && (kStrict || L_Q != 27))))
		return LBMX_BULK_MINBLOCKS < 3 ? LBMX_BULK_MINBLOCKS : 3;
	return MODE == S_AB ? LBMX_BULK_MINBLOCKS_AB : LBMX_BULK_MINBLOCKS;
}

// cells per thread: a thread of the bulk kernel handles CPT cells, BLOCK cells apart (so every warp access still covers 32
// consecutive cells).  All CPT x Q loads are issued before the first collision starts: more bytes in flight per SM at the same
// occupancy, which is what the narrow cases need (fp32: 128 B per warp access; D2Q9: 9 loads per cell).
// Chosen per lattice / precision / streaming mode from the tools/kbench sweep (profiles/kbench_r1_cpt.txt).
template <typename L, typename R, int MODE>
constexpr int bulk_cpt()
{
#ifdef LBMX_BULK_CPT
	return LBMX_BULK_CPT;
#else
	// round 2 (profiles/kbench_r2_cpt.txt), after the A-B kernels lost their face re-load: fp32 A-B is best with one cell per thread
	// (D3Q19 MRT 5.83 -> 6.35 TB/s, D3Q27 cumulant 5.25 -> 6.16 TB/s), D2Q9 fp32 A-A odd as well (4.83 -> 5.09 TB/s)
	if (L::Q >= 19)
		return sizeof(R) == 8 ? 1 : (MODE == S_AA_EVEN ? 2 : 1);
	return sizeof(R) == 8 ? (MODE == S_AB ? 1 : 2) : (MODE == S_AA_ODD ? 1 : 2);
#endif
}

// ARITH (= LBMX_STRICT of the object file) only makes the kernel symbols of the fast and the parity-arithmetic builds distinct
template <typename L, int KIND, typename R, int MODE, int ARITH = LBMX_STRICT>
__global__ void __launch_bounds__(LBMX_BULK_BLOCK, bulk_minblocks<KIND, R, MODE, L::Q>()) k_bulk(const LBMX_GRID_CONSTANT KParams<R> p)
{
	constexpr int CPT = bulk_cpt<L, R, MODE>();
	const int x = p.x_begin + blockIdx.y;
	const int yz0 = blockIdx.x * (LBMX_BULK_BLOCK * CPT) + threadIdx.x;
	R f[CPT][L::Q];
	Deltas d[CPT];
	int c[CPT], m[CPT];
	bool face[CPT];
	pdl_wait();		// everything the previous kernel of the chain wrote is visible from here on
	pdl_trigger();	// the next kernel may take the SM slots this grid frees while it drains
	if (p.inert != nullptr) {
		// maps with sizeable GEO_NOTHING regions (lbmx_map_upload decides): a CTA whose cells are all inert issues no population loads
		bool all_inert = true;
#pragma unroll
		for (int k = 0; k < CPT; k++) {
			const int first = blockIdx.x * (LBMX_BULK_BLOCK * CPT) + k * LBMX_BULK_BLOCK;
			if (first < p.YZ)
				all_inert = all_inert && p.inert[(long long) (x + p.ox) * p.inert_stride + first / LBMX_BULK_BLOCK] != 0;
		}
		if (all_inert) {
			if (p.out_mode != OUT_NONE) {
#pragma unroll
				for (int k = 0; k < CPT; k++)
					if (yz0 + k * LBMX_BULK_BLOCK < p.YZ)
						bulk_cold_cell<L, MODE>(p, x, yz0 + k * LBMX_BULK_BLOCK, L::NOTHING);
			}
			return;
		}
	}
	// ---- phase 1: every load of every cell of this thread
#pragma unroll
	for (int k = 0; k < CPT; k++) {
		const int yz = yz0 + k * LBMX_BULK_BLOCK;
		m[k] = -1;
		if (yz < p.YZ) {
			const int z = div_by_Y(p, yz);
			const int y = yz - z * p.Y;
			c[k] = (x + p.ox) * p.YZ + yz;
			m[k] = p.map[c[k]];
			// The neighbour offsets are formed with the periodic (wrapping) rule, which does not depend on the cell type and is
			// always in bounds: the Q population loads are in flight before the cell-type load has returned (one memory latency
			// off the critical path of a latency-bound kernel).  A-A: GEO_FLUID cells on an unghosted domain face thereby wrap
			// like GEO_PERIODIC ones; the reference leaves that case undefined (it steps out of the array: kernels.h:31-38,
			// SURVEY.md App. A).  A-B: such cells clamp in the reference (kernels.h:49-56) -- they belong to the boundary list.
			d[k] = neighbour_deltas<true>(p, true, x, y, z);
			face[k] = cell_on_face(L::NDIM, p.ox, p.X, p.Y, p.Z, x, y, z);
			stream_in<L, MODE, true>(p, f[k], c[k], d[k]);
			if constexpr (bulk_prefetches_macro<L, KIND>()) {
				if (p.out_mode >= OUT_MEAN)
					prefetch_macro_sums<L>(p, c[k]);
			}
		}
	}
	// ---- phase 2: collide and store, cell by cell
#pragma unroll
	for (int k = 0; k < CPT; k++) {
		if (m[k] < 0)
			continue;
		R rho, vx, vy, vz;
		if (L::bulk(m[k]) && ! (MODE == S_AB && face[k] && m[k] == L::FLUID)) {	 // (A-B: fluid cells on a face clamp -- boundary list)
#ifdef LBMX_EXP_NOCOLLIDE  // development experiment only: streaming without arithmetic = the memory-system ceiling of this access pattern
			rho = f[k][0];
			vx = vy = vz = R(0);
#else
			density_velocity<KIND == K_CUM_HP_RHO>(f[k], p.phys, rho, vx, vy, vz);
			collide<KIND, MODE != S_AB>(f[k], p.phys, p.eq, rho, vx, vy, vz);
#endif
		}
		else {
			if constexpr (MODE != S_AB || ! ab_whole_sectors<L, KIND, R>())
				continue;  // A-A: phase 3
			else {
				// A-B writes the OTHER array: a 32-byte sector that a warp's store covers only in part is not in L2 and costs DRAM a
				// read-modify-write -- two such lanes per lattice row (the wall / GEO_NOTHING skin of a duct) slow the whole kernel by
				// 8-14 % (profiles/solid_maps_r2_ab_features.txt).  So the obstacle and inert lanes store TOGETHER with the fluid lanes
				// of their warp: walls their bounced populations, inert cells the values the other array already holds there (which
				// the reference never touches: bc.h:53-60).
				if (m[k] == L::WALL && ! face[k])
					bounce_back<L>(f[k]);
				else if (m[k] == L::NOTHING)
					static_for<L::Q>([&](auto qc) { f[k][qc] = p.wr[qc][cell_index<true>(c[k])]; });
				else
					continue;  // a boundary-list cell
				rho = R(1);
				vx = vy = vz = R(0);
			}
		}
#ifdef LBMX_EXP_ODD_RECOMPUTE_OFFSETS  // tools/kbench experiment (no gain: ptxas already re-forms them; the barrier only cost registers -- cumulant fp64 odd 118 -> 128, fp32 72 -> 96):
		// the store addresses of the A-A odd step formed again from the six deltas instead of living through the collision
		if constexpr (MODE == S_AA_ODD)
			asm volatile("" : "+r"(d[k].xp), "+r"(d[k].xm), "+r"(d[k].yp), "+r"(d[k].ym), "+r"(d[k].zp), "+r"(d[k].zm), "+r"(c[k]));
#endif
		stream_out<L, MODE, true>(p, f[k], c[k], d[k]);
		output_macro<L>(p, c[k], rho, vx, vy, vz);
	}
	// ---- phase 3 (cold; A-A, where stores land on sectors the warp has just read): obstacle cells away from the faces, inert cells
#pragma unroll
	for (int k = 0; k < CPT; k++) {
		if (MODE == S_AB && ab_whole_sectors<L, KIND, R>())
			break;
		if (m[k] != L::WALL && m[k] != L::NOTHING)
			continue;
		const int yz = yz0 + k * LBMX_BULK_BLOCK;
		if constexpr (! bulk_cold_inline<L, R, MODE>())
			bulk_cold_cell<L, MODE>(p, x, yz, m[k]);
		else {
			// same thing in line, on the populations this thread already holds
			const int cc = (x + p.ox) * p.YZ + yz;
			if (m[k] == L::WALL) {
				const int z = div_by_Y(p, yz);
				const int y = yz - z * p.Y;
				if (cell_on_face(L::NDIM, p.ox, p.X, p.Y, p.Z, x, y, z))
					continue;
				const Deltas dd = neighbour_deltas<true>(p, true, x, y, z);
				bounce_back<L>(f[k]);
				stream_out<L, MODE, true>(p, f[k], cc, dd);
			}
			output_macro_at_rest<L, R>(p.macro, p.XYZ, p.out_mode, p.stat_counter, cc);
		}
	}
}

// =====================================================================================================================
// boundary kernel: every other cell type, run-time streaming mode (the slow, rare path)
// =====================================================================================================================
template <typename L, typename R>
LBMX_D R load_df(const KParams<R>& p, int q, int c)
{
	return p.cur[q * p.XYZ + c];
}

// moment inflow condition on the left face (Eichler 2024), d3q27/bc.h:82-136: rebuilds the nine +x populations
template <typename R>
LBMX_D void inflow_left_moments(R (&f)[27], R& rho, R vx, R vy, R vz)
{
	using L = D3Q27;
#define FQ(a, b, c) f[L::find(a, b, c)]
	const R ring0 = ((FQ(0, 1, 1) + FQ(0, -1, -1)) + (FQ(0, 1, -1) + FQ(0, -1, 1))) + ((FQ(0, 1, 0) + FQ(0, -1, 0)) + (FQ(0, 0, 1) + FQ(0, 0, -1)));
	const R ringm = ((FQ(-1, 1, 1) + FQ(-1, -1, -1)) + (FQ(-1, 1, -1) + FQ(-1, -1, 1))) + ((FQ(-1, 1, 0) + FQ(-1, -1, 0)) + (FQ(-1, 0, 1) + FQ(-1, 0, -1)));
	rho = R(1) / (R(1) - vx) * ((FQ(0, 0, 0) + ring0) + R(2) * (FQ(-1, 0, 0) + ringm));
	const R third = R(1.0 / 3.0);
	const R m100 = rho * vx, m010 = rho * vy, m001 = rho * vz;
	const R m011 = rho * (vy * vz);
	const R m020 = third * rho + rho * (vy * vy);
	const R m002 = third * rho + rho * (vz * vz);
	const R m021 = third * rho * vz + rho * ((vy * vy) * vz);
	const R m012 = third * rho * vy + rho * (vy * (vz * vz));
	const R m022 = R(1.0 / 9.0) * rho + third * rho * (vy * vy + vz * vz) + rho * (vy * vy) * (vz * vz);
	FQ(1, 0, 0) = (((m100 + (m022 - (m020 + m002))) + FQ(-1, 0, 0)) + ring0) + R(2) * ringm;
	FQ(1, 1, 0) = R(0.5) * ((m020 - m022) + (-m012 + m010)) - (FQ(-1, 1, 0) + FQ(0, 1, 0));
	FQ(1, -1, 0) = R(0.5) * ((m020 - m022) + (m012 - m010)) - (FQ(-1, -1, 0) + FQ(0, -1, 0));
	FQ(1, 0, 1) = R(0.5) * ((m002 - m022) + (-m021 + m001)) - (FQ(-1, 0, 1) + FQ(0, 0, 1));
	FQ(1, 0, -1) = R(0.5) * ((m002 - m022) + (m021 - m001)) - (FQ(-1, 0, -1) + FQ(0, 0, -1));
	FQ(1, 1, 1) = R(0.25) * ((m022 + m011) + (m021 + m012)) - (FQ(-1, 1, 1) + FQ(0, 1, 1));
	FQ(1, 1, -1) = R(0.25) * ((m022 - m011) + (-m021 + m012)) - (FQ(-1, 1, -1) + FQ(0, 1, -1));
	FQ(1, -1, 1) = R(0.25) * ((m022 - m011) + (m021 - m012)) - (FQ(-1, -1, 1) + FQ(0, -1, 1));
	FQ(1, -1, -1) = R(0.25) * ((m022 + m011) + (-m021 - m012)) - (FQ(-1, -1, -1) + FQ(0, -1, -1));
#undef FQ
}
template <typename R>
LBMX_D void inflow_left_moments(R (&)[9], R&, R, R, R)
{}
template <typename R>
LBMX_D void inflow_left_moments(R (&f)[19], R& rho, R vx, R vy, R vz)
{
	// the moment condition of d3q27/bc.h:82-136 is specific to 27 velocities; on D3Q19 GEO_INFLOW_LEFT imposes the equilibrium
	// of (rho = 1, inflow velocity), i.e. it behaves like GEO_INFLOW followed by a collision
	rho = R(1);
	R feq[19];
	equilibrium(feq, 0, rho, vx, vy, vz);
	static_for<19>([&](auto qc) { f[qc] = feq[qc]; });
}

// D2Q9 GEO_FLUID_NEAR_WALL: Bouzidi interpolated bounce-back (d2q9/bc.h:61-87,140-167), A-B only.  theta < 0: link does not hit
// a wall.  Runs after the ordinary pull has filled f.  REFERENCE QUIRK kept: the statements for the straight +-y links are
// written with the shadowed names zp/zm (bc.h:153,155) and end up assigning f[0], which the rest-particle line overwrites --
// so only the six links with an x component are interpolated.
template <typename R>
LBMX_D void bouzidi_near_wall(const KParams<R>& p, R (&f)[9], int c, const Deltas& d)
{
	using L = D2Q9;
	auto theta = [&](int dir) -> R { return p.bouzidi ? p.bouzidi[dir * p.XYZ + c] : R(-1); };
	auto fb = [&](R th, int k, int kbar, int offB, int offS) -> R {
		if (th < R(0))
			return p.cur[kbar * p.XYZ + (c + offS)];
		const R fA = p.cur[k * p.XYZ + c], fOppA = p.cur[kbar * p.XYZ + c], fB = p.cur[k * p.XYZ + (c + offB)];
		if (th <= R(0.5))
			return R(2) * th * fA + (R(1) - R(2) * th) * fB;
		const R w = R(0.5) / th;
		return (R(1) - w) * fOppA + w * fA;
	};
	// coefficient order: 0 E, 1 N, 2 W, 3 S, 4 NE, 5 NW, 6 SW, 7 SE
	f[L::find(1, 0)] = fb(theta(2), L::find(-1, 0), L::find(1, 0), d.xp, d.xm);
	f[L::find(-1, 0)] = fb(theta(0), L::find(1, 0), L::find(-1, 0), d.xm, d.xp);
	f[L::find(1, 1)] = fb(theta(6), L::find(-1, -1), L::find(1, 1), d.xp + d.yp, d.xm + d.ym);
	f[L::find(-1, 1)] = fb(theta(7), L::find(1, -1), L::find(-1, 1), d.xm + d.yp, d.xp + d.ym);
	f[L::find(-1, -1)] = fb(theta(4), L::find(1, 1), L::find(-1, -1), d.xm + d.ym, d.xp + d.yp);
	f[L::find(1, -1)] = fb(theta(5), L::find(-1, 1), L::find(1, -1), d.xp + d.ym, d.xm + d.yp);
	f[L::find(0, 0)] = p.cur[c];
}
template <typename R, int Q>
LBMX_D void bouzidi_near_wall(const KParams<R>&, R (&)[Q], int, const Deltas&)
{}

// symmetry planes: populations pointing in direction DST along AXIS take the value of their mirror image
// (d3q27/bc.h:172-237, d2q9/bc.h:168-191).  D2Q9 quirk: the straight +-y pair is addressed through shadowed names in the
// reference and ends up a no-op (see the wall rule below), so it is skipped here as well.
template <typename L, int AXIS, int DST, typename R>
LBMX_D void mirror_pops(R (&f)[L::Q])
{
	static_for<L::Q>([&](auto qc) {
		constexpr int q = qc;
		constexpr int cc = AXIS == 0 ? L::cx(q) : AXIS == 1 ? L::cy(q) : L::cz(q);
		if constexpr (cc == DST && ! (L::NDIM == 2 && AXIS == 1 && L::cx(q) == 0)) {
			constexpr int from = AXIS == 0 ? L::find(-L::cx(q), L::cy(q), L::cz(q)) : AXIS == 1 ? L::find(L::cx(q), -L::cy(q), L::cz(q)) : L::find(L::cx(q), L::cy(q), -L::cz(q));
			f[q] = f[from];
		}
	});
}

template <typename L, int KIND, typename R, int ARITH = LBMX_STRICT>
__global__ void __launch_bounds__(128, LBMX_BOUNDARY_MINBLOCKS) k_boundary(const KParams<R> p)
{
	const int i = p.nb_begin + blockIdx.x * blockDim.x + threadIdx.x;
	// Chained behind the bulk kernel of the same step on one stream.  pdl == 1: the two are independent (disjoint cells, one writer per
	// slot), so this kernel starts as soon as the bulk kernel's last CTAs are resident -- the bulk kernel triggers only after its own wait,
	// so everything older has completed -- and waits for it just before it exits: the next step's wait on THIS kernel then covers both.
	// Otherwise (A-A maps with GEO_OUTFLOW_RIGHT cells, which read a neighbour in place) it follows the whole bulk kernel.
	struct ChainEnd
	{
		bool at_end;
		LBMX_D ~ChainEnd()
		{
			if (at_end)
				pdl_wait();
		}
	} chain_end{p.pdl == 1};
	if (p.pdl != 1)
		pdl_wait();
	pdl_trigger();
	if (i >= p.nb_end)
		return;
	const int c = (int) p.blist[i];
	const int m = p.map[c];
	// recover (x,y,z) from the storage index
	const int xs = c / p.YZ;
	const int yz = c - xs * p.YZ;
	const int x = xs - p.ox;
	const int z = yz / p.Y;
	const int y = yz - z * p.Y;
	const bool aa = p.stream != S_AB;

	R rho = R(1), vx = R(0), vy = R(0), vz = R(0);
	if (m == L::NOTHING) {	// neither reads nor writes distributions (bc.h:53-60,252-253); reported rho=1, u=0
		output_macro<L>(p, c, rho, vx, vy, vz);
		return;
	}
	Deltas d = aa ? neighbour_deltas<true>(p, false, x, y, z) : neighbour_deltas<false>(p, false, x, y, z);
	const Deltas d_store = d;
	int c_load = c;
	if (m == L::OUTFLOW_RIGHT) {  // pull as if standing on the cell to the left: xp = x = xm (bc.h:63-65)
		c_load = c + d.xm;
		d.xp = 0;
		d.xm = 0;
	}
	R f[L::Q];
	if (m != L::OUTFLOW_RIGHT_INTERP) {
		if (p.stream == S_AB)
			stream_in<L, S_AB>(p, f, c_load, d);
		else if (p.stream == S_AA_EVEN)
			stream_in<L, S_AA_EVEN>(p, f, c_load, d);
		else
			stream_in<L, S_AA_ODD>(p, f, c_load, d);
	}
	else {
		// c_s-weighted interpolation of the populations entering through the right face (streaming_AB.h:209-242); A-B only
		constexpr R cs = R(0.5773502691896257);
		static_for<L::Q>([&](auto qc) {
			constexpr int q = qc;
			const int oyz = (L::cy(q) > 0 ? d.ym : L::cy(q) < 0 ? d.yp : 0) + (L::cz(q) > 0 ? d.zm : L::cz(q) < 0 ? d.zp : 0);
			if constexpr (L::cx(q) < 0)
				f[q] = cs * load_df<L>(p, q, c + d.xm + oyz) + (R(1) - cs) * load_df<L>(p, q, c + oyz);
			else
				f[q] = load_df<L>(p, q, c + (L::cx(q) > 0 ? d.xm : 0) + oyz);
		});
	}

	auto inflow_velocity = [&]() {
		if (p.inflow == 1) {
			vx = p.in_vx;
			vy = p.in_vy;
			if (L::NDIM == 3)
				vz = p.in_vz;
		}
		else if (p.inflow == 2) {
			vx = p.profile[y + z * p.profile_sy];
			vy = R(0);
			vz = R(0);
		}
		else if (p.inflow == 3) {  // Poiseuille profile over y (sim_2D/sim2d_3.cu:46-53); in_vy carries y0, in_vz carries 1 / (y1 - y0)
			R s = R(y - (int) p.in_vy) * p.in_vz;
			s = s < R(0) ? R(0) : (s > R(1) ? R(1) : s);
			vx = R(double(p.in_vx) * (4.0 * double(s) * (1.0 - double(s))));  // the reference's double literals: evaluated in double
			vy = R(0);
		}
		else {
			rho = R(1);
			vx = vy = vz = R(0);
		}
	};
	auto set_equilibrium = [&]() {
		R feq[L::Q];
		equilibrium_any(feq, p.eq, rho, vx, vy, vz);
		static_for<L::Q>([&](auto qc) { f[qc] = feq[qc]; });
	};
	if (m == L::INFLOW) {
		inflow_velocity();
		rho = R(1);
		set_equilibrium();
	}
	else if (m == L::INFLOW_LEFT) {
		inflow_velocity();
		inflow_left_moments(f, rho, vx, vy, vz);
	}
	else if (m == L::OUTFLOW_EQ) {
		density_velocity<KIND == K_CUM_HP_RHO>(f, p.phys, rho, vx, vy, vz);
		rho = R(1);
		set_equilibrium();
	}
	else if (m == L::OUTFLOW_RIGHT) {
		density_velocity<KIND == K_CUM_HP_RHO>(f, p.phys, rho, vx, vy, vz);
		rho = R(1);
	}
	else if (m == L::OUTFLOW_RIGHT_INTERP) {
		density_velocity<KIND == K_CUM_HP_RHO>(f, p.phys, rho, vx, vy, vz);
		R e1[L::Q], e0[L::Q];
		equilibrium_any(e1, p.eq, R(1), vx, vy, vz);
		equilibrium_any(e0, p.eq, rho, vx, vy, vz);
		static_for<L::Q>([&](auto qc) { f[qc] += e1[qc] - e0[qc]; });  // setEquilibriumDecomposition (common.h:94-124)
		rho = R(1);
	}
	else if (m == L::WALL)
		bounce_back<L>(f);
	else {
		if (m == L::SYM_TOP)
			mirror_pops<L, L::NDIM - 1, -1>(f);
		else if (m == L::SYM_BOTTOM)
			mirror_pops<L, L::NDIM - 1, +1>(f);
		else if (m == L::SYM_LEFT)
			mirror_pops<L, 0, +1>(f);
		else if (m == L::SYM_RIGHT)
			mirror_pops<L, 0, -1>(f);
		else if (m == L::SYM_BACK)
			mirror_pops<L, 1, +1>(f);
		else if (m == L::SYM_FRONT)
			mirror_pops<L, 1, -1>(f);
		else if (L::NDIM == 2 && m == 12 && p.stream == S_AB)  // D2Q9 GEO_FLUID_NEAR_WALL (d2q9/bc.h:29); plain fluid under A-A
			bouzidi_near_wall(p, f, c, d);
		density_velocity<KIND == K_CUM_HP_RHO>(f, p.phys, rho, vx, vy, vz);
	}

	if (L::collides(m))
		collide<KIND>(f, p.phys, p.eq, rho, vx, vy, vz);

	// stored at the true (x,y,z) also for OUTFLOW_RIGHT (kernels.h:97 passes the unmodified indices)
	if (p.stream == S_AB)
		stream_out<L, S_AB>(p, f, c, d_store);
	else if (p.stream == S_AA_EVEN)
		stream_out<L, S_AA_EVEN>(p, f, c, d_store);
	else
		stream_out<L, S_AA_ODD>(p, f, c, d_store);
	output_macro<L>(p, c, rho, vx, vy, vz);
}

// =====================================================================================================================
// initialisation / service kernels
// =====================================================================================================================
// LBM_BLOCK::setEquilibrium (lbm_block.hpp:219-250): every storage cell including ghosts; uniform state or per-cell fields
template <typename L, typename R, int ARITH = LBMX_STRICT>
__global__ void k_set_equilibrium(R* df, long long XYZ, long long n_cells, long long cell0, int eq, const double* rho, const double* vx, const double* vy,
								  const double* vz, double crho, double cvx, double cvy, double cvz)
{
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_cells)
		return;
	// the reference narrows the `real` (double) arguments to dreal at the call of EQ::eq_* (common.h:126-158)
	const R r = (R) (rho ? rho[i] : crho), ux = (R) (rho ? vx[i] : cvx), uy = (R) (rho ? vy[i] : cvy), uz = (R) (rho ? (vz ? vz[i] : 0.0) : cvz);
	R feq[L::Q];
	equilibrium_any(feq, eq, r, ux, uy, uz);
	static_for<L::Q>([&](auto qc) { df[qc * XYZ + cell0 + i] = feq[qc]; });
}

// LBM_BLOCK::computeInitialMacro (lbm_block.hpp:252-277): local read, force zeroed
template <typename L, typename R, int ARITH = LBMX_STRICT>
__global__ void k_initial_macro(const KParams<R> p)
{
	const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	const long long n = (long long) p.X * p.YZ;
	if (i >= n)
		return;
	const int c = (int) (i + (long long) p.ox * p.YZ);
	R f[L::Q];
	static_for<L::Q>([&](auto qc) { f[qc] = p.cur[qc * p.XYZ + c]; });
	Phys<R> ph = p.phys;
	ph.fx = ph.fy = ph.fz = R(0);
	R rho, vx, vy, vz;
	if (p.kahan_rho)
		density_velocity<true>(f, ph, rho, vx, vy, vz);
	else
		density_velocity<false>(f, ph, rho, vx, vy, vz);
	output_macro<L>(p, c, rho, vx, vy, vz);
}

// NaN scan of the density field (state.hpp:1170-1175)
template <typename R>
__global__ void k_has_nan(const R* rho, long long n, int* flag)
{
	bool bad = false;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		const R v = rho[i];
		bad |= (v != v);
	}
	if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0)
		atomicOr(flag, 1);
}

// copy `n_dirs` population planes (Y*Z reals each) between storage planes of one or two arrays: self halo exchange and packing
template <typename R>
__global__ void k_copy_planes(R* dst, const R* src, long long XYZ, int YZ, int n_dirs, const int* dirs, long long src_plane, long long dst_plane,
							  long long dst_XYZ = 0)  // dst_XYZ: component stride of the destination array when it is another slab's (0 = same as XYZ)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= YZ)
		return;
	const int q = dirs[blockIdx.y];
	dst[q * (dst_XYZ ? dst_XYZ : XYZ) + dst_plane * YZ + i] = src[q * XYZ + src_plane * YZ + i];
}

}  // namespace lbmx
#include "kernels_tma.cuh"
namespace lbmx {

// launcher table filled by the per-family translation units ---------------------------------------------------------
template <typename R>
struct StepKernels
{
	void (*bulk[3])(const KParams<R>);	// by StreamMode
	int cpt[3];						// cells per thread of the bulk kernels, by StreamMode
#if defined(__CUDACC__)
	void (*bulk_tma[3])(const KParams<R>);	// by StreamMode; A-A only ([S_AB] stays null)
#endif
	void (*boundary)(const KParams<R>);
	void (*initial_macro)(const KParams<R>);
	void (*set_equilibrium)(R*, long long, long long, long long, int, const double*, const double*, const double*, const double*, double, double, double, double);
};

template <typename L, int KIND, typename R>
StepKernels<R> make_step_kernels()
{
	StepKernels<R> k;
	k.bulk[S_AB] = k_bulk<L, KIND, R, S_AB>;
	k.bulk[S_AA_EVEN] = k_bulk<L, KIND, R, S_AA_EVEN>;
	k.bulk[S_AA_ODD] = k_bulk<L, KIND, R, S_AA_ODD>;
	k.cpt[S_AB] = bulk_cpt<L, R, S_AB>();
	k.cpt[S_AA_EVEN] = bulk_cpt<L, R, S_AA_EVEN>();
	k.cpt[S_AA_ODD] = bulk_cpt<L, R, S_AA_ODD>();
#if defined(__CUDACC__)
	k.bulk_tma[S_AB] = nullptr;
	k.bulk_tma[S_AA_EVEN] = k_bulk_tma<L, KIND, R, S_AA_EVEN>;
	k.bulk_tma[S_AA_ODD] = k_bulk_tma<L, KIND, R, S_AA_ODD>;
#endif
	k.boundary = k_boundary<L, KIND, R>;
	k.initial_macro = k_initial_macro<L, R>;
	k.set_equilibrium = k_set_equilibrium<L, R>;
	return k;
}

// one getter per (lattice, operator) family, defined in inst_*.cu; returns false if the precision is not R
bool get_kernels_d3q27_cum(StepKernels<float>&);
bool get_kernels_d3q27_cum(StepKernels<double>&);
bool get_kernels_d3q27_srt(StepKernels<float>&);
bool get_kernels_d3q27_srt(StepKernels<double>&);
bool get_kernels_d3q27_bgk(StepKernels<float>&);
bool get_kernels_d3q27_bgk(StepKernels<double>&);
bool get_kernels_d3q27_bgkgal(StepKernels<float>&);
bool get_kernels_d3q27_bgkgal(StepKernels<double>&);
bool get_kernels_d3q27_cumhp(StepKernels<float>&);
bool get_kernels_d3q27_cumhp(StepKernels<double>&);
bool get_kernels_d3q27_mrt(StepKernels<float>&);
bool get_kernels_d3q27_mrt(StepKernels<double>&);
bool get_kernels_d3q27_cum2017(StepKernels<float>&);
bool get_kernels_d3q27_cum2017(StepKernels<double>&);
bool get_kernels_d3q27_cum2017_strict(StepKernels<float>&);
bool get_kernels_d3q27_cum2017_strict(StepKernels<double>&);
bool get_kernels_d3q27_cumaa(StepKernels<float>&);
bool get_kernels_d3q27_cumaa(StepKernels<double>&);
bool get_kernels_d3q27_cumaa_strict(StepKernels<float>&);
bool get_kernels_d3q27_cumaa_strict(StepKernels<double>&);
bool get_kernels_d3q27_cum2017aa(StepKernels<float>&);
bool get_kernels_d3q27_cum2017aa(StepKernels<double>&);
bool get_kernels_d3q27_cum2017aa_strict(StepKernels<float>&);
bool get_kernels_d3q27_cum2017aa_strict(StepKernels<double>&);
bool get_kernels_d3q27_kbcn1(StepKernels<float>&);
bool get_kernels_d3q27_kbcn1(StepKernels<double>&);
bool get_kernels_d3q27_kbcn1_strict(StepKernels<float>&);
bool get_kernels_d3q27_kbcn1_strict(StepKernels<double>&);
bool get_kernels_d3q27_kbcn2(StepKernels<float>&);
bool get_kernels_d3q27_kbcn2(StepKernels<double>&);
bool get_kernels_d3q27_kbcn2_strict(StepKernels<float>&);
bool get_kernels_d3q27_kbcn2_strict(StepKernels<double>&);
bool get_kernels_d3q27_kbcn3(StepKernels<float>&);
bool get_kernels_d3q27_kbcn3(StepKernels<double>&);
bool get_kernels_d3q27_kbcn3_strict(StepKernels<float>&);
bool get_kernels_d3q27_kbcn3_strict(StepKernels<double>&);
bool get_kernels_d3q27_kbcn4(StepKernels<float>&);
bool get_kernels_d3q27_kbcn4(StepKernels<double>&);
bool get_kernels_d3q27_kbcn4_strict(StepKernels<float>&);
bool get_kernels_d3q27_kbcn4_strict(StepKernels<double>&);
bool get_kernels_d3q27_kbcc1(StepKernels<float>&);
bool get_kernels_d3q27_kbcc1(StepKernels<double>&);
bool get_kernels_d3q27_kbcc1_strict(StepKernels<float>&);
bool get_kernels_d3q27_kbcc1_strict(StepKernels<double>&);
bool get_kernels_d3q27_kbcc2(StepKernels<float>&);
bool get_kernels_d3q27_kbcc2(StepKernels<double>&);
bool get_kernels_d3q27_kbcc2_strict(StepKernels<float>&);
bool get_kernels_d3q27_kbcc2_strict(StepKernels<double>&);
bool get_kernels_d3q27_kbcc3(StepKernels<float>&);
bool get_kernels_d3q27_kbcc3(StepKernels<double>&);
bool get_kernels_d3q27_kbcc3_strict(StepKernels<float>&);
bool get_kernels_d3q27_kbcc3_strict(StepKernels<double>&);
bool get_kernels_d3q27_kbcc4(StepKernels<float>&);
bool get_kernels_d3q27_kbcc4(StepKernels<double>&);
bool get_kernels_d3q27_kbcc4_strict(StepKernels<float>&);
bool get_kernels_d3q27_kbcc4_strict(StepKernels<double>&);
bool get_kernels_d3q27_clbm(StepKernels<float>&);
bool get_kernels_d3q27_clbm(StepKernels<double>&);
bool get_kernels_d3q27_srtmf(StepKernels<float>&);
bool get_kernels_d3q27_srtmf(StepKernels<double>&);
// the same families in parity arithmetic (collide_strict.cuh, -fmad=false)
bool get_kernels_d3q27_clbm_strict(StepKernels<float>&);
bool get_kernels_d3q27_clbm_strict(StepKernels<double>&);
bool get_kernels_d3q27_srtmf_strict(StepKernels<float>&);
bool get_kernels_d3q27_srtmf_strict(StepKernels<double>&);
bool get_kernels_d3q27_cum_strict(StepKernels<float>&);
bool get_kernels_d3q27_cum_strict(StepKernels<double>&);
bool get_kernels_d3q27_srt_strict(StepKernels<float>&);
bool get_kernels_d3q27_srt_strict(StepKernels<double>&);
bool get_kernels_d3q27_bgk_strict(StepKernels<float>&);
bool get_kernels_d3q27_bgk_strict(StepKernels<double>&);
bool get_kernels_d3q27_bgkgal_strict(StepKernels<float>&);
bool get_kernels_d3q27_bgkgal_strict(StepKernels<double>&);
bool get_kernels_d3q27_cumhp_strict(StepKernels<float>&);
bool get_kernels_d3q27_cumhp_strict(StepKernels<double>&);
bool get_kernels_d3q27_mrt_strict(StepKernels<float>&);
bool get_kernels_d3q27_mrt_strict(StepKernels<double>&);
bool get_kernels_d2q9_srt_strict(StepKernels<float>&);
bool get_kernels_d2q9_srt_strict(StepKernels<double>&);
bool get_kernels_d2q9_clbm_strict(StepKernels<float>&);
bool get_kernels_d2q9_clbm_strict(StepKernels<double>&);
bool get_kernels_d3q19_srt(StepKernels<float>&);
bool get_kernels_d3q19_srt(StepKernels<double>&);
bool get_kernels_d3q19_mrt(StepKernels<float>&);
bool get_kernels_d3q19_mrt(StepKernels<double>&);
bool get_kernels_d2q9_srt(StepKernels<float>&);
bool get_kernels_d2q9_srt(StepKernels<double>&);
bool get_kernels_d2q9_clbm(StepKernels<float>&);
bool get_kernels_d2q9_clbm(StepKernels<double>&);

}  // namespace lbmx
