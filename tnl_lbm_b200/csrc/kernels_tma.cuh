// kernels_tma.cuh -- the A-A bulk kernel with its population traffic moved by the tensor memory accelerator (sm_100a).
//
// Why: under A-A streaming a cell reads and writes the SAME 27 slots within one step (even: slot q of its own cell, odd: slot q of
// the neighbour in direction c_q; d3q27/streaming_AA.h:12-116).  A CTA that owns TY consecutive cells of a row therefore owns, per
// population, one contiguous run of TY reals -- shifted by one element for the 18 populations with c_y != 0 on odd steps.  The plain
// kernel (k_bulk) issues those as per-lane accesses: a warp's 256 B then straddle three 128-byte lines instead of two, and every
// access carries its own address arithmetic (profiles/ncu_r1_kbulk_cum_f64_512.txt: 212 vs 108 integer instructions, 990 M vs 914 M
// L1 load sectors on the odd step).  Here one thread issues Q `cp.async.bulk.tensor` loads of [TY]-element boxes at element-granular
// coordinates (y0 + c_y, z + c_z, x + c_x, q) into shared memory, the cells are collided out of / back into those boxes, and one thread
// issues the Q box stores: no per-lane global addressing at all, whole-line DRAM/L2 traffic, out-of-range box parts are zero-filled on
// load and dropped on store by the hardware.
//
// What the hardware does not do is wrap: the populations that cross the periodic y faces are moved by the two edge lanes of a row with
// ordinary loads/stores (issued before the wait on the boxes, so their latency overlaps), and wrapped z / x coordinates are simply the
// box coordinates of another row / plane.
//
// Cells that are not GEO_FLUID / GEO_PERIODIC stay with k_boundary, which may run concurrently.  A box store rewrites every slot of
// the tile, so it is only legal when no other kernel changes a slot of the tile within the step: true for GEO_NOTHING (nobody touches
// those slots) and for 3-D GEO_WALL (full-way bounce-back under A-A leaves every slot as it was: swap after the pull, then the
// opposite store -- the boundary kernel rewrites identical bits).  A tile that holds any other cell type falls back, CTA-uniformly, to
// per-lane stores out of registers; its loads still come through the boxes (loads never race).
#pragma once
#if defined(__CUDACC__)
	#include <cuda.h>

	#ifndef LBMX_TMA_MINBLOCKS
		#define LBMX_TMA_MINBLOCKS 4
	#endif
	#ifndef LBMX_TMA_LD_POLICY
		#define LBMX_TMA_LD_POLICY 1  // 0: no hint, 1: L2 evict_first
	#endif
	#ifndef LBMX_TMA_ST_POLICY
		#define LBMX_TMA_ST_POLICY 1
	#endif

namespace lbmx {
namespace tma {

constexpr int TILE = 128;  // cells (= threads) per CTA

LBMX_D uint32_t smem_u32(const void* ptr)
{
	return (uint32_t) __cvta_generic_to_shared(ptr);
}
LBMX_D void mbar_init(uint64_t* bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
LBMX_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
LBMX_D void mbar_wait(uint64_t* bar, uint32_t parity)
{
	uint32_t done;
	long long t0 = 0;
	for (;;) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if (done)
			break;
		// watchdog, off the fast path: boxes that never land (a bad tensor map) must fail the launch loudly instead of hanging the GPU
		const long long now = clock64();
		if (t0 == 0)
			t0 = now;
		else if (now - t0 > (4ll << 30))
			__trap();
	}
}
LBMX_D uint64_t policy_evict_first()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
LBMX_D void load_box(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar, uint64_t pol)
{
	#if LBMX_TMA_LD_POLICY == 1
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(smem_u32(dst)),
				 "l"((uint64_t) tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(pol)
				 : "memory");
	#else
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)), "l"((uint64_t) tm),
				 "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
				 : "memory");
	#endif
}
LBMX_D void store_box(const CUtensorMap* tm, int c0, int c1, int c2, int c3, const void* src, uint64_t pol)
{
	#if LBMX_TMA_ST_POLICY == 1
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4}], [%5], %6;" ::"l"((uint64_t) tm), "r"(c0), "r"(c1), "r"(c2),
				 "r"(c3), "r"(smem_u32(src)), "l"(pol)
				 : "memory");
	#else
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"((uint64_t) tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
				 "r"(smem_u32(src))
				 : "memory");
	#endif
}
LBMX_D void stores_commit_and_drain()
{
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
	asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may be released once the engine has read it
}
LBMX_D void fence_generic_to_async_smem()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// the k-th population (k = 0 .. n-1) whose y component is SIGN, as a compile-time table
template <typename L, int SIGN>
LBMX_HD constexpr int y_mover(int k)
{
	int n = 0;
	for (int q = 0; q < L::Q; q++)
		if (L::cy(q) == SIGN) {
			if (n == k)
				return q;
			n++;
		}
	return -1;
}
template <typename L>
LBMX_HD constexpr int n_y_movers()
{
	int n = 0;
	for (int q = 0; q < L::Q; q++)
		n += L::cy(q) > 0;
	return n;
}

}  // namespace tma

// resident CTAs per SM: what the plain kernel of the same operator is sized for (bulk_minblocks), at most LBMX_TMA_MINBLOCKS
template <int KIND, typename R, int MODE>
constexpr int tma_minblocks()
{
	return bulk_minblocks<KIND, R, MODE>() < LBMX_TMA_MINBLOCKS ? bulk_minblocks<KIND, R, MODE>() : LBMX_TMA_MINBLOCKS;
}

// MODE: S_AA_EVEN or S_AA_ODD.  Launch: 128 threads, grid = ((Y / TY) * ceil(Z / TZ), planes); TY = p.tile_y (a power of two that
// divides Y, 8 <= TY <= 128), TZ = 128 / TY rows per CTA (1 on every lattice with Y a multiple of 128).
template <typename L, int KIND, typename R, int MODE, int ARITH = LBMX_STRICT>
__global__ void __launch_bounds__(tma::TILE, tma_minblocks<KIND, R, MODE>()) k_bulk_tma(const __grid_constant__ KParams<R> p, const __grid_constant__ CUtensorMap tm)
{
	constexpr int T = tma::TILE;
	constexpr bool ODD = MODE == S_AA_ODD;
	constexpr int NYM = tma::n_y_movers<L>();
	__shared__ alignas(128) R tile[L::Q][T];
	__shared__ alignas(8) uint64_t bar;
	const int tid = threadIdx.x;
	const int x = p.x_begin + blockIdx.y;
	const int TY = p.tile_y, ty_shift = p.tile_y_shift, TZ = T >> ty_shift;
	const int tiles_y = p.Y >> ty_shift;
	const int tzi = blockIdx.x / tiles_y;
	const int y0 = (blockIdx.x - tzi * tiles_y) << ty_shift, z0 = tzi * TZ;
	const int rows = min(TZ, p.Z - z0);
	if (tid == 0)
		tma::mbar_init(&bar, 1);
	__syncthreads();
	uint64_t pol = 0;
	#if LBMX_TMA_LD_POLICY == 1 || LBMX_TMA_ST_POLICY == 1
	pol = tma::policy_evict_first();
	#endif
	// box coordinates of population q for tile row r: (y, z, x-storage, q); odd steps address the neighbour in direction c_q
	auto box_z = [&](int q, int r) {
		int zz = z0 + r + (ODD ? L::cz(q) : 0);
		zz = zz < 0 ? zz + p.Z : (zz >= p.Z ? zz - p.Z : zz);
		return zz;
	};
	auto box_x = [&](int q) {
		int xx = x + (ODD ? L::cx(q) : 0);
		if (p.wrap)
			xx = xx < 0 ? xx + p.X : (xx >= p.X ? xx - p.X : xx);
		return xx + p.ox;
	};
	if (tid == 0) {
		tma::mbar_expect_tx(&bar, (uint32_t) (rows * L::Q * TY * (int) sizeof(R)));
		for (int r = 0; r < rows; r++)
			static_for<L::Q>([&](auto qc) {
				constexpr int q = qc;
				tma::load_box(&tile[q][r << ty_shift], &tm, y0 + (ODD ? L::cy(q) : 0), box_z(q, r), box_x(q), q, &bar, pol);
			});
	}
	// every thread: its cell, the cell type, and -- on the two lanes of a row that sit on a y face -- the wrapped populations
	const int ty = tid & (TY - 1), tz = tid >> ty_shift;
	const int y = y0 + ty, z = z0 + tz;
	const bool in = tz < rows;
	int c = 0, m = -1;
	Deltas d{};
	R edge[NYM];
	const bool lo = ODD && in && y == 0, hi = ODD && in && y == p.Y - 1;
	if (in) {
		c = (x + p.ox) * p.YZ + z * p.Y + y;
		m = p.map[c];
		d = neighbour_deltas<true>(p, true, x, y, z);
		if (lo)
			static_for<NYM>([&](auto kc) {
				constexpr int q = tma::y_mover<L, -1>(kc);
				edge[kc] = ld_df(p.rd[q] + cell_index<true>(c + dir_offset<L, 1>(d, q, +1)));
			});
		else if (hi)
			static_for<NYM>([&](auto kc) {
				constexpr int q = tma::y_mover<L, +1>(kc);
				edge[kc] = ld_df(p.rd[q] + cell_index<true>(c + dir_offset<L, 1>(d, q, +1)));
			});
	}
	const bool is_bulk = in && L::bulk(m);
	const bool rewritable = ! in || is_bulk || m == L::NOTHING || (L::NDIM == 3 && m == L::WALL);
	const int fallback = __syncthreads_or(! rewritable);
	tma::mbar_wait(&bar, 0);
	R f[L::Q];
	static_for<L::Q>([&](auto qc) {
		constexpr int q = qc;
		f[ODD ? L::opp(q) : q] = tile[q][tid];
	});
	if (lo)
		static_for<NYM>([&](auto kc) { f[L::opp(tma::y_mover<L, -1>(kc))] = edge[kc]; });
	else if (hi)
		static_for<NYM>([&](auto kc) { f[L::opp(tma::y_mover<L, +1>(kc))] = edge[kc]; });
	R rho = R(1), vx = R(0), vy = R(0), vz = R(0);
	if (is_bulk) {
		density_velocity(f, p.phys, rho, vx, vy, vz);
		collide<KIND>(f, p.phys, p.eq, rho, vx, vy, vz);
	}
	if (! fallback) {
		if (is_bulk) {
			static_for<L::Q>([&](auto qc) {
				constexpr int q = qc;
				tile[ODD ? q : L::opp(q)][tid] = f[q];
			});
			if (lo)
				static_for<NYM>([&](auto kc) {
					constexpr int q = tma::y_mover<L, -1>(kc);
					st_df(p.wr[q] + cell_index<true>(c + dir_offset<L, 2>(d, q, +1)), f[q]);
				});
			else if (hi)
				static_for<NYM>([&](auto kc) {
					constexpr int q = tma::y_mover<L, +1>(kc);
					st_df(p.wr[q] + cell_index<true>(c + dir_offset<L, 2>(d, q, +1)), f[q]);
				});
		}
		tma::fence_generic_to_async_smem();
		__syncthreads();
		if (tid == 0) {
			for (int r = 0; r < rows; r++)
				static_for<L::Q>([&](auto qc) {
					constexpr int q = qc;
					tma::store_box(&tm, y0 + (ODD ? L::cy(q) : 0), box_z(q, r), box_x(q), q, &tile[q][r << ty_shift], pol);
				});
			tma::stores_commit_and_drain();
		}
	}
	else if (is_bulk)
		stream_out<L, MODE, true>(p, f, c, d);
	if (is_bulk)
		output_macro<L>(p, c, rho, vx, vy, vz);
}

}  // namespace lbmx
#endif	// __CUDACC__
