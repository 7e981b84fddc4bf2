// kernels_tma.cuh -- the A-A bulk kernel with its population traffic moved by the TMA engine's bulk copies (sm_100a, UBLKCP).
//
// Why: under A-A streaming a cell reads and writes the SAME 27 slots within one step (even: slot q of its own cell, odd: slot q of
// the neighbour in direction c_q; d3q27/streaming_AA.h:12-116).  A CTA that owns TY consecutive cells of a row therefore owns, per
// population, one contiguous run of TY reals -- shifted by one element for the 18 populations with c_y != 0 on odd steps.  The plain
// kernel (k_bulk) issues those as per-lane accesses: a warp's 256 B then straddle three 128-byte lines instead of two, and every
// access carries its own address arithmetic (profiles/ncu_r1_kbulk_cum_f64_512.txt: 212 vs 108 integer instructions, 990 M vs 914 M
// L1 load sectors on the odd step).  Here one thread issues Q `cp.async.bulk` copies of the rows into shared memory, the cells are
// collided out of / back into those rows, and one thread issues the Q row stores.
//
// What the copy engine cannot do is start on an 8-byte boundary: bulk copies need 16-byte aligned addresses and sizes, and the
// tensor-box form (cp.async.bulk.tensor) faults on this GPU as soon as the box start is not 16-byte aligned -- measured,
// profiles/tma_alignment_probe_r2.txt.  So a row shifted by +-1 element travels as its 16-byte aligned middle part (TY - E elements,
// E = 16 / sizeof(real)) through the copy engine, and its E leftover elements at the two ends through ordinary per-lane accesses of
// the lanes that own them (lane 0 and lane TY-1 in fp64).  Those lanes are also the ones whose y neighbour may lie across a periodic
// face, so the wrap comes for free; wrapped z / x neighbours are simply other rows / planes.
//
// Shared memory, population q: element i holds the slot of global y = y0 + (i - r TY - PAD) of tile row r (PAD = E), i.e. the thread of
// tile cell (r, t) finds "its" slot at i = r TY + t + c_y(q) + PAD whatever the shift, bank-conflict free; every aligned middle part
// starts on a 16-byte boundary, and the leftover elements of neighbouring rows share their 16-byte granules only with each other.
//
// Cells that are not GEO_FLUID / GEO_PERIODIC stay with k_boundary, which may run concurrently.  A row store rewrites every slot of
// the tile, so it is only legal when no other kernel changes a slot of the tile within the step: true for GEO_NOTHING (nobody touches
// those slots) and for 3-D GEO_WALL (full-way bounce-back under A-A leaves every slot as it was: swap after the pull, then the
// opposite store -- the boundary kernel rewrites identical bits).  A tile that holds any other cell type falls back, CTA-uniformly, to
// per-lane stores out of registers; its loads still come through the copy engine (loads never race).
#pragma once
#if defined(__CUDACC__)

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
		// watchdog, off the fast path: rows that never land must fail the launch loudly instead of hanging the GPU
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
// global -> shared, completion counted in bytes on `bar`; src, dst and bytes are multiples of 16
LBMX_D void load_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol)
{
	#if LBMX_TMA_LD_POLICY == 1
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
				 "r"(smem_u32(bar)), "l"(pol)
				 : "memory");
	#else
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
	#endif
}
LBMX_D void store_row(void* dst, const void* src, uint32_t bytes, uint64_t pol)
{
	#if LBMX_TMA_ST_POLICY == 1
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes), "l"(pol) : "memory");
	#else
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
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

template <typename L>
LBMX_HD constexpr int n_y_movers()
{
	int n = 0;
	for (int q = 0; q < L::Q; q++)
		n += L::cy(q) != 0;
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
// divides Y, 4 E <= TY <= 128, tma_host.h), TZ = 128 / TY rows per CTA (1 on every lattice with Y a multiple of 128).
template <typename L, int KIND, typename R, int MODE, int ARITH = LBMX_STRICT>
__global__ void __launch_bounds__(tma::TILE, tma_minblocks<KIND, R, MODE>()) k_bulk_tma(const __grid_constant__ KParams<R> p)
{
	constexpr int T = tma::TILE;
	constexpr bool ODD = MODE == S_AA_ODD;
	constexpr int E = 16 / (int) sizeof(R);	 // elements per 16 bytes: the granule of the copy engine
	constexpr int PAD = E;
	constexpr int PITCH = T + 2 * E;
	__shared__ alignas(128) R tile[L::Q][PITCH];
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
	// Tile row r of population q: storage cell of its first aligned element, that element's place in shared memory, the length of the
	// aligned part.  Odd steps address the neighbour in direction c_q: wrapped z / x are other rows / planes (the periodic rule, as in
	// k_bulk); a row shifted by +1 in y starts E elements in, a row shifted by -1 ends E elements early.
	auto row_cell = [&](int q, int r) -> long long {
		int zz = z0 + r + (ODD ? L::cz(q) : 0);
		zz = zz < 0 ? zz + p.Z : (zz >= p.Z ? zz - p.Z : zz);
		int xx = x + (ODD ? L::cx(q) : 0);
		if (p.wrap)
			xx = xx < 0 ? xx + p.X : (xx >= p.X ? xx - p.X : xx);
		return (long long) (xx + p.ox) * p.YZ + (long long) zz * p.Y + y0 + ((ODD && L::cy(q) > 0) ? E : 0);
	};
	auto row_smem = [&](int q, int r) -> R* { return &tile[q][(r << ty_shift) + PAD + ((ODD && L::cy(q) > 0) ? E : 0)]; };
	auto row_len = [&](int q) -> int { return (ODD && L::cy(q) != 0) ? TY - E : TY; };
	if (tid == 0) {
		constexpr int NYM = ODD ? tma::n_y_movers<L>() : 0;
		tma::mbar_expect_tx(&bar, (uint32_t) (rows * ((L::Q - NYM) * TY + NYM * (TY - E)) * (int) sizeof(R)));
		for (int r = 0; r < rows; r++)
			static_for<L::Q>([&](auto qc) {
				constexpr int q = qc;
				tma::load_row(row_smem(q, r), p.rd[q] + row_cell(q, r), (uint32_t) (row_len(q) * (int) sizeof(R)), &bar, pol);
			});
	}
	// every thread: its cell, the cell type, and the leftover elements of the shifted rows (ordinary loads by the lanes that own them,
	// parked in the shared-memory places the copy engine does not write)
	const int ty = tid & (TY - 1), tz = tid >> ty_shift;
	const int y = y0 + ty, z = z0 + tz;
	const bool in = tz < rows;
	const int slot = (tz << ty_shift) + ty + PAD;
	int c = 0, m = -1;
	Deltas d{};
	// is this thread's slot of a population shifted by +1 / -1 outside the aligned middle part?
	const bool left_over_p = ODD && (ty + 1 < E || ty + 1 >= TY), left_over_m = ODD && (ty - 1 < 0 || ty - 1 >= TY - E);
	if (in) {
		c = (x + p.ox) * p.YZ + z * p.Y + y;
		m = p.map[c];
		d = neighbour_deltas<true>(p, true, x, y, z);
		if constexpr (ODD) {
			if (left_over_p || left_over_m)
				static_for<L::Q>([&](auto qc) {
					constexpr int q = qc;
					if constexpr (L::cy(q) != 0) {
						if (L::cy(q) > 0 ? left_over_p : left_over_m)
							tile[q][slot + L::cy(q)] = ld_df(p.rd[q] + cell_index<true>(c + dir_offset<L, 1>(d, q, +1)));
					}
				});
		}
	}
	const bool is_bulk = in && L::bulk(m);
	const bool rewritable = ! in || is_bulk || m == L::NOTHING || (L::NDIM == 3 && m == L::WALL);
	const int fallback = __syncthreads_or(! rewritable);
	tma::mbar_wait(&bar, 0);
	R f[L::Q];
	static_for<L::Q>([&](auto qc) {
		constexpr int q = qc;
		f[ODD ? L::opp(q) : q] = tile[q][slot + (ODD ? L::cy(q) : 0)];
	});
	R rho = R(1), vx = R(0), vy = R(0), vz = R(0);
	if (is_bulk) {
		density_velocity<KIND == K_CUM_HP_RHO>(f, p.phys, rho, vx, vy, vz);
		collide<KIND>(f, p.phys, p.eq, rho, vx, vy, vz);
	}
	if (! fallback) {
		if (is_bulk) {
			static_for<L::Q>([&](auto qc) {
				constexpr int q = qc;
				constexpr int dst = ODD ? q : L::opp(q);  // the slot population q is stored to (odd: at the neighbour in direction c_q)
				if constexpr (ODD && L::cy(dst) != 0) {
					if (L::cy(dst) > 0 ? left_over_p : left_over_m)
						st_df(p.wr[dst] + cell_index<true>(c + dir_offset<L, 2>(d, dst, +1)), f[q]);
					else
						tile[dst][slot + L::cy(dst)] = f[q];
				}
				else
					tile[dst][slot] = f[q];
			});
		}
		tma::fence_generic_to_async_smem();
		__syncthreads();
		if (tid == 0) {
			for (int r = 0; r < rows; r++)
				static_for<L::Q>([&](auto qc) {
					constexpr int q = qc;
					tma::store_row(p.wr[q] + row_cell(q, r), row_smem(q, r), (uint32_t) (row_len(q) * (int) sizeof(R)), pol);
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
