// engine.cu -- host side of liblbmx.so: the engine object behind the C ABI of include/lbmx.h.
//
// Replaces, for the hot path only, what LBM / LBM_BLOCK / State::SimUpdate do in the reference (include/lbm3d/lbm.hpp,
// lbm_block.hpp, state.hpp:980-1145): owns the device arrays of one x-slab, builds the launch plan from the cell-type
// map, advances time steps (even/odd parity, A-B rotation), and exchanges ghost planes with the neighbouring slabs
// (boundary planes first on a high-priority stream, NCCL send/recv on a communication stream, interior concurrently).
// No CPU fallback anywhere: without a CUDA device every call fails with LBMX_ERR_CUDA.
#include "../../include/lbmx.h"
#include "kernels.cuh"
#include "tma_host.h"

#include <cuda_runtime.h>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace lbmx;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
	g_err = msg;
	return code;
}

#define CU(call)                                                                                                                   \
	do {                                                                                                                           \
		cudaError_t err__ = (call);                                                                                                \
		if (err__ != cudaSuccess)                                                                                                  \
			return fail(LBMX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
	} while (0)

// ---- NCCL bound at run time so that the library loads (and the host-only helpers work) on machines without it, and so
//      that a process that already carries a libnccl.so.2 (e.g. torch's) shares that one copy --------------------------------
struct NcclApi
{
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool load(std::string& why)
	{
		if (handle)
			return true;
		const char* names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char* n : names) {
			handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
			if (handle)
				break;
		}
		if (! handle) {
			why = std::string("dlopen(libnccl.so.2): ") + dlerror();
			return false;
		}
#define SYM(field, name)                                   \
	field = (decltype(field)) dlsym(handle, name);         \
	if (! field) {                                         \
		why = std::string("dlsym ") + name + " failed";    \
		return false;                                      \
	}
		SYM(GetUniqueId, "ncclGetUniqueId")
		SYM(CommInitRank, "ncclCommInitRank")
		SYM(CommDestroy, "ncclCommDestroy")
		SYM(Send, "ncclSend")
		SYM(Recv, "ncclRecv")
		SYM(AllReduce, "ncclAllReduce")
		SYM(GroupStart, "ncclGroupStart")
		SYM(GroupEnd, "ncclGroupEnd")
		SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
		return true;
	}
};
NcclApi g_nccl;

#define NC(call)                                                                                           \
	do {                                                                                                   \
		ncclResult_t r__ = (call);                                                                         \
		if (r__ != ncclSuccess)                                                                            \
			return fail(LBMX_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r__));            \
	} while (0)

int q_of(int lattice)
{
	return lattice == LBMX_D2Q9 ? 9 : (lattice == LBMX_D3Q19 ? 19 : 27);
}
int n_macro_of(int lattice, int macro)
{
	if (macro == LBMX_MACRO_VOID)
		return 0;
	if (lattice == LBMX_D2Q9)
		return macro == LBMX_MACRO_DEFAULT ? 3 : (macro == LBMX_MACRO_MEAN ? 8 : 10);
	return macro == LBMX_MACRO_DEFAULT ? 4 : 13;
}

template <typename L>
int halo_dirs(int32_t* to_right, int32_t* to_left)
{
	int nr = 0, nl = 0;
	for (int q = 0; q < L::Q; q++) {
		if (L::cx(q) > 0)
			to_right[nr++] = q;
		if (L::cx(q) < 0)
			to_left[nl++] = q;
	}
	return nr;
}

}  // namespace

// ---- cell classification on the device (lbmx_map_upload): which cells go to the boundary list, how many per x-plane ----------
struct IsBoundaryCell
{
	const int16_t* map;
	int periodic, nothing, ndim, ox, X, Y, Z;
	bool ab;
	__host__ __device__ bool operator()(uint32_t c) const
	{
		const int m = map[c];
		if (m == periodic || m == nothing || (m == 0 && ! ab))
			return false;  // GEO_PERIODIC, GEO_NOTHING, GEO_FLUID under A-A
		if (m > 1)
			return true;
		// GEO_WALL (1 in every lattice) away from the lattice faces is bounced by the bulk kernel itself; GEO_FLUID (0) on a face clamps under A-B
		// (kernels.cuh: cell_in_boundary_list)
		const int YZ = Y * Z, xs = (int) (c / (uint32_t) YZ), yz = (int) (c - (uint32_t) xs * (uint32_t) YZ), z = yz / Y;
		return lbmx::cell_in_boundary_list(m, 0, periodic, 1, nothing, lbmx::cell_on_face(ndim, ox, X, Y, Z, xs - ox, yz - z * Y, z), ab);
	}
};

// per_plane[x] += boundary cells of plane x; per_plane[gridDim.y] += cells of type `reads_neighbour` (GEO_OUTFLOW_RIGHT: pulls the
// populations of the cell at x-1, bc.h:63-65 -- under A-A that is an in-place array another cell updates in the same step)
// per_plane[gridDim.y + 1] += (face_rule != 0) cells on a lattice face whose A-A neighbours lie outside the lattice: not GEO_NOTHING, not
// wrapped, on a y/z face or (face_rule == 2: no ghost planes) an x face -- kernels.h:30-37 takes +-1 unclamped there
__global__ void k_count_boundary_cells(const int16_t* map, long long first_cell, int YZ, int periodic, int reads_neighbour, unsigned* per_plane, int Y, int Z,
									   int face_rule, int nothing, int ndim, int ox)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	bool b = false, r = false, o = false, w = false, inert = false;
	if (i < YZ) {
		const int m = map[first_cell + (long long) blockIdx.y * YZ + i];
		b = lbmx::cell_in_boundary_list(m, 0, periodic, 1, nothing, lbmx::cell_on_face(ndim, ox, (int) gridDim.y, Y, Z, (int) blockIdx.y, i % Y, i / Y), face_rule == 0);
		r = m == reads_neighbour;
		w = m == 1 && ! b;	// GEO_WALL kept by the bulk kernel
		inert = m == nothing;
		if (face_rule && m != nothing && m != periodic) {  // periodic cells wrap (y, z always; x unless there are ghost planes)
			const int y = i % Y, z = i / Y;
			const bool yz_face = y == 0 || y == Y - 1 || (Z > 1 && (z == 0 || z == Z - 1));
			const bool x_face = face_rule == 2 && (blockIdx.y == 0 || blockIdx.y == gridDim.y - 1);
			o = yz_face || x_face;
		}
	}
	const unsigned n = __popc(__ballot_sync(0xffffffffu, b));
	const unsigned nr = __popc(__ballot_sync(0xffffffffu, r));
	const unsigned no = __popc(__ballot_sync(0xffffffffu, o));
	const unsigned nw = __popc(__ballot_sync(0xffffffffu, w));
	const unsigned ni = __popc(__ballot_sync(0xffffffffu, inert));
	if ((threadIdx.x & 31) == 0) {
		if (nw)
			atomicAdd(per_plane + gridDim.y + 2, nw);
		if (ni)
			atomicAdd(per_plane + gridDim.y + 3, ni);
		if (n)
			atomicAdd(per_plane + blockIdx.y, n);
		if (nr)
			atomicAdd(per_plane + gridDim.y, nr);
		if (no)
			atomicAdd(per_plane + gridDim.y + 1, no);
	}
}

// flags[plane][chunk] = 1 when the (up to) blockDim.x cells of that chunk of the plane are all GEO_NOTHING (KParams::inert)
__global__ void k_flag_inert_chunks(const int16_t* map, long long first_cell, int YZ, int nothing, uint8_t* flags)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool ok = i >= YZ || map[first_cell + (long long) blockIdx.y * YZ + i] == nothing;
	const int all = __syncthreads_and(ok);
	if (threadIdx.x == 0)
		flags[(size_t) blockIdx.y * gridDim.x + blockIdx.x] = all ? 1 : 0;
}

// =====================================================================================================================
// engine
// =====================================================================================================================
struct lbmx_engine
{
	lbmx_desc d{};
	lbmx_params prm{};
	int dev = 0;
	int Q = 27, NM = 4;
	size_t rs = 8;	// sizeof(real)
	int64_t X = 0, Y = 0, Z = 0, x0 = 0, ox = 0, YZ = 0, XYZ = 0;
	int left = -1, right = -1;	// neighbour ranks (-1: none)
	bool self_exchange = false;

	void* df[2] = {nullptr, nullptr};
	// The allocations behind df[]: a guard band of one x-plane + one row + one cell (zeroed) on either side of the populations.  The
	// A-A index rule takes x+-1, y+-1, z+-1 unclamped (kernels.h:30-37); at a face cell that is neither GEO_NOTHING nor periodic, on a
	// slab without ghost planes, the reference steps outside its array.  Here such a map stays inside this engine's own memory.
	void* df_alloc[2] = {nullptr, nullptr};
	size_t df_guard = 0;  // bytes
	void* macro = nullptr;
	int16_t* map = nullptr;
	uint32_t* blist = nullptr;
	void* profile = nullptr;
	void* bouzidi = nullptr;
	int64_t profile_sy = 0;
	int* d_flag = nullptr;
	int* d_dirs = nullptr;	// [2][9]: to_right, to_left
	int n_hdirs = 0;
	int32_t h_dirs[2][9];

	std::vector<int64_t> plane_start;  // boundary-list offsets per local x plane, size X+1
	int64_t nb = 0, n_bulk = 0;
	bool list_after_bulk = false;  // see lbmx_map_upload
	bool map_ready = false;

	unsigned ydiv_mul = 0, ydiv_shift = 0;
	int64_t iter = 0;
	cudaStream_t s_main = nullptr, s_edge = nullptr, s_comm = nullptr;
	cudaEvent_t ev_edge = nullptr, ev_comm = nullptr, ev_main = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
	// two consecutive steps (even + odd iteration) of the single-slab path as a CUDA graph, replayed inside long batches of a small
	// lattice where the gaps between dependent launches are a visible share of a 10-30 us step
	cudaGraphExec_t pair_exec = nullptr;
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	uint64_t state_version = 1, pair_version = 0;  // state_version: bumped by whatever changes the kernel parameters
	int pair_out_mode = -1, pair_launches = 0;
	bool graphs_enabled = true;
	bool pdl_enabled = true;  // LBMX_NO_PDL unsets it
	// single slab of a small lattice: the kernels of consecutive steps form one chain on the compute stream under programmatic dependent
	// launch (a step is tens of microseconds there; a dependent launch costs 2-3 of them, a fork/join through a second stream more)
	bool pdl_chain(bool) const { return pdl_enabled && ox == 0 && X * YZ <= (16ll << 20) && ! use_tma[S_AA_EVEN] && ! use_tma[S_AA_ODD]; }
	bool in_head_step = false;	// step_impl recursing for the odd head of a graph-replayed batch
	// peer-memory halo exchange (one node, NVLink): the neighbours' distribution arrays and arrival counters mapped through CUDA IPC
	struct Peer
	{
		void* df[2] = {nullptr, nullptr};  // the neighbour's df[0], df[1] in this process' address space
		void* df_base[2] = {nullptr, nullptr};	// what cudaIpcOpenMemHandle returned (the neighbour's allocation, guard band first)
		long long* flags = nullptr;		   // the neighbour's arrival counters: [0] bumped by its left neighbour, [1] by its right one
		int64_t X = 0, XYZ = 0;			   // the neighbour's slab
		bool mapped = false;
	};
	Peer peer_left, peer_right;
	// my_flags[0], [1]: exchanges that arrived from the left / right neighbour; [3], [4]: the last step batch the left / right neighbour
	// has declared itself ready to receive for (its arrays are no longer written from the host side)
	long long* my_flags = nullptr;
	bool p2p = false;
	int64_t xcount = 0;	 // halo exchanges enqueued so far (identical on every rank)
	int64_t batch = 0, batch_awaited = 0;  // lbmx_step calls so far (identical on every rank); the last one whose readiness was awaited
	// set by a halo wait that gave up (a neighbour that stopped stepping): host-mapped, so every later call can read it without a
	// stream synchronisation and fail -- the fields computed after it are stale
	long long* h_halo_error = nullptr;
	long long* d_halo_error = nullptr;
	long long halo_timeout_ns = 120ll * 1000 * 1000 * 1000;
	double* eq_stage[2] = {nullptr, nullptr};  // staging sets of lbmx_df_set_equilibrium_field
	int64_t eq_stage_cells = 0;
	ncclComm_t comm = nullptr;

	StepKernels<float> kf{};
	StepKernels<double> kd{};
	lbmx_stats stats{};
	// k_bulk_tma (A-A only): tile geometry, and which step parities go through it
	int tile_y = 0, tile_y_shift = 0;
	bool use_tma[3] = {false, false, false};  // by StreamMode
	int64_t inert_cells = 0;				  // GEO_NOTHING cells
	uint8_t* inert_flags = nullptr;			  // KParams::inert, allocated for maps with sizeable inert regions (lbmx_map_upload)
	int64_t walls_in_bulk = 0;				  // GEO_WALL cells away from the faces: bounced by the bulk kernel (kernels.cuh: cell_in_boundary_list)

	bool f64() const { return d.precision == LBMX_F64; }
	bool aa() const { return d.streaming == LBMX_STREAM_AA; }
	void* cur() const { return aa() ? df[0] : df[iter % 2]; }
	void* other() const { return aa() ? nullptr : df[(iter + 1) % 2]; }
};

namespace {

template <typename R>
KParams<R> make_params(const lbmx_engine* e)
{
	KParams<R> p{};
	p.cur = (R*) e->cur();
	p.out = (R*) e->other();
	p.macro = (R*) e->macro;
	p.map = e->map;
	p.profile = (const R*) e->profile;
	p.bouzidi = (const R*) e->bouzidi;
	p.blist = e->blist;
	p.inert = e->inert_flags;
	p.inert_stride = (int) ((e->YZ + LBMX_BULK_BLOCK - 1) / LBMX_BULK_BLOCK);
	p.XYZ = e->XYZ;
	p.X = (int) e->X;
	p.Y = (int) e->Y;
	p.Z = (int) e->Z;
	p.ox = (int) e->ox;
	p.YZ = (int) e->YZ;
	for (int q = 0; q < e->Q; q++) {
		p.rd[q] = p.cur + (size_t) q * e->XYZ;
		p.wr[q] = (e->aa() ? p.cur : p.out) + (size_t) q * e->XYZ;
	}
	p.ydiv_mul = e->ydiv_mul;
	p.ydiv_shift = e->ydiv_shift;
	p.tile_y = e->tile_y;
	p.tile_y_shift = e->tile_y_shift;
	p.x_begin = 0;
	p.nb_begin = 0;
	p.nb_end = (int) e->nb;
	p.wrap = e->ox == 0 ? 1 : 0;
	p.profile_sy = (int) e->profile_sy;
	p.eq = e->d.eq;
	p.inflow = e->d.inflow;
	p.stream = e->aa() ? (e->iter % 2 == 0 ? S_AA_EVEN : S_AA_ODD) : S_AB;
	p.out_mode = OUT_NONE;
	p.stat_counter = e->prm.stat_counter;
	p.kahan_rho = e->d.lattice == LBMX_D3Q27 && e->d.coll == LBMX_COLL_CUM_HP_RHO;
	const bool vm = e->d.macro == LBMX_MACRO_VOID;
	// MACRO_Void::copyQuantities is empty (d3q27/macro.h:174-188): the KernelStruct keeps lbmViscosity = 1, zero force
	p.phys.nu = vm ? R(1) : (R) e->prm.lbmViscosity;
	set_rates(p.phys);	// omega1 (IEEE division in precision R: the same bits the kernels used to compute per cell) and the 2017 cumulant rates
	p.phys.fx = vm ? R(0) : (R) e->prm.fx;
	p.phys.fy = vm ? R(0) : (R) e->prm.fy;
	p.phys.fz = (vm || e->d.lattice == LBMX_D2Q9) ? R(0) : (R) e->prm.fz;
	p.in_vx = (R) e->prm.inflow_vx;
	p.in_vy = (R) e->prm.inflow_vy;
	p.in_vz = (R) e->prm.inflow_vz;
	return p;
}

constexpr int BLOCK = LBMX_BULK_BLOCK;
#ifndef LBMX_TMA_DEFAULT
	#define LBMX_TMA_DEFAULT "0"
#endif

// launch the two step kernels over local planes [xb, xe): the bulk kernel on `st`, the boundary-list kernel on `st_list` (default:
// the same stream).  Within one step the two kernels touch disjoint cells and every population slot has exactly one writer, so
// they may run concurrently.
template <typename R>
cudaError_t launch_pdl(void (*kernel)(KParams<R>), dim3 grid, int block, cudaStream_t st, const KParams<R>& p)
{
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = grid;
	cfg.blockDim = dim3((unsigned) block);
	cfg.dynamicSmemBytes = 0;
	cfg.stream = st;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kernel, p);
}

// chain: both kernels on `st` under programmatic dependent launch (kernels.cuh: pdl_wait) -- the single-slab path of small lattices
template <typename R>
int launch_range(lbmx_engine* e, const StepKernels<R>& K, KParams<R> p, int xb, int xe, cudaStream_t st, cudaStream_t st_list = nullptr, bool chain = false)
{
	if (! st_list)
		st_list = st;
	if (xe <= xb)
		return LBMX_OK;
	p.x_begin = xb;
	p.nb_begin = (int) e->plane_start[xb];
	p.nb_end = (int) e->plane_start[xe];
	if (e->use_tma[p.stream] && e->walls_in_bulk == 0 && e->inert_cells == 0) {  // the experimental tile kernels know fluid cells only
		// one CTA per tile of tile_y x (128 / tile_y) cells; the populations travel as bulk copies of the TMA engine (kernels_tma.cuh)
		const int tz = tma::TILE / e->tile_y;
		dim3 grid((unsigned) ((e->Y / e->tile_y) * ((e->Z + tz - 1) / tz)), (unsigned) (xe - xb));
		K.bulk_tma[p.stream]<<<grid, tma::TILE, 0, st>>>(p);
		e->stats.tma_launches++;
	}
	else {
		const int per_cta = BLOCK * K.cpt[p.stream];
		dim3 grid((unsigned) ((e->YZ + per_cta - 1) / per_cta), (unsigned) (xe - xb));
		if (chain)
			CU(launch_pdl<R>(K.bulk[p.stream], grid, BLOCK, st, p));
		else
			K.bulk[p.stream]<<<grid, BLOCK, 0, st>>>(p);
	}
	e->stats.kernel_launches++;
	const int nbl = p.nb_end - p.nb_begin;
	if (nbl > 0) {
		if (chain) {
			p.pdl = e->list_after_bulk ? 2 : 1;
			CU(launch_pdl<R>(K.boundary, dim3((unsigned) ((nbl + BLOCK - 1) / BLOCK)), BLOCK, st, p));
		}
		else
			K.boundary<<<(nbl + BLOCK - 1) / BLOCK, BLOCK, 0, st_list>>>(p);
		e->stats.kernel_launches++;
	}
	CU(cudaGetLastError());
	return LBMX_OK;
}

size_t plane_bytes(const lbmx_engine* e)
{
	return (size_t) e->YZ * e->rs;
}

// peer-memory exchange: arrival counter in the receiver's memory, written after the planes (the copy kernel before it in the
// stream has completed, i.e. its stores are performed) and polled by the receiver before the kernels that read those planes
__global__ void k_signal_arrival(long long* peer_counter, long long value)
{
	__threadfence_system();
	*(volatile long long*) peer_counter = value;
	__threadfence_system();
}
__device__ __forceinline__ long long wall_ns()
{
	long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
// limit_ns <= 0: wait for ever, as an MPI receive would.  Otherwise a neighbour that is later than limit_ns (wall clock) makes the
// wait give up: the error word (host-mapped) is set, stays set, and every later call on this engine fails with LBMX_ERR_STATE.
__global__ void k_await_arrival(const long long* from_left, const long long* from_right, long long expected, long long* error_word, long long limit_ns, long long code)
{
	const long long t0 = wall_ns();
	for (int side = 0; side < 2; side++) {
		const long long* c = side == 0 ? from_left : from_right;
		if (! c)
			continue;
		while (*(volatile const long long*) c < expected) {
			if (limit_ns > 0 && wall_ns() - t0 > limit_ns) {
				*(volatile long long*) error_word = code;
				__threadfence_system();
				return;
			}
			__nanosleep(200);
		}
	}
	__threadfence_system();
}

// enqueue the halo exchange that follows the step at e->iter on s_comm (replaces synchronizeDFsAndMacroDevice, lbm.hpp:196-280)
template <typename R>
int exchange(lbmx_engine* e, void* arr)
{
	lbmx_halo_msg msgs[2];
	lbmx_halo_plan(e->d.lattice, e->d.streaming, e->iter, e->X, msgs);
	R* a = (R*) arr;
	const int nd = e->n_hdirs;
	dim3 grid((unsigned) ((e->YZ + 255) / 256), (unsigned) nd);
	if (e->self_exchange) {
		// single slab with ghost planes and periodic x: the neighbour on both sides is this slab
		for (int k = 0; k < 2; k++) {
			const lbmx_halo_msg& m = msgs[k];
			// which population slots travel depends on the streaming parity (A-A even steps carry the opposite slots), not on the direction
			const int* dirs = e->d_dirs + (m.dirs[0] == e->h_dirs[0][0] ? 0 : 9);
			k_copy_planes<R><<<grid, 256, 0, e->s_comm>>>(a, a, e->XYZ, (int) e->YZ, nd, dirs, m.src_plane, m.dst_plane);
			e->stats.kernel_launches++;
			e->stats.halo_bytes_sent += (int64_t) nd * plane_bytes(e);
		}
		CU(cudaGetLastError());
		return LBMX_OK;
	}
	if (e->d.nranks == 1)
		return LBMX_OK;	 // ghost planes without a periodic partner: nothing to exchange
	if (! e->comm)
		return fail(LBMX_ERR_STATE, "lbmx_step: nranks > 1 but lbmx_comm_init was not called");
	if (e->p2p) {
		// One copy kernel per direction stores the 9 planes straight into the neighbour's array over NVLink, then bumps the
		// neighbour's arrival counter.  Which of my arrays `arr` is tells which of the neighbour's it goes to (same rotation).
		const int which = arr == e->df[0] ? 0 : 1;
		e->xcount++;
		if (e->batch_awaited < e->batch) {
			// One-sided stores need a receiver that has stopped writing its arrays from the host side (initialisation, uploads, a
			// restored checkpoint): every rank announces "ready for batch b" on its compute stream when lbmx_step is entered, i.e.
			// after everything it did before; the first push of the batch waits for both neighbours' announcements.
			k_await_arrival<<<1, 1, 0, e->s_comm>>>(e->peer_left.mapped ? e->my_flags + 3 : nullptr, e->peer_right.mapped ? e->my_flags + 4 : nullptr, (long long) e->batch,
													e->d_halo_error, e->halo_timeout_ns, 2);
			e->stats.kernel_launches++;
			e->batch_awaited = e->batch;
		}
		for (int k = 0; k < 2; k++) {
			const lbmx_halo_msg& m = msgs[k];
			lbmx_engine::Peer& peer = m.to_right ? e->peer_right : e->peer_left;
			if (! peer.mapped)
				continue;
			lbmx_halo_msg pm[2];
			lbmx_halo_plan(e->d.lattice, e->d.streaming, e->iter, peer.X, pm);	// where the receiver keeps what arrives (its own layout)
			const int* dirs = e->d_dirs + (m.dirs[0] == e->h_dirs[0][0] ? 0 : 9);
			k_copy_planes<R><<<grid, 256, 0, e->s_comm>>>((R*) peer.df[which], a, e->XYZ, (int) e->YZ, nd, dirs, m.src_plane, pm[k].dst_plane, peer.XYZ);
			k_signal_arrival<<<1, 1, 0, e->s_comm>>>(peer.flags + (m.to_right ? 0 : 1), (long long) e->xcount);
			e->stats.kernel_launches += 2;
			e->stats.halo_bytes_sent += (int64_t) nd * plane_bytes(e);
		}
		CU(cudaGetLastError());
		return LBMX_OK;
	}
	// the 9 crossing populations of a plane are 9 separate contiguous runs of Y*Z reals (x is the slowest storage
	// dimension): they go out as 9 sends per direction inside one NCCL group -- no pack/unpack kernels at all
	NC(g_nccl.GroupStart());
	for (int k = 0; k < 2; k++) {
		const lbmx_halo_msg& m = msgs[k];
		const int peer_send = m.to_right ? e->right : e->left;
		const int peer_recv = m.to_right ? e->left : e->right;	// what arrives from the opposite side carries the same populations
		for (int i = 0; i < nd; i++) {
			const int q = m.dirs[i];
			if (peer_send >= 0) {
				NC(g_nccl.Send(a + q * e->XYZ + m.src_plane * e->YZ, plane_bytes(e), ncclInt8, peer_send, e->comm, e->s_comm));
				e->stats.halo_bytes_sent += plane_bytes(e);
			}
			if (peer_recv >= 0)
				NC(g_nccl.Recv(a + q * e->XYZ + m.dst_plane * e->YZ, plane_bytes(e), ncclInt8, peer_recv, e->comm, e->s_comm));
		}
	}
	NC(g_nccl.GroupEnd());
	e->stats.kernel_launches++;
	return LBMX_OK;
}

// peer-memory exchange: what the neighbours pushed after their last step has arrived before `st` continues
static void await_halo(lbmx_engine* e, cudaStream_t st)
{
	if (! e->p2p)
		return;
	k_await_arrival<<<1, 1, 0, st>>>(e->peer_left.mapped ? e->my_flags : nullptr, e->peer_right.mapped ? e->my_flags + 1 : nullptr, (long long) e->xcount, e->d_halo_error,
									 e->halo_timeout_ns, 1);
	e->stats.kernel_launches++;
}

// capture steps (iter, iter + 1) of the single-slab path, iter even, with every step using `out_mode`
template <typename R>
int capture_pair(lbmx_engine* e, const StepKernels<R>& K, int out_mode)
{
	if (e->pair_exec) {
		CU(cudaGraphExecDestroy(e->pair_exec));
		e->pair_exec = nullptr;
	}
	const int64_t launches0 = e->stats.kernel_launches, iter0 = e->iter;
	cudaGraph_t graph = nullptr;
	CU(cudaStreamBeginCapture(e->s_main, cudaStreamCaptureModeThreadLocal));
	int rc = LBMX_OK;
	for (int k = 0; k < 2 && rc == LBMX_OK; k++) {
		e->iter = iter0 + k;  // parity / rotation of this step
		KParams<R> p = make_params<R>(e);
		p.out_mode = out_mode;
		if (e->pdl_chain(true))
			rc = launch_range(e, K, p, 0, (int) e->X, e->s_main, nullptr, true);
		else if (e->nb == 0 || e->list_after_bulk)
			rc = launch_range(e, K, p, 0, (int) e->X, e->s_main);
		else {	// fork: the boundary-list kernel beside the bulk kernel; join before the next step
			cudaEventRecord(e->ev_fork, e->s_main);
			cudaStreamWaitEvent(e->s_edge, e->ev_fork, 0);
			rc = launch_range(e, K, p, 0, (int) e->X, e->s_main, e->s_edge);
			cudaEventRecord(e->ev_join, e->s_edge);
			cudaStreamWaitEvent(e->s_main, e->ev_join, 0);
		}
	}
	e->iter = iter0;
	e->pair_launches = (int) (e->stats.kernel_launches - launches0);
	e->stats.kernel_launches = launches0;  // nothing ran yet
	const cudaError_t ce = cudaStreamEndCapture(e->s_main, &graph);
	if (rc != LBMX_OK || ce != cudaSuccess || ! graph) {
		if (graph)
			cudaGraphDestroy(graph);
		cudaGetLastError();
		e->graphs_enabled = false;	// fall back to plain launches for the life of this engine
		return rc != LBMX_OK ? rc : LBMX_OK;
	}
	if (cudaGraphInstantiate(&e->pair_exec, graph, 0) != cudaSuccess) {
		cudaGetLastError();
		e->pair_exec = nullptr;
		e->graphs_enabled = false;
	}
	cudaGraphDestroy(graph);
	e->pair_version = e->state_version;
	e->pair_out_mode = out_mode;
	return LBMX_OK;
}

template <typename R>
int step_impl(lbmx_engine* e, const StepKernels<R>& K, int64_t nsteps)
{
	const bool ghosts = e->ox > 0;
	if (e->p2p && ! e->in_head_step) {
		// "ready for batch b": see exchange().  On the compute stream, behind every upload / initialisation kernel of this rank.
		e->batch++;
		if (e->peer_left.mapped)
			k_signal_arrival<<<1, 1, 0, e->s_main>>>(e->peer_left.flags + 4, (long long) e->batch);	 // I am my left neighbour's right neighbour
		if (e->peer_right.mapped)
			k_signal_arrival<<<1, 1, 0, e->s_main>>>(e->peer_right.flags + 3, (long long) e->batch);
		e->stats.kernel_launches += 2;
	}
	CU(cudaEventRecord(e->ev_main, e->s_main));	 // everything enqueued on the compute stream so far (uploads, initialisation) precedes the side streams' work
	int64_t first_plain = 0;
	// Small single-slab lattices (steps of tens of microseconds): replay step pairs as a graph.  Pairs start at even iterations (A-A
	// parity, A-B rotation); the last step of the batch stays a plain launch when it alone writes the macroscopic fields.
	constexpr int64_t kGraphMaxCells = 16ll << 20, kGraphMinSteps = 8;
	if (! ghosts && e->graphs_enabled && e->d.macro != LBMX_MACRO_MEAN && e->d.macro != LBMX_MACRO_WITH_MEAN_2D && nsteps >= kGraphMinSteps && e->X * e->YZ <= kGraphMaxCells) {
		const bool last_writes = e->d.macro == LBMX_MACRO_DEFAULT && e->d.macro_policy == LBMX_MACRO_LAST_STEP;
		const int out_mode = (e->d.macro == LBMX_MACRO_DEFAULT && e->d.macro_policy == LBMX_MACRO_EVERY_STEP) ? OUT_DEFAULT : OUT_NONE;
		const int64_t head = e->iter & 1, tail = last_writes ? 1 : 0;
		const int64_t pairs = (nsteps - head - tail) / 2;
		if (pairs > 0) {
			if (head) {	 // one plain step up to an even iteration: recurse with a batch below the graph threshold
				const lbmx_desc saved = e->d;
				const int saved_counter = e->prm.stat_counter;
				if (last_writes)
					e->d.macro_policy = LBMX_MACRO_NEVER;  // not the last step of the caller's batch
				e->in_head_step = true;
				const int rc1 = step_impl<R>(e, K, 1);
				e->in_head_step = false;
				e->d = saved;
				e->prm.stat_counter = saved_counter;
				if (rc1)
					return rc1;
				CU(cudaEventRecord(e->ev_main, e->s_main));
			}
			if (! e->pair_exec || e->pair_version != e->state_version || e->pair_out_mode != out_mode) {
				const int rc2 = capture_pair<R>(e, K, out_mode);
				if (rc2)
					return rc2;
			}
			if (e->pair_exec) {
				for (int64_t i = 0; i < pairs; i++) {
					CU(cudaGraphLaunch(e->pair_exec, e->s_main));
					e->iter += 2;
					e->stats.kernel_launches += e->pair_launches;
				}
				first_plain = head + 2 * pairs;
				CU(cudaEventRecord(e->ev_main, e->s_main));	 // what follows on the side stream waits for the replayed steps
			}
			else
				first_plain = head;
		}
	}
	for (int64_t s = first_plain; s < nsteps; s++) {
		KParams<R> p = make_params<R>(e);
		const bool last = s == nsteps - 1;
		if (e->d.macro == LBMX_MACRO_MEAN)
			p.out_mode = OUT_MEAN;
		else if (e->d.macro == LBMX_MACRO_DEFAULT && (e->d.macro_policy == LBMX_MACRO_EVERY_STEP || (e->d.macro_policy == LBMX_MACRO_LAST_STEP && last)))
			p.out_mode = OUT_DEFAULT;
		else if (e->d.macro == LBMX_MACRO_WITH_MEAN_2D) {
			// the sums advance every step while a gate is open; with both closed the instantaneous fields follow the policy
			const int gates = e->prm.macro_gates & (LBMX_GATE_MEANS | LBMX_GATE_FLUCS);
			if (e->d.macro_policy != LBMX_MACRO_NEVER && (gates || e->d.macro_policy == LBMX_MACRO_EVERY_STEP || last))
				p.out_mode = OUT_WITH_MEAN_2D + gates;
		}
		p.stat_counter = e->prm.stat_counter + (int) s;
		int rc;
		if (! ghosts) {
			if (e->pdl_chain(false)) {
				if ((rc = launch_range(e, K, p, 0, (int) e->X, e->s_main, nullptr, true)))
					return rc;
			}
			else if (e->nb == 0 || e->list_after_bulk) {
				if ((rc = launch_range(e, K, p, 0, (int) e->X, e->s_main)))
					return rc;
			}
			else {
				// the boundary-list kernel (a few thousand cells: a latency chain of its own) runs beside the bulk kernel on the second
				// stream; both kernels of a step wait for both kernels of the previous one
				CU(cudaStreamWaitEvent(e->s_edge, e->ev_main, 0));
				CU(cudaStreamWaitEvent(e->s_main, e->ev_edge, 0));
				if ((rc = launch_range(e, K, p, 0, (int) e->X, e->s_main, e->s_edge)))
					return rc;
				CU(cudaEventRecord(e->ev_main, e->s_main));
				CU(cudaEventRecord(e->ev_edge, e->s_edge));
			}
		}
		else if (e->list_after_bulk) {
			// A-A slab with GEO_OUTFLOW_RIGHT cells: those read a neighbour's populations in place, so the whole boundary list has to
			// follow the whole bulk kernel (as on a single slab) -- no edge-first overlap for this slab, the result stays reproducible
			CU(cudaStreamWaitEvent(e->s_main, e->ev_comm, 0));
			await_halo(e, e->s_main);
			if ((rc = launch_range(e, K, p, 0, (int) e->X, e->s_main)))
				return rc;
			CU(cudaEventRecord(e->ev_main, e->s_main));
			CU(cudaStreamWaitEvent(e->s_comm, e->ev_main, 0));
			if ((rc = exchange<R>(e, e->df[0])))
				return rc;
			CU(cudaEventRecord(e->ev_comm, e->s_comm));
		}
		else {
			// boundary planes first on the high-priority stream, then the exchange, interior concurrently (state.hpp:1060-1108)
			CU(cudaStreamWaitEvent(e->s_edge, e->ev_main, 0));	// previous step's interior (reads/writes next to the edge planes)
			CU(cudaStreamWaitEvent(e->s_edge, e->ev_comm, 0));	// previous exchange filled the ghost planes this step reads
			await_halo(e, e->s_edge);							// ... including what the neighbours pushed (peer-memory exchange)
			if ((rc = launch_range(e, K, p, 0, 1, e->s_edge)))
				return rc;
			if (e->X > 1 && (rc = launch_range(e, K, p, (int) e->X - 1, (int) e->X, e->s_edge)))
				return rc;
			CU(cudaEventRecord(e->ev_edge, e->s_edge));
			CU(cudaStreamWaitEvent(e->s_main, e->ev_comm, 0));
			if ((rc = launch_range(e, K, p, 1, (int) e->X - 1, e->s_main)))
				return rc;
			CU(cudaEventRecord(e->ev_main, e->s_main));
			CU(cudaStreamWaitEvent(e->s_comm, e->ev_edge, 0));
			// A-A odd steps also write the ghost planes from the planes next to the edge: those are edge-plane cells only
			void* arr = e->aa() ? e->df[0] : e->other();
			if ((rc = exchange<R>(e, arr)))
				return rc;
			CU(cudaEventRecord(e->ev_comm, e->s_comm));
		}
		e->iter++;
	}
	e->prm.stat_counter += (int) nsteps;
	// leave the engine in a state where s_main alone orders everything that was enqueued
	if (ghosts) {
		CU(cudaStreamWaitEvent(e->s_main, e->ev_comm, 0));
		await_halo(e, e->s_main);
	}
	if ((ghosts || e->nb > 0) && ! e->list_after_bulk)
		CU(cudaStreamWaitEvent(e->s_main, e->ev_edge, 0));
	return LBMX_OK;
}

int step_dispatch(lbmx_engine* e, int64_t nsteps)
{
	if (! e->map_ready)
		return fail(LBMX_ERR_STATE, "lbmx_step: upload a map first (lbmx_map_upload)");
	CU(cudaSetDevice(e->dev));
	return e->f64() ? step_impl<double>(e, e->kd, nsteps) : step_impl<float>(e, e->kf, nsteps);
}

bool pick_kernels(lbmx_engine* e)
{
	const lbmx_desc& d = e->d;
	if (d.flags & LBMX_FLAG_STRICT_ARITH) {	 // parity arithmetic (lattices checked by the caller)
		if (d.lattice == LBMX_D3Q27) {
			switch (d.coll) {
				case LBMX_COLL_CUM: return e->f64() ? get_kernels_d3q27_cum_strict(e->kd) : get_kernels_d3q27_cum_strict(e->kf);
				case LBMX_COLL_SRT: return e->f64() ? get_kernels_d3q27_srt_strict(e->kd) : get_kernels_d3q27_srt_strict(e->kf);
				case LBMX_COLL_BGK: return e->f64() ? get_kernels_d3q27_bgk_strict(e->kd) : get_kernels_d3q27_bgk_strict(e->kf);
				case LBMX_COLL_BGK_GALILEAN: return e->f64() ? get_kernels_d3q27_bgkgal_strict(e->kd) : get_kernels_d3q27_bgkgal_strict(e->kf);
				case LBMX_COLL_CUM_HP_RHO: return e->f64() ? get_kernels_d3q27_cumhp_strict(e->kd) : get_kernels_d3q27_cumhp_strict(e->kf);
				case LBMX_COLL_MRT_LES: return e->f64() ? get_kernels_d3q27_mrt_strict(e->kd) : get_kernels_d3q27_mrt_strict(e->kf);
				case LBMX_COLL_CLBM: return e->f64() ? get_kernels_d3q27_clbm_strict(e->kd) : get_kernels_d3q27_clbm_strict(e->kf);
				case LBMX_COLL_SRT_MODIF_FORCE: return e->f64() ? get_kernels_d3q27_srtmf_strict(e->kd) : get_kernels_d3q27_srtmf_strict(e->kf);
				case LBMX_COLL_KBC_N1: return e->f64() ? get_kernels_d3q27_kbcn1_strict(e->kd) : get_kernels_d3q27_kbcn1_strict(e->kf);
				case LBMX_COLL_KBC_N2: return e->f64() ? get_kernels_d3q27_kbcn2_strict(e->kd) : get_kernels_d3q27_kbcn2_strict(e->kf);
				case LBMX_COLL_KBC_N3: return e->f64() ? get_kernels_d3q27_kbcn3_strict(e->kd) : get_kernels_d3q27_kbcn3_strict(e->kf);
				case LBMX_COLL_KBC_N4: return e->f64() ? get_kernels_d3q27_kbcn4_strict(e->kd) : get_kernels_d3q27_kbcn4_strict(e->kf);
				case LBMX_COLL_KBC_C1: return e->f64() ? get_kernels_d3q27_kbcc1_strict(e->kd) : get_kernels_d3q27_kbcc1_strict(e->kf);
				case LBMX_COLL_KBC_C2: return e->f64() ? get_kernels_d3q27_kbcc2_strict(e->kd) : get_kernels_d3q27_kbcc2_strict(e->kf);
				case LBMX_COLL_KBC_C3: return e->f64() ? get_kernels_d3q27_kbcc3_strict(e->kd) : get_kernels_d3q27_kbcc3_strict(e->kf);
				case LBMX_COLL_KBC_C4: return e->f64() ? get_kernels_d3q27_kbcc4_strict(e->kd) : get_kernels_d3q27_kbcc4_strict(e->kf);
				case LBMX_COLL_CUM_2017: return e->f64() ? get_kernels_d3q27_cum2017_strict(e->kd) : get_kernels_d3q27_cum2017_strict(e->kf);
				case LBMX_COLL_CUM_ANTIALIAS: return e->f64() ? get_kernels_d3q27_cumaa_strict(e->kd) : get_kernels_d3q27_cumaa_strict(e->kf);
				case LBMX_COLL_CUM_2017_ANTIALIAS: return e->f64() ? get_kernels_d3q27_cum2017aa_strict(e->kd) : get_kernels_d3q27_cum2017aa_strict(e->kf);
			}
		}
		else if (d.lattice == LBMX_D2Q9) {
			if (d.coll == LBMX_COLL_SRT)
				return e->f64() ? get_kernels_d2q9_srt_strict(e->kd) : get_kernels_d2q9_srt_strict(e->kf);
			if (d.coll == LBMX_COLL_CLBM)
				return e->f64() ? get_kernels_d2q9_clbm_strict(e->kd) : get_kernels_d2q9_clbm_strict(e->kf);
		}
		return false;
	}
	if (d.lattice == LBMX_D3Q27) {
		switch (d.coll) {
			case LBMX_COLL_CUM: return e->f64() ? get_kernels_d3q27_cum(e->kd) : get_kernels_d3q27_cum(e->kf);
			case LBMX_COLL_SRT: return e->f64() ? get_kernels_d3q27_srt(e->kd) : get_kernels_d3q27_srt(e->kf);
			case LBMX_COLL_BGK: return e->f64() ? get_kernels_d3q27_bgk(e->kd) : get_kernels_d3q27_bgk(e->kf);
			case LBMX_COLL_BGK_GALILEAN: return e->f64() ? get_kernels_d3q27_bgkgal(e->kd) : get_kernels_d3q27_bgkgal(e->kf);
			case LBMX_COLL_CUM_HP_RHO: return e->f64() ? get_kernels_d3q27_cumhp(e->kd) : get_kernels_d3q27_cumhp(e->kf);
			case LBMX_COLL_MRT_LES: return e->f64() ? get_kernels_d3q27_mrt(e->kd) : get_kernels_d3q27_mrt(e->kf);
			case LBMX_COLL_CLBM: return e->f64() ? get_kernels_d3q27_clbm(e->kd) : get_kernels_d3q27_clbm(e->kf);
			case LBMX_COLL_SRT_MODIF_FORCE: return e->f64() ? get_kernels_d3q27_srtmf(e->kd) : get_kernels_d3q27_srtmf(e->kf);
			case LBMX_COLL_KBC_N1: return e->f64() ? get_kernels_d3q27_kbcn1(e->kd) : get_kernels_d3q27_kbcn1(e->kf);
			case LBMX_COLL_KBC_N2: return e->f64() ? get_kernels_d3q27_kbcn2(e->kd) : get_kernels_d3q27_kbcn2(e->kf);
			case LBMX_COLL_KBC_N3: return e->f64() ? get_kernels_d3q27_kbcn3(e->kd) : get_kernels_d3q27_kbcn3(e->kf);
			case LBMX_COLL_KBC_N4: return e->f64() ? get_kernels_d3q27_kbcn4(e->kd) : get_kernels_d3q27_kbcn4(e->kf);
			case LBMX_COLL_KBC_C1: return e->f64() ? get_kernels_d3q27_kbcc1(e->kd) : get_kernels_d3q27_kbcc1(e->kf);
			case LBMX_COLL_KBC_C2: return e->f64() ? get_kernels_d3q27_kbcc2(e->kd) : get_kernels_d3q27_kbcc2(e->kf);
			case LBMX_COLL_KBC_C3: return e->f64() ? get_kernels_d3q27_kbcc3(e->kd) : get_kernels_d3q27_kbcc3(e->kf);
			case LBMX_COLL_KBC_C4: return e->f64() ? get_kernels_d3q27_kbcc4(e->kd) : get_kernels_d3q27_kbcc4(e->kf);
			case LBMX_COLL_CUM_2017: return e->f64() ? get_kernels_d3q27_cum2017(e->kd) : get_kernels_d3q27_cum2017(e->kf);
			case LBMX_COLL_CUM_ANTIALIAS: return e->f64() ? get_kernels_d3q27_cumaa(e->kd) : get_kernels_d3q27_cumaa(e->kf);
			case LBMX_COLL_CUM_2017_ANTIALIAS: return e->f64() ? get_kernels_d3q27_cum2017aa(e->kd) : get_kernels_d3q27_cum2017aa(e->kf);
		}
	}
	else if (d.lattice == LBMX_D3Q19) {
		if (d.coll == LBMX_COLL_SRT)
			return e->f64() ? get_kernels_d3q19_srt(e->kd) : get_kernels_d3q19_srt(e->kf);
		if (d.coll == LBMX_COLL_MRT_LES)
			return e->f64() ? get_kernels_d3q19_mrt(e->kd) : get_kernels_d3q19_mrt(e->kf);
	}
	else if (d.lattice == LBMX_D2Q9) {
		if (d.coll == LBMX_COLL_SRT)
			return e->f64() ? get_kernels_d2q9_srt(e->kd) : get_kernels_d2q9_srt(e->kf);
		if (d.coll == LBMX_COLL_CLBM)
			return e->f64() ? get_kernels_d2q9_clbm(e->kd) : get_kernels_d2q9_clbm(e->kf);
	}
	return false;
}

// A halo wait that gave up leaves the ghost planes stale: whatever was computed afterwards is wrong.  The word is host-mapped and
// sticky; every entry point that advances or reads the state checks it (after its own stream synchronisation where it has one).
int halo_error(const lbmx_engine* e, const char* who)
{
	if (! e->h_halo_error)
		return LBMX_OK;
	const long long code = *(volatile const long long*) e->h_halo_error;
	if (code == 0)
		return LBMX_OK;
	return fail(LBMX_ERR_STATE, std::string(who) + ": the peer-memory halo exchange gave up waiting for a neighbour (" +
									(code == 2 ? "it never entered the same lbmx_step batch" : "its planes of an earlier step never arrived") + ", limit " +
									std::to_string(e->halo_timeout_ns / 1000000000ll) + " s, LBMX_HALO_TIMEOUT_S); the fields of this engine are stale from that step on");
}

// host <-> device copies of [ncomp][X(+2ox)][Z][Y] arrays with or without the ghost planes
int copy_components(lbmx_engine* e, void* dev, void* host, int ncomp, size_t elem, bool with_ghosts, bool to_device)
{
	// Every copy is enqueued on the compute stream and waited for there.  A plain cudaMemcpy from pageable host memory runs on the
	// legacy default stream, which the engine's non-blocking streams do not synchronise with, and may return once the data sit in the
	// driver's staging buffer -- before the DMA has landed: a kernel launched on s_main right afterwards could still read the old
	// contents (seen as a miscounted boundary list when several processes share one GPU).
	CU(cudaSetDevice(e->dev));
	const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
	const size_t comp_dev = (size_t) e->XYZ * elem;
	if (with_ghosts || e->ox == 0) {
		const size_t bytes = comp_dev * ncomp;
		CU(to_device ? cudaMemcpyAsync(dev, host, bytes, kind, e->s_main) : cudaMemcpyAsync(host, dev, bytes, kind, e->s_main));
		CU(cudaStreamSynchronize(e->s_main));
		return halo_error(e, "lbmx copy");
	}
	const size_t comp_host = (size_t) e->X * e->YZ * elem;
	const size_t skip = (size_t) e->ox * e->YZ * elem;
	for (int c = 0; c < ncomp; c++) {  // the interior planes of one component are contiguous on both sides
		char* dp = (char*) dev + (size_t) c * comp_dev + skip;
		char* hp = (char*) host + (size_t) c * comp_host;
		CU(to_device ? cudaMemcpyAsync(dp, hp, comp_host, kind, e->s_main) : cudaMemcpyAsync(hp, dp, comp_host, kind, e->s_main));
	}
	CU(cudaStreamSynchronize(e->s_main));
	return halo_error(e, "lbmx copy");
}

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

const char* lbmx_last_error(void)
{
	return g_err.c_str();
}

int lbmx_version(void)
{
	return LBMX_VERSION;
}

int lbmx_decompose_x(int64_t X, int32_t nranks, int32_t rank, int64_t* x_offset, int64_t* x_local)
{
	if (X < 1 || nranks < 1 || rank < 0 || rank >= nranks || nranks > X)
		return fail(LBMX_ERR_ARG, "lbmx_decompose_x: need 0 <= rank < nranks <= X");
	// contiguous x-planes, remainder spread over the first ranks (decomposeLattice_D1Q3 splits the same way)
	const int64_t base = X / nranks, rem = X % nranks;
	const int64_t off = rank * base + std::min<int64_t>(rank, rem);
	if (x_offset)
		*x_offset = off;
	if (x_local)
		*x_local = base + (rank < rem ? 1 : 0);
	return LBMX_OK;
}

int lbmx_halo_directions(int32_t lattice, int32_t* to_right, int32_t* to_left)
{
	int32_t r[9], l[9];
	int n;
	if (lattice == LBMX_D3Q27)
		n = halo_dirs<D3Q27>(r, l);
	else if (lattice == LBMX_D3Q19)
		n = halo_dirs<D3Q19>(r, l);
	else if (lattice == LBMX_D2Q9)
		n = halo_dirs<D2Q9>(r, l);
	else
		return 0;
	if (to_right)
		std::memcpy(to_right, r, n * sizeof(int32_t));
	if (to_left)
		std::memcpy(to_left, l, n * sizeof(int32_t));
	return n;
}

int lbmx_halo_plan(int32_t lattice, int32_t streaming, int64_t iteration, int64_t X_local, lbmx_halo_msg msgs[2])
{
	int32_t r[9], l[9];
	const int n = lbmx_halo_directions(lattice, r, l);
	if (n == 0)
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_halo_plan: unknown lattice");
	// storage planes with one ghost plane per side: 0 = left ghost, 1 = first interior, X_local = last interior, X_local+1 = right ghost
	const int64_t gL = 0, first = 1, last = X_local, gR = X_local + 1;
	lbmx_halo_msg& toR = msgs[0];
	lbmx_halo_msg& toL = msgs[1];
	toR.to_right = 1;
	toL.to_right = 0;
	toR.n_dirs = toL.n_dirs = n;
	const bool aa = streaming == LBMX_STREAM_AA;
	const bool even = (iteration % 2) == 0;
	if (! aa) {
		// A-B: df_out holds post-collision values at their source cell; the +x movers of my last plane are what the right
		// neighbour pulls from its left ghost plane (lbm_block.hpp:423-450 with df_sync_directions, defs.h:309-340)
		std::memcpy(toR.dirs, r, sizeof(r));
		std::memcpy(toL.dirs, l, sizeof(l));
		toR.src_plane = last;
		toR.dst_plane = gL;
		toL.src_plane = first;
		toL.dst_plane = gR;
	}
	else if (even) {
		// A-A even step: post-collision f_q sits in slot opp(q) of its own cell, so what must travel right (the +x movers)
		// is found in the -x slots of my last plane, and lands in the same slots of the neighbour's left ghost plane
		// (lbm_block.hpp:428-436: "opposite direction on even steps")
		std::memcpy(toR.dirs, l, sizeof(l));
		std::memcpy(toL.dirs, r, sizeof(r));
		toR.src_plane = last;
		toR.dst_plane = gL;
		toL.src_plane = first;
		toL.dst_plane = gR;
	}
	else {
		// A-A odd step: the +x movers of my last plane were written into MY right ghost plane (slot q at x + c_q); they
		// belong to the first interior plane of the right neighbour (lbm_block.hpp:437-442: buffer offset 1)
		std::memcpy(toR.dirs, r, sizeof(r));
		std::memcpy(toL.dirs, l, sizeof(l));
		toR.src_plane = gR;
		toR.dst_plane = first;
		toL.src_plane = gL;
		toL.dst_plane = last;
	}
	return LBMX_OK;
}

int lbmx_device_count(int32_t* count)
{
	if (! count)
		return fail(LBMX_ERR_ARG, "lbmx_device_count: null argument");
	int n = 0;
	CU(cudaGetDeviceCount(&n));
	*count = n;
	return LBMX_OK;
}

int lbmx_create(const lbmx_desc* desc, lbmx_engine** out)
{
	if (! desc || ! out)
		return fail(LBMX_ERR_ARG, "lbmx_create: null argument");
	*out = nullptr;
	const lbmx_desc& d = *desc;
	if (d.X < 1 || d.Y < 1 || d.Z < 1)
		return fail(LBMX_ERR_ARG, "lbmx_create: lattice sizes must be positive");
	if (d.lattice != LBMX_D3Q27 && d.lattice != LBMX_D3Q19 && d.lattice != LBMX_D2Q9)
		return fail(LBMX_ERR_ARG, "lbmx_create: lattice selector");
	if (d.lattice == LBMX_D3Q19 && d.eq != LBMX_EQ_STD)
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_create: D3Q19 has only the polynomial equilibrium (the product form needs 27 velocities)");
	if (d.lattice == LBMX_D2Q9 && d.Z != 1)
		return fail(LBMX_ERR_ARG, "lbmx_create: D2Q9 needs Z == 1 (the reference's X x Y x 1 lattice, sim_2D/sim2d_1.cu:134)");
	if (d.precision != LBMX_F32 && d.precision != LBMX_F64)
		return fail(LBMX_ERR_ARG, "lbmx_create: precision");
	if (d.streaming != LBMX_STREAM_AB && d.streaming != LBMX_STREAM_AA)
		return fail(LBMX_ERR_ARG, "lbmx_create: streaming");
	if (d.macro < LBMX_MACRO_VOID || d.macro > LBMX_MACRO_WITH_MEAN_2D || d.inflow < LBMX_INFLOW_NONE || d.inflow > LBMX_INFLOW_PARABOLIC_Y)
		return fail(LBMX_ERR_ARG, "lbmx_create: macro / inflow selector");
	if (d.macro == LBMX_MACRO_WITH_MEAN_2D && d.lattice != LBMX_D2Q9)
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_create: LBMX_MACRO_WITH_MEAN_2D is the macro class of sim_2D/sim2d_2.cu (D2Q9)");
	if (d.inflow == LBMX_INFLOW_PARABOLIC_Y && d.lattice != LBMX_D2Q9)
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_create: LBMX_INFLOW_PARABOLIC_Y is the 2-D inflow of sim_2D/sim2d_3.cu (D2Q9)");
	if (d.eq != LBMX_EQ_STD && d.eq != LBMX_EQ_INV_CUM && d.eq != LBMX_EQ_ENTROPIC)
		return fail(LBMX_ERR_ARG, "lbmx_create: eq selector");
	if (d.eq == LBMX_EQ_ENTROPIC && ! (d.lattice == LBMX_D3Q27 && d.coll >= LBMX_COLL_KBC_N1 && d.coll <= LBMX_COLL_KBC_C4))
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_create: LBMX_EQ_ENTROPIC is available with the D3Q27 KBC operators (where only initialisation and boundary cells evaluate it)");
	if (d.lattice == LBMX_D2Q9 && d.eq != LBMX_EQ_STD)
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_create: D2Q9 has only the polynomial equilibrium (d2q9/eq.h)");
	if (d.nranks < 1 || d.rank < 0 || d.rank >= d.nranks)
		return fail(LBMX_ERR_ARG, "lbmx_create: rank / nranks");
	if ((d.flags & LBMX_FLAG_STRICT_ARITH) && d.lattice == LBMX_D3Q19)
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_create: LBMX_FLAG_STRICT_ARITH follows the reference's arithmetic, and the reference has no D3Q19");

	lbmx_engine* e = new lbmx_engine();
	e->d = d;
	if (e->d.nranks > 1)
		e->d.ghost_x = 1;
	int rc = lbmx_decompose_x(d.X, d.nranks, d.rank, &e->x0, &e->X);
	if (rc) {
		delete e;
		return rc;
	}
	e->Y = d.Y;
	e->Z = d.Z;
	e->ox = e->d.ghost_x ? 1 : 0;
	e->YZ = e->Y * e->Z;
	e->XYZ = (e->X + 2 * e->ox) * e->YZ;
	e->Q = q_of(d.lattice);
	e->NM = n_macro_of(d.lattice, d.macro);
	e->rs = d.precision == LBMX_F64 ? 8 : 4;
	if (e->XYZ >= (int64_t) 1 << 31) {
		delete e;
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_create: more than 2^31 storage cells per slab (cell indices are 32-bit); use more slabs");
	}
	if (! pick_kernels(e)) {
		delete e;
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_create: no kernel family for this lattice / collision combination");
	}
	if (e->d.nranks > 1) {
		e->left = e->d.rank > 0 ? e->d.rank - 1 : (e->d.periodic_x ? e->d.nranks - 1 : -1);
		e->right = e->d.rank < e->d.nranks - 1 ? e->d.rank + 1 : (e->d.periodic_x ? 0 : -1);
	}
	else if (e->ox) {
		e->self_exchange = e->d.periodic_x != 0;
	}
	e->n_hdirs = lbmx_halo_directions(d.lattice, e->h_dirs[0], e->h_dirs[1]);
	if (e->Y > 1) {
		// division by an invariant (Granlund-Montgomery, N = 31): L = ceil(log2 Y), mul = ceil(2^(31+L) / Y) < 2^32, exact for n < 2^31
		unsigned L = 0;
		while (((int64_t) 1 << L) < e->Y)
			L++;
		const unsigned __int128 one = 1;
		e->ydiv_mul = (unsigned) (((one << (31 + L)) + (unsigned __int128) e->Y - 1) / (unsigned __int128) e->Y);
		e->ydiv_shift = L - 1;
	}

#define CUX(call)                                                                                             \
	do {                                                                                                      \
		cudaError_t err__ = (call);                                                                           \
		if (err__ != cudaSuccess) {                                                                           \
			fail(LBMX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));                       \
			lbmx_destroy(e);                                                                                  \
			return LBMX_ERR_CUDA;                                                                             \
		}                                                                                                     \
	} while (0)
	if (d.device >= 0)
		CUX(cudaSetDevice(d.device));
	CUX(cudaGetDevice(&e->dev));
	int lo = 0, hi = 0;
	CUX(cudaDeviceGetStreamPriorityRange(&lo, &hi));
	CUX(cudaStreamCreateWithPriority(&e->s_main, cudaStreamNonBlocking, lo));
	CUX(cudaStreamCreateWithPriority(&e->s_edge, cudaStreamNonBlocking, hi));
	CUX(cudaStreamCreateWithPriority(&e->s_comm, cudaStreamNonBlocking, hi));
	CUX(cudaEventCreateWithFlags(&e->ev_edge, cudaEventDisableTiming));
	CUX(cudaEventCreateWithFlags(&e->ev_comm, cudaEventDisableTiming));
	CUX(cudaEventCreateWithFlags(&e->ev_main, cudaEventDisableTiming));
	CUX(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
	CUX(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
	e->graphs_enabled = std::getenv("LBMX_NO_GRAPH") == nullptr;
	e->pdl_enabled = std::getenv("LBMX_NO_PDL") == nullptr;
	CUX(cudaEventCreate(&e->ev_t0));
	CUX(cudaEventCreate(&e->ev_t1));
	const size_t df_bytes = (size_t) e->Q * e->XYZ * e->rs;
	const int ncopies = e->aa() ? 1 : 2;
	e->df_guard = (((size_t) (e->YZ + e->Y + 1) * e->rs) + 511) / 512 * 512;
	for (int i = 0; i < ncopies; i++) {
		CUX(cudaMalloc(&e->df_alloc[i], df_bytes + 2 * e->df_guard));
		CUX(cudaMemsetAsync(e->df_alloc[i], 0, df_bytes + 2 * e->df_guard, e->s_main));
		e->df[i] = (char*) e->df_alloc[i] + e->df_guard;
	}
	if (e->aa()) {
		// LBMX_TMA = 0 | even | odd | both: which A-A parities run k_bulk_tma (default: see DESIGN.md section 3.3)
		const char* want = std::getenv("LBMX_TMA");
		const std::string w = want ? want : LBMX_TMA_DEFAULT;
		e->tile_y = tma_tile_y(e->Y, (int) e->rs);
		if (w != "0" && e->tile_y > 0) {
			while ((1 << e->tile_y_shift) < e->tile_y)
				e->tile_y_shift++;
			e->use_tma[S_AA_EVEN] = w == "even" || w == "both";
			e->use_tma[S_AA_ODD] = w == "odd" || w == "both";
		}
		else
			e->tile_y = 0;
	}
	if (e->NM > 0) {
		CUX(cudaMalloc(&e->macro, (size_t) e->NM * e->XYZ * e->rs));
		CUX(cudaMemsetAsync(e->macro, 0, (size_t) e->NM * e->XYZ * e->rs, e->s_main));
	}
	CUX(cudaMalloc(&e->map, (size_t) e->XYZ * sizeof(int16_t)));
	CUX(cudaMemsetAsync(e->map, 0, (size_t) e->XYZ * sizeof(int16_t), e->s_main));
	CUX(cudaMalloc(&e->d_flag, sizeof(int)));
	CUX(cudaMalloc(&e->d_dirs, sizeof(int) * 18));
	CUX(cudaMemcpyAsync(e->d_dirs, e->h_dirs, sizeof(int) * 18, cudaMemcpyHostToDevice, e->s_main));
	CUX(cudaStreamSynchronize(e->s_main));
	// registers of the step kernels, for the measurement harness
	cudaFuncAttributes fa{};
	const int mode = e->aa() ? S_AA_EVEN : S_AB;
	const void* kb = e->f64() ? (const void*) e->kd.bulk[mode] : (const void*) e->kf.bulk[mode];
	const void* kq = e->f64() ? (const void*) e->kd.boundary : (const void*) e->kf.boundary;
	if (cudaFuncGetAttributes(&fa, kb) == cudaSuccess)
		e->stats.bulk_regs = fa.numRegs;
	if (cudaFuncGetAttributes(&fa, kq) == cudaSuccess)
		e->stats.boundary_regs = fa.numRegs;
	e->stats.bulk_block = BLOCK;
#undef CUX
	e->prm.lbmViscosity = 0.01;
	*out = e;
	return LBMX_OK;
}

static void release_peer_memory(lbmx_engine* e);

int lbmx_destroy(lbmx_engine* e)
{
	if (! e)
		return LBMX_OK;
	cudaSetDevice(e->dev);
	cudaDeviceSynchronize();
	release_peer_memory(e);
	if (e->comm && g_nccl.CommDestroy)
		g_nccl.CommDestroy(e->comm);
	if (e->my_flags)
		cudaFree(e->my_flags);
	if (e->h_halo_error)
		cudaFreeHost(e->h_halo_error);
	for (void* p : {e->df_alloc[0], e->df_alloc[1], e->macro, (void*) e->map, (void*) e->blist, (void*) e->inert_flags, e->profile, e->bouzidi, (void*) e->d_flag, (void*) e->d_dirs})
		if (p)
			cudaFree(p);
	if (e->pair_exec)
		cudaGraphExecDestroy(e->pair_exec);
	for (double* p : e->eq_stage)
		if (p)
			cudaFree(p);
	for (cudaEvent_t ev : {e->ev_edge, e->ev_comm, e->ev_main, e->ev_t0, e->ev_t1, e->ev_fork, e->ev_join})
		if (ev)
			cudaEventDestroy(ev);
	for (cudaStream_t s : {e->s_main, e->s_edge, e->s_comm})
		if (s)
			cudaStreamDestroy(s);
	delete e;
	return LBMX_OK;
}

int lbmx_get_layout(const lbmx_engine* e, lbmx_layout* out)
{
	if (! e || ! out)
		return fail(LBMX_ERR_ARG, "lbmx_get_layout: null argument");
	out->X_local = e->X;
	out->Y = e->Y;
	out->Z = e->Z;
	out->x_offset = e->x0;
	out->ghost_x = e->ox;
	out->XYZ = e->XYZ;
	out->Q = e->Q;
	out->n_macro = e->NM;
	out->sizeof_real = (int32_t) e->rs;
	out->dfmax = e->aa() ? 1 : 2;
	return LBMX_OK;
}

int lbmx_comm_unique_id(void* id128)
{
	std::string why;
	if (! g_nccl.load(why))
		return fail(LBMX_ERR_NCCL, why);
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	NC(g_nccl.GetUniqueId(&id));
	std::memcpy(id128, &id, 128);
	return LBMX_OK;
}

static void release_peer_memory(lbmx_engine* e)
{
	const bool shared = e->peer_right.mapped && e->peer_left.mapped && e->peer_right.flags == e->peer_left.flags;
	for (lbmx_engine::Peer* peer : {&e->peer_left, &e->peer_right}) {
		if (peer == &e->peer_right && shared) {
			*peer = lbmx_engine::Peer{};
			continue;
		}
		for (void* p : {peer->df_base[0], peer->df_base[1], (void*) peer->flags})
			if (p)
				cudaIpcCloseMemHandle(p);
		*peer = lbmx_engine::Peer{};
	}
	cudaGetLastError();
}

// Peer-memory halo exchange: map the neighbours' distribution arrays and arrival counters into this process (CUDA IPC; one node,
// GPUs connected by NVLink / NVSwitch).  The handles travel over the NCCL communicator that was just created.  Every rank must
// succeed, otherwise all of them keep the NCCL send/recv exchange (also selectable with LBMX_HALO=nccl).
static int setup_peer_memory(lbmx_engine* e)
{
	struct Packet
	{
		cudaIpcMemHandle_t df[2];
		cudaIpcMemHandle_t flags;
		int64_t X, XYZ, guard;
		int32_t ok, pad;
	};
	const char* mode = std::getenv("LBMX_HALO");
	int ok = ! (mode && std::strcmp(mode, "nccl") == 0);
	CU(cudaMalloc(&e->my_flags, 8 * sizeof(long long)));  // [0], [1]: arrival counters; [3], [4]: the neighbours' ready-for-batch announcements
	CU(cudaMemsetAsync(e->my_flags, 0, 8 * sizeof(long long), e->s_main));
	if (! e->h_halo_error) {
		CU(cudaHostAlloc((void**) &e->h_halo_error, sizeof(long long), cudaHostAllocMapped));
		*e->h_halo_error = 0;
		CU(cudaHostGetDevicePointer((void**) &e->d_halo_error, e->h_halo_error, 0));
	}
	if (const char* v = std::getenv("LBMX_HALO_TIMEOUT_S"))	 // 0: wait for ever (what an MPI receive does); default 120 s
		e->halo_timeout_ns = (long long) (std::atof(v) * 1e9);
	Packet mine{};
	if (ok) {
		ok = cudaIpcGetMemHandle(&mine.df[0], e->df_alloc[0]) == cudaSuccess && cudaIpcGetMemHandle(&mine.flags, e->my_flags) == cudaSuccess;
		if (ok && e->df[1])
			ok = cudaIpcGetMemHandle(&mine.df[1], e->df_alloc[1]) == cudaSuccess;
		cudaGetLastError();
	}
	mine.X = e->X;
	mine.XYZ = e->XYZ;
	mine.guard = (int64_t) e->df_guard;
	mine.ok = ok;
	Packet* d_pk = nullptr;	 // [0] mine, [1] from the left neighbour, [2] from the right neighbour
	CU(cudaMalloc(&d_pk, 3 * sizeof(Packet)));
	CU(cudaMemsetAsync(d_pk, 0, 3 * sizeof(Packet), e->s_main));  // stream-ordered before the sends below (see copy_components)
	CU(cudaMemcpyAsync(d_pk, &mine, sizeof(Packet), cudaMemcpyHostToDevice, e->s_main));
	CU(cudaStreamSynchronize(e->s_main));
	NC(g_nccl.GroupStart());  // same posting order as exchange_full_planes (left and right may be the same rank)
	if (e->right >= 0)
		NC(g_nccl.Send(d_pk, sizeof(Packet), ncclInt8, e->right, e->comm, e->s_main));
	if (e->left >= 0)
		NC(g_nccl.Recv(d_pk + 1, sizeof(Packet), ncclInt8, e->left, e->comm, e->s_main));
	if (e->left >= 0)
		NC(g_nccl.Send(d_pk, sizeof(Packet), ncclInt8, e->left, e->comm, e->s_main));
	if (e->right >= 0)
		NC(g_nccl.Recv(d_pk + 2, sizeof(Packet), ncclInt8, e->right, e->comm, e->s_main));
	NC(g_nccl.GroupEnd());
	CU(cudaStreamSynchronize(e->s_main));
	Packet got[3];
	CU(cudaMemcpyAsync(got, d_pk, 3 * sizeof(Packet), cudaMemcpyDeviceToHost, e->s_main));
	CU(cudaStreamSynchronize(e->s_main));
	ok = ok && (e->left < 0 || got[1].ok) && (e->right < 0 || got[2].ok);
	auto map_peer = [&](lbmx_engine::Peer& peer, const Packet& pk) {
		if (pk.guard != (int64_t) e->df_guard)	// same Y, Z and precision on every slab
			return false;
		if (cudaIpcOpenMemHandle(&peer.df_base[0], pk.df[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
			return false;
		peer.df[0] = (char*) peer.df_base[0] + e->df_guard;
		if (e->df[1]) {
			if (cudaIpcOpenMemHandle(&peer.df_base[1], pk.df[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
				return false;
			peer.df[1] = (char*) peer.df_base[1] + e->df_guard;
		}
		if (cudaIpcOpenMemHandle((void**) &peer.flags, pk.flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
			return false;
		peer.X = pk.X;
		peer.XYZ = pk.XYZ;
		peer.mapped = true;
		return true;
	};
	if (ok && e->left >= 0)
		ok = map_peer(e->peer_left, got[1]);
	if (ok && e->right >= 0) {
		if (e->right == e->left)
			e->peer_right = e->peer_left;  // two slabs, periodic: one neighbour on both sides, one mapping
		else
			ok = map_peer(e->peer_right, got[2]);
	}
	cudaGetLastError();
	// agreement: the exchange protocol must be the same on every rank
	int* d_ok = (int*) d_pk;
	CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, e->s_main));
	NC(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, e->comm, e->s_main));
	CU(cudaStreamSynchronize(e->s_main));
	int all_ok = 0;
	CU(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, e->s_main));
	CU(cudaStreamSynchronize(e->s_main));
	CU(cudaFree(d_pk));
	e->p2p = all_ok != 0;
	if (! e->p2p)
		release_peer_memory(e);
	return LBMX_OK;
}

int lbmx_comm_init(lbmx_engine* e, const void* id128)
{
	if (! e || ! id128)
		return fail(LBMX_ERR_ARG, "lbmx_comm_init: null argument");
	if (e->d.nranks < 2)
		return LBMX_OK;
	std::string why;
	if (! g_nccl.load(why))
		return fail(LBMX_ERR_NCCL, why);
	CU(cudaSetDevice(e->dev));
	ncclUniqueId id;
	std::memcpy(&id, id128, 128);
	NC(g_nccl.CommInitRank(&e->comm, e->d.nranks, id, e->d.rank));
	return setup_peer_memory(e);
}

static int exchange_full_planes(lbmx_engine* e, void* arr, int ncomp, size_t elem)
{
	// every component: my first / last interior plane -> the neighbours' ghost planes (map and initial-state synchronisation)
	if (e->ox == 0)
		return LBMX_OK;
	CU(cudaSetDevice(e->dev));
	const size_t pb = (size_t) e->YZ * elem;
	char* a = (char*) arr;
	auto plane = [&](int comp, int64_t xs) { return a + ((size_t) comp * e->XYZ + (size_t) xs * e->YZ) * elem; };
	if (e->d.nranks == 1) {
		if (! e->self_exchange)
			return LBMX_OK;
		for (int c = 0; c < ncomp; c++) {
			CU(cudaMemcpyAsync(plane(c, 0), plane(c, e->X), pb, cudaMemcpyDeviceToDevice, e->s_main));
			CU(cudaMemcpyAsync(plane(c, e->X + 1), plane(c, 1), pb, cudaMemcpyDeviceToDevice, e->s_main));
		}
		return LBMX_OK;
	}
	if (! e->comm)
		return fail(LBMX_ERR_STATE, "ghost-plane synchronisation needs lbmx_comm_init first");
	NC(g_nccl.GroupStart());
	// Posting order matters when both neighbours are the same rank (2 slabs, periodic): NCCL pairs the sends and receives of
	// a rank pair in posting order, so "what goes right" must be posted together with "what comes from the left".
	for (int c = 0; c < ncomp; c++) {
		if (e->right >= 0)
			NC(g_nccl.Send(plane(c, e->X), pb, ncclInt8, e->right, e->comm, e->s_main));
		if (e->left >= 0)
			NC(g_nccl.Recv(plane(c, 0), pb, ncclInt8, e->left, e->comm, e->s_main));
		if (e->left >= 0)
			NC(g_nccl.Send(plane(c, 1), pb, ncclInt8, e->left, e->comm, e->s_main));
		if (e->right >= 0)
			NC(g_nccl.Recv(plane(c, e->X + 1), pb, ncclInt8, e->right, e->comm, e->s_main));
	}
	NC(g_nccl.GroupEnd());
	return LBMX_OK;
}

int lbmx_map_upload(lbmx_engine* e, const int16_t* host_map, int with_ghosts)
{
	if (! e || ! host_map)
		return fail(LBMX_ERR_ARG, "lbmx_map_upload: null argument");
	int rc = copy_components(e, e->map, (void*) host_map, 1, sizeof(int16_t), with_ghosts != 0, true);
	if (rc)
		return rc;
	if (! with_ghosts && e->ox) {
		// LBM::synchronizeMapDevice (lbm.hpp:283-289); ghost planes without a neighbour keep GEO_FLUID (0)
		if ((rc = exchange_full_planes(e, e->map, 1, sizeof(int16_t))))
			return rc;
		CU(cudaStreamSynchronize(e->s_main));
	}
	// launch plan: cells that are neither GEO_FLUID nor GEO_PERIODIC nor a GEO_WALL away from the faces go to the boundary list, ordered by storage index
	// (so list ranges per x-plane are contiguous and neighbouring entries are neighbouring cells).  Counted and compacted on
	// the device: the map never travels back to the host.
	const int periodic = e->d.lattice == LBMX_D2Q9 ? (int) D2Q9::PERIODIC : (int) D3Q27::PERIODIC;	// D3Q19 shares the D3Q27 cell types
	const long long first_cell = (long long) e->ox * e->YZ, n_cells = (long long) e->X * e->YZ;
	const int outflow_right = e->d.lattice == LBMX_D2Q9 ? (int) D2Q9::OUTFLOW_RIGHT : (int) D3Q27::OUTFLOW_RIGHT;
	unsigned* d_counts = nullptr;
	CU(cudaMalloc(&d_counts, sizeof(unsigned) * (size_t) (e->X + 4)));
	CU(cudaMemsetAsync(d_counts, 0, sizeof(unsigned) * (size_t) (e->X + 4), e->s_main));
	const int nothing = e->d.lattice == LBMX_D2Q9 ? (int) D2Q9::NOTHING : (int) D3Q27::NOTHING;
	const int ndim = e->d.lattice == LBMX_D2Q9 ? 2 : 3;
	const int face_rule = ! e->aa() ? 0 : (e->ox == 0 ? 2 : 1);
	k_count_boundary_cells<<<dim3((unsigned) ((e->YZ + 255) / 256), (unsigned) e->X), 256, 0, e->s_main>>>(e->map, first_cell, (int) e->YZ, periodic, outflow_right, d_counts,
																											(int) e->Y, (int) e->Z, face_rule, nothing, ndim, (int) e->ox);
	e->stats.kernel_launches++;
	std::vector<unsigned> counts((size_t) e->X + 4);
	CU(cudaMemcpyAsync(counts.data(), d_counts, sizeof(unsigned) * (size_t) (e->X + 4), cudaMemcpyDeviceToHost, e->s_main));
	CU(cudaStreamSynchronize(e->s_main));
	CU(cudaFree(d_counts));
	// Under A-A such a cell reads, in place, populations that the cell to its left rewrites in the same step (the reference has the
	// same race).  Keeping the boundary list strictly after the bulk kernel makes the outcome reproducible.
	// The same holds for A-A cells on a bare lattice face (counts[X + 1], the reference steps out of its array there): their
	// stores land in slots other cells own.
	e->list_after_bulk = e->aa() && (counts[(size_t) e->X] > 0 || counts[(size_t) e->X + 1] > 0);
	e->plane_start.assign((size_t) e->X + 1, 0);
	for (int64_t x = 0; x < e->X; x++)
		e->plane_start[(size_t) x + 1] = e->plane_start[(size_t) x] + counts[(size_t) x];
	e->nb = e->plane_start[(size_t) e->X];
	e->n_bulk = n_cells - e->nb;
	if (e->blist) {
		CU(cudaFree(e->blist));
		e->blist = nullptr;
	}
	if (e->nb > 0) {
		CU(cudaMalloc(&e->blist, (size_t) e->nb * sizeof(uint32_t)));
		cub::CountingInputIterator<uint32_t> cells((uint32_t) first_cell);
		IsBoundaryCell pred{e->map, periodic, nothing, ndim, (int) e->ox, (int) e->X, (int) e->Y, (int) e->Z, ! e->aa()};
		int* d_selected = nullptr;
		void* d_temp = nullptr;
		size_t temp_bytes = 0;
		CU(cudaMalloc(&d_selected, sizeof(int)));
		CU(cub::DeviceSelect::If(nullptr, temp_bytes, cells, e->blist, d_selected, (int) n_cells, pred, e->s_main));
		CU(cudaMalloc(&d_temp, temp_bytes));
		CU(cub::DeviceSelect::If(d_temp, temp_bytes, cells, e->blist, d_selected, (int) n_cells, pred, e->s_main));	 // stable: output is sorted
		e->stats.kernel_launches++;
		int selected = 0;
		CU(cudaMemcpyAsync(&selected, d_selected, sizeof(int), cudaMemcpyDeviceToHost, e->s_main));
		CU(cudaStreamSynchronize(e->s_main));
		CU(cudaFree(d_temp));
		CU(cudaFree(d_selected));
		if (selected != e->nb)
			return fail(LBMX_ERR_STATE, "lbmx_map_upload: boundary-list compaction disagrees with the per-plane counts");
	}
	e->stats.boundary_cells = e->nb;
	e->stats.bulk_cells = e->n_bulk;
	e->stats.aa_cells_reaching_outside = counts[(size_t) e->X + 1];
	e->walls_in_bulk = counts[(size_t) e->X + 2];
	e->inert_cells = counts[(size_t) e->X + 3];
	// maps with sizeable GEO_NOTHING regions: one flag per 128 consecutive cells of a plane, so that CTAs of the bulk kernel whose cells
	// are all inert skip their (speculative) population loads
	if (e->inert_flags) {
		CU(cudaFree(e->inert_flags));
		e->inert_flags = nullptr;
	}
	if (e->inert_cells * 16 >= n_cells) {
		const int stride = (int) ((e->YZ + BLOCK - 1) / BLOCK);
		CU(cudaMalloc(&e->inert_flags, (size_t) stride * (size_t) (e->X + 2 * e->ox)));
		CU(cudaMemsetAsync(e->inert_flags, 0, (size_t) stride * (size_t) (e->X + 2 * e->ox), e->s_main));
		k_flag_inert_chunks<<<dim3((unsigned) stride, (unsigned) e->X), BLOCK, 0, e->s_main>>>(e->map, first_cell, (int) e->YZ, nothing, e->inert_flags + (size_t) e->ox * stride);
		e->stats.kernel_launches++;
		CU(cudaStreamSynchronize(e->s_main));
	}
	e->map_ready = true;
	e->state_version++;
	return LBMX_OK;
}

int lbmx_map_download(lbmx_engine* e, int16_t* host_map, int with_ghosts)
{
	if (! e || ! host_map)
		return fail(LBMX_ERR_ARG, "lbmx_map_download: null argument");
	return copy_components(e, e->map, host_map, 1, sizeof(int16_t), with_ghosts != 0, false);
}

static int set_eq_common(lbmx_engine* e, const double* rho, const double* vx, const double* vy, const double* vz, double crho, double cvx, double cvy, double cvz)
{
	CU(cudaSetDevice(e->dev));
	const bool field = rho != nullptr;
	const int64_t n = field ? e->X * e->YZ : e->XYZ;
	const int64_t cell0 = field ? e->ox * e->YZ : 0;
	const double* src[4] = {rho, vx, vy, vz};
	int rc = LBMX_OK;
	auto launch = [&](int64_t cells, int64_t first, const double* f0, const double* f1, const double* f2, const double* f3) {
		const unsigned blocks = (unsigned) ((cells + 127) / 128);
		if (e->f64())
			e->kd.set_equilibrium<<<blocks, 128, 0, e->s_main>>>((double*) e->df[0], e->XYZ, cells, first, e->d.eq, f0, f1, f2, f3, crho, cvx, cvy, cvz);
		else
			e->kf.set_equilibrium<<<blocks, 128, 0, e->s_main>>>((float*) e->df[0], e->XYZ, cells, first, e->d.eq, f0, f1, f2, f3, crho, cvx, cvy, cvz);
		e->stats.kernel_launches++;
	};
	if (! field)
		launch(n, cell0, nullptr, nullptr, nullptr, nullptr);
	else {
		// The four host fields stream through two staging sets of `chunk` cells: the copy of chunk c+1 (on the communication stream)
		// overlaps the equilibrium kernel of chunk c, and no lattice-sized temporary is allocated.
		int64_t chunk_cells = 8ll << 20;
		if (const char* v = std::getenv("LBMX_EQ_CHUNK"))
			chunk_cells = std::max<int64_t>(1, std::atoll(v));
		const int64_t chunk = std::min<int64_t>(n, chunk_cells);
		// the staging sets live as long as the engine (0.5 GB at the default chunk): a solver that re-initialises pays for them once
		if (e->eq_stage_cells != chunk) {
			for (int b = 0; b < 2; b++) {
				if (e->eq_stage[b])
					cudaFree(e->eq_stage[b]);
				e->eq_stage[b] = nullptr;
			}
			e->eq_stage_cells = 0;
			if (cudaMalloc(&e->eq_stage[0], (size_t) chunk * 4 * sizeof(double)) != cudaSuccess || cudaMalloc(&e->eq_stage[1], (size_t) chunk * 4 * sizeof(double)) != cudaSuccess)
				rc = fail(LBMX_ERR_CUDA, std::string("lbmx_df_set_equilibrium_field: ") + cudaGetErrorString(cudaGetLastError()));
			else
				e->eq_stage_cells = chunk;
		}
		double* stage[2] = {e->eq_stage[0], e->eq_stage[1]};
		cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
		for (int b = 0; b < 2 && rc == LBMX_OK; b++)
			if (cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&consumed[b], cudaEventDisableTiming) != cudaSuccess)
				rc = fail(LBMX_ERR_CUDA, std::string("lbmx_df_set_equilibrium_field: ") + cudaGetErrorString(cudaGetLastError()));
		if (rc == LBMX_OK) {
			cudaEventRecord(consumed[0], e->s_main);  // also orders the copies after whatever the compute stream did before
			cudaEventRecord(consumed[1], e->s_main);
		}
		for (int64_t off = 0, c = 0; off < n && rc == LBMX_OK; off += chunk, c++) {
			const int b = (int) (c & 1);
			const int64_t len = std::min(chunk, n - off);
			cudaStreamWaitEvent(e->s_comm, consumed[b], 0);
			for (int k = 0; k < 4; k++)
				if (src[k] && cudaMemcpyAsync(stage[b] + (size_t) k * chunk, src[k] + off, (size_t) len * sizeof(double), cudaMemcpyHostToDevice, e->s_comm) != cudaSuccess)
					rc = fail(LBMX_ERR_CUDA, std::string("lbmx_df_set_equilibrium_field: ") + cudaGetErrorString(cudaGetLastError()));
			cudaEventRecord(copied[b], e->s_comm);
			cudaStreamWaitEvent(e->s_main, copied[b], 0);
			launch(len, cell0 + off, stage[b], stage[b] + chunk, stage[b] + 2 * chunk, src[3] ? stage[b] + 3 * chunk : nullptr);
			cudaEventRecord(consumed[b], e->s_main);
		}
		if (cudaStreamSynchronize(e->s_comm) != cudaSuccess || cudaStreamSynchronize(e->s_main) != cudaSuccess)
			rc = rc ? rc : fail(LBMX_ERR_CUDA, "lbmx_df_set_equilibrium_field: transfer failed");
		for (int b = 0; b < 2; b++) {
			if (copied[b])
				cudaEventDestroy(copied[b]);
			if (consumed[b])
				cudaEventDestroy(consumed[b]);
		}
	}
	if (rc == LBMX_OK && (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(e->s_main) != cudaSuccess))
		rc = fail(LBMX_ERR_CUDA, "lbmx_df_set_equilibrium: kernel failed");
	if (rc)
		return rc;
	if (field && e->ox) {
		if ((rc = exchange_full_planes(e, e->df[0], e->Q, e->rs)))
			return rc;
	}
	// "copy the initialized DFs so that they are not overridden" (lbm_block.hpp:247-249)
	if (! e->aa())
		CU(cudaMemcpyAsync(e->df[1], e->df[0], (size_t) e->Q * e->XYZ * e->rs, cudaMemcpyDeviceToDevice, e->s_main));
	CU(cudaStreamSynchronize(e->s_main));
	return LBMX_OK;
}

int lbmx_df_set_equilibrium(lbmx_engine* e, double rho, double vx, double vy, double vz)
{
	if (! e)
		return fail(LBMX_ERR_ARG, "lbmx_df_set_equilibrium: null engine");
	return set_eq_common(e, nullptr, nullptr, nullptr, nullptr, rho, vx, vy, vz);
}

int lbmx_df_set_equilibrium_field(lbmx_engine* e, const double* rho, const double* vx, const double* vy, const double* vz)
{
	if (! e || ! rho || ! vx || ! vy)
		return fail(LBMX_ERR_ARG, "lbmx_df_set_equilibrium_field: null argument");
	return set_eq_common(e, rho, vx, vy, vz, 0, 0, 0, 0);
}

int lbmx_df_upload(lbmx_engine* e, int which, const void* host_df, int with_ghosts)
{
	if (! e || ! host_df || which < 0 || which > 1)
		return fail(LBMX_ERR_ARG, "lbmx_df_upload: bad argument");
	void* dst = which == 0 ? e->cur() : e->other();
	if (! dst)
		return fail(LBMX_ERR_ARG, "lbmx_df_upload: the A-A pattern has a single array");
	return copy_components(e, dst, (void*) host_df, e->Q, e->rs, with_ghosts != 0, true);
}

int lbmx_df_download(lbmx_engine* e, int which, void* host_df, int with_ghosts)
{
	if (! e || ! host_df || which < 0 || which > 1)
		return fail(LBMX_ERR_ARG, "lbmx_df_download: bad argument");
	void* src = which == 0 ? e->cur() : e->other();
	if (! src)
		return fail(LBMX_ERR_ARG, "lbmx_df_download: the A-A pattern has a single array");
	return copy_components(e, src, host_df, e->Q, e->rs, with_ghosts != 0, false);
}

int lbmx_df_sync_ghosts(lbmx_engine* e)
{
	if (! e)
		return fail(LBMX_ERR_ARG, "lbmx_df_sync_ghosts: null engine");
	int rc = exchange_full_planes(e, e->cur(), e->Q, e->rs);
	if (rc)
		return rc;
	CU(cudaStreamSynchronize(e->s_main));
	return LBMX_OK;
}

static int init_out_mode(const lbmx_engine* e)
{
	if (e->d.macro == LBMX_MACRO_WITH_MEAN_2D)	// computeInitialMacro calls the same outputMacro, open gates included
		return OUT_WITH_MEAN_2D + (e->prm.macro_gates & (LBMX_GATE_MEANS | LBMX_GATE_FLUCS));
	return e->d.macro == LBMX_MACRO_MEAN ? OUT_MEAN : OUT_DEFAULT;
}

int lbmx_macro_init(lbmx_engine* e)
{
	if (! e)
		return fail(LBMX_ERR_ARG, "lbmx_macro_init: null engine");
	if (e->NM == 0)
		return LBMX_OK;
	CU(cudaSetDevice(e->dev));
	const int64_t n = e->X * e->YZ;
	const unsigned blocks = (unsigned) ((n + 127) / 128);
	if (e->f64()) {
		KParams<double> p = make_params<double>(e);
		p.out_mode = init_out_mode(e);
		e->kd.initial_macro<<<blocks, 128, 0, e->s_main>>>(p);
	}
	else {
		KParams<float> p = make_params<float>(e);
		p.out_mode = init_out_mode(e);
		e->kf.initial_macro<<<blocks, 128, 0, e->s_main>>>(p);
	}
	e->stats.kernel_launches++;
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(e->s_main));
	return LBMX_OK;
}

int lbmx_macro_download(lbmx_engine* e, void* host_macro, int with_ghosts)
{
	if (! e || ! host_macro)
		return fail(LBMX_ERR_ARG, "lbmx_macro_download: null argument");
	if (e->NM == 0)
		return LBMX_OK;
	return copy_components(e, e->macro, host_macro, e->NM, e->rs, with_ghosts != 0, false);
}

int lbmx_macro_upload(lbmx_engine* e, const void* host_macro, int with_ghosts)
{
	if (! e || ! host_macro)
		return fail(LBMX_ERR_ARG, "lbmx_macro_upload: null argument");
	if (e->NM == 0)
		return LBMX_OK;
	return copy_components(e, e->macro, (void*) host_macro, e->NM, e->rs, with_ghosts != 0, true);
}

int lbmx_set_params(lbmx_engine* e, const lbmx_params* p)
{
	if (! e || ! p)
		return fail(LBMX_ERR_ARG, "lbmx_set_params: null argument");
	if (p->lbmViscosity == 0.0)
		return fail(LBMX_ERR_ARG, "lbmx_set_params: lbmViscosity must not be 0 (state.hpp:985-990 aborts on it)");
	if (std::memcmp(&e->prm, p, sizeof(lbmx_params)) != 0)
		e->state_version++;	 // captured step pairs carry the parameters by value
	e->prm = *p;
	return LBMX_OK;
}

int lbmx_set_inflow_profile(lbmx_engine* e, const void* host_profile, int64_t size_y, int64_t size_z)
{
	if (! e || ! host_profile || size_y < 1 || size_z < 1)
		return fail(LBMX_ERR_ARG, "lbmx_set_inflow_profile: bad argument");
	if (size_y < e->Y || size_z < e->Z)	 // the kernels read profile[y + z * size_y] for every inflow cell (y, z) of the lattice
		return fail(LBMX_ERR_ARG, "lbmx_set_inflow_profile: the profile must cover the lattice's (y, z) cross-section");
	CU(cudaSetDevice(e->dev));
	CU(cudaStreamSynchronize(e->s_main));
	if (e->profile)
		CU(cudaFree(e->profile));
	e->profile = nullptr;
	CU(cudaMalloc(&e->profile, (size_t) (size_y * size_z) * e->rs));
	CU(cudaMemcpyAsync(e->profile, host_profile, (size_t) (size_y * size_z) * e->rs, cudaMemcpyHostToDevice, e->s_main));	 // stream-ordered, see copy_components
	CU(cudaStreamSynchronize(e->s_main));
	e->profile_sy = size_y;
	e->state_version++;
	return LBMX_OK;
}

int lbmx_bouzidi_upload(lbmx_engine* e, const void* host_coeff)
{
	if (! e || ! host_coeff)
		return fail(LBMX_ERR_ARG, "lbmx_bouzidi_upload: null argument");
	if (e->d.lattice != LBMX_D2Q9 || e->aa())
		return fail(LBMX_ERR_UNSUPPORTED, "lbmx_bouzidi_upload: the near-wall interpolation exists for D2Q9 with A-B streaming only (d2q9/bc.h:140-167)");
	CU(cudaSetDevice(e->dev));
	if (! e->bouzidi) {
		CU(cudaMalloc(&e->bouzidi, (size_t) 8 * e->XYZ * e->rs));
		CU(cudaMemsetAsync(e->bouzidi, 0xbf, (size_t) 8 * e->XYZ * e->rs, e->s_main));	// ghost planes: a negative value (0xbfbf... < 0 in both precisions)
		e->state_version++;
	}
	return copy_components(e, e->bouzidi, (void*) host_coeff, 8, e->rs, false, true);
}

int lbmx_step(lbmx_engine* e, int64_t nsteps)
{
	if (! e || nsteps < 0)
		return fail(LBMX_ERR_ARG, "lbmx_step: bad argument");
	if (e->d.inflow == LBMX_INFLOW_PROFILE_YZ && ! e->profile)
		return fail(LBMX_ERR_STATE, "lbmx_step: inflow profile selected but lbmx_set_inflow_profile was not called");
	if (int rc = halo_error(e, "lbmx_step"))
		return rc;
	return step_dispatch(e, nsteps);
}

int lbmx_sync(lbmx_engine* e)
{
	if (! e)
		return fail(LBMX_ERR_ARG, "lbmx_sync: null engine");
	CU(cudaSetDevice(e->dev));
	CU(cudaStreamSynchronize(e->s_main));
	CU(cudaStreamSynchronize(e->s_edge));
	CU(cudaStreamSynchronize(e->s_comm));
	return halo_error(e, "lbmx_sync");
}

int lbmx_step_timed(lbmx_engine* e, int64_t nsteps, float* elapsed_ms)
{
	if (! e || ! elapsed_ms)
		return fail(LBMX_ERR_ARG, "lbmx_step_timed: null argument");
	int rc = lbmx_sync(e);
	if (rc)
		return rc;
	CU(cudaEventRecord(e->ev_t0, e->s_main));
	if ((rc = lbmx_step(e, nsteps)))
		return rc;
	CU(cudaEventRecord(e->ev_t1, e->s_main));  // s_main has waited for the edge and communication streams
	CU(cudaEventSynchronize(e->ev_t1));
	CU(cudaEventElapsedTime(elapsed_ms, e->ev_t0, e->ev_t1));
	return lbmx_sync(e);
}

int lbmx_halo_time(lbmx_engine* e, int32_t reps, float* ms_per_exchange)
{
	if (! e || ! ms_per_exchange || reps < 1)
		return fail(LBMX_ERR_ARG, "lbmx_halo_time: bad argument");
	if (e->ox == 0 || e->iter == 0)
		return fail(LBMX_ERR_STATE, "lbmx_halo_time: needs ghost planes and at least one completed step");
	int rc = lbmx_sync(e);
	if (rc)
		return rc;
	CU(cudaEventRecord(e->ev_t0, e->s_comm));
	for (int32_t r = 0; r < reps; r++) {
		e->iter--;	// the exchange that follows step (iter - 1): same planes, same slots, same values -> idempotent
		void* arr = e->aa() ? e->df[0] : e->other();
		rc = e->f64() ? exchange<double>(e, arr) : exchange<float>(e, arr);
		e->iter++;
		if (rc)
			return rc;
	}
	await_halo(e, e->s_comm);  // peer-memory exchange: the neighbours' planes have arrived as well
	CU(cudaEventRecord(e->ev_t1, e->s_comm));
	CU(cudaEventSynchronize(e->ev_t1));
	float ms = 0;
	CU(cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1));
	*ms_per_exchange = ms / (float) reps;
	return lbmx_sync(e);
}

int lbmx_get_iterations(const lbmx_engine* e, int64_t* it)
{
	if (! e || ! it)
		return fail(LBMX_ERR_ARG, "lbmx_get_iterations: null argument");
	*it = e->iter;
	return LBMX_OK;
}

int lbmx_set_iterations(lbmx_engine* e, int64_t it)
{
	if (! e || it < 0)
		return fail(LBMX_ERR_ARG, "lbmx_set_iterations: bad argument");
	e->iter = it;
	return LBMX_OK;
}

int lbmx_has_nan(lbmx_engine* e, int32_t* flag)
{
	if (! e || ! flag)
		return fail(LBMX_ERR_ARG, "lbmx_has_nan: null argument");
	*flag = 0;
	if (e->NM == 0)
		return LBMX_OK;
	CU(cudaSetDevice(e->dev));
	CU(cudaMemsetAsync(e->d_flag, 0, sizeof(int), e->s_main));
	if (e->f64())
		k_has_nan<double><<<592, 256, 0, e->s_main>>>((const double*) e->macro, e->XYZ, e->d_flag);
	else
		k_has_nan<float><<<592, 256, 0, e->s_main>>>((const float*) e->macro, e->XYZ, e->d_flag);
	e->stats.kernel_launches++;
	CU(cudaGetLastError());
	int h = 0;
	CU(cudaMemcpyAsync(&h, e->d_flag, sizeof(int), cudaMemcpyDeviceToHost, e->s_main));
	CU(cudaStreamSynchronize(e->s_main));
	*flag = h;
	return halo_error(e, "lbmx_has_nan");
}

int lbmx_get_device_ptrs(lbmx_engine* e, lbmx_ptrs* out)
{
	if (! e || ! out)
		return fail(LBMX_ERR_ARG, "lbmx_get_device_ptrs: null argument");
	out->dfs[0] = e->cur();
	out->dfs[1] = e->other();
	out->dmacro = e->macro;
	out->dmap = e->map;
	out->even_iter = (e->iter % 2) == 0;
	out->reserved = 0;
	return LBMX_OK;
}

int lbmx_get_stats(lbmx_engine* e, lbmx_stats* out)
{
	if (! e || ! out)
		return fail(LBMX_ERR_ARG, "lbmx_get_stats: null argument");
	*out = e->stats;
	out->halo_peer_memory = e->p2p ? 1 : 0;
	return LBMX_OK;
}

}  // extern "C"
