// engine_host.cpp -- TEST TOOL: the engine's CUDA kernels (tnl_lbm_b200/csrc/kernels.cuh: k_bulk, k_boundary, k_set_equilibrium,
// k_initial_macro, with everything they include) compiled for the HOST and run thread by thread over the grid the engine would
// launch, behind the C interface of the CPU checkers (oracle/oracle_api.h).  tests/test_kernels_on_host.py compares the result with
// the restatement of the reference on every golden case, so streaming offsets, the cell-type dispatch, the boundary list, the
// invariant division and the macro modes are checked where there is no GPU.  Never loaded by the product (tnl_lbm_b200/binding.py
// knows only liblbmx.so, which has no CPU path).
//
// Why running the threads one after another is a faithful emulation: under both streaming patterns every population slot is read
// and written by exactly one cell per step, so the result of a launch does not depend on the order of its threads.
//
// Built by tests/test_kernels_on_host.py: one object per kernel family (-DHK_LAT= -DHK_KIND= -DHK_NAME=), one with -DHK_MAIN,
// g++ -ffp-contract=off; -DLBMX_STRICT=1 selects the parity arithmetic (bit-identical to the reference's strict build), 0 the default.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../oracle/oracle_api.h"

#ifndef HK_MAIN
// ---- what kernels.cuh needs from the CUDA language ----------------------------------------------------------------------------
	#define __host__
	#define __device__
	#define __global__
	#define __forceinline__ inline
	#define __launch_bounds__(...)
	#define __noinline__
using std::sqrt;
struct HkDim3
{
	unsigned x = 1, y = 1, z = 1;
};
// Set by the launch loops below, one "thread" at a time.  One instance for the whole library (inline variables): kernels whose
// template arguments do not name the operator (k_set_equilibrium, k_initial_macro) are merged across the family objects by the linker.
inline HkDim3 blockIdx, threadIdx, blockDim, gridDim;
namespace {
template <typename T>
inline T __ldg(const T* p)
{
	return *p;
}
template <typename T>
inline T __ldcg(const T* p)
{
	return *p;
}
template <typename T>
inline T __ldcs(const T* p)
{
	return *p;
}
template <typename T>
inline T __ldlu(const T* p)
{
	return *p;
}
template <typename T>
inline void __stcs(T* p, T v)
{
	*p = v;
}
template <typename T>
inline void __stcg(T* p, T v)
{
	*p = v;
}
template <typename T>
inline void __stwt(T* p, T v)
{
	*p = v;
}
inline unsigned __umulhi(unsigned a, unsigned b)
{
	return (unsigned) (((unsigned long long) a * b) >> 32);
}
inline bool __any_sync(unsigned, bool b)
{
	return b;
}
inline int atomicOr(int* p, int v)
{
	const int old = *p;
	*p |= v;
	return old;
}
}  // namespace
	#ifndef LBMX_STRICT
		#define LBMX_STRICT 1
	#endif
	#include "../../tnl_lbm_b200/csrc/kernels.cuh"
using namespace lbmx;

namespace {


// what lbmx_create / make_params / launch_range do for a single slab (tnl_lbm_b200/csrc/engine.cu), restated for the emulation
template <typename L, typename R>
KParams<R> params_for(const oracle_desc* d, const oracle_params* op, R* cur, R* out, R* macro, const int16_t* map, const uint32_t* blist, int nb, int64_t iteration)
{
	KParams<R> p{};
	const bool aa = d->streaming == ORC_STREAM_AA;
	p.cur = cur;
	p.out = out;
	p.macro = macro;
	p.map = map;
	p.profile = (const R*) op->vx_profile;
	p.bouzidi = (const R*) op->bouzidi_coeff;
	p.blist = blist;
	p.X = (int) d->X;
	p.Y = (int) d->Y;
	p.Z = (int) d->Z;
	p.ox = (int) d->ox;
	p.YZ = p.Y * p.Z;
	p.XYZ = (long long) (d->X + 2 * d->ox) * p.YZ;
	for (int q = 0; q < L::Q; q++) {
		p.rd[q] = cur + (size_t) q * p.XYZ;
		p.wr[q] = (aa ? cur : out) + (size_t) q * p.XYZ;
	}
	if (p.Y > 1) {	// division by the invariant Y, as lbmx_create prepares it
		unsigned lg = 0;
		while (((int64_t) 1 << lg) < p.Y)
			lg++;
		const unsigned __int128 one = 1;
		p.ydiv_mul = (unsigned) (((one << (31 + lg)) + (unsigned __int128) p.Y - 1) / (unsigned __int128) p.Y);
		p.ydiv_shift = lg - 1;
	}
	p.x_begin = 0;
	p.nb_begin = 0;
	p.nb_end = nb;
	p.wrap = d->nproc == 1 ? 1 : 0;
	p.profile_sy = (int) op->profile_size_y;
	p.eq = d->eq;
	p.inflow = d->inflow;
	p.stream = aa ? (iteration % 2 == 0 ? S_AA_EVEN : S_AA_ODD) : S_AB;
	p.out_mode = OUT_NONE;
	p.stat_counter = op->stat_counter;
	p.kahan_rho = d->lattice == ORC_D3Q27 && d->coll == ORC_COLL_CUM_HP_RHO;
	const bool vm = d->macro == ORC_MACRO_VOID;	 // MACRO_Void::copyQuantities is empty: viscosity 1, no force (d3q27/macro.h:174-188)
	p.phys.nu = vm ? R(1) : (R) op->lbmViscosity;
	set_rates(p.phys);
	p.phys.fx = vm ? R(0) : (R) op->fx;
	p.phys.fy = vm ? R(0) : (R) op->fy;
	p.phys.fz = (vm || d->lattice == ORC_D2Q9) ? R(0) : (R) op->fz;
	p.in_vx = (R) op->inflow_vx;
	p.in_vy = (R) op->inflow_vy;
	p.in_vz = (R) op->inflow_vz;
	return p;
}

template <typename Kernel, typename P>
void launch(Kernel k, unsigned gx, unsigned gy, unsigned block, const P& p)
{
	gridDim.x = gx;
	gridDim.y = gy;
	blockDim.x = block;
	for (unsigned by = 0; by < gy; by++)
		for (unsigned bx = 0; bx < gx; bx++)
			for (unsigned t = 0; t < block; t++) {
				blockIdx.x = bx;
				blockIdx.y = by;
				threadIdx.x = t;
				k(p);
			}
}

int out_mode_of(const oracle_desc* d, const oracle_params* op)
{
	switch (d->macro) {
		case ORC_MACRO_DEFAULT: return OUT_DEFAULT;	 // the reference writes rho, u every step
		case ORC_MACRO_MEAN: return OUT_MEAN;
		case ORC_MACRO_WITH_MEAN_2D: return OUT_WITH_MEAN_2D + (op->macro_gates & 3);
		default: return OUT_NONE;
	}
}

template <typename L, int KIND, typename R>
int step_family(const oracle_desc* d, const oracle_params* op, void* df_a, void* df_b, void* macro, const int16_t* map, int64_t iteration, int32_t nsteps)
{
	const bool aa = d->streaming == ORC_STREAM_AA;
	const int YZ = (int) (d->Y * d->Z);
	// boundary list in storage order over the interior planes (lbmx_map_upload)
	std::vector<uint32_t> blist;
	for (long long c = (long long) d->ox * YZ; c < (long long) (d->ox + d->X) * YZ; c++) {
		const int xs = (int) (c / YZ), yz = (int) (c - (long long) xs * YZ), z = yz / (int) d->Y;
		if (cell_in_boundary_list(map[c], (int) L::FLUID, (int) L::PERIODIC, (int) L::WALL, (int) L::NOTHING, cell_on_face(L::NDIM, (int) d->ox, (int) d->X, (int) d->Y, (int) d->Z, xs - (int) d->ox, yz - z * (int) d->Y, z), ! aa))
			blist.push_back((uint32_t) c);
	}
	// inert-chunk flags (lbmx_map_upload builds them for maps with sizeable GEO_NOTHING regions; here: whenever such a cell exists)
	const int stride = (YZ + LBMX_BULK_BLOCK - 1) / LBMX_BULK_BLOCK;
	std::vector<uint8_t> inert((size_t) stride * (size_t) (d->X + 2 * d->ox), 0);
	bool any_inert = false;
	for (long long xs = d->ox; xs < d->ox + d->X; xs++)
		for (int ch = 0; ch < stride; ch++) {
			bool all = true;
			for (int i = ch * LBMX_BULK_BLOCK; i < std::min(YZ, (ch + 1) * LBMX_BULK_BLOCK); i++) {
				all = all && map[xs * YZ + i] == (int) L::NOTHING;
				any_inert = any_inert || map[xs * YZ + i] == (int) L::NOTHING;
			}
			inert[(size_t) xs * stride + ch] = all ? 1 : 0;
		}
	const StepKernels<R> K = make_step_kernels<L, KIND, R>();
	for (int32_t s = 0; s < nsteps; s++) {
		const int64_t it = iteration + s;
		R* cur = (R*) ((aa || it % 2 == 0) ? df_a : df_b);
		R* out = aa ? nullptr : (R*) (it % 2 == 0 ? df_b : df_a);
		KParams<R> p = params_for<L, R>(d, op, cur, out, (R*) macro, map, blist.data(), (int) blist.size(), it);
		p.out_mode = out_mode_of(d, op);
		p.stat_counter = op->stat_counter + s;
		p.inert = any_inert ? inert.data() : nullptr;
		p.inert_stride = stride;
		const int per_cta = LBMX_BULK_BLOCK * K.cpt[p.stream];
		launch(K.bulk[p.stream], (unsigned) ((YZ + per_cta - 1) / per_cta), (unsigned) d->X, LBMX_BULK_BLOCK, p);
		if (! blist.empty())
			launch(K.boundary, (unsigned) ((blist.size() + LBMX_BULK_BLOCK - 1) / LBMX_BULK_BLOCK), 1, LBMX_BULK_BLOCK, p);
	}
	return 0;
}

template <typename L, int KIND, typename R>
int set_eq_family(const oracle_desc* d, void* df, const double* rho, const double* vx, const double* vy, const double* vz, double crho, double cvx, double cvy, double cvz)
{
	if (rho && d->ox != 0)
		return 2;  // the engine's field variant takes the interior cells only; the emulation covers unghosted lattices
	const long long XYZ = (long long) (d->X + 2 * d->ox) * d->Y * d->Z;
	const StepKernels<R> K = make_step_kernels<L, KIND, R>();
	gridDim.x = (unsigned) ((XYZ + 127) / 128);
	blockDim.x = 128;
	for (unsigned b = 0; b < gridDim.x; b++)
		for (unsigned t = 0; t < 128; t++) {
			blockIdx.x = b;
			threadIdx.x = t;
			K.set_equilibrium((R*) df, XYZ, XYZ, 0, d->eq, rho, vx, vy, vz, crho, cvx, cvy, cvz);
		}
	return 0;
}

template <typename L, int KIND, typename R>
int initial_macro_family(const oracle_desc* d, const oracle_params* op, void* df, void* macro)
{
	KParams<R> p = params_for<L, R>(d, op, (R*) df, nullptr, (R*) macro, nullptr, nullptr, 0, 0);
	p.out_mode = out_mode_of(d, op);
	if (p.out_mode == OUT_NONE)
		return 0;
	const long long n = (long long) d->X * d->Y * d->Z;
	const StepKernels<R> K = make_step_kernels<L, KIND, R>();
	gridDim.x = (unsigned) ((n + 127) / 128);
	blockDim.x = 128;
	for (unsigned b = 0; b < gridDim.x; b++)
		for (unsigned t = 0; t < 128; t++) {
			blockIdx.x = b;
			threadIdx.x = t;
			K.initial_macro(p);
		}
	return 0;
}

}  // namespace

	#ifdef HK_SERVICE
// ---- the plane copies of the halo exchange (k_has_nan votes across a warp: not reproducible one thread at a time) ----------------
extern "C" {
// k_copy_planes as the engine launches it (exchange<R> in engine.cu): grid (ceil(YZ / 256), n_dirs), 256 threads
int hk_copy_planes(int f64, void* dst, const void* src, long long XYZ, int YZ, int n_dirs, const int* dirs, long long src_plane, long long dst_plane, long long dst_XYZ)
{
	gridDim.x = (unsigned) ((YZ + 255) / 256);
	gridDim.y = (unsigned) n_dirs;
	blockDim.x = 256;
	for (unsigned by = 0; by < gridDim.y; by++)
		for (unsigned bx = 0; bx < gridDim.x; bx++)
			for (unsigned t = 0; t < 256; t++) {
				blockIdx.x = bx;
				blockIdx.y = by;
				threadIdx.x = t;
				if (f64)
					k_copy_planes<double>((double*) dst, (const double*) src, XYZ, YZ, n_dirs, dirs, src_plane, dst_plane, dst_XYZ);
				else
					k_copy_planes<float>((float*) dst, (const float*) src, XYZ, YZ, n_dirs, dirs, src_plane, dst_plane, dst_XYZ);
			}
	return 0;
}
}
	#else
	#define HK_CAT_(a, b) a##b
	#define HK_CAT(a, b) HK_CAT_(a, b)
extern "C" {
int HK_CAT(hk_step_, HK_NAME)(const oracle_desc* d, const oracle_params* p, void* a, void* b, void* mac, const int16_t* map, int64_t it, int32_t n)
{
	return d->precision == ORC_F64 ? step_family<HK_LAT, HK_KIND, double>(d, p, a, b, mac, map, it, n) : step_family<HK_LAT, HK_KIND, float>(d, p, a, b, mac, map, it, n);
}
int HK_CAT(hk_set_eq_, HK_NAME)(const oracle_desc* d, void* df, const double* rho, const double* vx, const double* vy, const double* vz, double crho, double cvx, double cvy, double cvz)
{
	return d->precision == ORC_F64 ? set_eq_family<HK_LAT, HK_KIND, double>(d, df, rho, vx, vy, vz, crho, cvx, cvy, cvz)
								   : set_eq_family<HK_LAT, HK_KIND, float>(d, df, rho, vx, vy, vz, crho, cvx, cvy, cvz);
}
int HK_CAT(hk_initial_macro_, HK_NAME)(const oracle_desc* d, const oracle_params* p, void* df, void* mac)
{
	return d->precision == ORC_F64 ? initial_macro_family<HK_LAT, HK_KIND, double>(d, p, df, mac) : initial_macro_family<HK_LAT, HK_KIND, float>(d, p, df, mac);
}
}
	#endif	// HK_SERVICE

#else  // HK_MAIN: the oracle_api.h entry points, dispatching on (lattice, operator) to the family objects that were linked in
	#include <dlfcn.h>

extern "C" const char* oracle_kind(void);
namespace {
struct Family
{
	int (*step)(const oracle_desc*, const oracle_params*, void*, void*, void*, const int16_t*, int64_t, int32_t) = nullptr;
	int (*set_eq)(const oracle_desc*, void*, const double*, const double*, const double*, const double*, double, double, double, double) = nullptr;
	int (*initial_macro)(const oracle_desc*, const oracle_params*, void*, void*) = nullptr;
};

// family objects are named <lattice>_<operator number>: found by symbol name, so the set that is linked in may be any subset
bool find_family(const oracle_desc* d, Family& f)
{
	const char* lat = d->lattice == ORC_D3Q27 ? "d3q27" : d->lattice == ORC_D2Q9 ? "d2q9" : d->lattice == ORC_D3Q19 ? "d3q19" : nullptr;
	if (! lat)
		return false;
	char name[96];
	static void* self = nullptr;
	if (! self) {  // this library (loaded with local scope by ctypes), not the main program
		Dl_info info;
		if (! dladdr((void*) &oracle_kind, &info) || ! (self = dlopen(info.dli_fname, RTLD_NOW | RTLD_NOLOAD)))
			return false;
	}
	std::snprintf(name, sizeof name, "hk_step_%s_%d", lat, (int) d->coll);
	f.step = (decltype(f.step)) dlsym(self, name);
	std::snprintf(name, sizeof name, "hk_set_eq_%s_%d", lat, (int) d->coll);
	f.set_eq = (decltype(f.set_eq)) dlsym(self, name);
	std::snprintf(name, sizeof name, "hk_initial_macro_%s_%d", lat, (int) d->coll);
	f.initial_macro = (decltype(f.initial_macro)) dlsym(self, name);
	return f.step && f.set_eq && f.initial_macro;
}
}  // namespace

extern "C" {
const char* oracle_kind(void)
{
	return "engine_host";
}
int oracle_supported(const oracle_desc* d)
{
	Family f;
	return find_family(d, f) ? 0 : 1;
}
int oracle_step(const oracle_desc* d, const oracle_params* p, void* df_a, void* df_b, void* macro, const int16_t* map, int64_t iteration, int32_t nsteps, int32_t)
{
	Family f;
	return find_family(d, f) ? f.step(d, p, df_a, df_b, macro, map, iteration, nsteps) : 1;
}
int oracle_set_equilibrium(const oracle_desc* d, void* df, double rho, double vx, double vy, double vz)
{
	Family f;
	return find_family(d, f) ? f.set_eq(d, df, nullptr, nullptr, nullptr, nullptr, rho, vx, vy, vz) : 1;
}
int oracle_set_equilibrium_field(const oracle_desc* d, void* df, const double* rho, const double* vx, const double* vy, const double* vz)
{
	Family f;
	return find_family(d, f) ? f.set_eq(d, df, rho, vx, vy, vz, 0, 0, 0, 0) : 1;
}
int oracle_initial_macro(const oracle_desc* d, const oracle_params* p, void* df, void* macro)
{
	Family f;
	return find_family(d, f) ? f.initial_macro(d, p, df, macro) : 1;
}
}
#endif
