// lbmx_host.h -- host-side mirror of TNL-LBM's solver-facing interface on top of the C ABI (include/lbmx.h).
//
// The reference is a header-only C++ template framework: a solver composes LBM_CONFIG<TRAITS, KernelStruct, DATA, COLL, EQ,
// STREAMING, BC, MACRO> (include/lbm3d/defs.h:169-250), derives StateLocal from State<NSE> (include/lbm3d/state.h:89-330),
// paints the map through nse.setBoundaryX/Y/Z / setMap, sets block.data.* and calls execute(state) (include/lbm3d/core.h:38-101).
// This header keeps those names, argument meanings and error behaviour (exceptions), but the trait classes are *tags*: they
// carry no device code, they select a kernel family of liblbmx.so through the lbmx_desc enums.  Everything that touches the GPU
// goes through the C ABI.  Out of scope here: writers (VTK/ADIOS), checkpoints, IBM, MPI -- the hooks exist and
// are no-ops, so that unmodified solvers compile and run.
//
// Device-side user code cannot cross a C ABI.  The finite set the reference's own solvers use is recognised structurally:
//   DATA with member `vx_profile`           -> LBMX_INFLOW_PROFILE_YZ   (NSE_Data_XProfileInflow, sim_NSE/sim_2.cu:16-33)
//   DATA with member `inflow_vx`            -> LBMX_INFLOW_CONST        (NSE_Data_ConstInflow lbm_data.h:98-115, NSE2D_Data_ConstInflow)
//   otherwise                               -> LBMX_INFLOW_NONE         (NSE_Data_NoInflow lbm_data.h:117-131)
// Custom MACRO / COLL / BC classes are rejected at compile time (they lack the lbmx_* tag constants).
#pragma once

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "lbmx.h"

// The reference's state.h pulls fmt and spdlog in for its solvers (state.h:10,13); unmodified solvers rely on that.
#include <iostream>
#if __has_include(<fmt/core.h>)
	#include <fmt/core.h>
#endif
#if __has_include(<spdlog/spdlog.h>)
	#include <spdlog/spdlog.h>
#endif

#if ! defined(AB_PATTERN) && ! defined(AA_PATTERN)
	#define AB_PATTERN	// the reference's default (defs.h:3-9)
#endif
#ifndef CUDA_HOSTDEV
	#define CUDA_HOSTDEV
#endif
#ifndef __cuda_callable__
	#define __cuda_callable__
#endif

// ---------------------------------------------------------------------------------------------------------------------------
// the sliver of TNL the solvers name themselves (StaticVector, MPI::Comm, sqr); single process, no MPI underneath
// ---------------------------------------------------------------------------------------------------------------------------
#ifndef LBMX_HAVE_REAL_TNL
namespace TNL {
namespace Containers {
template <int N, typename T>
struct StaticVector
{
	T d[N]{};
	StaticVector() = default;
	StaticVector(T fill)
	{
		for (int i = 0; i < N; i++)
			d[i] = fill;
	}
	StaticVector(T a, T b, T c) : d{a, b, c} {}
	T& x() { return d[0]; }
	T& y() { return d[1]; }
	T& z() { return d[2]; }
	const T& x() const { return d[0]; }
	const T& y() const { return d[1]; }
	const T& z() const { return d[2]; }
	T& operator[](int i) { return d[i]; }
	const T& operator[](int i) const { return d[i]; }
};
}  // namespace Containers
namespace Devices {
struct Host {};
struct Cuda {};
}  // namespace Devices
namespace MPI {
struct Comm
{
	int dummy = 0;
};
struct ScopedInitializer
{
	ScopedInitializer(int&, char**&) {}
};
inline int GetSize(const Comm&) { return 1; }
inline int GetRank(const Comm&) { return 0; }
template <typename T, typename Op>
inline T reduce(T v, Op, const Comm&)
{
	return v;
}
template <typename T>
inline void Bcast(T*, int, int, const Comm&)
{}
}  // namespace MPI
template <typename T>
inline T sqr(T v)
{
	return v * v;
}
struct Timer
{
	std::chrono::steady_clock::time_point t0;
	double acc = 0;
	bool running = false;
	void start()
	{
		t0 = std::chrono::steady_clock::now();
		running = true;
	}
	void stop()
	{
		if (running)
			acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		running = false;
	}
	double getRealTime() const { return acc + (running ? std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() : 0.0); }
};
}  // namespace TNL
	#ifndef MPI_COMM_WORLD
static const TNL::MPI::Comm MPI_COMM_WORLD{};
enum lbmx_mpi_op { MPI_SUM, MPI_LOR, MPI_LAND, MPI_MAX, MPI_MIN };
	#endif
#endif
using TNLMPI_INIT = TNL::MPI::ScopedInitializer;

namespace lbmx_host {
inline void check(int rc, const char* what)
{
	if (rc != LBMX_OK)
		throw std::runtime_error(std::string(what) + ": " + lbmx_last_error());
}
inline void log_info(const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	std::vfprintf(stdout, fmt, ap);
	va_end(ap);
	std::fputc('\n', stdout);
	std::fflush(stdout);
}
template <typename T, typename = void>
struct has_vx_profile : std::false_type {};
template <typename T>
struct has_vx_profile<T, std::void_t<decltype(std::declval<T&>().vx_profile)>> : std::true_type {};
template <typename T, typename = void>
struct has_inflow_vx : std::false_type {};
template <typename T>
struct has_inflow_vx<T, std::void_t<decltype(std::declval<T&>().inflow_vx)>> : std::true_type {};
template <typename T, typename = void>
struct has_inflow_vy : std::false_type {};
template <typename T>
struct has_inflow_vy<T, std::void_t<decltype(std::declval<T&>().inflow_vy)>> : std::true_type {};
template <typename T, typename = void>
struct has_inflow_vz : std::false_type {};
template <typename T>
struct has_inflow_vz<T, std::void_t<decltype(std::declval<T&>().inflow_vz)>> : std::true_type {};
}  // namespace lbmx_host

// ---------------------------------------------------------------------------------------------------------------------------
// defs.h: traits, DF roles, direction enums, KernelStruct tags, LBM_CONFIG
// ---------------------------------------------------------------------------------------------------------------------------
#if defined(AB_PATTERN)
enum : std::uint8_t { df_cur, df_out, DFMAX };
#else
enum : std::uint8_t { df_cur, DFMAX };
#endif

template <typename _dreal = float, typename _real = double, typename _idx = long int, typename _map_t = short int>
struct Traits
{
	using real = _real;
	using dreal = _dreal;
	using idx = _idx;
	using map_t = _map_t;
	using point_t = TNL::Containers::StaticVector<3, real>;
	using idx3d = TNL::Containers::StaticVector<3, idx>;
	static_assert(std::is_same<_dreal, float>::value || std::is_same<_dreal, double>::value, "dreal must be float or double");
	static_assert(sizeof(_map_t) == 2, "the engine stores cell types as 16-bit integers (defs.h:75)");
	static constexpr int lbmx_precision = std::is_same<_dreal, double>::value ? LBMX_F64 : LBMX_F32;
};
using TraitsSP = Traits<float>;
using TraitsDP = Traits<double>;

enum : std::uint8_t { zz = 0, pz = 1, mz = 2, zp = 3, zm = 4, pp = 5, mm = 6, pm = 7, mp = 8 };	 // defs.h:257-270
enum : std::uint8_t {  // defs.h:273-305
	zzz = 0, pzz = 1, mzz = 2, zpz = 3, zmz = 4, zzp = 5, zzm = 6, ppz = 7, mmz = 8, pmz = 9, mpz = 10, pzp = 11, mzm = 12, pzm = 13, mzp = 14,
	zpp = 15, zmm = 16, zpm = 17, zmp = 18, ppp = 19, mmm = 20, ppm = 21, mmp = 22, pmp = 23, mpm = 24, pmm = 25, mpp = 26
};

template <typename REAL>
struct D2Q9_KernelStruct
{
	static constexpr int D = 2;
	static constexpr int Q = 9;
	static constexpr int lbmx_lattice = LBMX_D2Q9;
};
template <typename REAL>
struct D3Q27_KernelStruct
{
	static constexpr int Q = 27;
	static constexpr int lbmx_lattice = LBMX_D3Q27;
};

// equilibria (tags)
template <typename TRAITS>
struct D3Q27_EQ { static constexpr int lbmx_eq = LBMX_EQ_STD; };
template <typename TRAITS>
struct D3Q27_EQ_INV_CUM { static constexpr int lbmx_eq = LBMX_EQ_INV_CUM; };
template <typename TRAITS>
struct D2Q9_EQ { static constexpr int lbmx_eq = LBMX_EQ_STD; };

// collision operators (tags; `id` strings as in the reference, e.g. col_cum.h:11)
#define LBMX_COLL_TAG(NAME, DEFAULT_EQ, COLL, ID)                 \
	template <typename TRAITS, typename LBM_EQ = DEFAULT_EQ<TRAITS>> \
	struct NAME                                                   \
	{                                                             \
		using EQ = LBM_EQ;                                        \
		static constexpr const char* id = ID;                     \
		static constexpr int lbmx_coll = COLL;                    \
	};
LBMX_COLL_TAG(D3Q27_CUM, D3Q27_EQ, LBMX_COLL_CUM, "CUM")
LBMX_COLL_TAG(D3Q27_SRT, D3Q27_EQ, LBMX_COLL_SRT, "SRT")
LBMX_COLL_TAG(D3Q27_BGK, D3Q27_EQ, LBMX_COLL_BGK, "BGK")
LBMX_COLL_TAG(D3Q27_MRT, D3Q27_EQ, LBMX_COLL_MRT_LES, "MRT_LES")
LBMX_COLL_TAG(D2Q9_SRT, D2Q9_EQ, LBMX_COLL_SRT, "SRT")
LBMX_COLL_TAG(D2Q9_CLBM, D2Q9_EQ, LBMX_COLL_CLBM, "CLBM")
#undef LBMX_COLL_TAG

// streaming (tag): the pattern is the reference's preprocessor choice (defs.h:3-9)
template <typename TRAITS>
struct D3Q27_STREAMING
{
#ifdef AA_PATTERN
	static constexpr int lbmx_streaming = LBMX_STREAM_AA;
#else
	static constexpr int lbmx_streaming = LBMX_STREAM_AB;
#endif
};
template <typename TRAITS>
struct D2Q9_STREAMING : D3Q27_STREAMING<TRAITS> {};

// cell-type sets (d3q27/bc.h:17-49, d2q9/bc.h:16-57)
template <typename CONFIG>
struct D3Q27_BC_All
{
	using map_t = typename CONFIG::TRAITS::map_t;
	enum GEO : map_t { GEO_FLUID, GEO_WALL, GEO_INFLOW, GEO_INFLOW_LEFT, GEO_OUTFLOW_EQ, GEO_OUTFLOW_RIGHT, GEO_OUTFLOW_RIGHT_INTERP, GEO_PERIODIC, GEO_NOTHING,
					   GEO_SYM_TOP, GEO_SYM_BOTTOM, GEO_SYM_LEFT, GEO_SYM_RIGHT, GEO_SYM_BACK, GEO_SYM_FRONT };
	static bool isPeriodic(map_t m) { return m == GEO_PERIODIC; }
	static bool isFluid(map_t m) { return m == GEO_FLUID; }
	static bool isWall(map_t m) { return m == GEO_WALL; }
	static bool doCollision(map_t m) { return isFluid(m) || isPeriodic(m) || m == GEO_OUTFLOW_RIGHT || m == GEO_OUTFLOW_RIGHT_INTERP || m == GEO_INFLOW_LEFT; }
	static constexpr int lbmx_bc = 3;
};
template <typename CONFIG>
struct D2Q9_BC_All
{
	using map_t = typename CONFIG::TRAITS::map_t;
	enum GEO : map_t { GEO_FLUID, GEO_WALL, GEO_INFLOW, GEO_OUTFLOW_EQ, GEO_OUTFLOW_RIGHT, GEO_OUTFLOW_RIGHT_INTERP, GEO_PERIODIC, GEO_NOTHING, GEO_SYM_TOP,
					   GEO_SYM_BOTTOM, GEO_SYM_LEFT, GEO_SYM_RIGHT, GEO_FLUID_NEAR_WALL, GEO_TRANSFER_FS, GEO_TRANSFER_SF, GEO_TRANSFER_SW };
	static bool isPeriodic(map_t m) { return m == GEO_PERIODIC; }
	static bool isFluid(map_t m) { return m == GEO_FLUID || m == GEO_FLUID_NEAR_WALL; }
	static bool isWall(map_t m) { return m == GEO_WALL; }
	static bool isSolid(map_t) { return false; }
	static bool doCollision(map_t m) { return isFluid(m) || isPeriodic(m) || m == GEO_OUTFLOW_RIGHT || m == GEO_OUTFLOW_RIGHT_INTERP; }
	static constexpr int lbmx_bc = 2;
};

// macroscopic output policies (d3q27/macro.h:50-188, d2q9/macro.h:49-140)
template <typename TRAITS>
struct D3Q27_MACRO_Default
{
	enum { e_rho, e_vx, e_vy, e_vz, N };
	static const bool use_syncMacro = false;
	static constexpr int overlap_width = 1;
	static constexpr int lbmx_macro = LBMX_MACRO_DEFAULT;
};
template <typename TRAITS>
struct D3Q27_MACRO_Mean
{
	enum { e_rho, e_vx, e_vy, e_vz, e_vm_x, e_vm_y, e_vm_z, e_vm2_xx, e_vm2_yy, e_vm2_zz, e_vm2_xy, e_vm2_xz, e_vm2_yz, N };
	static const bool use_syncMacro = false;
	static constexpr int overlap_width = 1;
	static constexpr int lbmx_macro = LBMX_MACRO_MEAN;
};
template <typename TRAITS>
struct D3Q27_MACRO_Void
{
	static const int N = 0;
	static const bool use_syncMacro = false;
	static constexpr int overlap_width = 1;
	static constexpr int lbmx_macro = LBMX_MACRO_VOID;
};
template <typename TRAITS>
struct D2Q9_MACRO_Default
{
	enum { e_rho, e_vx, e_vy, N };
	static const bool use_syncMacro = false;
	static constexpr int overlap_width = 1;
	static constexpr int lbmx_macro = LBMX_MACRO_DEFAULT;
};
template <typename TRAITS>
struct D2Q9_MACRO_Mean
{
	enum { e_rho, e_vx, e_vy, e_vm_x, e_vm_y, e_vm2_xx, e_vm2_yy, e_vm2_xy, N };
	static const bool use_syncMacro = false;
	static constexpr int overlap_width = 1;
	static constexpr int lbmx_macro = LBMX_MACRO_MEAN;
};
template <typename TRAITS>
struct D2Q9_MACRO_Void
{
	static const int N = 0;
	static const bool use_syncMacro = false;
	static constexpr int overlap_width = 1;
	static constexpr int lbmx_macro = LBMX_MACRO_VOID;
};

// block.data: the per-step scalars of lbm_data.h:7-131 (the array pointers are filled from lbmx_get_device_ptrs)
template <typename TRAITS>
struct LBM_Data
{
	using idx = typename TRAITS::idx;
	using dreal = typename TRAITS::dreal;
	using map_t = typename TRAITS::map_t;
	bool even_iter = true;
	idx XYZ = 0;
	idx sizes[3] = {0, 0, 0};
	dreal lbmViscosity = 0;
	int stat_counter = 0;
	dreal* dfs[DFMAX] = {};
	dreal* dmacro = nullptr;
	map_t* dmap = nullptr;
	dreal* bouzidi_coeff_ptr = nullptr;	 // informational on the host side: the engine owns the device copy (lbmx_bouzidi_upload)
	idx X() const { return sizes[0]; }
	idx Y() const { return sizes[1]; }
	idx Z() const { return sizes[2]; }
};
template <typename TRAITS>
struct NSE_Data : LBM_Data<TRAITS>
{
	using dreal = typename LBM_Data<TRAITS>::dreal;
	dreal fx = 0;
	dreal fy = 0;
	dreal fz = 0;
};
template <typename TRAITS>
struct NSE_Data_ConstInflow : NSE_Data<TRAITS>
{
	using dreal = typename TRAITS::dreal;
	dreal inflow_vx = 0;
	dreal inflow_vy = 0;
	dreal inflow_vz = 0;
};
template <typename TRAITS>
struct NSE_Data_NoInflow : NSE_Data<TRAITS>
{};

template <typename _TRAITS, template <typename> class _KERNEL_STRUCT, typename _DATA, typename _COLL, typename _EQ, typename _STREAMING,
		  template <typename> class _BC, typename _MACRO>
struct LBM_CONFIG
{
	using TRAITS = _TRAITS;
	template <typename REAL>
	using KernelStruct = _KERNEL_STRUCT<REAL>;
	using DATA = _DATA;
	using COLL = _COLL;
	using EQ = _EQ;
	using STREAMING = _STREAMING;
	using BC = _BC<LBM_CONFIG>;
	using MACRO = _MACRO;
	static constexpr int Q = KernelStruct<typename TRAITS::dreal>::Q;

	// trait classes -> descriptor enums; a class without the tag is a custom device-side trait, which cannot cross the C ABI
	static constexpr int lbmx_lattice = KernelStruct<typename TRAITS::dreal>::lbmx_lattice;
	static constexpr int lbmx_coll = COLL::lbmx_coll;
	static constexpr int lbmx_eq = EQ::lbmx_eq;
	static constexpr int lbmx_streaming = STREAMING::lbmx_streaming;
	static constexpr int lbmx_macro = MACRO::lbmx_macro;
	static constexpr int lbmx_precision = TRAITS::lbmx_precision;
	static constexpr int lbmx_inflow =
		lbmx_host::has_vx_profile<DATA>::value ? LBMX_INFLOW_PROFILE_YZ : (lbmx_host::has_inflow_vx<DATA>::value ? LBMX_INFLOW_CONST : LBMX_INFLOW_NONE);
	static_assert(std::is_base_of<NSE_Data<TRAITS>, DATA>::value, "DATA must derive from NSE_Data<TRAITS> (lbm_data.h:87-96)");
};

// ---------------------------------------------------------------------------------------------------------------------------
// lattice.h: physical <-> lattice units (include/lbm3d/lattice.h:15-156)
// ---------------------------------------------------------------------------------------------------------------------------
template <int D_ = 3, typename real = float, typename idx = int>
struct Lattice
{
	using RealType = real;
	using GlobalIndexType = idx;
	using PointType = TNL::Containers::StaticVector<D_, real>;
	using CoordinatesType = TNL::Containers::StaticVector<D_, idx>;
	static constexpr int D = D_;
	CoordinatesType global = 0;
	PointType physOrigin = 0;
	RealType physDl = 0;
	real physDt = 0;
	real physViscosity = 0;
	real lbmViscosity() const { return phys2lbmViscosity(physViscosity); }
	real phys2lbmViscosity(real v) const { return physDt / physDl / physDl * v; }
	real lbm2physViscosity(real v) const { return physDl * physDl / physDt * v; }
	PointType lbm2physPoint(idx x, idx y, idx z) const { return PointType(lbm2physX(x), lbm2physY(y), lbm2physZ(z)); }
	real lbm2physX(idx x) const { return physOrigin.x() + (x - (real) 0.5) * physDl; }
	real lbm2physY(idx y) const { return physOrigin.y() + (y - (real) 0.5) * physDl; }
	real lbm2physZ(idx z) const { return physOrigin.z() + (z - (real) 0.5) * physDl; }
	real phys2lbmX(real x) const { return (x - physOrigin.x()) / physDl + (real) 0.5; }
	real phys2lbmY(real y) const { return (y - physOrigin.y()) / physDl + (real) 0.5; }
	real phys2lbmZ(real z) const { return (z - physOrigin.z()) / physDl + (real) 0.5; }
	PointType phys2lbmPoint(PointType p) const { return PointType(phys2lbmX(p.x()), phys2lbmY(p.y()), phys2lbmZ(p.z())); }
	real lbm2physVelocity(real v) const { return v / physDt * physDl; }
	real phys2lbmVelocity(real v) const { return v * physDt / physDl; }
	real lbm2physForce(real f) const { return f * physDl / physDt / physDt; }
	real phys2lbmForce(real f) const { return f / physDl * physDt * physDt; }
	static constexpr int getMeshDimension() { return D; }
	static constexpr int getDimension() { return D; }
	const CoordinatesType& size() const { return global; }
	const CoordinatesType& getDimensions() const { return global; }
	const PointType& getOrigin() const { return physOrigin; }
};

// ---------------------------------------------------------------------------------------------------------------------------
// lbm_block.h: one sub-domain = one engine handle + the host mirrors a solver reads and paints
// ---------------------------------------------------------------------------------------------------------------------------
template <typename CONFIG>
struct LBM_BLOCK
{
	using TRAITS = typename CONFIG::TRAITS;
	using MACRO = typename CONFIG::MACRO;
	using idx = typename TRAITS::idx;
	using dreal = typename TRAITS::dreal;
	using real = typename TRAITS::real;
	using map_t = typename TRAITS::map_t;
	using idx3d = typename TRAITS::idx3d;

	// host arrays addressed with GLOBAL lattice indices, storage order of the reference (x, z, y), y fastest
	struct HostMap
	{
		std::vector<map_t> v;
		idx3d off, loc;
		map_t& operator()(idx x, idx y, idx z) { return v[(size_t) (((x - off.x()) * loc.z() + (z - off.z())) * loc.y() + (y - off.y()))]; }
		const map_t& operator()(idx x, idx y, idx z) const { return v[(size_t) (((x - off.x()) * loc.z() + (z - off.z())) * loc.y() + (y - off.y()))]; }
		void setValue(map_t value) { std::fill(v.begin(), v.end(), value); }
	};
	struct HostField
	{
		std::vector<dreal> v;
		idx3d off, loc;
		size_t n = 0;
		dreal& operator()(int k, idx x, idx y, idx z) { return v[k * n + (size_t) (((x - off.x()) * loc.z() + (z - off.z())) * loc.y() + (y - off.y()))]; }
		const dreal& operator()(int k, idx x, idx y, idx z) const
		{
			return v[k * n + (size_t) (((x - off.x()) * loc.z() + (z - off.z())) * loc.y() + (y - off.y()))];
		}
		dreal* getData() { return v.empty() ? nullptr : v.data(); }
		const dreal* getData() const { return v.empty() ? nullptr : v.data(); }
		void setValue(dreal value) { std::fill(v.begin(), v.end(), value); }
	};
	// device-side Bouzidi array of the reference (`block.dBouzidi = block.hBouzidi;`, sim_2D/sim2d_2.cu:328): a proxy whose
	// assignment uploads through the C ABI
	struct DeviceBouzidi
	{
		LBM_BLOCK* owner = nullptr;
		bool uploaded = false;
		DeviceBouzidi& operator=(const HostField& h)
		{
			if (owner && owner->engine && ! h.v.empty()) {
				lbmx_host::check(lbmx_bouzidi_upload(owner->engine, h.v.data()), "lbmx_bouzidi_upload");
				uploaded = true;
			}
			return *this;
		}
		dreal* getData() const { return uploaded ? reinterpret_cast<dreal*>(owner) : nullptr; }	// non-null token: the pointer itself is not usable on the host
	};

	typename CONFIG::DATA data;
	idx3d global, local, offset;
	int rank = 0, nproc = 1, id = 0;
	HostMap hmap;
	HostField hmacro;
	HostField hfs[DFMAX];  // host copies of the distributions, filled by copyDFsToHost only
	HostField hBouzidi;	   // D2Q9 near-wall coefficients [8][local], allocated by allocateBouzidiCoeffArrays (lbm_block.hpp:740-770)
	DeviceBouzidi dBouzidi;
	lbmx_engine* engine = nullptr;

	LBM_BLOCK(idx3d global_, idx3d local_, idx3d offset_) : global(global_), local(local_), offset(offset_) { dBouzidi.owner = this; }
	LBM_BLOCK(const LBM_BLOCK&) = delete;
	LBM_BLOCK(LBM_BLOCK&& o) noexcept
	: data(o.data), global(o.global), local(o.local), offset(o.offset), rank(o.rank), nproc(o.nproc), id(o.id), hmap(std::move(o.hmap)),
	  hmacro(std::move(o.hmacro)), hBouzidi(std::move(o.hBouzidi)), engine(o.engine)
	{
		o.engine = nullptr;
		dBouzidi.owner = this;
		dBouzidi.uploaded = o.dBouzidi.uploaded;
	}
	~LBM_BLOCK()
	{
		if (engine)
			lbmx_destroy(engine);
	}

	bool isLocalIndex(idx x, idx y, idx z) const { return isLocalX(x) && isLocalY(y) && isLocalZ(z); }
	bool isLocalX(idx x) const { return x >= offset.x() && x < offset.x() + local.x(); }
	bool isLocalY(idx y) const { return y >= offset.y() && y < offset.y() + local.y(); }
	bool isLocalZ(idx z) const { return z >= offset.z() && z < offset.z() + local.z(); }

	// map painting in global indices, last writer wins (lbm_block.hpp:304-342)
	void setMap(idx x, idx y, idx z, map_t value)
	{
		if (isLocalIndex(x, y, z))
			hmap(x, y, z) = value;
	}
	void setBoundaryX(idx x, map_t value)
	{
		if (isLocalX(x))
			for (idx y = offset.y(); y < offset.y() + local.y(); y++)
				for (idx z = offset.z(); z < offset.z() + local.z(); z++)
					hmap(x, y, z) = value;
	}
	void setBoundaryY(idx y, map_t value)
	{
		if (isLocalY(y))
			for (idx x = offset.x(); x < offset.x() + local.x(); x++)
				for (idx z = offset.z(); z < offset.z() + local.z(); z++)
					hmap(x, y, z) = value;
	}
	void setBoundaryZ(idx z, map_t value)
	{
		if (isLocalZ(z))
			for (idx x = offset.x(); x < offset.x() + local.x(); x++)
				for (idx y = offset.y(); y < offset.y() + local.y(); y++)
					hmap(x, y, z) = value;
	}
	void resetMap(map_t geo_type) { hmap.setValue(geo_type); }

	void allocateHostData()
	{
		const size_t n = (size_t) local.x() * local.y() * local.z();
		hmap.v.assign(n, 0);
		hmap.off = offset;
		hmap.loc = local;
		hmacro.v.assign(n * (MACRO::N > 0 ? MACRO::N : 1), 0);
		hmacro.off = offset;
		hmacro.loc = local;
		hmacro.n = n;
		data.sizes[0] = local.x();
		data.sizes[1] = local.y();
		data.sizes[2] = local.z();
		data.XYZ = (idx) n;
	}
	// LBM_BLOCK::allocateDeviceData (lbm_block.hpp:525-595) -> lbmx_create
	void allocateDeviceData(bool periodic_lattice)
	{
		if (engine)
			return;
		lbmx_desc d{};
		d.lattice = CONFIG::lbmx_lattice;
		d.coll = CONFIG::lbmx_coll;
		d.eq = CONFIG::lbmx_eq;
		d.streaming = CONFIG::lbmx_streaming;
		d.macro = CONFIG::lbmx_macro;
		d.inflow = CONFIG::lbmx_inflow;
		d.precision = CONFIG::lbmx_precision;
		d.macro_policy = LBMX_MACRO_EVERY_STEP;	 // the drop-in keeps the reference's observable behaviour; solvers may relax it
		d.X = global.x();
		d.Y = global.y();
		d.Z = global.z();
		d.rank = rank;
		d.nranks = nproc;
		d.device = -1;
		d.ghost_x = nproc > 1;
		d.periodic_x = periodic_lattice;
		lbmx_host::check(lbmx_create(&d, &engine), "lbmx_create");
		refreshPointers();
	}
	void refreshPointers()
	{
		lbmx_ptrs p{};
		lbmx_host::check(lbmx_get_device_ptrs(engine, &p), "lbmx_get_device_ptrs");
		data.dfs[df_cur] = (dreal*) p.dfs[0];
#if defined(AB_PATTERN)
		data.dfs[df_out] = (dreal*) p.dfs[1];
#endif
		data.dmacro = (dreal*) p.dmacro;
		data.dmap = (map_t*) p.dmap;
		data.even_iter = p.even_iter != 0;
	}
	void allocateBouzidiCoeffArrays()
	{
		const size_t n = (size_t) local.x() * local.y() * local.z();
		hBouzidi.v.assign(8 * n, (dreal) -1);
		hBouzidi.off = offset;
		hBouzidi.loc = local;
		hBouzidi.n = n;
	}
	// also moves the Bouzidi coefficients when they are allocated, like the reference (lbm_block.hpp:355-364)
	void copyMapToDevice()
	{
		lbmx_host::check(lbmx_map_upload(engine, hmap.v.data(), 0), "lbmx_map_upload");
		if (hBouzidi.getData() != nullptr)
			dBouzidi = hBouzidi;
	}
	void copyMapToHost() { lbmx_host::check(lbmx_map_download(engine, hmap.v.data(), 0), "lbmx_map_download"); }
	void copyMacroToHost()
	{
		if (MACRO::N > 0)
			lbmx_host::check(lbmx_macro_download(engine, hmacro.v.data(), 0), "lbmx_macro_download");
	}
	void copyMacroToDevice()
	{
		if (MACRO::N > 0)
			lbmx_host::check(lbmx_macro_upload(engine, hmacro.v.data(), 0), "lbmx_macro_upload");
	}
	void copyDFsToHost(uint8_t dftype)
	{
		HostField& h = hfs[dftype];
		const size_t n = (size_t) local.x() * local.y() * local.z();
		h.v.resize(n * CONFIG::Q);
		h.off = offset;
		h.loc = local;
		h.n = n;
		lbmx_host::check(lbmx_df_download(engine, dftype == df_cur ? 0 : 1, h.v.data(), 0), "lbmx_df_download");
	}
	void copyDFsToDevice(uint8_t dftype) { lbmx_host::check(lbmx_df_upload(engine, dftype == df_cur ? 0 : 1, hfs[dftype].v.data(), 0), "lbmx_df_upload"); }
	void setEquilibrium(real rho, real vx, real vy, real vz) { lbmx_host::check(lbmx_df_set_equilibrium(engine, rho, vx, vy, vz), "lbmx_df_set_equilibrium"); }
	void computeInitialMacro() { lbmx_host::check(lbmx_macro_init(engine), "lbmx_macro_init"); }

	// block.data -> lbmx_params (what passing the POD by value to the kernel did in the reference, state.hpp:1039)
	void pushParams()
	{
		lbmx_params p{};
		p.lbmViscosity = data.lbmViscosity;
		p.fx = data.fx;
		p.fy = data.fy;
		p.fz = data.fz;
		if constexpr (lbmx_host::has_inflow_vx<typename CONFIG::DATA>::value)
			p.inflow_vx = data.inflow_vx;
		if constexpr (lbmx_host::has_inflow_vy<typename CONFIG::DATA>::value)
			p.inflow_vy = data.inflow_vy;
		if constexpr (lbmx_host::has_inflow_vz<typename CONFIG::DATA>::value)
			p.inflow_vz = data.inflow_vz;
		p.stat_counter = data.stat_counter;
		lbmx_host::check(lbmx_set_params(engine, &p), "lbmx_set_params");
	}
};

// ---------------------------------------------------------------------------------------------------------------------------
// lbm.h: the blocks of this rank, fan-out of the block methods, iteration counter (include/lbm3d/lbm.h, lbm.hpp)
// ---------------------------------------------------------------------------------------------------------------------------
template <typename CONFIG>
struct LBM
{
	using MACRO = typename CONFIG::MACRO;
	using TRAITS = typename CONFIG::TRAITS;
	using BLOCK = LBM_BLOCK<CONFIG>;
	using idx = typename TRAITS::idx;
	using dreal = typename TRAITS::dreal;
	using real = typename TRAITS::real;
	using map_t = typename TRAITS::map_t;
	using point_t = typename TRAITS::point_t;
	using idx3d = typename TRAITS::idx3d;
	using lat_t = Lattice<3, real, idx>;

	TNL::MPI::Comm communicator;
	int rank = 0;
	int nproc = 1;
	lat_t lat;
	std::vector<BLOCK> blocks;
	int total_blocks = 0;
	bool periodic_lattice = false;

	real physCharLength;
	real physFinalTime = 1e10;
	real physStartTime = 0;
	int iterations = 0;
	int startIterations = 0;
	bool terminate = false;

	LBM() = delete;
	LBM(const LBM&) = delete;
	LBM(LBM&&) = default;
	LBM(const TNL::MPI::Comm& comm, lat_t lat_, bool periodic_lattice_ = false) : communicator(comm), lat(lat_), periodic_lattice(periodic_lattice_)
	{
		rank = TNL::MPI::GetRank(comm);
		nproc = TNL::MPI::GetSize(comm);
		// one x-slab per rank (lbmx_decompose_x replaces decomposeLattice_D1Q3 / decomposeLattice_D3Q27)
		int64_t x0 = 0, xl = 0;
		lbmx_host::check(lbmx_decompose_x(lat.global.x(), nproc, rank, &x0, &xl), "lbmx_decompose_x");
		blocks.emplace_back(lat.global, idx3d((idx) xl, lat.global.y(), lat.global.z()), idx3d((idx) x0, 0, 0));
		blocks.back().rank = rank;
		blocks.back().nproc = nproc;
		total_blocks = nproc;
		physCharLength = lat.physDl * (real) lat.global.y();
	}

	real Re(real physvel) { return std::fabs(physvel) * physCharLength / lat.physViscosity; }
	real physTime() { return lat.physDt * (real) iterations; }

#define LBMX_FANOUT(NAME)        \
	void NAME()                  \
	{                            \
		for (auto& b : blocks)   \
			b.NAME();            \
	}
	LBMX_FANOUT(copyMapToHost)
	LBMX_FANOUT(copyMapToDevice)
	LBMX_FANOUT(copyMacroToHost)
	LBMX_FANOUT(copyMacroToDevice)
	LBMX_FANOUT(allocateHostData)
	LBMX_FANOUT(computeInitialMacro)
	LBMX_FANOUT(allocateBouzidiCoeffArrays)
#undef LBMX_FANOUT
	void copyDFsToHost(uint8_t t)
	{
		for (auto& b : blocks)
			b.copyDFsToHost(t);
	}
	void copyDFsToDevice(uint8_t t)
	{
		for (auto& b : blocks)
			b.copyDFsToDevice(t);
	}
	void copyDFsToHost()
	{
		for (uint8_t t = 0; t < DFMAX; t++)
			copyDFsToHost(t);
	}
	void copyDFsToDevice()
	{
		for (uint8_t t = 0; t < DFMAX; t++)
			copyDFsToDevice(t);
	}
	void allocateDeviceData()
	{
		for (auto& b : blocks)
			b.allocateDeviceData(periodic_lattice);
	}
	bool isAnyLocalIndex(idx x, idx y, idx z)
	{
		for (auto& b : blocks)
			if (b.isLocalIndex(x, y, z))
				return true;
		return false;
	}
	void setMap(idx x, idx y, idx z, map_t v)
	{
		for (auto& b : blocks)
			b.setMap(x, y, z, v);
	}
	void setBoundaryX(idx x, map_t v)
	{
		for (auto& b : blocks)
			b.setBoundaryX(x, v);
	}
	void setBoundaryY(idx y, map_t v)
	{
		for (auto& b : blocks)
			b.setBoundaryY(y, v);
	}
	void setBoundaryZ(idx z, map_t v)
	{
		for (auto& b : blocks)
			b.setBoundaryZ(z, v);
	}
	void resetMap(map_t v)
	{
		for (auto& b : blocks)
			b.resetMap(v);
	}
	void setEquilibrium(real rho, real vx, real vy, real vz)
	{
		for (auto& b : blocks)
			b.setEquilibrium(rho, vx, vy, vz);
	}
	// LBM::updateKernelData (lbm.hpp:314-330): parity and A-B rotation live in the engine; mirror them into block.data
	void updateKernelData()
	{
		for (auto& b : blocks) {
			lbmx_host::check(lbmx_set_iterations(b.engine, iterations), "lbmx_set_iterations");
			b.refreshPointers();
		}
	}
	template <typename F>
	void forLocalLatticeSites(F f)
	{
		for (auto& b : blocks)
			for (idx x = b.offset.x(); x < b.offset.x() + b.local.x(); x++)
				for (idx z = b.offset.z(); z < b.offset.z() + b.local.z(); z++)
					for (idx y = b.offset.y(); y < b.offset.y() + b.local.y(); y++)
						f(b, x, y, z);
	}
	template <typename F>
	void forAllLatticeSites(F f)
	{
		forLocalLatticeSites(f);
	}
};

// ---------------------------------------------------------------------------------------------------------------------------
// state.h: the driver object a solver derives from (include/lbm3d/state.h:89-330, state.hpp)
// ---------------------------------------------------------------------------------------------------------------------------
enum : std::uint8_t { STAT_RESET, STAT2_RESET, PRINT, VTK1D, VTK2D, VTK3D, PROBE1, PROBE2, PROBE3, SAVESTATE, VTK3DCUT, MAX_COUNTER };

template <typename REAL>
struct counter
{
	int count = 0;
	REAL period = -1.0;
	bool action(REAL time) { return period > 0 && time >= count * period; }
};

template <typename NSE>
struct State
{
	using TRAITS = typename NSE::TRAITS;
	using MACRO = typename NSE::MACRO;
	using BLOCK_NSE = LBM_BLOCK<NSE>;
	using map_t = typename TRAITS::map_t;
	using idx = typename TRAITS::idx;
	using dreal = typename TRAITS::dreal;
	using real = typename TRAITS::real;
	using point_t = typename TRAITS::point_t;
	using idx3d = typename TRAITS::idx3d;
	using lat_t = typename LBM<NSE>::lat_t;
	using T_COUNTER = counter<real>;

	std::string id;
	LBM<NSE> nse;
	T_COUNTER cnt[MAX_COUNTER];
	int n_cuts = 0;	 // cuts registered through add*cut*: the writers behind them are out of scope, the calls are accepted

	virtual void probe1() {}
	virtual void probe2() {}
	virtual void probe3() {}
	virtual void statReset() {}
	virtual void stat2Reset() {}

	template <typename real1, typename real2>
	bool vtk_helper(const char* iid, real1 ivalue, int idofs, char* id_, real2& value, int& dofs)
	{
		std::snprintf(id_, 500, "%s", iid);
		dofs = idofs;
		value = ivalue;
		return true;
	}
	virtual void writeVTKs_2D() {}
	virtual void writeVTKs_3D() {}
	virtual void writeVTKs_3Dcut() {}
	virtual void writeVTKs_1D() {}
	template <typename... ARGS>
	void add2Dcut_X(idx, const char*, ARGS...) { n_cuts++; }
	template <typename... ARGS>
	void add2Dcut_Y(idx, const char*, ARGS...) { n_cuts++; }
	template <typename... ARGS>
	void add2Dcut_Z(idx, const char*, ARGS...) { n_cuts++; }
	template <typename... ARGS>
	void add3Dcut(idx, idx, idx, idx, idx, idx, idx, const char*, ARGS...) { n_cuts++; }
	template <typename... ARGS>
	void add1Dcut(point_t, point_t, const char*, ARGS...) { n_cuts++; }
	template <typename... ARGS>
	void add1Dcut_X(real, real, const char*, ARGS...) { n_cuts++; }
	template <typename... ARGS>
	void add1Dcut_Y(real, real, const char*, ARGS...) { n_cuts++; }
	template <typename... ARGS>
	void add1Dcut_Z(real, real, const char*, ARGS...) { n_cuts++; }

	virtual bool outputData(const BLOCK_NSE&, int, int, char*, idx, idx, idx, real&, int&) { return false; }

	virtual bool estimateMemoryDemands() { return true; }
	// State::reset (state.hpp:880-896)
	virtual void reset()
	{
		resetDFs();
		nse.resetMap(NSE::BC::GEO_FLUID);
		setupBoundaries();
		nse.copyMapToDevice();
		nse.computeInitialMacro();
		nse.copyMacroToHost();
	}
	virtual void resetDFs() { nse.setEquilibrium(1, 0, 0, 0); }
	virtual void setupBoundaries() {}
	// State::SimInit (state.hpp:907-977) without the checkpoint branch
	virtual void SimInit()
	{
		timer_SimInit.start();
		nse.allocateDeviceData();
		nse.iterations = 0;
		for (auto& c : cnt)
			c.count = 0;
		reset();
		lbmx_host::log_info("lbmx: %s lattice %ld x %ld x %ld, %s, %s, %s", NSE::COLL::id, (long) nse.lat.global.x(), (long) nse.lat.global.y(),
							(long) nse.lat.global.z(), NSE::lbmx_precision == LBMX_F64 ? "fp64" : "fp32", NSE::lbmx_streaming == LBMX_STREAM_AA ? "A-A" : "A-B",
							n_cuts ? "output cuts registered but writers are out of scope (DESIGN.md)" : "no output cuts");
		timer_SimInit.stop();
	}
	// State::updateKernelData (state.hpp:1314-1321)
	virtual void updateKernelData()
	{
		nse.updateKernelData();
		for (auto& b : nse.blocks)
			b.data.lbmViscosity = (dreal) nse.lat.lbmViscosity();
	}
	virtual void updateKernelVelocities() {}
	// State::SimUpdate (state.hpp:980-1145): one time step
	virtual void SimUpdate()
	{
		timer_SimUpdate.start();
		if (nse.lat.lbmViscosity() == 0) {	// state.hpp:985-990
			lbmx_host::log_info("error: LBM viscosity is 0");
			nse.terminate = true;
			timer_SimUpdate.stop();
			return;
		}
		computeBeforeLBMKernel();
		for (auto& b : nse.blocks) {
			b.pushParams();
			lbmx_host::check(lbmx_step(b.engine, 1), "lbmx_step");
		}
		computeAfterLBMKernel();
		nse.iterations++;
		bool doCopy = false;
		for (int c = 0; c < MAX_COUNTER; c++)
			if (c != PRINT && c != SAVESTATE)
				doCopy |= cnt[c].action(nse.physTime());
		if (doCopy)
			nse.copyMacroToHost();
		timer_SimUpdate.stop();
	}
	// State::AfterSimUpdate (state.hpp:1148-1278): cadence-driven hooks, NaN scan, GLUPS line
	virtual void AfterSimUpdate()
	{
		const real t = nse.physTime();
		if (cnt[PRINT].action(t)) {
			int32_t nan = 0;
			lbmx_host::check(lbmx_has_nan(nse.blocks.front().engine, &nan), "lbmx_has_nan");
			if (nan) {
				lbmx_host::log_info("nan detected");
				nse.terminate = true;
			}
			for (auto& b : nse.blocks)
				lbmx_host::check(lbmx_sync(b.engine), "lbmx_sync");
			const double now = timer_total.getRealTime();
			const double glups = (nse.iterations - glups_prev_iterations) / (now - glups_prev_time + 1e-30) * (double) nse.lat.global.x() * (double) nse.lat.global.y()
							   * (double) nse.lat.global.z() * 1e-9;
			lbmx_host::log_info("GLUPS=%.3f iter=%d t=%1.3fs dt=%1.2e lbmVisc=%1.2e WT=%.0fs", glups, nse.iterations, (double) t, (double) nse.lat.physDt,
								(double) nse.lat.lbmViscosity(), now);
			glups_prev_iterations = nse.iterations;
			glups_prev_time = now;
			cnt[PRINT].count++;
		}
		if (cnt[STAT_RESET].action(t)) {
			statReset();
			cnt[STAT_RESET].count++;
		}
		if (cnt[STAT2_RESET].action(t)) {
			stat2Reset();
			cnt[STAT2_RESET].count++;
		}
		if (cnt[PROBE1].action(t)) {
			probe1();
			cnt[PROBE1].count++;
		}
		if (cnt[PROBE2].action(t)) {
			probe2();
			cnt[PROBE2].count++;
		}
		if (cnt[PROBE3].action(t)) {
			probe3();
			cnt[PROBE3].count++;
		}
		if (cnt[VTK1D].action(t)) {
			writeVTKs_1D();
			cnt[VTK1D].count++;
		}
		if (cnt[VTK2D].action(t)) {
			writeVTKs_2D();
			cnt[VTK2D].count++;
		}
		if (cnt[VTK3D].action(t)) {
			writeVTKs_3D();
			cnt[VTK3D].count++;
		}
		if (cnt[VTK3DCUT].action(t)) {
			writeVTKs_3Dcut();
			cnt[VTK3DCUT].count++;
		}
	}
	virtual void AfterSimFinished()
	{
		for (auto& b : nse.blocks)
			lbmx_host::check(lbmx_sync(b.engine), "lbmx_sync");
		const double total = timer_total.getRealTime(), upd = timer_SimUpdate.getRealTime();
		const double cells = (double) nse.lat.global.x() * (double) nse.lat.global.y() * (double) nse.lat.global.z();
		lbmx_host::log_info("total walltime: %.1f s, SimInit time: %.1f s, SimUpdate time: %.1f s", total, timer_SimInit.getRealTime(), upd);
		lbmx_host::log_info("final GLUPS: average (based on total time) %.3f", cells * nse.iterations / (total + 1e-30) * 1e-9);
	}
	virtual void computeBeforeLBMKernel() {}
	virtual void computeAfterLBMKernel() {}
	virtual void copyAllToDevice()
	{
		nse.copyMapToDevice();
		nse.copyMacroToDevice();
	}
	virtual void copyAllToHost()
	{
		nse.copyMapToHost();
		nse.copyMacroToHost();
	}

	bool canCompute() { return true; }	// result-directory locking and restart flags (state.hpp:13-66) are out of scope
	void flagCreate(const char*) {}
	void flagDelete(const char*) {}
	bool flagExists(const char*) { return false; }
	virtual void checkpointStateLocal(int) {}
	void saveState() {}
	void loadState() {}

	TNL::Timer timer_total;
	long wallTime = -1;
	bool wallTimeReached() { return wallTime > 0 && timer_total.getRealTime() >= (double) wallTime; }
	double getWallTime(bool = false) { return timer_total.getRealTime(); }
	int glups_prev_iterations = 0;
	double glups_prev_time = 0;
	TNL::Timer timer_SimInit, timer_SimUpdate, timer_AfterSimUpdate, timer_compute, timer_compute_overlaps, timer_wait_communication, timer_wait_computation;

	template <typename... ARGS>
	State(const std::string& id_, const TNL::MPI::Comm& communicator, lat_t lat, ARGS&&... args) : id(id_), nse(communicator, lat, std::forward<ARGS>(args)...)
	{
		nse.allocateHostData();
		timer_total.start();
	}
	virtual ~State() = default;
};

// ---------------------------------------------------------------------------------------------------------------------------
// core.h: the time loop (include/lbm3d/core.h:38-101)
// ---------------------------------------------------------------------------------------------------------------------------
template <typename STATE>
void execute(STATE& state)
{
	state.SimInit();
	state.AfterSimUpdate();	 // snapshot of the initial condition
	bool quit = false;
	while (! quit) {
		state.updateKernelData();
		state.updateKernelVelocities();
		state.SimUpdate();
		state.AfterSimUpdate();
		if (state.wallTimeReached()) {
			state.copyAllToHost();
			lbmx_host::log_info("maximum wall time reached");
			state.saveState();
			quit = true;
		}
		else if (state.cnt[SAVESTATE].action(state.getWallTime(true))) {
			state.copyAllToHost();
			state.saveState();
			state.cnt[SAVESTATE].count++;
		}
		if (state.nse.physTime() > state.nse.physFinalTime) {
			lbmx_host::log_info("physFinalTime reached");
			quit = true;
			state.flagCreate("finished");
			state.flagDelete("loadstate");
		}
		if (state.nse.terminate) {
			lbmx_host::log_info("terminate flag triggered");
			quit = true;
			state.flagCreate("terminated");
			state.flagDelete("loadstate");
		}
	}
	state.AfterSimFinished();
}
