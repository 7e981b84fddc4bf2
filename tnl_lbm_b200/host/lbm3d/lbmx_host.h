// lbmx_host.h -- host-side mirror of TNL-LBM's solver-facing interface on top of the C ABI (include/lbmx.h).
//
// The reference is a header-only C++ template framework: a solver composes LBM_CONFIG<TRAITS, KernelStruct, DATA, COLL, EQ,
// STREAMING, BC, MACRO> (include/lbm3d/defs.h:169-250), derives StateLocal from State<NSE> (include/lbm3d/state.h:89-330),
// paints the map through nse.setBoundaryX/Y/Z / setMap, sets block.data.* and calls execute(state) (include/lbm3d/core.h:38-101).
// This header keeps those names, argument meanings and error behaviour (exceptions), but the trait classes are *tags*: they
// carry no device code, they select a kernel family of liblbmx.so through the lbmx_desc enums.  Everything that touches the GPU
// goes through the C ABI.  ADIOS2 is not a dependency: checkpoints and the 3-D / 2-D cut writers keep the reference's variable
// names, 1-D local-storage shapes and (z,y,x) float ordering, but land in a raw-dump directory (one .bin per variable plus a
// text index, see CheckpointManager / RawWriter below).  Out of scope: VTK/BP encoders, IBM, MPI.
//
// Device-side user code cannot cross a C ABI.  The finite set the reference's own solvers use is recognised structurally:
//   DATA with member `vx_profile`           -> LBMX_INFLOW_PROFILE_YZ   (NSE_Data_XProfileInflow, sim_NSE/sim_2.cu:16-33)
//   DATA with members `u_max_lbm, y0, inv_den` -> LBMX_INFLOW_PARABOLIC_Y (NSE2D_Data_ParabolicInflow, sim_2D/sim2d_3.cu:36-55)
//   DATA with member `inflow_vx`            -> LBMX_INFLOW_CONST        (NSE_Data_ConstInflow lbm_data.h:98-115, NSE2D_Data_ConstInflow)
//   otherwise                               -> LBMX_INFLOW_NONE         (NSE_Data_NoInflow lbm_data.h:117-131)
// Custom MACRO / COLL / BC classes are rejected at compile time (they lack the lbmx_* tag constants).
#pragma once

#include <chrono>
#include <cmath>
#include <algorithm>
#include <cerrno>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <arpa/inet.h>
#include <dirent.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <math.h>    // the C++ wrappers of <math.h> / <stdlib.h> put the float / double overloads of abs, sqrt, fabs ... into the global
#include <stdlib.h>  // namespace, as the CUDA headers do for the reference's solvers (sim_NSE/sim_2.cu:243 calls abs() on a real)
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include "lbmx.h"

// The reference's state.h pulls fmt and spdlog in for its solvers (state.h:10,13); unmodified solvers rely on that.
#include <iostream>
#if __has_include(<fmt/core.h>)
	#include <fmt/core.h>
	#define LBMX_HAVE_FMT 1
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
#ifndef PI
	#define PI 3.1415926535897932384	 // lbm_common/ciselnik.h:4 (solvers use it, e.g. sim_NSE/sim_2.cu:76)
#endif
#ifndef __cuda_callable__
	#define __cuda_callable__
#endif

// ---------------------------------------------------------------------------------------------------------------------------
// Process group without MPI (the reference runs one MPI rank per GPU, lbm.h:38-45; MPI is not a dependency here).
// Rank, size and local rank come from the launcher's environment -- torchrun (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR /
// MASTER_PORT), mpirun or srun if one is used, or LBMX_RANK / LBMX_WORLD_SIZE / LBMX_LOCAL_RANK set by hand.  Host-side
// collectives (the few reductions and broadcasts of the control plane, and the 128-byte NCCL id) travel over TCP in a star on
// rank 0; the data path is NCCL inside the engine.  Example, 4 GPUs of one node:
//     python -m torch.distributed.run --no-python --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29500 ./sim_1 4
// ---------------------------------------------------------------------------------------------------------------------------
namespace lbmx_host {
struct World
{
	int rank = 0, size = 1, local_rank = 0;
	bool connected = false;
	std::vector<int> peers;	 // rank 0: socket of every other rank (index = rank); other ranks: peers[0] = socket to rank 0

	static World& get()
	{
		static World w = make();
		return w;
	}
	static int env_int(std::initializer_list<const char*> names, int fallback)
	{
		for (const char* n : names)
			if (const char* v = std::getenv(n))
				if (*v)
					return std::atoi(v);
		return fallback;
	}
	static World make()
	{
		World w;
		w.size = env_int({"LBMX_WORLD_SIZE", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"}, 1);
		w.rank = env_int({"LBMX_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"}, 0);
		w.local_rank = env_int({"LBMX_LOCAL_RANK", "LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"}, w.rank);
		if (w.size < 1 || w.rank < 0 || w.rank >= w.size)
			throw std::runtime_error("lbmx: inconsistent rank / world size in the environment");
		return w;
	}
	static void send_all(int fd, const void* buf, size_t n)
	{
		const char* p = (const char*) buf;
		while (n > 0) {
			const ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
			if (k <= 0)
				throw std::runtime_error("lbmx: lost the connection to a peer rank (send)");
			p += k;
			n -= (size_t) k;
		}
	}
	static void recv_all(int fd, void* buf, size_t n)
	{
		char* p = (char*) buf;
		while (n > 0) {
			const ssize_t k = ::recv(fd, p, n, 0);
			if (k <= 0)
				throw std::runtime_error("lbmx: lost the connection to a peer rank (recv)");
			p += k;
			n -= (size_t) k;
		}
	}
	// rendezvous: rank 0 listens on LBMX_MASTER_PORT (default: MASTER_PORT + 1, torchrun keeps MASTER_PORT for its own store)
	void connect()
	{
		if (connected || size == 1)
			return;
		const char* addr = std::getenv("LBMX_MASTER_ADDR");
		if (! addr)
			addr = std::getenv("MASTER_ADDR");
		if (! addr)
			addr = "127.0.0.1";
		const int port = env_int({"LBMX_MASTER_PORT"}, env_int({"MASTER_PORT"}, 29576) + 1);
		const int one = 1;
		if (rank == 0) {
			const int ls = ::socket(AF_INET, SOCK_STREAM, 0);
			::setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
			sockaddr_in sa{};
			sa.sin_family = AF_INET;
			sa.sin_addr.s_addr = htonl(INADDR_ANY);
			sa.sin_port = htons((uint16_t) port);
			if (::bind(ls, (sockaddr*) &sa, sizeof sa) != 0 || ::listen(ls, size) != 0)
				throw std::runtime_error("lbmx: rank 0 cannot listen on port " + std::to_string(port) + ": " + std::strerror(errno));
			peers.assign((size_t) size, -1);
			for (int k = 1; k < size; k++) {
				const int fd = ::accept(ls, nullptr, nullptr);
				if (fd < 0)
					throw std::runtime_error(std::string("lbmx: accept failed: ") + std::strerror(errno));
				::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
				int32_t r = -1;
				recv_all(fd, &r, sizeof r);
				if (r < 1 || r >= size || peers[(size_t) r] != -1)
					throw std::runtime_error("lbmx: unexpected rank announced itself at the rendezvous");
				peers[(size_t) r] = fd;
			}
			::close(ls);
		}
		else {
			addrinfo hints{}, *res = nullptr;
			hints.ai_family = AF_INET;
			hints.ai_socktype = SOCK_STREAM;
			if (::getaddrinfo(addr, std::to_string(port).c_str(), &hints, &res) != 0 || ! res)
				throw std::runtime_error(std::string("lbmx: cannot resolve ") + addr);
			int fd = -1;
			for (int attempt = 0; attempt < 600 && fd < 0; attempt++) {	 // up to 60 s for rank 0 to come up
				fd = ::socket(AF_INET, SOCK_STREAM, 0);
				if (::connect(fd, res->ai_addr, res->ai_addrlen) != 0) {
					::close(fd);
					fd = -1;
					::usleep(100000);
				}
			}
			::freeaddrinfo(res);
			if (fd < 0)
				throw std::runtime_error(std::string("lbmx: cannot reach rank 0 at ") + addr + ":" + std::to_string(port));
			::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
			const int32_t r = rank;
			send_all(fd, &r, sizeof r);
			peers.assign(1, fd);
		}
		connected = true;
	}
	enum Op { SUM, MAX, MIN };
	void allreduce(double* v, int n, Op op)
	{
		if (size == 1)
			return;
		connect();
		if (rank == 0) {
			std::vector<double> in((size_t) n);
			for (int r = 1; r < size; r++) {
				recv_all(peers[(size_t) r], in.data(), sizeof(double) * (size_t) n);
				for (int i = 0; i < n; i++)
					v[i] = op == SUM ? v[i] + in[(size_t) i] : (op == MAX ? std::max(v[i], in[(size_t) i]) : std::min(v[i], in[(size_t) i]));
			}
			for (int r = 1; r < size; r++)
				send_all(peers[(size_t) r], v, sizeof(double) * (size_t) n);
		}
		else {
			send_all(peers[0], v, sizeof(double) * (size_t) n);
			recv_all(peers[0], v, sizeof(double) * (size_t) n);
		}
	}
	void bcast(void* buf, size_t bytes, int root = 0)
	{
		if (size == 1)
			return;
		connect();
		if (root != 0) {  // relay through rank 0
			if (rank == root)
				send_all(peers[0], buf, bytes);
			else if (rank == 0)
				recv_all(peers[(size_t) root], buf, bytes);
		}
		if (rank == 0) {
			for (int r = 1; r < size; r++)
				if (r != root)
					send_all(peers[(size_t) r], buf, bytes);
		}
		else if (rank != root)
			recv_all(peers[0], buf, bytes);
	}
	void barrier()
	{
		double x = 0;
		allreduce(&x, 1, SUM);
	}
};
}  // namespace lbmx_host

// ---------------------------------------------------------------------------------------------------------------------------
// the sliver of TNL the solvers name themselves (StaticVector, MPI::Comm, sqr); TNL::MPI maps onto lbmx_host::World
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
	ScopedInitializer(int&, char**&) { lbmx_host::World::get().connect(); }
};
inline int GetSize(const Comm& = Comm{}) { return lbmx_host::World::get().size; }
inline int GetRank(const Comm& = Comm{}) { return lbmx_host::World::get().rank; }
// Op is one of the MPI_* constants below; values travel as double (the control plane reduces a handful of scalars)
template <typename T, typename Op>
inline T reduce(T v, Op op, const Comm& = Comm{})
{
	double x = (double) v;
	const int o = (int) op;	 // MPI_SUM, MPI_LOR, MPI_LAND, MPI_MAX, MPI_MIN
	lbmx_host::World::get().allreduce(&x, 1, o == 0 ? lbmx_host::World::SUM : ((o == 1 || o == 3) ? lbmx_host::World::MAX : lbmx_host::World::MIN));
	return (T) x;
}
template <typename T>
inline void Bcast(T* data, int count, int root, const Comm& = Comm{})
{
	lbmx_host::World::get().bcast(data, sizeof(T) * (size_t) count, root);
}
inline void Barrier(const Comm& = Comm{}) { lbmx_host::World::get().barrier(); }
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
struct has_u_max_lbm : std::false_type {};
template <typename T>
struct has_u_max_lbm<T, std::void_t<decltype(std::declval<T&>().u_max_lbm), decltype(std::declval<T&>().inv_den), decltype(std::declval<T&>().y0)>> : std::true_type {};
template <typename T, typename = void>
struct has_accumulate_gates : std::false_type {};
template <typename T>
struct has_accumulate_gates<T, std::void_t<decltype(std::declval<T&>().accumulate_means), decltype(std::declval<T&>().accumulate_flucs)>> : std::true_type {};
// MACRO -> descriptor enum: the mirror's own classes carry a tag; a solver-defined class is recognised by its channel list where the engine
// has that output built in (D2Q9_MACRO_WithMean, sim_2D/sim2d_2.cu:53-104); anything else is device code that cannot cross the C ABI
template <typename M, typename = void>
struct has_macro_tag : std::false_type {};
template <typename M>
struct has_macro_tag<M, std::void_t<decltype(M::lbmx_macro)>> : std::true_type {};
template <typename M, typename = void>
struct is_with_mean_2d : std::false_type {};
template <typename M>
struct is_with_mean_2d<M, std::void_t<decltype(M::e_svx), decltype(M::e_svy), decltype(M::e_mean_vx_frozen), decltype(M::e_mean_vy_frozen),
										  decltype(M::e_smag_uprime), decltype(M::e_suprime2_sum), decltype(M::e_svprime2_sum)>>
: std::integral_constant<bool, (int) M::e_rho == 0 && (int) M::e_vx == 1 && (int) M::e_vy == 2 && (int) M::e_svx == 3 && (int) M::e_svy == 4
									   && (int) M::e_mean_vx_frozen == 5 && (int) M::e_mean_vy_frozen == 6 && (int) M::e_smag_uprime == 7
									   && (int) M::e_suprime2_sum == 8 && (int) M::e_svprime2_sum == 9 && (int) M::N == 10>
{};
template <typename M, bool TAGGED = has_macro_tag<M>::value>
struct macro_kind
{
	static constexpr int lbmx_macro = M::lbmx_macro;
};
template <typename M>
struct macro_kind<M, false>
{
	static_assert(is_with_mean_2d<M>::value, "MACRO has no lbmx_macro tag: a solver-defined macro class is device code and cannot cross the C ABI");
	static constexpr int lbmx_macro = LBMX_MACRO_WITH_MEAN_2D;
};
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
	// host-side NDArray of the reference's layout (defs.h:85-100: permutation (x, z, y), y fastest) for solver-owned fields such as
	// the analytical profile cache of sim_NSE/sim_2.cu:52-113; device-side arrays live behind the C ABI
	template <typename T, typename Device = TNL::Devices::Host>
	struct array3d
	{
		std::vector<T> v;
		idx nx = 0, ny = 0, nz = 0;
		void setSizes(idx x, idx y, idx z)
		{
			nx = x, ny = y, nz = z;
			v.assign((size_t) x * y * z, T());
		}
		void allocate() {}
		T* getData() { return v.empty() ? nullptr : v.data(); }
		const T* getData() const { return v.empty() ? nullptr : v.data(); }
		T& operator()(idx x, idx y, idx z) { return v[((size_t) x * nz + z) * ny + y]; }
		const T& operator()(idx x, idx y, idx z) const { return v[((size_t) x * nz + z) * ny + y]; }
		void setValue(T value) { std::fill(v.begin(), v.end(), value); }
		idx getStorageSize() const { return (idx) v.size(); }
	};
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
struct D3Q27_EQ_ENTROPIC { static constexpr int lbmx_eq = LBMX_EQ_ENTROPIC; };
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
// the reference's cumulant switches (defs.h:254-255) select a kernel family instead of an #ifdef branch
#if defined(USE_GEIER_CUM_2017) && defined(USE_GEIER_CUM_ANTIALIAS)
LBMX_COLL_TAG(D3Q27_CUM, D3Q27_EQ, LBMX_COLL_CUM_2017_ANTIALIAS, "CUM")
#elif defined(USE_GEIER_CUM_2017)
LBMX_COLL_TAG(D3Q27_CUM, D3Q27_EQ, LBMX_COLL_CUM_2017, "CUM")
#elif defined(USE_GEIER_CUM_ANTIALIAS)
LBMX_COLL_TAG(D3Q27_CUM, D3Q27_EQ, LBMX_COLL_CUM_ANTIALIAS, "CUM")
#elif defined(USE_HIGH_PRECISION_RHO)  // defs.h:252: Kahan-summed density (the engine has that build of the default cumulant operator; not of the others)
LBMX_COLL_TAG(D3Q27_CUM, D3Q27_EQ, LBMX_COLL_CUM_HP_RHO, "CUM")
#else
LBMX_COLL_TAG(D3Q27_CUM, D3Q27_EQ, LBMX_COLL_CUM, "CUM")
#endif
LBMX_COLL_TAG(D3Q27_SRT, D3Q27_EQ, LBMX_COLL_SRT, "SRT")
#if defined(USE_GALILEAN_CORRECTION)  // defs.h:253: the solver was built with the switch, the engine has that build of the operator too
LBMX_COLL_TAG(D3Q27_BGK, D3Q27_EQ, LBMX_COLL_BGK_GALILEAN, "BGK")
#else
LBMX_COLL_TAG(D3Q27_BGK, D3Q27_EQ, LBMX_COLL_BGK, "BGK")
#endif
LBMX_COLL_TAG(D3Q27_MRT, D3Q27_EQ, LBMX_COLL_MRT_LES, "MRT_LES")
LBMX_COLL_TAG(D3Q27_CLBM, D3Q27_EQ, LBMX_COLL_CLBM, "CLBM")
LBMX_COLL_TAG(D3Q27_SRT_MODIF_FORCE, D3Q27_EQ, LBMX_COLL_SRT_MODIF_FORCE, "SRT_MRT_MODIF_FORCE")
LBMX_COLL_TAG(D3Q27_KBC_N1, D3Q27_EQ, LBMX_COLL_KBC_N1, "KBC_N1")
LBMX_COLL_TAG(D3Q27_KBC_N2, D3Q27_EQ_ENTROPIC, LBMX_COLL_KBC_N2, "KBC_N2")
LBMX_COLL_TAG(D3Q27_KBC_N3, D3Q27_EQ_ENTROPIC, LBMX_COLL_KBC_N3, "KBC_N3")
LBMX_COLL_TAG(D3Q27_KBC_N4, D3Q27_EQ_ENTROPIC, LBMX_COLL_KBC_N4, "KBC_N4")
LBMX_COLL_TAG(D3Q27_KBC_C1, D3Q27_EQ, LBMX_COLL_KBC_C1, "KBC_C1")
LBMX_COLL_TAG(D3Q27_KBC_C2, D3Q27_EQ_ENTROPIC, LBMX_COLL_KBC_C2, "KBC_C2")
LBMX_COLL_TAG(D3Q27_KBC_C3, D3Q27_EQ_ENTROPIC, LBMX_COLL_KBC_C3, "KBC_C3")
LBMX_COLL_TAG(D3Q27_KBC_C4, D3Q27_EQ_ENTROPIC, LBMX_COLL_KBC_C4, "KBC_C4")
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
// base of solver-defined macro classes (d2q9/macro.h:5-46); the derived class supplies the channel enum and is recognised structurally
template <typename TRAITS>
struct D2Q9_MACRO_Base
{
	static const bool use_syncMacro = false;
	static constexpr int overlap_width = 1;
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
	static constexpr int lbmx_macro = lbmx_host::macro_kind<MACRO>::lbmx_macro;
	static constexpr int lbmx_precision = TRAITS::lbmx_precision;
	static constexpr int lbmx_inflow =
		lbmx_host::has_vx_profile<DATA>::value
			? LBMX_INFLOW_PROFILE_YZ
			: (lbmx_host::has_u_max_lbm<DATA>::value ? LBMX_INFLOW_PARABOLIC_Y : (lbmx_host::has_inflow_vx<DATA>::value ? LBMX_INFLOW_CONST : LBMX_INFLOW_NONE));
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
// checkpoint.h: CheckpointManager with the reference's method names (include/lbm3d/checkpoint.h:6-135), raw-dump backend.
// A checkpoint "file" is a directory (as a .bp is): attributes.txt (name <tab> value), variables.txt (name dtype count) and
// <name>.bin per variable -- the same variable names ("LBM_df_0_block_0", ...) and 1-D local-storage shape (ghost planes
// included) the reference hands to ADIOS2 (checkpoint.h:58-101), so the arrays are interchangeable with a BP dump of them.
// ---------------------------------------------------------------------------------------------------------------------------
#ifndef ADIOS2_H_
namespace adios2 {
enum class Mode { Undefined, Write, Read, Append };
struct ADIOS {};
}  // namespace adios2
#endif

namespace lbmx_host {
inline bool file_exists(const std::string& path)
{
	struct stat st;
	return ::stat(path.c_str(), &st) == 0;
}
inline void make_dirs(const std::string& path)
{
	for (size_t i = 1; i <= path.size(); i++)
		if (i == path.size() || path[i] == '/')
			::mkdir(path.substr(0, i).c_str(), 0777);
}
template <typename T>
constexpr const char* dtype_name()
{
	if (std::is_same<T, double>::value) return "float64";
	if (std::is_same<T, float>::value) return "float32";
	if (std::is_same<T, short>::value) return "int16";
	if (std::is_same<T, int>::value) return "int32";
	return sizeof(T) == 8 ? "int64" : "bytes";
}
}  // namespace lbmx_host

class CheckpointManager
{
	std::string dir;
	adios2::Mode mode = adios2::Mode::Undefined;
	std::map<std::string, std::string> attributes;
	std::map<std::string, std::pair<std::string, size_t>> variables;  // name -> (dtype, count)

	// every rank keeps the index of its own variables; rank 0 also writes the attributes (identical on all ranks)
	static std::string index_name()
	{
		const lbmx_host::World& w = lbmx_host::World::get();
		return w.size > 1 ? "variables_rank_" + std::to_string(w.rank) + ".txt" : "variables.txt";
	}

public:
	CheckpointManager() = default;
	explicit CheckpointManager(adios2::ADIOS&) {}

	void start(const std::string filename, adios2::Mode m)
	{
		dir = filename;
		mode = m;
		attributes.clear();
		variables.clear();
		if (m == adios2::Mode::Write) {
			lbmx_host::make_dirs(dir);
			return;
		}
		std::ifstream fa(dir + "/attributes.txt"), fv(dir + "/" + index_name());
		if (! fa || ! fv)
			throw std::runtime_error("CheckpointManager: cannot open checkpoint " + dir);
		std::string line;
		while (std::getline(fa, line)) {
			const size_t t = line.find('\t');
			if (t != std::string::npos)
				attributes[line.substr(0, t)] = line.substr(t + 1);
		}
		std::string name, dtype;
		size_t count;
		while (fv >> name >> dtype >> count)
			variables[name] = {dtype, count};
	}
	void performDeferred() {}
	void finalize()
	{
		if (mode == adios2::Mode::Write) {
			std::ofstream fv(dir + "/" + index_name());
			for (const auto& v : variables)
				fv << v.first << ' ' << v.second.first << ' ' << v.second.second << '\n';
			bool ok = (bool) fv;
			if (lbmx_host::World::get().rank == 0) {
				std::ofstream fa(dir + "/attributes.txt");
				for (const auto& a : attributes)
					fa << a.first << '\t' << a.second << '\n';
				ok = ok && (bool) fa;
			}
			if (! ok)
				throw std::runtime_error("CheckpointManager: cannot write the index of " + dir);
		}
		mode = adios2::Mode::Undefined;
	}
	// removes a checkpoint directory written by this class (used when a staged checkpoint replaces the previous one)
	static void discard(const std::string& path)
	{
		if (! lbmx_host::file_exists(path + "/attributes.txt"))
			return;
		if (DIR* d = ::opendir(path.c_str())) {
			while (dirent* ent = ::readdir(d)) {
				const std::string n = ent->d_name;
				if (n != "." && n != "..")
					::remove((path + "/" + n).c_str());
			}
			::closedir(d);
		}
		::rmdir(path.c_str());
	}

	template <typename T, typename CastToType = T>
	void saveLoadAttribute(const std::string& name, T& variable)
	{
		if (mode == adios2::Mode::Write) {
			char buf[64];
			if (std::is_floating_point<CastToType>::value)
				std::snprintf(buf, sizeof buf, "%.17g", (double) static_cast<CastToType>(variable));
			else
				std::snprintf(buf, sizeof buf, "%lld", (long long) static_cast<CastToType>(variable));
			attributes[name] = buf;
		}
		else {
			auto it = attributes.find(name);
			if (it == attributes.end())
				throw std::runtime_error("CheckpointManager: attribute " + name + " is not in " + dir);
			if (std::is_floating_point<CastToType>::value)
				variable = static_cast<T>(static_cast<CastToType>(std::strtod(it->second.c_str(), nullptr)));
			else
				variable = static_cast<T>(static_cast<CastToType>(std::strtoll(it->second.c_str(), nullptr, 10)));
		}
	}

	// raw 1-D array under its final name
	template <typename T>
	void saveLoadRaw(const std::string& name, T* data, size_t count)
	{
		const std::string path = dir + "/" + name + ".bin";
		if (mode == adios2::Mode::Write) {
			std::FILE* f = std::fopen(path.c_str(), "wb");
			if (! f || std::fwrite(data, sizeof(T), count, f) != count) {
				if (f)
					std::fclose(f);
				throw std::runtime_error("CheckpointManager: cannot write " + path);
			}
			std::fclose(f);
			variables[name] = {lbmx_host::dtype_name<T>(), count};
		}
		else {
			auto it = variables.find(name);
			if (it == variables.end() || it->second.second != count || it->second.first != lbmx_host::dtype_name<T>())
				throw std::runtime_error("CheckpointManager: variable " + name + " is missing from " + dir + " or has another shape/type");
			std::FILE* f = std::fopen(path.c_str(), "rb");
			if (! f || std::fread(data, sizeof(T), count, f) != count) {
				if (f)
					std::fclose(f);
				throw std::runtime_error("CheckpointManager: cannot read " + path);
			}
			std::fclose(f);
		}
	}
	// checkpoint.h:58-101: one variable per block, name + "_block_<id>", shape = local storage size
	template <typename BLOCK, typename Array>
	void saveLoadVariable(std::string name, BLOCK& block, Array& array)
	{
		saveLoadRaw(name + "_block_" + std::to_string(block.id), array.data(), array.size());
	}
	// checkpoint.h:106-133: per-rank independent arrays (std::vector or anything with data()/size())
	template <typename Array>
	void saveLoadLocalArray(std::string name, int rank, Array& array)
	{
		saveLoadRaw(name + "_rank_" + std::to_string(rank), array.data(), array.size());
	}
	adios2::Mode getMode() const { return mode; }
};

// ---------------------------------------------------------------------------------------------------------------------------
// ADIOSWriter stand-in for the 3-D / cut writers (lbm_block.hpp:800-846, adios_writer.h): same variable names ("wall", "<id>",
// "<id>X/Y/Z", "TIME") and (z,y,x) ordering, float32 / int32 payloads
// ---------------------------------------------------------------------------------------------------------------------------
struct RawWriter
{
	std::ofstream index;
	std::FILE* payload = nullptr;
	std::string base;
	size_t offset = 0;
	// two files per (file name, cycle): <filename>.<cycle>.txt = index (name dtype count dofs byte_offset), .bin = payloads
	RawWriter(const std::string& filename, int cycle, const long (&global)[3], const long (&local)[3], const long (&off)[3], double dl)
	{
		const lbmx_host::World& w = lbmx_host::World::get();
		base = filename + "." + std::to_string(cycle) + (w.size > 1 ? ".rank" + std::to_string(w.rank) : "");  // one piece per rank; the header says where it sits
		const size_t slash = base.rfind('/');
		if (slash != std::string::npos)
			lbmx_host::make_dirs(base.substr(0, slash));
		index.open(base + ".txt");
		payload = std::fopen((base + ".bin").c_str(), "wb");
		if (! index || ! payload)
			throw std::runtime_error("RawWriter: cannot create " + base);
		index << "# global " << global[0] << ' ' << global[1] << ' ' << global[2] << " local " << local[0] << ' ' << local[1] << ' ' << local[2] << " offset "
			  << off[0] << ' ' << off[1] << ' ' << off[2] << " physDl " << dl << " order zyx\n";
	}
	RawWriter(const RawWriter&) = delete;
	~RawWriter()
	{
		if (payload)
			std::fclose(payload);
	}
	template <typename T>
	void write(const std::string& name, const std::vector<T>& v, int dofs)
	{
		if (std::fwrite(v.data(), sizeof(T), v.size(), payload) != v.size())
			throw std::runtime_error("RawWriter: cannot write " + base + ".bin");
		index << name << ' ' << lbmx_host::dtype_name<T>() << ' ' << v.size() << ' ' << dofs << ' ' << offset << '\n';
		offset += v.size() * sizeof(T);
	}
	void write(const std::string& name, double value) { index << name << " scalar " << value << '\n'; }
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
				lbmx_host::check(lbmx_bouzidi_upload(owner->eng(), h.v.data()), "lbmx_bouzidi_upload");
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
	const dreal* uploaded_profile = nullptr;  // inflow profile last handed to the engine (pushParams)
	int64_t uploaded_profile_sy = 0;
	bool profile_uploaded = false;

	// Deferred stepping.  The reference launches one kernel per SimUpdate and copies the macroscopic fields to the host only on the
	// cadence of its counters (state.hpp:1029-1042, 1134-1142); between two such points nothing on the host can see the device.
	// step(defer = true) therefore only RECORDS the step; the recorded steps go to the engine as ONE lbmx_step(n) batch (graph replay
	// on small lattices, rho/u written by the last step only -- LBMX_MACRO_LAST_STEP) as soon as anything needs the device state:
	// every method of this block that touches the engine goes through eng(), which flushes first.  A step whose parameters differ
	// from the recorded ones (viscosity, force, inflow, gates; stat_counter is expected to advance by one per step, as the engine
	// advances it) closes the batch.  Values at every point the host can observe are bit-identical to stepping one at a time.
	int64_t pending = 0;		 // steps recorded and not yet enqueued
	lbmx_params pending_prm{};	 // their parameters (stat_counter: of the first recorded step)
	int64_t steps_enqueued = 0;	 // diagnostics: steps / batches handed to lbmx_step
	int64_t batches_enqueued = 0;
	static constexpr int64_t max_pending = 256;	 // bounds how far the host loop (wall-time checks, flag files) runs ahead of the device
	void flush()
	{
		if (pending > 0) {
			const int64_t n = pending;
			pending = 0;
			lbmx_host::check(lbmx_step(engine, n), "lbmx_step");
			steps_enqueued += n;
			batches_enqueued++;
		}
	}
	lbmx_engine* eng()
	{
		flush();
		return engine;
	}

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
		d.macro_policy = LBMX_MACRO_LAST_STEP;	 // rho,u by the last step of each batch; a batch is one step unless execute() defers (LBM_BLOCK::pending)
		if (const char* v = std::getenv("LBMX_STRICT_ARITH"))  // run an unmodified solver in the reference's own rounding (lbmx.h: LBMX_FLAG_STRICT_ARITH)
			if (v[0] && v[0] != '0')
				d.flags |= LBMX_FLAG_STRICT_ARITH;
		d.X = global.x();
		d.Y = global.y();
		d.Z = global.z();
		d.rank = rank;
		d.nranks = nproc;
		d.device = -1;
		if (nproc > 1) {  // one process per GPU: the launcher's local rank picks the device
			int32_t ndev = 0;
			lbmx_host::check(lbmx_device_count(&ndev), "lbmx_device_count");
			if (ndev < 1)
				throw std::runtime_error("lbmx: no CUDA device visible to rank " + std::to_string(rank));
			d.device = lbmx_host::World::get().local_rank % ndev;
		}
		d.ghost_x = nproc > 1;
		d.periodic_x = periodic_lattice;
		lbmx_host::check(lbmx_create(&d, &engine), "lbmx_create");
		if (nproc > 1) {
			// the communicator of the reference's synchronisers (lbm_block.hpp:410-473) becomes the engine's NCCL communicator:
			// rank 0 draws the id, the process group broadcasts its 128 bytes
			unsigned char id128[128] = {};
			if (rank == 0)
				lbmx_host::check(lbmx_comm_unique_id(id128), "lbmx_comm_unique_id");
			lbmx_host::World::get().bcast(id128, sizeof id128, 0);
			lbmx_host::check(lbmx_comm_init(engine, id128), "lbmx_comm_init");
		}
		refreshPointers();
	}
	void refreshPointers()
	{
		lbmx_ptrs p{};
		lbmx_host::check(lbmx_get_device_ptrs(engine, &p), "lbmx_get_device_ptrs");
		// with steps recorded but not yet enqueued: parity and A-B rotation as they will be once those have run (lbm.hpp:320-327)
		const bool flip = (pending & 1) != 0;
		data.dfs[df_cur] = (dreal*) p.dfs[0];
#if defined(AB_PATTERN)
		data.dfs[df_cur] = (dreal*) p.dfs[flip ? 1 : 0];
		data.dfs[df_out] = (dreal*) p.dfs[flip ? 0 : 1];
#endif
		data.dmacro = (dreal*) p.dmacro;
		data.dmap = (map_t*) p.dmap;
		data.even_iter = (p.even_iter != 0) != flip;
	}
	// LBM::updateKernelData sets the iteration the next step runs as; a no-op while the recorded steps account for the difference
	void setIterations(int64_t it)
	{
		int64_t have = 0;
		lbmx_host::check(lbmx_get_iterations(engine, &have), "lbmx_get_iterations");
		if (have + pending == it)
			return;
		flush();
		lbmx_host::check(lbmx_set_iterations(engine, it), "lbmx_set_iterations");
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
		lbmx_host::check(lbmx_map_upload(eng(), hmap.v.data(), 0), "lbmx_map_upload");
		if (hBouzidi.getData() != nullptr)
			dBouzidi = hBouzidi;
		lbmx_stats st{};
		if (lbmx_get_stats(engine, &st) == LBMX_OK && st.aa_cells_reaching_outside > 0 && ! warned_aa_faces) {
			warned_aa_faces = true;	 // once per block: the map is uploaded again after every repaint
			lbmx_host::log_info("warning: %lld cells on the lattice faces take A-A neighbours outside the lattice (kernels.h:30-37 is unclamped: undefined in the "
								"reference, meaningless values there here); give the faces a GEO_NOTHING or periodic skin",
								(long long) st.aa_cells_reaching_outside);
		}
	}
	bool warned_aa_faces = false;
	void copyMapToHost() { lbmx_host::check(lbmx_map_download(eng(), hmap.v.data(), 0), "lbmx_map_download"); }
	void copyMacroToHost()
	{
		if (MACRO::N > 0)
			lbmx_host::check(lbmx_macro_download(eng(), hmacro.v.data(), 0), "lbmx_macro_download");
	}
	void copyMacroToDevice()
	{
		if (MACRO::N > 0)
			lbmx_host::check(lbmx_macro_upload(eng(), hmacro.v.data(), 0), "lbmx_macro_upload");
	}
	void copyDFsToHost(uint8_t dftype)
	{
		HostField& h = hfs[dftype];
		const size_t n = (size_t) local.x() * local.y() * local.z();
		h.v.resize(n * CONFIG::Q);
		h.off = offset;
		h.loc = local;
		h.n = n;
		lbmx_host::check(lbmx_df_download(eng(), dftype == df_cur ? 0 : 1, h.v.data(), 0), "lbmx_df_download");
	}
	void copyDFsToDevice(uint8_t dftype) { lbmx_host::check(lbmx_df_upload(eng(), dftype == df_cur ? 0 : 1, hfs[dftype].v.data(), 0), "lbmx_df_upload"); }
	void setEquilibrium(real rho, real vx, real vy, real vz) { lbmx_host::check(lbmx_df_set_equilibrium(eng(), rho, vx, vy, vz), "lbmx_df_set_equilibrium"); }
	void computeInitialMacro() { lbmx_host::check(lbmx_macro_init(eng()), "lbmx_macro_init"); }

	// Checkpoint variables of this block (state.hpp:712-727): map, every DF copy and the macro array in LOCAL STORAGE shape, i.e.
	// ghost x-planes included, moved straight between the engine and the checkpoint (no second host copy of 29 GB of DFs).
	void checkpoint(CheckpointManager& ck)
	{
		lbmx_layout L{};
		lbmx_host::check(lbmx_get_layout(eng(), &L), "lbmx_get_layout");
		const bool save = ck.getMode() == adios2::Mode::Write;
		const std::string blk = "_block_" + std::to_string(id);
		{
			std::vector<map_t> m((size_t) L.XYZ);
			if (save)
				lbmx_host::check(lbmx_map_download(eng(), m.data(), 1), "lbmx_map_download");
			ck.saveLoadRaw("LBM_map" + blk, m.data(), m.size());
			if (! save) {
				lbmx_host::check(lbmx_map_upload(eng(), m.data(), 1), "lbmx_map_upload");
				copyMapToHost();
			}
		}
		std::vector<dreal> a((size_t) L.XYZ * std::max<int>(L.Q, L.n_macro));
		for (int dfty = 0; dfty < L.dfmax; dfty++) {
			const size_t n = (size_t) L.XYZ * L.Q;
			if (save)
				lbmx_host::check(lbmx_df_download(eng(), dfty, a.data(), 1), "lbmx_df_download");
			ck.saveLoadRaw("LBM_df_" + std::to_string(dfty) + blk, a.data(), n);
			if (! save)
				lbmx_host::check(lbmx_df_upload(eng(), dfty, a.data(), 1), "lbmx_df_upload");
		}
		if (MACRO::N > 0) {
			const size_t n = (size_t) L.XYZ * L.n_macro;
			if (save)
				lbmx_host::check(lbmx_macro_download(eng(), a.data(), 1), "lbmx_macro_download");
			ck.saveLoadRaw("LBM_macro" + blk, a.data(), n);
			if (! save) {
				lbmx_host::check(lbmx_macro_upload(eng(), a.data(), 1), "lbmx_macro_upload");
				copyMacroToHost();
			}
		}
	}

	// Output of the solver's outputData() hook over a box of this block in the reference's (z, y, x) order and variable naming
	// (writeVTK_3D / writeVTK_3Dcut / writeVTK_2DcutX/Y/Z, lbm_block.hpp:800-1110): "wall" = map as int32, then every field the
	// hook enumerates, scalar "<id>" or vector components "<id>X/Y/Z", float32; box = origin o, extent g, stride `step`.
	template <typename LAT, typename Output>
	void writeBox(const LAT& lat, Output&& outputData, const std::string& filename, real time, int cycle, idx ox, idx oy, idx oz, idx gx, idx gy, idx gz, idx step) const
	{
		const idx x0 = std::max(ox, offset.x()), x1 = std::min(ox + gx, offset.x() + local.x());
		const idx y0 = std::max(oy, offset.y()), y1 = std::min(oy + gy, offset.y() + local.y());
		const idx z0 = std::max(oz, offset.z()), z1 = std::min(oz + gz, offset.z() + local.z());
		if (x0 >= x1 || y0 >= y1 || z0 >= z1)
			return;
		auto cnt = [step](idx n) { return (long) ((n + step - 1) / step); };
		const long G[3] = {cnt(gx), cnt(gy), cnt(gz)}, Lc[3] = {cnt(x1 - x0), cnt(y1 - y0), cnt(z1 - z0)};
		const long O[3] = {cnt(x0 - ox), cnt(y0 - oy), cnt(z0 - oz)};
		RawWriter out(filename, cycle, G, Lc, O, (double) lat.physDl * (double) step);
		std::vector<int> wall;
		for (idx z = z0; z < z1; z += step)
			for (idx y = y0; y < y1; y += step)
				for (idx x = x0; x < x1; x += step)
					wall.push_back(hmap(x, y, z));
		out.write("wall", wall, 1);
		char idd[500];
		real value;
		int dofs = 1;
		std::vector<float> field;
		for (int index = 0; outputData(*this, index, 0, idd, offset.x(), offset.y(), offset.z(), value, dofs); index++) {
			const std::string name(idd);
			const int n = dofs;
			for (int dof = 0; dof < n; dof++) {
				field.clear();
				for (idx z = z0; z < z1; z += step)
					for (idx y = y0; y < y1; y += step)
						for (idx x = x0; x < x1; x += step) {
							outputData(*this, index, dof, idd, x, y, z, value, dofs);
							field.push_back((float) value);
						}
				static const char* suffix[3] = {"X", "Y", "Z"};
				out.write(n > 1 && dof < 3 ? name + suffix[dof] : name, field, n);
			}
		}
		out.write("TIME", (double) time);
	}
	template <typename LAT, typename Output>
	void writeVTK_3D(const LAT& lat, Output&& o, const std::string& f, real time, int cycle) const
	{
		writeBox(lat, o, f, time, cycle, 0, 0, 0, global.x(), global.y(), global.z(), 1);
	}
	template <typename LAT, typename Output>
	void writeVTK_3Dcut(const LAT& lat, Output&& o, const std::string& f, real time, int cycle, idx ox, idx oy, idx oz, idx gx, idx gy, idx gz, idx step) const
	{
		writeBox(lat, o, f, time, cycle, ox, oy, oz, gx, gy, gz, step);
	}
	template <typename LAT, typename Output>
	void writeVTK_2DcutX(const LAT& lat, Output&& o, const std::string& f, real time, int cycle, idx X) const
	{
		writeBox(lat, o, f, time, cycle, X, 0, 0, 1, global.y(), global.z(), 1);
	}
	template <typename LAT, typename Output>
	void writeVTK_2DcutY(const LAT& lat, Output&& o, const std::string& f, real time, int cycle, idx Y) const
	{
		writeBox(lat, o, f, time, cycle, 0, Y, 0, global.x(), 1, global.z(), 1);
	}
	template <typename LAT, typename Output>
	void writeVTK_2DcutZ(const LAT& lat, Output&& o, const std::string& f, real time, int cycle, idx Z) const
	{
		writeBox(lat, o, f, time, cycle, 0, 0, Z, global.x(), global.y(), 1, 1);
	}

	// block.data -> lbmx_params (what passing the POD by value to the kernel did in the reference, state.hpp:1039)
	lbmx_params makeParams() const
	{
		lbmx_params p{};
		p.lbmViscosity = data.lbmViscosity;
		p.fx = data.fx;
		p.fy = data.fy;
		p.fz = data.fz;
		if constexpr (lbmx_host::has_u_max_lbm<typename CONFIG::DATA>::value) {	 // NSE2D_Data_ParabolicInflow (sim_2D/sim2d_3.cu:36-55)
			p.inflow_vx = data.u_max_lbm;
			p.inflow_vy = (double) data.y0;
			p.inflow_vz = data.inv_den;
		}
		if constexpr (lbmx_host::has_inflow_vx<typename CONFIG::DATA>::value)
			p.inflow_vx = data.inflow_vx;
		if constexpr (lbmx_host::has_inflow_vy<typename CONFIG::DATA>::value)
			p.inflow_vy = data.inflow_vy;
		if constexpr (lbmx_host::has_inflow_vz<typename CONFIG::DATA>::value)
			p.inflow_vz = data.inflow_vz;
		p.stat_counter = data.stat_counter;
		if constexpr (lbmx_host::has_accumulate_gates<typename CONFIG::DATA>::value)  // gates of D2Q9_MACRO_WithMean (sim_2D/sim2d_2.cu:121-122)
			p.macro_gates = (data.accumulate_means ? LBMX_GATE_MEANS : 0) | (data.accumulate_flucs ? LBMX_GATE_FLUCS : 0);
		return p;
	}
	// NSE_Data_XProfileInflow (sim_NSE/sim_2.cu:16-33): the solver owns a host array dreal[y + z * size_y]; it is uploaded when the
	// pointer or its size changes.  Without a profile (periodic runs with forcing) the inflow cells do not exist: a zero profile.
	bool profileChanged() const
	{
		if constexpr (lbmx_host::has_vx_profile<typename CONFIG::DATA>::value) {
			const int64_t sy = data.vx_profile ? (int64_t) data.size_y : (int64_t) local.y();
			return data.vx_profile != uploaded_profile || sy != uploaded_profile_sy || ! profile_uploaded;
		}
		return false;
	}
	void pushProfile()
	{
		if constexpr (lbmx_host::has_vx_profile<typename CONFIG::DATA>::value) {
			if (! profileChanged())
				return;
			const dreal* src = data.vx_profile;
			const int64_t sy = src ? (int64_t) data.size_y : (int64_t) local.y();
			std::vector<dreal> zeros;
			if (! src) {
				zeros.assign((size_t) local.y() * local.z(), (dreal) 0);
				src = zeros.data();
			}
			lbmx_host::check(lbmx_set_inflow_profile(eng(), src, sy, (int64_t) local.z()), "lbmx_set_inflow_profile");
			uploaded_profile = data.vx_profile;
			uploaded_profile_sy = sy;
			profile_uploaded = true;
		}
	}
	void pushParams()
	{
		const lbmx_params p = makeParams();
		lbmx_host::check(lbmx_set_params(eng(), &p), "lbmx_set_params");
		pushProfile();
	}
	// One time step with block.data as it stands now (the kernel launch of state.hpp:1034-1108).  defer: record it (see `pending`).
	void step(bool defer)
	{
		const lbmx_params p = makeParams();
		if (pending > 0) {
			// stat_counter is read by the running means of MACRO_Mean only (d3q27/macro.h:120): there the engine advances it by one per
			// step of a batch, so the solver's value has to advance the same way for the batch to continue; elsewhere it is unused
			lbmx_params expect = pending_prm;
			if (CONFIG::lbmx_macro == LBMX_MACRO_MEAN)
				expect.stat_counter += (int) pending;
			else
				expect.stat_counter = p.stat_counter;
			if (std::memcmp(&p, &expect, sizeof p) != 0 || profileChanged())
				flush();
		}
		if (pending == 0) {
			lbmx_host::check(lbmx_set_params(engine, &p), "lbmx_set_params");
			pushProfile();
			pending_prm = p;
		}
		pending++;
		if (! defer)
			flush();
		else if (pending >= max_pending) {
			lbmx_host::check(lbmx_sync(engine), "lbmx_sync");  // at most one batch in flight: the host loop stays within max_pending steps of the device
			flush();
		}
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
	bool defer_steps = false;  // set by execute() when the solver overrides none of the per-step hooks that could look at the device (LBM_BLOCK::pending)

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
		blocks.back().id = rank;
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
			b.setIterations(iterations);
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
	using lbmx_state_base = State;	// lets execute() tell inherited hooks from overridden ones

	std::string id;
	LBM<NSE> nse;
	T_COUNTER cnt[MAX_COUNTER];
	CheckpointManager checkpoint;

	struct T_PROBE2DCUT
	{
		std::string name;
		int type = 0;  // 0,1,2 = plane normal to x,y,z
		int cycle = 0;
		idx position = 0;
	};
	struct T_PROBE3DCUT
	{
		std::string name;
		idx ox = 0, oy = 0, oz = 0, lx = 0, ly = 0, lz = 0, step = 1;
		int cycle = 0;
	};
	struct T_PROBE1DCUT
	{
		std::string name;
		int type = 0;  // 0,1,2 = the line runs along x,y,z
		idx pos1 = 0, pos2 = 0;
		int cycle = 0;
	};
	struct T_PROBE1DLINECUT
	{
		std::string name;
		point_t from, to;
		int cycle = 0;
	};
	std::vector<T_PROBE2DCUT> probe2Dvec;
	std::vector<T_PROBE3DCUT> probe3Dvec;
	std::vector<T_PROBE1DCUT> probe1Dvec;
	std::vector<T_PROBE1DLINECUT> probe1Dlinevec;

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
	template <typename... ARGS>
	static std::string cut_name(const char* fmts, ARGS... args)
	{
#ifdef LBMX_HAVE_FMT
		return fmt::format(fmts, args...);
#else
		return fmts;
#endif
	}
	auto output_hook()
	{
		return [this](const BLOCK_NSE& block, int index, int dof, char* desc, idx x, idx y, idx z, real& value, int& dofs)
		{
			return this->outputData(block, index, dof, desc, x, y, z, value, dofs);
		};
	}
	// state.hpp:383-400, 410-470, 473-542: raw dumps results_<id>/output_3D.<cycle>.{txt,bin}, output_3Dcut_<name>.<cycle>.*,
	// output_2D_<name>.<cycle>.* (RawWriter)
	virtual void writeVTKs_3D()
	{
		lbmx_host::make_dirs("results_" + id);
		for (const auto& block : nse.blocks)
			block.writeVTK_3D(nse.lat, output_hook(), "results_" + id + "/output_3D", nse.physTime(), cnt[VTK3D].count);
	}
	virtual void writeVTKs_3Dcut()
	{
		lbmx_host::make_dirs("results_" + id);
		for (auto& p : probe3Dvec) {
			for (const auto& block : nse.blocks)
				block.writeVTK_3Dcut(nse.lat, output_hook(), "results_" + id + "/output_3Dcut_" + p.name, nse.physTime(), p.cycle, p.ox, p.oy, p.oz, p.lx, p.ly, p.lz, p.step);
			p.cycle++;
		}
	}
	virtual void writeVTKs_2D()
	{
		lbmx_host::make_dirs("results_" + id);
		for (auto& p : probe2Dvec) {
			const std::string fname = "results_" + id + "/output_2D_" + p.name;
			for (const auto& block : nse.blocks) {
				if (p.type == 0)
					block.writeVTK_2DcutX(nse.lat, output_hook(), fname, nse.physTime(), p.cycle, p.position);
				else if (p.type == 1)
					block.writeVTK_2DcutY(nse.lat, output_hook(), fname, nse.physTime(), p.cycle, p.position);
				else
					block.writeVTK_2DcutZ(nse.lat, output_hook(), fname, nse.physTime(), p.cycle, p.position);
			}
			p.cycle++;
		}
	}
	// 1-D cuts along an axis (add1Dcut_X/Y/Z) and along a physical line (add1Dcut): text tables of the outputData() fields in the
	// reference's format, results_<id>/probes1D/<name>_rank<rrr>_<cycle> (state.hpp:175-340)
	virtual void writeVTKs_1D()
	{
		for (auto& pr : probe1Dvec) {
			char suffix[64];
			std::snprintf(suffix, sizeof suffix, "_rank%03d_%06d", nse.rank, pr.cycle);
			const std::string fname = "results_" + id + "/probes1D/" + pr.name + suffix;
			const size_t slash = fname.rfind('/');
			lbmx_host::make_dirs(fname.substr(0, slash));
			write1Dcut_axis(pr.type, pr.pos1, pr.pos2, fname);
			pr.cycle++;
		}
		for (auto& pr : probe1Dlinevec) {
			char suffix[64];
			std::snprintf(suffix, sizeof suffix, "_rank%03d_%06d", nse.rank, pr.cycle);
			const std::string fname = "results_" + id + "/probes1D/" + pr.name + suffix;
			const size_t slash = fname.rfind('/');
			lbmx_host::make_dirs(fname.substr(0, slash));
			write1Dcut(pr.from, pr.to, fname);
			pr.cycle++;
		}
	}
	// header "#time ...", "#1:<coordinate>  2:<field> ..." as the reference writes it (state.hpp:225-238)
	void write_1d_header(std::FILE* fout, const char* coordinate)
	{
		char idd[500];
		real value;
		int dofs = 1;
		std::fprintf(fout, "#time %f s\n", (double) nse.physTime());
		std::fprintf(fout, "#1:%s", coordinate);
		const BLOCK_NSE& b0 = nse.blocks.front();
		int count = 2;
		for (int index = 0; outputData(b0, index, 0, idd, b0.offset.x(), b0.offset.y(), b0.offset.z(), value, dofs); index++) {
			if (dofs == 1)
				std::fprintf(fout, "\t%d:%s", count++, idd);
			else
				for (int i = 0; i < dofs; i++)
					std::fprintf(fout, "\t%d:%s[%d]", count++, idd, i);
		}
		std::fprintf(fout, "\n");
	}
	void write_1d_row(std::FILE* fout, const BLOCK_NSE& block, double coordinate, idx x, idx y, idx z)
	{
		char idd[500];
		real value;
		int dofs = 1;
		std::fprintf(fout, "%e", coordinate);
		for (int index = 0; outputData(block, index, 0, idd, block.offset.x(), block.offset.y(), block.offset.z(), value, dofs); index++) {
			const int n = dofs;
			for (int dof = 0; dof < n; dof++) {
				outputData(block, index, dof, idd, x, y, z, value, dofs);
				std::fprintf(fout, "\t%e", (double) value);
			}
		}
		std::fprintf(fout, "\n");
	}
	// type 0/1/2: the line runs along x/y/z through the two given lattice coordinates of the other axes (state.hpp:262-340)
	void write1Dcut_axis(int type, idx pos1, idx pos2, const std::string& fname)
	{
		std::FILE* fout = std::fopen(fname.c_str(), "wt");
		if (! fout)
			throw std::runtime_error("write1Dcut: cannot create " + fname);
		write_1d_header(fout, type == 0 ? "x" : (type == 1 ? "y" : "z"));
		for (const auto& block : nse.blocks) {
			const idx lo = type == 0 ? block.offset.x() : (type == 1 ? block.offset.y() : block.offset.z());
			const idx n = type == 0 ? block.local.x() : (type == 1 ? block.local.y() : block.local.z());
			for (idx i = lo; i < lo + n; i++) {
				const idx x = type == 0 ? i : pos1, y = type == 1 ? i : (type == 0 ? pos1 : pos2), z = type == 2 ? i : pos2;
				if (! block.isLocalIndex(x, y, z))
					continue;
				const double coordinate = type == 0 ? nse.lat.lbm2physX(i) : (type == 1 ? nse.lat.lbm2physY(i) : nse.lat.lbm2physZ(i));
				write_1d_row(fout, block, coordinate, x, y, z);
			}
		}
		std::fclose(fout);
	}
	void write1Dcut_X(idx y, idx z, const std::string& fname) { write1Dcut_axis(0, y, z, fname); }
	void write1Dcut_Y(idx x, idx z, const std::string& fname) { write1Dcut_axis(1, x, z, fname); }
	void write1Dcut_Z(idx x, idx y, const std::string& fname) { write1Dcut_axis(2, x, y, fname); }
	// sampling along a physical line (state.hpp:210-259)
	void write1Dcut(point_t from, point_t to, const std::string& fname)
	{
		std::FILE* fout = std::fopen(fname.c_str(), "wt");
		if (! fout)
			throw std::runtime_error("write1Dcut: cannot create " + fname);
		const point_t i = nse.lat.phys2lbmPoint(from), f = nse.lat.phys2lbmPoint(to);
		const real dx = i.x() - f.x(), dy = i.y() - f.y(), dz = i.z() - f.z();
		const real dist = std::sqrt(dx * dx + dy * dy + dz * dz);
		real ds = (real) 1.0 / (dist * (real) 2.0);
		if ((i[0] == f[0] && i[1] == f[1]) || (i[1] == f[1] && i[2] == f[2]) || (i[0] == f[0] && i[2] == f[2]))
			ds = (real) 1.0 / dist;	 // sampling along an axis: one sample per cell
		write_1d_header(fout, "rel_pos");
		for (real sv = 0; sv <= (real) 1.0; sv += ds) {
			const real px = i.x() + sv * (f.x() - i.x()), py = i.y() + sv * (f.y() - i.y()), pz = i.z() + sv * (f.z() - i.z());
			for (const auto& block : nse.blocks)
				if (block.isLocalIndex((idx) px, (idx) py, (idx) pz))
					write_1d_row(fout, block, (double) ((sv * dist - (real) 0.5) * nse.lat.physDl), (idx) px, (idx) py, (idx) pz);
		}
		std::fclose(fout);
	}
	template <typename... ARGS>
	void add2Dcut(int type, idx pos, const char* fmts, ARGS... args)
	{
		T_PROBE2DCUT p;
		p.name = cut_name(fmts, args...);
		p.type = type;
		p.position = pos;
		probe2Dvec.push_back(p);
	}
	template <typename... ARGS>
	void add2Dcut_X(idx x, const char* fmts, ARGS... args) { add2Dcut(0, x, fmts, args...); }
	template <typename... ARGS>
	void add2Dcut_Y(idx y, const char* fmts, ARGS... args) { add2Dcut(1, y, fmts, args...); }
	template <typename... ARGS>
	void add2Dcut_Z(idx z, const char* fmts, ARGS... args) { add2Dcut(2, z, fmts, args...); }
	template <typename... ARGS>
	void add3Dcut(idx ox, idx oy, idx oz, idx lx, idx ly, idx lz, idx step, const char* fmts, ARGS... args)
	{
		T_PROBE3DCUT p;
		p.name = cut_name(fmts, args...);
		p.ox = ox, p.oy = oy, p.oz = oz, p.lx = lx, p.ly = ly, p.lz = lz, p.step = step;
		probe3Dvec.push_back(p);
	}
	// state.hpp:72-170: a line between two physical points, or along an axis through two physical coordinates of the other axes
	template <typename... ARGS>
	void add1Dcut(point_t from, point_t to, const char* fmts, ARGS... args)
	{
		T_PROBE1DLINECUT p;
		p.name = cut_name(fmts, args...);
		p.from = from;
		p.to = to;
		probe1Dlinevec.push_back(p);
	}
	template <typename... ARGS>
	void add1Dcut_axis(int type, idx pos1, idx pos2, const char* fmts, ARGS... args)
	{
		T_PROBE1DCUT p;
		p.name = cut_name(fmts, args...);
		p.type = type;
		p.pos1 = pos1;
		p.pos2 = pos2;
		probe1Dvec.push_back(p);
	}
	template <typename... ARGS>
	void add1Dcut_X(real y, real z, const char* fmts, ARGS... args) { add1Dcut_axis(0, nse.lat.phys2lbmY(y), nse.lat.phys2lbmZ(z), fmts, args...); }
	template <typename... ARGS>
	void add1Dcut_Y(real x, real z, const char* fmts, ARGS... args) { add1Dcut_axis(1, nse.lat.phys2lbmX(x), nse.lat.phys2lbmZ(z), fmts, args...); }
	template <typename... ARGS>
	void add1Dcut_Z(real x, real y, const char* fmts, ARGS... args) { add1Dcut_axis(2, nse.lat.phys2lbmX(x), nse.lat.phys2lbmY(y), fmts, args...); }

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
	// State::SimInit (state.hpp:907-977)
	virtual void SimInit()
	{
		timer_SimInit.start();
		nse.allocateDeviceData();
		nse.iterations = 0;
		for (auto& c : cnt)
			c.count = 0;
		cnt[SAVESTATE].count = 1;  // skip the initial save of state
		if (flagExists("loadstate"))
			loadState();  // engine arrays and host mirrors come from results_<id>/checkpoint.bp
		else
			reset();
		lbmx_host::log_info("lbmx: %s lattice %ld x %ld x %ld, %s, %s, %s", NSE::COLL::id, (long) nse.lat.global.x(), (long) nse.lat.global.y(),
							(long) nse.lat.global.z(), NSE::lbmx_precision == LBMX_F64 ? "fp64" : "fp32", NSE::lbmx_streaming == LBMX_STREAM_AA ? "A-A" : "A-B",
							"raw-dump writers");
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
		for (auto& b : nse.blocks)
			b.step(nse.defer_steps);
		nse.iterations++;
		bool doCopy = false;
		for (int c = 0; c < MAX_COUNTER; c++)
			if (c != PRINT && c != SAVESTATE)
				doCopy |= cnt[c].action(nse.physTime());
		if (doCopy)
			nse.copyMacroToHost();
		timer_SimUpdate.stop();
	}
	// State::AfterSimUpdate (state.hpp:1148-1278): hook, NaN scan on the cadence of the other actions, probes and writers in the
	// reference's order, statistics resets last (followed by a macro upload), GLUPS line on rank 0
	virtual void AfterSimUpdate()
	{
		computeAfterLBMKernel();
		const real t = nse.physTime();
		bool write_info = false;
		for (int c : {(int) PRINT, (int) VTK1D, (int) VTK2D, (int) VTK3D, (int) VTK3DCUT, (int) PROBE1, (int) PROBE2, (int) PROBE3})
			write_info |= cnt[c].action(t);
		if (write_info)
			cnt[PRINT].count++;
		bool nan_detected = false;
		if (nse.iterations > 1 && write_info && MACRO::N > 0) {
			int32_t nan = 0;
			lbmx_host::check(lbmx_has_nan(nse.blocks.front().eng(), &nan), "lbmx_has_nan");
			nan_detected = TNL::MPI::reduce(nan != 0, MPI_LOR, nse.communicator);  // every rank takes the same decision (state.hpp:1166-1188)
			if (nan_detected) {
				lbmx_host::log_info("Detected NaN, terminating the simulation.");
				nse.terminate = true;
				nse.copyMacroToHost();
			}
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
		if (cnt[VTK3D].action(t) || nan_detected) {
			writeVTKs_3D();
			cnt[VTK3D].count++;
		}
		if (cnt[VTK3DCUT].action(t)) {
			writeVTKs_3Dcut();
			cnt[VTK3DCUT].count++;
		}
		if (cnt[VTK2D].action(t) || nan_detected) {
			writeVTKs_2D();
			cnt[VTK2D].count++;
		}
		if (cnt[VTK1D].action(t)) {
			writeVTKs_1D();
			cnt[VTK1D].count++;
		}
		// statReset is called after all probes and output; the cleared statistics go back to the device (state.hpp:1231-1242)
		if (cnt[STAT_RESET].action(t)) {
			statReset();
			nse.copyMacroToDevice();
			cnt[STAT_RESET].count++;
		}
		if (cnt[STAT2_RESET].action(t)) {
			stat2Reset();
			nse.copyMacroToDevice();
			cnt[STAT2_RESET].count++;
		}
		if (write_info && nse.iterations > 1 && nse.rank == 0) {
			for (auto& b : nse.blocks)
				lbmx_host::check(lbmx_sync(b.eng()), "lbmx_sync");
			const double now = timer_total.getRealTime();
			const double glups = (nse.iterations - glups_prev_iterations) / std::max(1e-6, now - glups_prev_time) * (double) nse.lat.global.x()
							   * (double) nse.lat.global.y() * (double) nse.lat.global.z() * 1e-9;
			lbmx_host::log_info("GLUPS=%.3f iter=%d t=%1.3fs dt=%1.2e lbmVisc=%1.2e WT=%.0fs", glups, nse.iterations, (double) t, (double) nse.lat.physDt,
								(double) nse.lat.lbmViscosity(), now);
			glups_prev_iterations = nse.iterations;
			glups_prev_time = now;
		}
	}
	virtual void AfterSimFinished()
	{
		for (auto& b : nse.blocks)
			lbmx_host::check(lbmx_sync(b.eng()), "lbmx_sync");
		const double total = timer_total.getRealTime(), upd = timer_SimUpdate.getRealTime();
		const double cells = (double) nse.lat.global.x() * (double) nse.lat.global.y() * (double) nse.lat.global.z();
		lbmx_host::log_info("total walltime: %.1f s, SimInit time: %.1f s, SimUpdate time: %.1f s", total, timer_SimInit.getRealTime(), upd);
		lbmx_host::log_info("final GLUPS: average (based on total time) %.3f", cells * nse.iterations / (total + 1e-30) * 1e-9);
		if (! nse.blocks.empty())
			lbmx_host::log_info("lbmx: %lld steps enqueued in %lld batches (%s)", (long long) nse.blocks.front().steps_enqueued, (long long) nse.blocks.front().batches_enqueued,
								nse.defer_steps ? "steps between host-observable points travel together" : "one batch per SimUpdate");
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

	// results-directory flags (state.hpp:13-66); directory locking is out of scope
	std::string flag_path(const char* flagname) const { return "results_" + id + "/flag." + flagname; }
	void flagCreate(const char* flagname)
	{
		if (nse.rank != 0)
			return;
		lbmx_host::make_dirs("results_" + id);
		std::ofstream(flag_path(flagname)).put('\n');
	}
	void flagDelete(const char* flagname)
	{
		if (nse.rank == 0)
			::remove(flag_path(flagname).c_str());
	}
	bool flagExists(const char* flagname) { return lbmx_host::file_exists(flag_path(flagname)); }
	bool canCompute()
	{
		int result = 1;
		if (nse.rank == 0) {
			if (flagExists("loadstate"))
				result = 1;
			else if (flagExists("finished") || flagExists("terminated")) {
				lbmx_host::log_info("results_%s is in finished/terminated state, there is nothing to compute", id.c_str());
				result = 0;
			}
		}
		TNL::MPI::Bcast(&result, 1, 0, nse.communicator);
		return result != 0;
	}

	// State::checkpointState (state.hpp:677-738): same attribute and variable names
	void checkpointState(adios2::Mode mode)
	{
		checkpoint.saveLoadAttribute("LBM_total_blocks", nse.total_blocks);
		checkpoint.saveLoadAttribute("LBM_physCharLength", nse.physCharLength);
		checkpoint.saveLoadAttribute("LBM_physFinalTime", nse.physFinalTime);
		checkpoint.saveLoadAttribute("LBM_iterations", nse.iterations);
		for (int c = 0; c < MAX_COUNTER; c++) {
			const std::string name = "State_counter_" + std::to_string(c);
			checkpoint.saveLoadAttribute(name + "_count", cnt[c].count);
			checkpoint.saveLoadAttribute(name + "_period", cnt[c].period);
		}
		for (std::size_t i = 0; i < probe3Dvec.size(); i++)
			checkpoint.saveLoadAttribute("State_probe3D_" + std::to_string(i) + "_cycle", probe3Dvec[i].cycle);
		for (std::size_t i = 0; i < probe2Dvec.size(); i++)
			checkpoint.saveLoadAttribute("State_probe2D_" + std::to_string(i) + "_cycle", probe2Dvec[i].cycle);
		for (std::size_t i = 0; i < probe1Dvec.size(); i++)
			checkpoint.saveLoadAttribute("State_probe1D_" + std::to_string(i) + "_cycle", probe1Dvec[i].cycle);
		for (std::size_t i = 0; i < probe1Dlinevec.size(); i++)
			checkpoint.saveLoadAttribute("State_probe1Dline_" + std::to_string(i) + "_cycle", probe1Dlinevec[i].cycle);
		for (auto& block : nse.blocks) {
			if (mode == adios2::Mode::Read)	 // "df_cur" / "df_out" are roles that rotate with the iteration count (lbm.hpp:314-330): restore it first
				block.setIterations(nse.iterations);
			block.checkpoint(checkpoint);
		}
		if (mode == adios2::Mode::Read) {
			nse.physStartTime = nse.physTime();
			nse.startIterations = nse.iterations;
			glups_prev_iterations = nse.startIterations;
			glups_prev_time = timer_total.getRealTime();
			nse.updateKernelData();	 // A-A parity and A-B rotation follow the restored iteration count
		}
	}
	virtual void checkpointStateLocal(adios2::Mode) {}
	// State::saveState / loadState (state.hpp:740-781): stage, swap, flag
	void saveState()
	{
		const std::string tmp = "results_" + id + "/checkpoint_tmp.bp", fin = "results_" + id + "/checkpoint.bp";
		lbmx_host::log_info("Saving checkpoint in %s", tmp.c_str());
		if (nse.rank == 0)
			CheckpointManager::discard(tmp);  // leftovers of an interrupted save
		TNL::MPI::Barrier(nse.communicator);
		checkpoint.start(tmp, adios2::Mode::Write);
		checkpointState(adios2::Mode::Write);
		checkpointStateLocal(adios2::Mode::Write);
		checkpoint.finalize();
		TNL::MPI::Barrier(nse.communicator);  // every rank's piece is on disk before the staged directory replaces the old one
		if (nse.rank == 0) {
			CheckpointManager::discard(fin);
			if (::rename(tmp.c_str(), fin.c_str()) != 0)
				throw std::runtime_error("saveState: cannot move " + tmp + " to " + fin + ": " + std::strerror(errno));
		}
		flagCreate("loadstate");
		TNL::MPI::Barrier(nse.communicator);
	}
	void loadState()
	{
		const std::string fin = "results_" + id + "/checkpoint.bp";
		lbmx_host::log_info("Loading data from checkpoint in %s", fin.c_str());
		checkpoint.start(fin, adios2::Mode::Read);
		checkpointState(adios2::Mode::Read);
		checkpointStateLocal(adios2::Mode::Read);
		checkpoint.finalize();
	}

	TNL::Timer timer_total;
	long wallTime = -1;
	bool wallTimeReached()	// collective: all ranks stop together (state.hpp:783-795)
	{
		const bool local = wallTime > 0 && timer_total.getRealTime() >= (double) wallTime;
		return TNL::MPI::reduce(local, MPI_LOR, nse.communicator);
	}
	double getWallTime(bool collective = false)
	{
		const double t = timer_total.getRealTime();
		return collective ? TNL::MPI::reduce(t, MPI_MAX, nse.communicator) : t;
	}
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
namespace lbmx_host {
// true when STATE inherits State<NSE>::NAME unchanged: &STATE::NAME then still has the base's member-pointer type
#define LBMX_NOT_OVERRIDDEN(STATE, NAME) std::is_same<decltype(&STATE::NAME), void (STATE::lbmx_state_base::*)()>::value
template <typename STATE>
constexpr bool steps_can_be_deferred()
{
	// hooks that run between two kernel launches and may look at the device through raw pointers (block.data.dmacro / dfs) or launch
	// kernels of their own; updateKernelData / updateKernelVelocities only set block.data on the host and are called every step anyway
	return LBMX_NOT_OVERRIDDEN(STATE, SimUpdate) && LBMX_NOT_OVERRIDDEN(STATE, AfterSimUpdate) && LBMX_NOT_OVERRIDDEN(STATE, computeBeforeLBMKernel)
		&& LBMX_NOT_OVERRIDDEN(STATE, computeAfterLBMKernel);
}
#undef LBMX_NOT_OVERRIDDEN
}  // namespace lbmx_host

template <typename STATE>
void execute(STATE& state)
{
	// Steps between two points where the host looks at the device travel as one batch (LBM_BLOCK::pending), unless the solver overrides a
	// hook that could observe the device in between, or LBMX_HOST_BATCH=0 asks for one launch per SimUpdate as in the reference.
	{
		const char* v = std::getenv("LBMX_HOST_BATCH");
		state.nse.defer_steps = lbmx_host::steps_can_be_deferred<STATE>() && ! (v && v[0] == '0');
	}
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
