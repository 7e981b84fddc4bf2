/* lbmx.h -- C ABI of the B200-native lattice-Boltzmann time-stepping engine (liblbmx.so).
 *
 * This is the drop-in boundary for the one hot path of TNL-LBM (buresjan/tnl-lbm): the fused
 * collide-and-stream update over a uniform structure-of-arrays lattice.  The reference has no FFI of
 * its own; its seam is the trivially-copyable `block.data` POD handed by value to cudaLBMKernel
 * (include/lbm3d/lbm_data.h:7-131, launch site include/lbm3d/state.hpp:1034-1040) plus the
 * LBM_BLOCK / LBM methods that prepare it.  Every entry point below names the reference interface
 * it replaces (paths relative to the reference's include/lbm3d/ unless stated).
 *
 * Conventions: C linkage, opaque handle, plain pointers and sizes, `int` status (0 = LBMX_OK),
 * lbmx_last_error() for the message of the last failure on the calling thread.  A handle is not
 * thread-safe: one driver thread per engine (the reference drives one rank from one thread too).
 * There is no CPU fallback: every call that needs the GPU fails with LBMX_ERR_CUDA without one.
 *
 * Host array layout at this boundary is the reference's (lbm_data.h:49-57, defs.h:85-86):
 *     cell(x,y,z)   = (x*Z + z)*Y + y                     y fastest, then z, then x
 *     df(q,x,y,z)   = q*XYZ + cell(x,y,z)                 structure of arrays over q
 * for the LOCAL slab of this rank.  With `with_ghosts` = 1 the x extent is X_local + 2*ghost_x and
 * plane 0 is the left ghost plane (the reference's storage including overlaps, as its checkpoints
 * save it: checkpoint.h:58-101).
 */
#ifndef LBMX_H
#define LBMX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBMX_VERSION 100

enum lbmx_status { LBMX_OK = 0, LBMX_ERR_ARG = 1, LBMX_ERR_UNSUPPORTED = 2, LBMX_ERR_CUDA = 3, LBMX_ERR_NCCL = 4, LBMX_ERR_STATE = 5 };

/* lattice / trait selectors: the template arguments of LBM_CONFIG (defs.h:169-250) as enums */
enum lbmx_lattice {
	LBMX_D3Q27 = 0,
	LBMX_D2Q9 = 1,
	LBMX_D3Q19 = 2 /* directions 0..18 of the D3Q27 numbering (defs.h:273-295), SRT and MRT_LES only; the reference has no such
					  lattice (BASELINE.json names it): parity unpinned by construction */
};
enum lbmx_coll {
	LBMX_COLL_CUM = 0,	   /* D3Q27_CUM      d3q27/col_cum.h:14-485 */
	LBMX_COLL_SRT = 1,	   /* D3Q27_SRT      d3q27/col_srt.h:16-108   | D2Q9_SRT  d2q9/col_srt.h:16-44 */
	LBMX_COLL_BGK = 2,	   /* D3Q27_BGK      d3q27/col_bgk.h:16-145 */
	LBMX_COLL_MRT_LES = 3, /* D3Q27_MRT      d3q27/col_mrt.h:13-141 */
	LBMX_COLL_CLBM = 4,	   /* D2Q9_CLBM      d2q9/col_clbm.h:13-89    | D3Q27_CLBM d3q27/col_clbm.h:6-447 (by lattice) */
	LBMX_COLL_SRT_MODIF_FORCE = 5, /* D3Q27_SRT_MODIF_FORCE  d3q27/col_srt_modif_force.h:9-120 */
	/* 6-9: the *_WELL operators -- not instantiable in the reference itself (D3Q27_COMMON_WELL lacks setEquilibriumLat /
	 * setEquilibriumDecomposition, which lbm_block.hpp:243 and d3q27/bc.h:141 require), so there is no behaviour to match */
	LBMX_COLL_CUM_2017 = 10,		   /* D3Q27_CUM built with -DUSE_GEIER_CUM_2017 (defs.h:254; col_cum.h:177-208,258-276) */
	LBMX_COLL_CUM_ANTIALIAS = 11,	   /* D3Q27_CUM built with -DUSE_GEIER_CUM_ANTIALIAS (defs.h:255; col_cum.h:215-229) */
	LBMX_COLL_CUM_2017_ANTIALIAS = 12, /* both switches */
	LBMX_COLL_KBC_N1 = 13,			   /* D3Q27_KBC_N1..N4  d3q27/col_kbc_n.h:254-1272 (shear part from raw moments: D, D+T, D+Q, D+T+Q) */
	LBMX_COLL_KBC_N2 = 14,
	LBMX_COLL_KBC_N3 = 15,
	LBMX_COLL_KBC_N4 = 16,
	LBMX_COLL_KBC_C1 = 17,			   /* D3Q27_KBC_C1..C4  d3q27/col_kbc_c.h:283-1301 (shear part from central moments) */
	LBMX_COLL_KBC_C2 = 18,
	LBMX_COLL_KBC_C3 = 19,
	LBMX_COLL_KBC_C4 = 20,
	LBMX_COLL_BGK_GALILEAN = 21,	   /* D3Q27_BGK built with -DUSE_GALILEAN_CORRECTION (defs.h:253; col_bgk.h:20-45) */
	LBMX_COLL_CUM_HP_RHO = 22		   /* D3Q27_CUM built with -DUSE_HIGH_PRECISION_RHO (defs.h:252; d3q27/common.h:19-29: the density is a Kahan sum over the 27 populations) */
};
enum lbmx_eq {
	LBMX_EQ_STD = 0,	  /* D3Q27_EQ eq.h:8-130, D2Q9_EQ */
	LBMX_EQ_INV_CUM = 1,  /* D3Q27_EQ_INV_CUM eq_inv_cum.h:13-137 */
	/* 2: D3Q27_EQ_WELL belongs to the *_WELL operators (not instantiable in the reference, see above) */
	LBMX_EQ_ENTROPIC = 3  /* D3Q27_EQ_ENTROPIC eq_entropic.h:11-211; initialisation and boundary cells of the KBC operators */
};
enum lbmx_streaming { LBMX_STREAM_AB = 0 /* streaming_AB.h */, LBMX_STREAM_AA = 1 /* streaming_AA.h */ };
enum lbmx_macro {
	LBMX_MACRO_VOID = 0,
	LBMX_MACRO_DEFAULT = 1,
	LBMX_MACRO_MEAN = 2, /* d3q27/macro.h:50-188, d2q9/macro.h */
	/* D2Q9 only, 10 channels: rho, vx, vy, sum vx, sum vy, frozen <vx>, frozen <vy>, sum |u'|, sum u'^2, sum v'^2
	 * (D2Q9_MACRO_WithMean, sim_2D/sim2d_2.cu:53-104); the sums advance only while the gates in lbmx_params.macro_gates are set */
	LBMX_MACRO_WITH_MEAN_2D = 3
};
enum lbmx_inflow {
	LBMX_INFLOW_NONE = 0,	  /* NSE_Data_NoInflow      lbm_data.h:117-131 */
	LBMX_INFLOW_CONST = 1,	  /* NSE_Data_ConstInflow   lbm_data.h:98-115, NSE2D_Data_ConstInflow sim_2D/sim2d_1.cu:20-35 */
	LBMX_INFLOW_PROFILE_YZ = 2, /* NSE_Data_XProfileInflow sim_NSE/sim_2.cu:16-33 */
	LBMX_INFLOW_PARABOLIC_Y = 3 /* NSE2D_Data_ParabolicInflow sim_2D/sim2d_3.cu:36-55 (D2Q9): vx = u_max * 4 s (1 - s), s = clamp((y - y0) * inv_den, 0, 1);
								  lbmx_params carries u_max in inflow_vx, y0 in inflow_vy and inv_den in inflow_vz */
};
enum lbmx_precision { LBMX_F32 = 0 /* TraitsSP */, LBMX_F64 = 1 /* TraitsDP */ }; /* defs.h:118-119 */

/* when the per-cell macroscopic fields are written */
enum lbmx_macro_policy {
	LBMX_MACRO_EVERY_STEP = 0, /* as the reference: every step (d3q27/macro.h:64-71) */
	LBMX_MACRO_LAST_STEP = 1,  /* only by the last step of each lbmx_step() batch: same values at every point where the
								  host can observe them (the reference copies them out on output cadence only, state.hpp:1134-1142) */
	LBMX_MACRO_NEVER = 2
};

typedef struct lbmx_desc
{
	int32_t lattice, coll, eq, streaming, macro, inflow, precision;
	int32_t macro_policy;
	int64_t X, Y, Z;  /* GLOBAL lattice size (D2Q9: Z = 1) */
	int32_t rank;	  /* this process' x-slab ...                                       (lattice_decomposition.h:16-55) */
	int32_t nranks;	  /* ... out of nranks contiguous x-slabs, one GPU each */
	int32_t device;	  /* CUDA device ordinal, -1 = current */
	int32_t ghost_x;  /* 1: one ghost x-plane per side and the reference's nproc>1 index rule (kernels.h:21-29,39-48);
						 forced to 1 when nranks > 1; with nranks == 1 the exchange is a periodic self-exchange */
	int32_t periodic_x; /* slab 0 and slab nranks-1 are neighbours (State ctor `periodic_lattice`, lattice_decomposition.h:148-162) */
	int32_t flags;		/* LBMX_FLAG_* */
	int32_t reserved[2];
} lbmx_desc;

/* "Parity arithmetic": kernels that evaluate every expression in the reference's floating-point association, with true divisions
 * and without FMA contraction (tnl_lbm_b200/csrc/collide_strict.cuh, nvcc -fmad=false).  With this flag the engine reproduces the
 * reference's strict-IEEE CPU build BIT FOR BIT in fp32 and fp64 (tests/test_gpu_parity.py: every golden case and 1000-step runs).
 * Why it exists: in fp32 the reference is not reproducible to the 1e-5 tolerance against itself across compilers' contraction
 * choices (2e-5 in velocity after 1000 steps between its strict and its FMA-contracted CPU build), so "within tolerance of the
 * reference" needs a pinned arithmetic to be checkable.  The default kernels (one reciprocal, pruned transforms, FMA) stay within
 * 1e-12 (fp64) of it and are faster.  D3Q27 and D2Q9 only: D3Q19 has no reference. */
#define LBMX_FLAG_STRICT_ARITH 1

/* per-step scalars: the non-pointer members of block.data (lbm_data.h:12-30,87-115), set by
 * State::updateKernelData (state.hpp:1314-1321) and the solver's updateKernelVelocities() */
typedef struct lbmx_params
{
	double lbmViscosity;
	double fx, fy, fz;
	double inflow_vx, inflow_vy, inflow_vz;
	int32_t stat_counter; /* MACRO_Mean sample index; lbmx_step(n>1) increments it per step */
	int32_t macro_gates;  /* LBMX_MACRO_WITH_MEAN_2D: LBMX_GATE_* bits = block.data.accumulate_means / accumulate_flucs
							 (sim_2D/sim2d_2.cu:121-122); 0 otherwise */
} lbmx_params;
#define LBMX_GATE_MEANS 1
#define LBMX_GATE_FLUCS 2

typedef struct lbmx_layout
{
	int64_t X_local, Y, Z; /* slab size without ghost planes */
	int64_t x_offset;	   /* global x of local plane 0 */
	int64_t ghost_x;	   /* ghost planes per side */
	int64_t XYZ;		   /* storage cells per component = (X_local + 2*ghost_x)*Y*Z */
	int32_t Q, n_macro, sizeof_real, dfmax; /* dfmax: 2 for A-B, 1 for A-A (defs.h:40-63) */
} lbmx_layout;

/* raw device pointers mirroring block.data for code that wants them (checkpoint / writers): lbm_data.h:24-30 */
typedef struct lbmx_ptrs
{
	void* dfs[2]; /* dfs[0] = df_cur, dfs[1] = df_out of the NEXT step (A-A: dfs[1] = NULL) */
	void* dmacro;
	int16_t* dmap;
	int32_t even_iter;
	int32_t reserved;
} lbmx_ptrs;

typedef struct lbmx_engine lbmx_engine;

const char* lbmx_last_error(void);
int lbmx_version(void);

/* Host-only helpers (no GPU needed) -------------------------------------------------------------------------------------- */

/* 1-D slab decomposition along x; replaces decomposeLattice_D1Q3 (lattice_decomposition.h:16-55) */
int lbmx_decompose_x(int64_t X, int32_t nranks, int32_t rank, int64_t* x_offset, int64_t* x_local);

/* Populations that cross an x-face: the +x movers go right, the -x movers go left (df_sync_directions, defs.h:309-340).
 * Writes up to 9 direction indices into each list, returns the count. */
int lbmx_halo_directions(int32_t lattice, int32_t* to_right, int32_t* to_left);

/* One halo transfer of the exchange that follows the step at `iteration` (replaces LBM_BLOCK::startDrealArraySynchronization,
 * lbm_block.hpp:410-451, and LBM::synchronizeDFsAndMacroDevice, lbm.hpp:196-280).  Planes are storage x indices of the local
 * slab including ghosts (0 = left ghost, ghost_x .. ghost_x+X_local-1 = interior). */
typedef struct lbmx_halo_msg
{
	int32_t to_right;	/* 1: goes to the right neighbour, 0: to the left neighbour */
	int32_t n_dirs;		/* number of populations in dirs[] */
	int32_t dirs[9];	/* population (q) indices, the same slot on both sides */
	int64_t src_plane;	/* storage x-plane read on the sender */
	int64_t dst_plane;	/* storage x-plane written on the receiver */
} lbmx_halo_msg;
int lbmx_halo_plan(int32_t lattice, int32_t streaming, int64_t iteration, int64_t X_local, lbmx_halo_msg msgs[2]);

/* Engine life cycle ------------------------------------------------------------------------------------------------------ */

/* replaces LBM ctor + LBM_BLOCK::allocateDeviceData (lbm.hpp:6-22, lbm_block.hpp:525-595) */
/* number of CUDA devices this process sees (a launcher that starts one process per GPU picks device = local rank % count) */
int lbmx_device_count(int32_t* count);
int lbmx_create(const lbmx_desc* desc, lbmx_engine** out);
int lbmx_destroy(lbmx_engine* e);
int lbmx_get_layout(const lbmx_engine* e, lbmx_layout* out);

/* Halo transport.  After lbmx_comm_init the ranks try to map each other's distribution arrays (CUDA IPC): if every rank succeeds, the
 * exchange of a step is two copy kernels that store the 9 crossing populations of a boundary plane directly into the neighbour's
 * ghost (or, on A-A odd steps, boundary) plane over NVLink, each followed by a bump of an arrival counter in the neighbour's memory;
 * the neighbour's edge-plane kernels of the next step poll that counter.  Otherwise (or with LBMX_HALO=nccl) the exchange is
 * ncclSend/ncclRecv, 9 plane messages per direction in one group.  Both are bit-identical. */
/* Multi-GPU: rank 0 makes an id (128 bytes), the host distributes it (any broadcast the launcher offers), every rank joins.
 * Replaces the MPI communicator of DistributedNDArraySynchronizer.  NCCL is bound at run time (dlopen "libnccl.so.2"). */
int lbmx_comm_unique_id(void* id128);
int lbmx_comm_init(lbmx_engine* e, const void* id128);

/* State upload / download ------------------------------------------------------------------------------------------------- */

/* LBM_BLOCK::copyMapToDevice (+ LBM::synchronizeMapDevice for the ghost planes): lbm_block.hpp:344-350,462-473.
 * Also classifies the cells for the launch plan (bulk kernel vs. boundary list). */
int lbmx_map_upload(lbmx_engine* e, const int16_t* host_map, int with_ghosts);
int lbmx_map_download(lbmx_engine* e, int16_t* host_map, int with_ghosts);

/* LBM_BLOCK::setEquilibrium: every site incl. ghost planes, every DF copy (lbm_block.hpp:219-250) */
int lbmx_df_set_equilibrium(lbmx_engine* e, double rho, double vx, double vy, double vz);
/* per-cell variant (forLocalLatticeSites + setEquilibriumLat, common.h:126-158): double[X_local*Z*Y] fields, vz may be NULL */
int lbmx_df_set_equilibrium_field(lbmx_engine* e, const double* rho, const double* vx, const double* vy, const double* vz);
/* LBM_BLOCK::copyDFsToDevice / copyDFsToHost (lbm_block.hpp:352-376); host type = the engine's precision.
 * which = 0: the array the next step reads (df_cur), 1: the other A-B copy. */
int lbmx_df_upload(lbmx_engine* e, int which, const void* host_df, int with_ghosts);
int lbmx_df_download(lbmx_engine* e, int which, void* host_df, int with_ghosts);
/* fill the ghost planes of df_cur from the neighbours' boundary planes, all Q populations (SimInit's first
 * synchronizeDFsAndMacroDevice(df_cur), state.hpp:966-972) */
int lbmx_df_sync_ghosts(lbmx_engine* e);

/* LBM_BLOCK::computeInitialMacro (lbm_block.hpp:252-277) */
int lbmx_macro_init(lbmx_engine* e);
/* LBM_BLOCK::copyMacroToHost / copyMacroToDevice (lbm_block.hpp:378-392): [n_macro][X_local*Z*Y] */
int lbmx_macro_download(lbmx_engine* e, void* host_macro, int with_ghosts);
int lbmx_macro_upload(lbmx_engine* e, const void* host_macro, int with_ghosts);

int lbmx_set_params(lbmx_engine* e, const lbmx_params* p);
/* NSE_Data_XProfileInflow::vx_profile (sim_NSE/sim_2.cu:16-33): real[size_y*size_z] in the engine's precision, read as
 * profile[y + z*size_y]; size_y >= Y and size_z >= Z (LBMX_ERR_ARG otherwise) */
int lbmx_set_inflow_profile(lbmx_engine* e, const void* host_profile, int64_t size_y, int64_t size_z);

/* D2Q9 GEO_FLUID_NEAR_WALL (Bouzidi interpolated bounce-back, d2q9/bc.h:61-87,140-167; A-B streaming only): the coefficient array
 * of LBM_BLOCK::allocateBouzidiCoeffArrays / block.data.bouzidi_coeff_ptr (lbm_data.h:69-83): real[8][X_local*Z*Y], direction
 * order E,N,W,S,NE,NW,SW,SE, negative = the link does not hit a wall.  Without it every coefficient reads -1. */
int lbmx_bouzidi_upload(lbmx_engine* e, const void* host_coeff);

/* Time stepping ------------------------------------------------------------------------------------------------------------ */

/* State::SimUpdate + LBM::updateKernelData (state.hpp:980-1145, lbm.hpp:314-330): advances `nsteps` iterations.
 * Handles even/odd parity, the A-B pointer rotation, boundary-planes-first ordering and the halo exchange overlapped
 * with the interior update.  Asynchronous: returns after enqueueing; lbmx_sync() waits. */
int lbmx_step(lbmx_engine* e, int64_t nsteps);
int lbmx_sync(lbmx_engine* e);
/* same, bracketed by CUDA events on the engine's own compute stream; synchronises; elapsed device time in ms */
int lbmx_step_timed(lbmx_engine* e, int64_t nsteps, float* elapsed_ms);

/* Measurement aid: repeat the halo exchange of the last completed step `reps` times, ALONE on the communication stream, and
 * return the device time per exchange (CUDA events).  Idempotent on the data; collective over the ranks of the communicator.
 * Reported by bench.py next to the NVLink-bound time of the same bytes. */
int lbmx_halo_time(lbmx_engine* e, int32_t reps, float* ms_per_exchange);
int lbmx_get_iterations(const lbmx_engine* e, int64_t* it);
int lbmx_set_iterations(lbmx_engine* e, int64_t it); /* checkpoint restore: parity travels with the raw arrays */

/* the NaN scan of State::AfterSimUpdate (state.hpp:1166-1188): OR over rho != rho of the macro array */
int lbmx_has_nan(lbmx_engine* e, int32_t* flag);

/* escape hatch mirroring block.data (lbm_block.hpp:583-593).  The engine works on its own non-blocking streams: call lbmx_sync()
 * before touching these arrays from other streams (the legacy default stream included), and finish that work before the next
 * lbmx_step / upload / download. */
int lbmx_get_device_ptrs(lbmx_engine* e, lbmx_ptrs* out);

/* Introspection for the measurement harness -------------------------------------------------------------------------------- */
typedef struct lbmx_stats
{
	int64_t kernel_launches;   /* kernels of this library launched since creation */
	int64_t halo_bytes_sent;   /* bytes handed to NCCL / copied for the self-exchange */
	int64_t boundary_cells;	   /* cells handled by the boundary-list kernel */
	int64_t bulk_cells;		   /* cells handled by the bulk kernel */
	int32_t bulk_regs, boundary_regs; /* registers per thread of the two step kernels (cudaFuncGetAttributes) */
	int32_t bulk_block;
	int32_t halo_peer_memory; /* 1: halos are stored straight into the neighbours' arrays over NVLink (CUDA IPC peer mappings, arrival counters);
								 0: NCCL send/recv (multi-node, IPC unavailable, or LBMX_HALO=nccl in the environment) */
	int64_t aa_cells_reaching_outside; /* A-A only, set by lbmx_map_upload: cells on a lattice face (y, z; x too on a slab without ghost planes)
										  that are neither GEO_NOTHING nor wrapped.  The A-A index rule is unclamped (kernels.h:30-37), so their
										  neighbours lie outside the lattice: undefined behaviour in the reference.  Here the arrays carry a zeroed
										  guard band, so the accesses stay inside the engine's memory, but the values at those cells are as
										  meaningless as there -- give A-A lattices a GEO_NOTHING (or periodic) skin, as sim_2.cu:125-138 does. */
	int64_t tma_launches;	   /* of kernel_launches: bulk kernels whose populations travelled as bulk copies of the TMA engine (k_bulk_tma, A-A only) */
} lbmx_stats;
int lbmx_get_stats(lbmx_engine* e, lbmx_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* LBMX_H */
