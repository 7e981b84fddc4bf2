/* oracle_api.h -- C interface shared by the two CPU checkers of this repository.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load these libraries, and only as the checker or the reported CPU baseline.
 *
 * Two libraries export exactly this interface:
 *   oracle/liboracle_port.so    the plain C++ restatement (oracle/lbm_oracle.cpp), travels as source
 *   oracle/_ref/libref_{ab,aa}.so  the reference's own per-cell code (include/lbm3d/kernels.h:60-100
 *                                  and the trait headers it instantiates) compiled through
 *                                  oracle/ref_shim from /root/reference/include by oracle/Makefile
 *
 * Array layout is the reference's (lbm_data.h:49-67, defs.h:85-86):
 *   idx(q,x,y,z) = q*XYZ + ((x+ox)*Z + z)*Y + y,  XYZ = (X+2*ox)*Y*Z, y fastest.
 */
#ifndef LBM_ORACLE_API_H
#define LBM_ORACLE_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_D3Q27 = 0, ORC_D2Q9 = 1, ORC_D3Q19 = 2 };
enum {
	ORC_COLL_CUM = 0,
	ORC_COLL_SRT = 1,
	ORC_COLL_BGK = 2,
	ORC_COLL_MRT_LES = 3,
	ORC_COLL_CLBM = 4,			   /* D2Q9_CLBM or D3Q27_CLBM, by lattice */
	ORC_COLL_SRT_MODIF_FORCE = 5,  /* D3Q27 only from here on */
	ORC_COLL_SRT_WELL = 6,
	ORC_COLL_BGK_WELL = 7,
	ORC_COLL_CLBM_WELL = 8,
	ORC_COLL_CUM_WELL = 9,
	ORC_COLL_CUM_2017 = 10,			 /* D3Q27_CUM compiled with -DUSE_GEIER_CUM_2017 (defs.h:254) */
	ORC_COLL_CUM_ANTIALIAS = 11,	 /* ... with -DUSE_GEIER_CUM_ANTIALIAS (defs.h:255) */
	ORC_COLL_CUM_2017_ANTIALIAS = 12, /* ... with both */
	ORC_COLL_KBC_N1 = 13, /* D3Q27_KBC_N1..N4 (col_kbc_n.h), C1..C4 (col_kbc_c.h): 13..20 */
	ORC_COLL_KBC_N2 = 14,
	ORC_COLL_KBC_N3 = 15,
	ORC_COLL_KBC_N4 = 16,
	ORC_COLL_KBC_C1 = 17,
	ORC_COLL_KBC_C2 = 18,
	ORC_COLL_KBC_C3 = 19,
	ORC_COLL_KBC_C4 = 20,
	ORC_COLL_BGK_GALILEAN = 21, /* D3Q27_BGK compiled with -DUSE_GALILEAN_CORRECTION (defs.h:253, col_bgk.h:20-45) */
	ORC_COLL_CUM_HP_RHO = 22 /* D3Q27_CUM compiled with -DUSE_HIGH_PRECISION_RHO (defs.h:252, d3q27/common.h:19-29) */
};
enum { ORC_EQ_STD = 0, ORC_EQ_INV_CUM = 1, ORC_EQ_WELL = 2, ORC_EQ_ENTROPIC = 3 };
enum { ORC_STREAM_AB = 0, ORC_STREAM_AA = 1 };
enum { ORC_MACRO_VOID = 0, ORC_MACRO_DEFAULT = 1, ORC_MACRO_MEAN = 2, ORC_MACRO_WITH_MEAN_2D = 3 /* sim_2D/sim2d_2.cu:53-104, D2Q9 */ };
enum { ORC_GATE_MEANS = 1, ORC_GATE_FLUCS = 2 }; /* block.data.accumulate_means / accumulate_flucs (sim2d_2.cu:121-122) */
enum { ORC_INFLOW_NONE = 0, ORC_INFLOW_CONST = 1, ORC_INFLOW_PROFILE_YZ = 2,
	   ORC_INFLOW_PARABOLIC_Y = 3 /* sim_2D/sim2d_3.cu:36-55; inflow_vx = u_max_lbm, inflow_vy = y0, inflow_vz = inv_den */ };
enum { ORC_F32 = 0, ORC_F64 = 1 };

typedef struct oracle_desc
{
	int32_t lattice, coll, eq, streaming, macro, inflow, precision;
	int32_t nproc;	 /* the `nproc` argument of the reference kernel: 1 = wrap rule, >1 = ghost-plane rule (kernels.h:21-29) */
	int64_t X, Y, Z; /* local lattice size without overlaps */
	int64_t ox;		 /* ghost x-planes on each side (0 or 1) */
} oracle_desc;

typedef struct oracle_params
{
	double lbmViscosity;
	double fx, fy, fz;
	double inflow_vx, inflow_vy, inflow_vz;
	const void* vx_profile; /* ORC_INFLOW_PROFILE_YZ: dreal[y + z*profile_size_y] (sim_NSE/sim_2.cu:16-33) */
	int64_t profile_size_y;
	int32_t stat_counter; /* MACRO_Mean sample index (d3q27/macro.h:117) */
	int32_t macro_gates; /* ORC_GATE_* bits, ORC_MACRO_WITH_MEAN_2D only */
	const void* bouzidi_coeff; /* D2Q9 GEO_FLUID_NEAR_WALL: dreal[8][XYZ], direction order E,N,W,S,NE,NW,SW,SE, < 0 = link does not hit a wall
								  (lbm_data.h:69-83); NULL = every coefficient reads -1 */
} oracle_params;

/* 0 = ok, nonzero = combination not available in this library */
int oracle_supported(const oracle_desc* d);

/* Advance `nsteps` steps starting at iteration `iteration` (parity and A-B ping-pong follow
 * lbm.hpp:314-330: even iteration reads df_a and writes df_b; A-A uses df_a only).
 * Cell visiting order is x, z, y (state.hpp:1116-1121); nthreads>1 uses OpenMP over (x,z). */
int oracle_step(const oracle_desc* d, const oracle_params* p, void* df_a, void* df_b, void* macro, const int16_t* map, int64_t iteration,
				int32_t nsteps, int32_t nthreads);

/* Every site, ghost planes included, := EQ(rho, v) (lbm_block.hpp:219-250). */
int oracle_set_equilibrium(const oracle_desc* d, void* df, double rho, double vx, double vy, double vz);

/* Per-cell equilibrium from fields given in the a1 layout of one macro component (XYZ reals each, type double). */
int oracle_set_equilibrium_field(const oracle_desc* d, void* df, const double* rho, const double* vx, const double* vy, const double* vz);

/* rho,u of the stored DFs with the force zeroed (lbm_block.hpp:252-277). */
int oracle_initial_macro(const oracle_desc* d, const oracle_params* p, void* df, void* macro);

const char* oracle_kind(void); /* "reference" or "port" */

#ifdef __cplusplus
}
#endif
#endif
