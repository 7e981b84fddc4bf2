// ref_common.h -- glue that drives the REFERENCE's own per-cell code on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle_api.h).  This file contains no LBM arithmetic of
// its own: it includes the reference headers read-only from /root/reference/include
// (through the TNL stand-in in oracle/ref_shim) and instantiates
//   LBMKernel<NSE>                       include/lbm3d/kernels.h:60-100
//   COLL::setEquilibriumLat              include/lbm3d/d3q27/common.h:126-158, d2q9/common.h:74-86
//   computeInitialMacro's lambda body    include/lbm3d/lbm_block.hpp:252-277
// The streaming pattern is a preprocessor choice in the reference (defs.h:3-9), hence one
// shared library per pattern.  HAVE_MPI is always defined: with zero overlaps it is identical
// to the plain build, with ox=1 it gives the ghost-plane index rule (kernels.h:39-48).
#pragma once

#if ! defined(AB_PATTERN) && ! defined(AA_PATTERN)
	#error "compile with -DAB_PATTERN or -DAA_PATTERN"
#endif
#ifndef HAVE_MPI
	#define HAVE_MPI
#endif

#include "lbm3d/defs.h"
#include "lbm3d/lbm_data.h"
#include "lbm3d/kernels.h"

#include "oracle_api.h"

#ifdef _OPENMP
	#include <omp.h>
#endif

// ---- DATA flavours that the reference defines inside its solver .cu files (not includable) ----

// sim_NSE/sim_2.cu:16-33
template <typename TRAITS>
struct Ref_Data_ProfileInflow : NSE_Data<TRAITS>
{
	using idx = typename TRAITS::idx;
	using dreal = typename TRAITS::dreal;
	dreal* vx_profile = nullptr;
	idx size_y = 0;
	template <typename LBM_KS>
	void inflow(LBM_KS& KS, idx x, idx y, idx z)
	{
		KS.vx = vx_profile[y + z * size_y];
		KS.vy = 0;
		KS.vz = 0;
	}
};

// sim_2D/sim2d_1.cu:20-35
template <typename TRAITS>
struct Ref_Data2D_ConstInflow : NSE_Data<TRAITS>
{
	using idx = typename TRAITS::idx;
	using dreal = typename TRAITS::dreal;
	dreal inflow_vx = 0;
	dreal inflow_vy = 0;
	template <typename LBM_KS>
	void inflow(LBM_KS& KS, idx x, idx y, idx z)
	{
		KS.vx = inflow_vx;
		KS.vy = inflow_vy;
	}
};

template <typename TRAITS>
struct Ref_Data2D_NoInflow : NSE_Data<TRAITS>
{
	using idx = typename TRAITS::idx;
	template <typename LBM_KS>
	void inflow(LBM_KS& KS, idx x, idx y, idx z)
	{
		KS.rho = 1;
		KS.vx = 0;
		KS.vy = 0;
	}
};

template <typename T>
inline void ref_bind_inflow(NSE_Data_ConstInflow<T>& SD, const oracle_params* p)
{
	using dreal = typename T::dreal;
	SD.inflow_vx = (dreal) p->inflow_vx;
	SD.inflow_vy = (dreal) p->inflow_vy;
	SD.inflow_vz = (dreal) p->inflow_vz;
}
template <typename T>
inline void ref_bind_inflow(NSE_Data_NoInflow<T>&, const oracle_params*)
{}
template <typename T>
inline void ref_bind_inflow(Ref_Data_ProfileInflow<T>& SD, const oracle_params* p)
{
	SD.vx_profile = (typename T::dreal*) p->vx_profile;
	SD.size_y = p->profile_size_y;
}
template <typename T>
inline void ref_bind_inflow(Ref_Data2D_ConstInflow<T>& SD, const oracle_params* p)
{
	using dreal = typename T::dreal;
	SD.inflow_vx = (dreal) p->inflow_vx;
	SD.inflow_vy = (dreal) p->inflow_vy;
}
template <typename T>
inline void ref_bind_inflow(Ref_Data2D_NoInflow<T>&, const oracle_params*)
{}

// sim_2D/sim2d_3.cu:36-55 (and sim2d_2.cu:108-133): Poiseuille profile over y computed per cell
template <typename TRAITS>
struct Ref_Data2D_ParabolicInflow : NSE_Data<TRAITS>
{
	using idx = typename TRAITS::idx;
	using dreal = typename TRAITS::dreal;
	dreal u_max_lbm = 0;
	idx y0 = 1;
	idx y1 = 1;
	dreal inv_den = 1;
	bool accumulate_means = false;	// gates read by the solver's macro class (sim2d_2.cu:121-122)
	bool accumulate_flucs = false;
	template <typename LBM_KS>
	void inflow(LBM_KS& KS, idx, idx y, idx)
	{
		dreal s = (dreal) (y - y0) * inv_den;
		if (s < 0)
			s = 0;
		else if (s > 1)
			s = 1;
		KS.vx = u_max_lbm * (4.0 * s * (1.0 - s));
		KS.vy = 0;
	}
};
template <typename T>
inline void ref_bind_inflow(Ref_Data2D_ParabolicInflow<T>& SD, const oracle_params* p)
{
	using dreal = typename T::dreal;
	SD.u_max_lbm = (dreal) p->inflow_vx;
	SD.y0 = (typename T::idx) p->inflow_vy;
	SD.inv_den = (dreal) p->inflow_vz;
	SD.accumulate_means = (p->macro_gates & ORC_GATE_MEANS) != 0;
	SD.accumulate_flucs = (p->macro_gates & ORC_GATE_FLUCS) != 0;
}

template <typename NSE>
inline typename NSE::DATA ref_make_data(const oracle_desc* d, const oracle_params* p)
{
	using dreal = typename NSE::TRAITS::dreal;
	typename NSE::DATA SD;
	SD.indexer.s[0] = d->X;
	SD.indexer.s[1] = d->Y;
	SD.indexer.s[2] = d->Z;
	SD.indexer.o[0] = d->ox;
	SD.indexer.o[1] = 0;
	SD.indexer.o[2] = 0;
	SD.XYZ = SD.indexer.getStorageSize();
	if (p) {
		SD.lbmViscosity = (dreal) p->lbmViscosity;
		SD.stat_counter = p->stat_counter;
		SD.fx = (dreal) p->fx;
		SD.fy = (dreal) p->fy;
		SD.fz = (dreal) p->fz;
		SD.bouzidi_coeff_ptr = (dreal*) p->bouzidi_coeff;
		ref_bind_inflow(SD, p);
	}
	return SD;
}

#ifndef USE_CUDA  // with USE_CUDA kernels.h declares the __global__ cudaLBMKernel instead of the per-cell host function (kernels.h:60-66)
// one or more full time steps: State::SimUpdate host branch (state.hpp:1114-1123) + LBM::updateKernelData (lbm.hpp:314-330)
template <typename NSE>
int ref_step(const oracle_desc* d, const oracle_params* p, void* df_a, void* df_b, void* macro, const int16_t* map, int64_t iteration, int nsteps,
			 int nthreads)
{
	using dreal = typename NSE::TRAITS::dreal;
	using idx = typename NSE::TRAITS::idx;
	auto SD = ref_make_data<NSE>(d, p);
	SD.dmacro = (dreal*) macro;
	SD.dmap = (typename NSE::TRAITS::map_t*) map;
	dreal* dfs[2] = {(dreal*) df_a, (dreal*) df_b};
	const short nproc = (short) d->nproc;
	const idx X = d->X, Y = d->Y, Z = d->Z;
	for (int64_t it = iteration; it < iteration + nsteps; it++) {
		SD.even_iter = (it % 2) == 0;
		const int i = (int) (it % DFMAX);
		for (int k = 0; k < DFMAX; k++) {
			int knew = (k - i) <= 0 ? (k - i + DFMAX) % DFMAX : k - i;
			SD.dfs[k] = dfs[knew];
		}
#pragma omp parallel for schedule(static) collapse(2) num_threads(nthreads)
		for (idx x = 0; x < X; x++)
			for (idx z = 0; z < Z; z++)
				for (idx y = 0; y < Y; y++)
					LBMKernel<NSE>(SD, x, y, z, nproc);
	}
	return 0;
}

#endif

template <typename dreal, typename idx>
struct RefLatView
{
	dreal* p;
	TNL::Containers::Indexer3<idx> ix;
	idx XYZ;
	dreal& operator()(int q, idx x, idx y, idx z)
	{
		return p[q * XYZ + ix.getStorageIndex(x, y, z)];
	}
};

template <typename NSE>
int ref_set_eq(const oracle_desc* d, void* df, const double* rho, const double* vx, const double* vy, const double* vz, double crho, double cvx,
			   double cvy, double cvz)
{
	using dreal = typename NSE::TRAITS::dreal;
	using idx = typename NSE::TRAITS::idx;
	auto SD = ref_make_data<NSE>(d, nullptr);
	RefLatView<dreal, idx> view{(dreal*) df, SD.indexer, SD.XYZ};
	for (idx x = -d->ox; x < d->X + d->ox; x++)
		for (idx z = 0; z < d->Z; z++)
			for (idx y = 0; y < d->Y; y++) {
				if (rho) {
					const idx i = SD.indexer.getStorageIndex(x, y, z);
					NSE::COLL::setEquilibriumLat(view, x, y, z, rho[i], vx[i], vy[i], vz ? vz[i] : 0.0);
				}
				else
					NSE::COLL::setEquilibriumLat(view, x, y, z, crho, cvx, cvy, cvz);
			}
	return 0;
}

template <typename NSE>
int ref_init_macro(const oracle_desc* d, const oracle_params* p, void* df, void* macro)
{
	using dreal = typename NSE::TRAITS::dreal;
	using idx = typename NSE::TRAITS::idx;
	auto SD = ref_make_data<NSE>(d, p);
	SD.dmacro = (dreal*) macro;
	SD.dfs[df_cur] = (dreal*) df;
	for (idx x = 0; x < d->X; x++)
		for (idx z = 0; z < d->Z; z++)
			for (idx y = 0; y < d->Y; y++) {
				typename NSE::template KernelStruct<dreal> KS;
				for (int i = 0; i < NSE::Q; i++)
					KS.f[i] = SD.df(df_cur, i, x, y, z);
				NSE::MACRO::copyQuantities(SD, KS, x, y, z);
				NSE::MACRO::zeroForcesInKS(KS);
				NSE::COLL::computeDensityAndVelocity(KS);
				NSE::MACRO::outputMacro(SD, KS, x, y, z);
			}
	return 0;
}

// operation selector passed through the per-family dispatchers
struct RefCall
{
	int op;	 // 0 = query, 1 = step, 2 = set_eq, 3 = init_macro
	const oracle_desc* d;
	const oracle_params* p;
	void *df_a, *df_b, *macro;
	const int16_t* map;
	int64_t iteration;
	int nsteps, nthreads;
	const double *rho, *vx, *vy, *vz;
	double crho, cvx, cvy, cvz;
};

template <typename NSE>
int ref_invoke(const RefCall& c)
{
	switch (c.op) {
		case 0:
			return 0;
#ifndef USE_CUDA
		case 1:
			return ref_step<NSE>(c.d, c.p, c.df_a, c.df_b, c.macro, c.map, c.iteration, c.nsteps, c.nthreads);
#endif
		case 2:
			return ref_set_eq<NSE>(c.d, c.df_a, c.rho, c.vx, c.vy, c.vz, c.crho, c.cvx, c.cvy, c.cvz);
		case 3:
			return ref_init_macro<NSE>(c.d, c.p, c.df_a, c.macro);
	}
	return -2;
}

// per-family dispatchers (one translation unit each, so `make -j` parallelises the heavy instantiations);
// each returns -1 when the descriptor is not one of its combinations
int ref_dispatch_d3q27_cum(const RefCall& c);
int ref_dispatch_d3q27_srt(const RefCall& c);
int ref_dispatch_d3q27_bgk(const RefCall& c);
int ref_dispatch_d3q27_bgk_gal(const RefCall& c);
int ref_dispatch_d3q27_cum_hprho(const RefCall& c);
int ref_dispatch_d3q27_mrt(const RefCall& c);
int ref_dispatch_d3q27_clbm(const RefCall& c);
int ref_dispatch_d3q27_srtmf(const RefCall& c);
int ref_dispatch_d3q27_cum2017(const RefCall& c);
int ref_dispatch_d3q27_cumaa(const RefCall& c);
int ref_dispatch_d3q27_cum2017aa(const RefCall& c);
int ref_dispatch_d3q27_kbc_n(const RefCall& c);
int ref_dispatch_d3q27_kbc_c(const RefCall& c);
int ref_dispatch_d2q9(const RefCall& c);
