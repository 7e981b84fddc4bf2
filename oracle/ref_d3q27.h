// ref_d3q27.h -- instantiation table for the reference's D3Q27 configurations (TEST INFRASTRUCTURE ONLY).
// Mirrors the composition a solver writes in its run<TRAITS>() (sim_NSE/sim_1.cu:155-172).
#pragma once
#include "ref_common.h"

#include "lbm3d/d3q27/macro.h"
#include "lbm3d/d3q27/bc.h"
#include "lbm3d/d3q27/eq.h"
#include "lbm3d/d3q27/eq_inv_cum.h"
#include "lbm3d/d3q27/eq_entropic.h"
#ifdef AA_PATTERN
	#include "lbm3d/d3q27/streaming_AA.h"
#else
	#include "lbm3d/d3q27/streaming_AB.h"
#endif

// The reference's A-A streaming header lacks streamingInterpRight although D3Q27_BC_All names it
// unconditionally (streaming_AA.h:118-119 vs bc.h:138-139): add an empty member so that it compiles.
// GEO_OUTFLOW_RIGHT_INTERP is therefore undefined under A-A and never used by the tests.
template <typename TRAITS>
struct RefStreaming3 : D3Q27_STREAMING<TRAITS>
{
#ifdef AA_PATTERN
	using idx = typename TRAITS::idx;
	template <typename LBM_DATA, typename LBM_KS>
	static void streamingInterpRight(LBM_DATA&, LBM_KS&, idx, idx, idx, idx, idx, idx, idx, idx, idx)
	{}
#endif
};

template <typename TRAITS, template <typename, typename> class COLLT, template <typename> class EQT, typename DATA, typename MACRO>
using RefCfg3 = LBM_CONFIG<TRAITS, D3Q27_KernelStruct, DATA, COLLT<TRAITS, EQT<TRAITS>>, EQT<TRAITS>, RefStreaming3<TRAITS>, D3Q27_BC_All, MACRO>;

template <typename TRAITS, template <typename, typename> class COLLT, template <typename> class EQT, bool EXTRAS>
int ref_dispatch3_te(const RefCall& c)
{
	const int m = c.d->macro, f = c.d->inflow;
	if (m == ORC_MACRO_DEFAULT && f == ORC_INFLOW_CONST)
		return ref_invoke<RefCfg3<TRAITS, COLLT, EQT, NSE_Data_ConstInflow<TRAITS>, D3Q27_MACRO_Default<TRAITS>>>(c);
	if constexpr (EXTRAS) {
		if (m == ORC_MACRO_VOID && f == ORC_INFLOW_CONST)
			return ref_invoke<RefCfg3<TRAITS, COLLT, EQT, NSE_Data_ConstInflow<TRAITS>, D3Q27_MACRO_Void<TRAITS>>>(c);
		if (m == ORC_MACRO_MEAN && f == ORC_INFLOW_CONST)
			return ref_invoke<RefCfg3<TRAITS, COLLT, EQT, NSE_Data_ConstInflow<TRAITS>, D3Q27_MACRO_Mean<TRAITS>>>(c);
		if (m == ORC_MACRO_DEFAULT && f == ORC_INFLOW_PROFILE_YZ)
			return ref_invoke<RefCfg3<TRAITS, COLLT, EQT, Ref_Data_ProfileInflow<TRAITS>, D3Q27_MACRO_Default<TRAITS>>>(c);
		if (m == ORC_MACRO_DEFAULT && f == ORC_INFLOW_NONE)
			return ref_invoke<RefCfg3<TRAITS, COLLT, EQT, NSE_Data_NoInflow<TRAITS>, D3Q27_MACRO_Default<TRAITS>>>(c);
	}
	return -1;
}

// EXTRAS_INV: also instantiate the non-default MACRO / inflow flavours for the EQ_INV_CUM composition
template <template <typename, typename> class COLLT, bool EXTRAS_INV, bool ENTROPIC = false>
int ref_dispatch3(const RefCall& c)
{
	if (c.d->lattice != ORC_D3Q27)
		return -1;
#ifdef AA_PATTERN
	if (c.d->streaming != ORC_STREAM_AA)
		return -1;
#else
	if (c.d->streaming != ORC_STREAM_AB)
		return -1;
#endif
	const bool dp = c.d->precision == ORC_F64;
	if (c.d->eq == ORC_EQ_STD)
		return dp ? ref_dispatch3_te<TraitsDP, COLLT, D3Q27_EQ, false>(c) : ref_dispatch3_te<TraitsSP, COLLT, D3Q27_EQ, false>(c);
	if (c.d->eq == ORC_EQ_INV_CUM)
		return dp ? ref_dispatch3_te<TraitsDP, COLLT, D3Q27_EQ_INV_CUM, EXTRAS_INV>(c)
				  : ref_dispatch3_te<TraitsSP, COLLT, D3Q27_EQ_INV_CUM, EXTRAS_INV>(c);
	if constexpr (ENTROPIC) {
		if (c.d->eq == ORC_EQ_ENTROPIC)
			return dp ? ref_dispatch3_te<TraitsDP, COLLT, D3Q27_EQ_ENTROPIC, false>(c) : ref_dispatch3_te<TraitsSP, COLLT, D3Q27_EQ_ENTROPIC, false>(c);
	}
	return -1;
}
