// reference D2Q9 configurations (include/lbm3d/d2q9/*.h; composition as sim_2D/sim2d_1.cu:161-178) -- TEST INFRASTRUCTURE ONLY
#include "ref_common.h"
#include "lbm_common/ciselnik.h"  // d2q9/bc.h uses the constant macros without including them (core.h pulls them in earlier)

#include "lbm3d/d2q9/macro.h"
#include "lbm3d/d2q9/bc.h"
#include "lbm3d/d2q9/eq.h"
#ifdef AA_PATTERN
	#include "lbm3d/d2q9/streaming_AA.h"
#else
	#include "lbm3d/d2q9/streaming_AB.h"
#endif
#include "lbm3d/d2q9/col_srt.h"
#include "lbm3d/d2q9/col_clbm.h"

// same gap as in 3-D: d2q9/bc.h:124 names streamingInterpRight, d2q9/streaming_AA.h has none
template <typename TRAITS>
struct RefStreaming2 : D2Q9_STREAMING<TRAITS>
{
#ifdef AA_PATTERN
	using idx = typename TRAITS::idx;
	template <typename LBM_DATA, typename LBM_KS>
	static void streamingInterpRight(LBM_DATA&, LBM_KS&, idx, idx, idx, idx, idx, idx, idx, idx, idx)
	{}
#endif
};

// The macro class sim_2D/sim2d_2.cu:53-104 defines inside the solver (a main program, so it cannot be included): three instantaneous
// channels, two velocity sums gated by DATA::accumulate_means, and three fluctuation sums about a frozen mean gated by
// DATA::accumulate_flucs.  Built on the reference's own D2Q9_MACRO_Base and driven by the reference's kernel like any other MACRO.
template <typename TRAITS>
struct Ref_MACRO2D_WithMean : D2Q9_MACRO_Base<TRAITS>
{
	using dreal = typename TRAITS::dreal;
	using idx = typename TRAITS::idx;
	enum { e_rho, e_vx, e_vy, e_svx, e_svy, e_mean_vx_frozen, e_mean_vy_frozen, e_smag_uprime, e_suprime2_sum, e_svprime2_sum, N };

	template <typename LBM_DATA, typename LBM_KS>
	static void outputMacro(LBM_DATA& SD, LBM_KS& KS, idx x, idx y, idx z)
	{
		SD.macro(e_rho, x, y, z) = KS.rho;
		SD.macro(e_vx, x, y, z) = KS.vx;
		SD.macro(e_vy, x, y, z) = KS.vy;
		if (SD.accumulate_means) {
			SD.macro(e_svx, x, y, z) += KS.vx;
			SD.macro(e_svy, x, y, z) += KS.vy;
		}
		if (SD.accumulate_flucs) {
			const dreal dux = KS.vx - SD.macro(e_mean_vx_frozen, x, y, z);
			const dreal duy = KS.vy - SD.macro(e_mean_vy_frozen, x, y, z);
			const dreal mag = sqrt(dux * dux + duy * duy);	// unqualified, as the solver writes it
			SD.macro(e_smag_uprime, x, y, z) += mag;
			SD.macro(e_suprime2_sum, x, y, z) += dux * dux;
			SD.macro(e_svprime2_sum, x, y, z) += duy * duy;
		}
	}
	template <typename LBM_DATA, typename LBM_KS>
	static void copyQuantities(LBM_DATA& SD, LBM_KS& KS, idx, idx, idx)
	{
		KS.lbmViscosity = SD.lbmViscosity;
		KS.fx = SD.fx;
		KS.fy = SD.fy;
	}
};

template <typename TRAITS, template <typename, typename> class COLLT, typename DATA, typename MACRO>
using RefCfg2 = LBM_CONFIG<TRAITS, D2Q9_KernelStruct, DATA, COLLT<TRAITS, D2Q9_EQ<TRAITS>>, D2Q9_EQ<TRAITS>, RefStreaming2<TRAITS>, D2Q9_BC_All, MACRO>;

template <typename TRAITS, template <typename, typename> class COLLT>
int ref_dispatch2_t(const RefCall& c)
{
	const int m = c.d->macro, f = c.d->inflow;
	if (m == ORC_MACRO_DEFAULT && f == ORC_INFLOW_CONST)
		return ref_invoke<RefCfg2<TRAITS, COLLT, Ref_Data2D_ConstInflow<TRAITS>, D2Q9_MACRO_Default<TRAITS>>>(c);
	if (m == ORC_MACRO_VOID && f == ORC_INFLOW_CONST)
		return ref_invoke<RefCfg2<TRAITS, COLLT, Ref_Data2D_ConstInflow<TRAITS>, D2Q9_MACRO_Void<TRAITS>>>(c);
	if (m == ORC_MACRO_MEAN && f == ORC_INFLOW_CONST)
		return ref_invoke<RefCfg2<TRAITS, COLLT, Ref_Data2D_ConstInflow<TRAITS>, D2Q9_MACRO_Mean<TRAITS>>>(c);
	if (m == ORC_MACRO_DEFAULT && f == ORC_INFLOW_NONE)
		return ref_invoke<RefCfg2<TRAITS, COLLT, Ref_Data2D_NoInflow<TRAITS>, D2Q9_MACRO_Default<TRAITS>>>(c);
	if (m == ORC_MACRO_DEFAULT && f == ORC_INFLOW_PARABOLIC_Y)
		return ref_invoke<RefCfg2<TRAITS, COLLT, Ref_Data2D_ParabolicInflow<TRAITS>, D2Q9_MACRO_Default<TRAITS>>>(c);
	if (m == ORC_MACRO_WITH_MEAN_2D && f == ORC_INFLOW_PARABOLIC_Y)  // sim_2D/sim2d_2.cu:869-875
		return ref_invoke<RefCfg2<TRAITS, COLLT, Ref_Data2D_ParabolicInflow<TRAITS>, Ref_MACRO2D_WithMean<TRAITS>>>(c);
	return -1;
}

int ref_dispatch_d2q9(const RefCall& c)
{
	if (c.d->lattice != ORC_D2Q9 || c.d->eq != ORC_EQ_STD || c.d->Z != 1)
		return -1;
#ifdef AA_PATTERN
	if (c.d->streaming != ORC_STREAM_AA)
		return -1;
#else
	if (c.d->streaming != ORC_STREAM_AB)
		return -1;
#endif
	const bool dp = c.d->precision == ORC_F64;
	if (c.d->coll == ORC_COLL_SRT)
		return dp ? ref_dispatch2_t<TraitsDP, D2Q9_SRT>(c) : ref_dispatch2_t<TraitsSP, D2Q9_SRT>(c);
	if (c.d->coll == ORC_COLL_CLBM)
		return dp ? ref_dispatch2_t<TraitsDP, D2Q9_CLBM>(c) : ref_dispatch2_t<TraitsSP, D2Q9_CLBM>(c);
	return -1;
}
