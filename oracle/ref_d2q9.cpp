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
