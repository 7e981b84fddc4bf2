// reference D3Q27_CUM (include/lbm3d/d3q27/col_cum.h:14-485) on top of D3Q27_COMMON compiled with the switch of defs.h:252
// (d3q27/common.h:19-29: Kahan-summed density) -- TEST INFRASTRUCTURE ONLY.
// The class templates keep their names whatever the switch is, so this variant of common.h and the operator built on it live in a
// namespace of their own; everything common.h and col_cum.h include themselves is already in (ref_d3q27.h), their own first inclusion
// in this translation unit happens inside the namespace.
#include "ref_d3q27.h"
#define USE_HIGH_PRECISION_RHO
namespace ref_cum_hp_rho {
#include "lbm3d/d3q27/common.h"
#include "lbm3d/d3q27/col_cum.h"
}
int ref_dispatch_d3q27_cum_hprho(const RefCall& c)
{
	return c.d->coll == ORC_COLL_CUM_HP_RHO ? ref_dispatch3<ref_cum_hp_rho::D3Q27_CUM, true>(c) : -1;
}
