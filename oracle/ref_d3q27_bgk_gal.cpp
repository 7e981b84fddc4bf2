// reference D3Q27_BGK (include/lbm3d/d3q27/col_bgk.h:16-145) compiled with the switch of defs.h:253 -- TEST INFRASTRUCTURE ONLY.
// The class template keeps its name whatever the switch is, so this variant lives in a namespace of its own.
#include "ref_d3q27.h"
#include "lbm3d/d3q27/common.h"
#define USE_GALILEAN_CORRECTION
namespace ref_bgk_galilean {
#include "lbm3d/d3q27/col_bgk.h"
}
int ref_dispatch_d3q27_bgk_gal(const RefCall& c)
{
	return c.d->coll == ORC_COLL_BGK_GALILEAN ? ref_dispatch3<ref_bgk_galilean::D3Q27_BGK, false>(c) : -1;
}
