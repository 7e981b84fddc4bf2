// reference D3Q27_CUM (include/lbm3d/d3q27/col_cum.h:14-485) compiled with the switches of defs.h:254-255 -- TEST INFRASTRUCTURE ONLY.
// The class template keeps its name whatever the switches are, so this variant lives in a namespace of its own.
#include "ref_d3q27.h"
#include "lbm3d/d3q27/common.h"
#define USE_GEIER_CUM_2017
#define USE_GEIER_CUM_ANTIALIAS
namespace ref_cum_2017aa {
#include "lbm3d/d3q27/col_cum.h"
}
int ref_dispatch_d3q27_cum2017aa(const RefCall& c)
{
	return c.d->coll == ORC_COLL_CUM_2017_ANTIALIAS ? ref_dispatch3<ref_cum_2017aa::D3Q27_CUM, false>(c) : -1;
}
