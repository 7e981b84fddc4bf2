// reference D3Q27_CUM (include/lbm3d/d3q27/col_cum.h:14-485) -- TEST INFRASTRUCTURE ONLY
#include "ref_d3q27.h"
#include "lbm3d/d3q27/col_cum.h"
int ref_dispatch_d3q27_cum(const RefCall& c)
{
	return c.d->coll == ORC_COLL_CUM ? ref_dispatch3<D3Q27_CUM, true>(c) : -1;
}
