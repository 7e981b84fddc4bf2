// reference D3Q27_BGK (include/lbm3d/d3q27/col_bgk.h:16-145) -- TEST INFRASTRUCTURE ONLY
#include "ref_d3q27.h"
#include "lbm3d/d3q27/col_bgk.h"
int ref_dispatch_d3q27_bgk(const RefCall& c)
{
	return c.d->coll == ORC_COLL_BGK ? ref_dispatch3<D3Q27_BGK, false>(c) : -1;
}
