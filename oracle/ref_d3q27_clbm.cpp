// reference D3Q27_CLBM (include/lbm3d/d3q27/col_clbm.h:6-447) -- TEST INFRASTRUCTURE ONLY
#include "ref_d3q27.h"
#include "lbm3d/d3q27/col_clbm.h"
int ref_dispatch_d3q27_clbm(const RefCall& c)
{
	return c.d->coll == ORC_COLL_CLBM ? ref_dispatch3<D3Q27_CLBM, false>(c) : -1;
}
