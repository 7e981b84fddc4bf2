// reference D3Q27_MRT ("MRT_LES", include/lbm3d/d3q27/col_mrt.h:13-141) -- TEST INFRASTRUCTURE ONLY
#include "ref_d3q27.h"
#include "lbm3d/d3q27/col_mrt.h"
int ref_dispatch_d3q27_mrt(const RefCall& c)
{
	return c.d->coll == ORC_COLL_MRT_LES ? ref_dispatch3<D3Q27_MRT, false>(c) : -1;
}
