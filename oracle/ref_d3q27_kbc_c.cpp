// reference D3Q27_KBC_C1..C4 (include/lbm3d/d3q27/col_kbc_c.h:283-1301) -- TEST INFRASTRUCTURE ONLY
#include "ref_d3q27.h"
#include "lbm3d/d3q27/col_kbc_c.h"
int ref_dispatch_d3q27_kbc_c(const RefCall& c)
{
	switch (c.d->coll) {
		case ORC_COLL_KBC_C1: return ref_dispatch3<D3Q27_KBC_C1, false, true>(c);
		case ORC_COLL_KBC_C2: return ref_dispatch3<D3Q27_KBC_C2, false, true>(c);
		case ORC_COLL_KBC_C3: return ref_dispatch3<D3Q27_KBC_C3, false, true>(c);
		case ORC_COLL_KBC_C4: return ref_dispatch3<D3Q27_KBC_C4, false, true>(c);
	}
	return -1;
}
