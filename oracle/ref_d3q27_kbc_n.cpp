// reference D3Q27_KBC_N1..N4 (include/lbm3d/d3q27/col_kbc_n.h:254-1272) -- TEST INFRASTRUCTURE ONLY
#include "ref_d3q27.h"
#include "lbm3d/d3q27/col_kbc_n.h"
int ref_dispatch_d3q27_kbc_n(const RefCall& c)
{
	switch (c.d->coll) {
		case ORC_COLL_KBC_N1: return ref_dispatch3<D3Q27_KBC_N1, false, true>(c);
		case ORC_COLL_KBC_N2: return ref_dispatch3<D3Q27_KBC_N2, false, true>(c);
		case ORC_COLL_KBC_N3: return ref_dispatch3<D3Q27_KBC_N3, false, true>(c);
		case ORC_COLL_KBC_N4: return ref_dispatch3<D3Q27_KBC_N4, false, true>(c);
	}
	return -1;
}
