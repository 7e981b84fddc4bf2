// reference D3Q27_SRT (include/lbm3d/d3q27/col_srt.h:16-108) -- TEST INFRASTRUCTURE ONLY
#include "ref_d3q27.h"
#include "lbm3d/d3q27/col_srt.h"
int ref_dispatch_d3q27_srt(const RefCall& c)
{
	return c.d->coll == ORC_COLL_SRT ? ref_dispatch3<D3Q27_SRT, false>(c) : -1;
}
