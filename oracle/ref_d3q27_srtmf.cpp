// reference D3Q27_SRT_MODIF_FORCE (include/lbm3d/d3q27/col_srt_modif_force.h:9-120) -- TEST INFRASTRUCTURE ONLY
#include "ref_d3q27.h"
#include "lbm3d/d3q27/col_srt_modif_force.h"
int ref_dispatch_d3q27_srtmf(const RefCall& c)
{
	return c.d->coll == ORC_COLL_SRT_MODIF_FORCE ? ref_dispatch3<D3Q27_SRT_MODIF_FORCE, false>(c) : -1;
}
