// inst.cu -- one explicit instantiation family per object file (compiled 12x by build.py with different -D flags so that
// the heavy unrolled kernels build in parallel):  -DLBMX_FAMILY=d3q27_cum -DLBMX_LAT=D3Q27 -DLBMX_KIND=K_CUM -DLBMX_REAL=double
#include "kernels.cuh"

#define LBMX_CAT_(a, b) a##b
#define LBMX_CAT(a, b) LBMX_CAT_(a, b)

namespace lbmx {
bool LBMX_CAT(get_kernels_, LBMX_FAMILY)(StepKernels<LBMX_REAL>& k)
{
	k = make_step_kernels<LBMX_LAT, LBMX_KIND, LBMX_REAL>();
	return true;
}
}  // namespace lbmx
