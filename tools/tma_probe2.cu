// tma_probe2.cu -- the CUDA C++ Programming Guide's own tensor-map example (2-D int tile through cuda::barrier and the libcu++
// cp_async_bulk_tensor wrappers), to tell "this box cannot run tensor-map TMA from a user binary" from "my PTX is wrong".
//   tma_probe2 <variant>   0: guide example   1: same, launched through cudaLaunchKernelEx with a 1x1x1 cluster   2: L2 promotion 128B + box row of 128 bytes
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("variant %d: CUDA error '%s' at line %d\n", g_v, cudaGetErrorString(e_), __LINE__); return 1;} } while (0)
static int g_v = 0;
constexpr int GW = 256, GH = 64, SW = 32, SH = 8;

__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out)
{
	__shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
	__shared__ barrier bar;
	if (threadIdx.x == 0) {
		init(&bar, blockDim.x);
		cde::fence_proxy_async_shared_cta();
	}
	__syncthreads();
	barrier::arrival_token token;
	if (threadIdx.x == 0) {
		cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
		token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
	}
	else
		token = bar.arrive();
	bar.wait(std::move(token));
	if (threadIdx.x < SW)
		out[threadIdx.x] = smem_buffer[1][threadIdx.x];
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
								CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv)
{
	g_v = argc > 1 ? atoi(argv[1]) : 0;
	std::vector<int> h(GW * GH);
	for (int i = 0; i < GW * GH; i++)
		h[i] = i;
	int *d, *out;
	CK(cudaMalloc(&d, sizeof(int) * GW * GH));
	CK(cudaMalloc(&out, sizeof(int) * 64));
	CK(cudaMemcpy(d, h.data(), sizeof(int) * GW * GH, cudaMemcpyHostToDevice));
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult qres;
	CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
	CUtensorMap tm{};
	uint64_t size[2] = {GW, GH};
	uint64_t stride[1] = {GW * sizeof(int)};
	uint32_t box[2] = {SW, SH};
	uint32_t estr[2] = {1, 1};
	CUresult r = ((EncodeTiled) fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
									g_v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		printf("variant %d: encode failed %d\n", g_v, (int) r);
		return 1;
	}
	if (g_v == 1) {
		cudaLaunchConfig_t cfg{};
		cfg.gridDim = dim3(1);
		cfg.blockDim = dim3(128);
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeClusterDimension;
		at[0].val.clusterDim.x = 1;
		at[0].val.clusterDim.y = 1;
		at[0].val.clusterDim.z = 1;
		cfg.attrs = at;
		cfg.numAttrs = 1;
		CK(cudaLaunchKernelEx(&cfg, kernel, tm, 64, 16, out));
	}
	else
		kernel<<<1, 128>>>(tm, 64, 16, out);
	CK(cudaGetLastError());
	CK(cudaDeviceSynchronize());
	int o[4];
	CK(cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost));
	printf("variant %d: row 1 of the tile = %d %d %d (expected %d %d %d)\n", g_v, o[0], o[1], o[2], 17 * GW + 64, 17 * GW + 65, 17 * GW + 66);
	return 0;
}
