// tma_probe.cu -- development probe: which piece of a TMA box copy does this GPU / driver reject?  One stage per process.
//   tma_probe <stage>
//   0: mbarrier only (init, arrive with expect_tx 0, try_wait)          1: 1-D bulk copy (cp.async.bulk, no descriptor)
//   2: 4-D fp64 tensor map in kernel-parameter space                    3: 4-D fp32-typed view of the same array (coordinates x2)
//   4: 2-D fp32-typed view                                               5: 4-D fp64 tensor map read from GLOBAL memory
//   6: 2-D fp64                                                          7: stage 3 + box store back
#include "../tnl_lbm_b200/csrc/kernels.cuh"
#include "../tnl_lbm_b200/csrc/tma_host.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace lbmx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("stage %d: CUDA error '%s' at line %d\n", g_stage, cudaGetErrorString(e_), __LINE__); return 1;} } while (0)
static int g_stage = 0;

__global__ void k_probe(const __grid_constant__ CUtensorMap tm, const CUtensorMap* tm_global, const double* src, double* out, int stage, int bytes)
{
	__shared__ alignas(128) double tile[128];
	__shared__ alignas(8) uint64_t bar;
	const int tid = threadIdx.x;
	if (tid == 0)
		tma::mbar_init(&bar, 1);
	__syncthreads();
	if (tid == 0) {
		const uint32_t dst = tma::smem_u32(tile), mb = tma::smem_u32(&bar);
		if (stage == 0)
			tma::mbar_expect_tx(&bar, 0);
		else {
			tma::mbar_expect_tx(&bar, (uint32_t) bytes);
			if (stage == 1)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src + 16), "r"(bytes), "r"(mb) : "memory");
			else if (stage == 2)
				asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"((uint64_t) &tm), "r"(1), "r"(1), "r"(1), "r"(1), "r"(mb) : "memory");
			else if (stage == 3 || stage == 7)
				asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"((uint64_t) &tm), "r"(2), "r"(1), "r"(1), "r"(1), "r"(mb) : "memory");
			else if (stage == 4)
				asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"((uint64_t) &tm), "r"(2), "r"(1), "r"(mb) : "memory");
			else if (stage == 5)
				asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"((uint64_t) tm_global), "r"(1), "r"(1), "r"(1), "r"(1), "r"(mb) : "memory");
			else if (stage == 6)
				asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"((uint64_t) &tm), "r"(1), "r"(1), "r"(mb) : "memory");
		}
	}
	uint32_t done = 0;
	for (int spin = 0; spin < (1 << 22) && ! done; spin++)
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(tma::smem_u32(&bar)), "r"(0) : "memory");
	if (tid == 0)
		out[200] = done ? 1.0 : -1.0;
	if (tid < 32)
		out[tid] = done ? tile[tid] : -7.0;
	if (stage == 7 && done) {
		if (tid < 32)
			tile[tid] = tile[tid] + 1000.0;
		tma::fence_generic_to_async_smem();
		__syncthreads();
		if (tid == 0) {
			asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"((uint64_t) &tm), "r"(2), "r"(1), "r"(1), "r"(1), "r"(tma::smem_u32(tile)) : "memory");
			tma::stores_commit_and_drain();
		}
	}
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
								CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
	g_stage = argc > 1 ? atoi(argv[1]) : 0;
	const int Y = 64, Z = 4, X = 3, Q = 27, TY = 32;
	const long long n = (long long) Y * Z * X * Q;
	std::vector<double> h(n);
	for (long long i = 0; i < n; i++)
		h[i] = (double) i;
	double *d, *out;
	CK(cudaMalloc(&d, n * 8));
	CK(cudaMalloc(&out, 256 * 8));
	CK(cudaMemset(out, 0, 256 * 8));
	CK(cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice));
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult qres;
	CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
	EncodeTiled encode = (EncodeTiled) fn;
	alignas(64) CUtensorMap tm;
	memset(&tm, 0, sizeof(tm));
	const cuuint32_t estr[4] = {1, 1, 1, 1};
	CUresult r = CUDA_SUCCESS;
	long long first = ((1ll * X + 1) * Z + 1) * Y + 1;	// element (y=1, z=1, x=1, q=1)
	if (g_stage == 2 || g_stage == 5) {
		const cuuint64_t dims[4] = {Y, Z, X, Q}, strides[3] = {Y * 8, Y * Z * 8, Y * Z * X * 8};
		const cuuint32_t box[4] = {TY, 1, 1, 1};
		r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	}
	else if (g_stage == 3 || g_stage == 7) {
		const cuuint64_t dims[4] = {2 * Y, Z, X, Q}, strides[3] = {Y * 8, Y * Z * 8, Y * Z * X * 8};
		const cuuint32_t box[4] = {2 * TY, 1, 1, 1};
		r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	}
	else if (g_stage == 4) {
		const cuuint64_t dims[2] = {2 * Y, (cuuint64_t) Z * X * Q}, strides[1] = {Y * 8};
		const cuuint32_t box[2] = {2 * TY, 1};
		r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		first = Y + 1;
	}
	else if (g_stage == 6) {
		const cuuint64_t dims[2] = {Y, (cuuint64_t) Z * X * Q}, strides[1] = {Y * 8};
		const cuuint32_t box[2] = {TY, 1};
		r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		first = Y + 1;
	}
	else if (g_stage == 1)
		first = 16;
	if (r != CUDA_SUCCESS) {
		printf("stage %d: encode failed with %d\n", g_stage, (int) r);
		return 1;
	}
	CUtensorMap* tm_global = nullptr;
	CK(cudaMalloc(&tm_global, sizeof(CUtensorMap)));
	CK(cudaMemcpy(tm_global, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
	printf("stage %d: descriptor words:", g_stage);
	for (int i = 0; i < 16; i++)
		printf(" %016llx", (unsigned long long) ((const uint64_t*) &tm)[i]);
	printf("\n");
	fflush(stdout);
	k_probe<<<1, 128>>>(tm, tm_global, d, out, g_stage, TY * 8);
	CK(cudaGetLastError());
	CK(cudaDeviceSynchronize());
	std::vector<double> o(256);
	CK(cudaMemcpy(o.data(), out, 256 * 8, cudaMemcpyDeviceToHost));
	printf("stage %d: wait %s; tile[0..2] = %.0f %.0f %.0f (expected %lld %lld %lld)\n", g_stage, o[200] > 0 ? "completed" : "TIMED OUT", o[0], o[1], o[2], first, first + 1, first + 2);
	if (g_stage == 7) {
		CK(cudaMemcpy(h.data(), d, n * 8, cudaMemcpyDeviceToHost));
		printf("stage %d: after store, global[first] = %.0f (expected %lld)\n", g_stage, h[first], first + 1000);
	}
	return 0;
}
