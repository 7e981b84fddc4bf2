// tma_probe3.cu -- one factor at a time away from the programming guide's working example.
//   tma_probe3 <dtype: 0 i32 | 1 f32 | 2 f64> <rank: 2|4> <box_inner elems> <box_outer> <x coordinate (elements)> <barrier: 0 cuda::barrier | 1 raw mbarrier, count 1>
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("  -> CUDA error '%s' at line %d\n", cudaGetErrorString(e_), __LINE__); return 1;} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void kernel(const __grid_constant__ CUtensorMap tm, int rank, int x, int y, int bytes, int style, unsigned* out)
{
	__shared__ alignas(128) unsigned char smem[16384];
#pragma nv_diag_suppress static_var_with_dynamic_init
	__shared__ barrier bar;
	__shared__ alignas(8) uint64_t raw;
	if (style == 0) {
		if (threadIdx.x == 0) {
			init(&bar, blockDim.x);
			cde::fence_proxy_async_shared_cta();
		}
		__syncthreads();
		barrier::arrival_token token;
		if (threadIdx.x == 0) {
			const uint32_t mb = s32(cuda::device::barrier_native_handle(bar));
			if (rank == 2)
				asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(smem)), "l"((uint64_t) &tm), "r"(x), "r"(y), "r"(mb) : "memory");
			else
				asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(s32(smem)), "l"((uint64_t) &tm), "r"(x), "r"(y), "r"(1), "r"(1), "r"(mb) : "memory");
			token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
		}
		else
			token = bar.arrive();
		bar.wait(std::move(token));
	}
	else {
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&raw)), "r"(1) : "memory");
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&raw)), "r"(bytes) : "memory");
			if (rank == 2)
				asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(smem)), "l"((uint64_t) &tm), "r"(x), "r"(y), "r"(s32(&raw)) : "memory");
			else
				asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(s32(smem)), "l"((uint64_t) &tm), "r"(x), "r"(y), "r"(1), "r"(1), "r"(s32(&raw)) : "memory");
		}
		uint32_t done = 0;
		for (int spin = 0; spin < (1 << 22) && ! done; spin++)
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(&raw)), "r"(0) : "memory");
		if (threadIdx.x == 0)
			out[63] = done;
	}
	if (threadIdx.x < 8)
		out[threadIdx.x] = ((const unsigned*) smem)[threadIdx.x];
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
								CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv)
{
	const int dtype = atoi(argv[1]), rank = atoi(argv[2]), bi = atoi(argv[3]), bo = atoi(argv[4]), x = atoi(argv[5]), style = atoi(argv[6]);
	const int es = dtype == 2 ? 8 : 4;
	const int ROWB = 1024, ROWS = 64;  // bytes per row, rows (x 2 x 2 for the 4-D case)
	const size_t total = (size_t) ROWB * ROWS * 4;
	std::vector<unsigned> h(total / 4);
	for (size_t i = 0; i < h.size(); i++)
		h[i] = (unsigned) i;
	unsigned char* d;
	unsigned* out;
	CK(cudaMalloc(&d, total));
	CK(cudaMalloc(&out, 64 * 4));
	CK(cudaMemset(out, 0, 64 * 4));
	CK(cudaMemcpy(d, h.data(), total, cudaMemcpyHostToDevice));
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult qres;
	CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
	CUtensorMap tm{};
	const cuuint64_t dims[4] = {(cuuint64_t) (ROWB / es), (cuuint64_t) ROWS, 2, 2};
	const cuuint64_t strides[3] = {(cuuint64_t) ROWB, (cuuint64_t) ROWB * ROWS, (cuuint64_t) ROWB * ROWS * 2};
	const cuuint32_t box[4] = {(cuuint32_t) bi, (cuuint32_t) bo, 1, 1};
	const cuuint32_t estr[4] = {1, 1, 1, 1};
	const CUtensorMapDataType dt = dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : (dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
	printf("dtype=%d rank=%d box=%dx%d (%d B inner) x=%d (%d B offset) barrier=%d\n", dtype, rank, bi, bo, bi * es, x, x * es, style);
	fflush(stdout);
	CUresult r = ((EncodeTiled) fn)(&tm, dt, rank, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		printf("  -> encode failed %d\n", (int) r);
		return 1;
	}
	const int ycoord = 3;
	kernel<<<1, 128>>>(tm, rank, x, ycoord, bi * bo * es, style, out);
	CK(cudaGetLastError());
	CK(cudaDeviceSynchronize());
	unsigned o[64];
	CK(cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost));
	const unsigned expect = (unsigned) (((rank == 4 ? (size_t) ROWB * ROWS * 3 : 0) + (size_t) ycoord * ROWB + (size_t) x * es) / 4);
	printf("  -> ok: first words %u %u (expected %u %u)%s\n", o[0], o[1], expect, expect + 1, style == 1 ? (o[63] ? " wait completed" : " WAIT TIMED OUT") : "");
	return 0;
}
