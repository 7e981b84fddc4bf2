// tma_host.h -- host side of k_bulk_tma: the tensor map over one distribution array and the tile geometry.
// The driver entry point is looked up at run time (the library links the runtime statically and does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

namespace lbmx {

// cells of one row per CTA: the largest power of two that divides Y, at most 128; 0 if k_bulk_tma cannot be used on this lattice
// (a box row must be at least 128 bytes for the shared-memory alignment of the boxes, and the row pitch a multiple of 16 bytes)
inline int tma_tile_y(long long Y, int sizeof_real)
{
	int t = 1;
	while (t < 128 && Y % (2 * t) == 0)
		t *= 2;
	if ((long long) t * sizeof_real < 128 || (Y * sizeof_real) % 16 != 0)
		return 0;
	return t;
}

// 4-D tensor map (y, z, x-storage, q) over a [Q][Xs][Z][Y] array of reals, box = tile_y x 1 x 1 x 1, no swizzle, zero fill
inline bool make_df_tensor_map(CUtensorMap* tm, void* base, int sizeof_real, long long Y, long long Z, long long Xs, int Q, int tile_y, std::string* why)
{
	typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
									CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static EncodeTiled encode = nullptr;
	if (! encode) {
		void* fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || ! fn) {
			cudaGetLastError();
			if (why)
				*why = "cuTensorMapEncodeTiled is not available from this driver";
			return false;
		}
		encode = (EncodeTiled) fn;
	}
	const cuuint64_t dims[4] = {(cuuint64_t) Y, (cuuint64_t) Z, (cuuint64_t) Xs, (cuuint64_t) Q};
	const cuuint64_t strides[3] = {(cuuint64_t) (Y * sizeof_real), (cuuint64_t) (Y * Z * sizeof_real), (cuuint64_t) (Y * Z * Xs * sizeof_real)};
	const cuuint32_t box[4] = {(cuuint32_t) tile_y, 1, 1, 1};
	const cuuint32_t estr[4] = {1, 1, 1, 1};
	const CUresult r = encode(tm, sizeof_real == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr,
							  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		if (why)
			*why = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int) r);
		return false;
	}
	return true;
}

}  // namespace lbmx
