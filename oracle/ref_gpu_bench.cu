// ref_gpu_bench.cu -- the REFERENCE's own CUDA kernel, cudaLBMKernel<NSE> (include/lbm3d/kernels.h:60-100), recompiled for
// sm_100a through the TNL stand-in (oracle/ref_shim) and timed on the same synthetic box as bench.py: the "recompiled baseline"
// the engine is supposed to beat (SURVEY.md §8d "second baseline").  MEASUREMENT INFRASTRUCTURE ONLY (see oracle_api.h): built by
// `make -C oracle refgpu` where /root/reference exists, into oracle/_ref/ (binary only, travels to the GPU box).
//
// Launch geometry is the reference's: threads along y, block (1,128,1) for fp64 / (1,256,1) for fp32
// (lbm_block.hpp:190-217, block_size_optimizer.h:86-98), one launch + streamSynchronize per step (state.hpp:1034-1042).
//   ref_gpu_bench_{ab,aa} [size] [steps] [f32]
#define USE_CUDA
#include "ref_d3q27.h"
#include "lbm3d/d3q27/col_cum.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

template <typename TRAITS>
int run(int S, int steps)
{
	using NSE = RefCfg3<TRAITS, D3Q27_CUM, D3Q27_EQ_INV_CUM, NSE_Data_ConstInflow<TRAITS>, D3Q27_MACRO_Default<TRAITS>>;
	using dreal = typename TRAITS::dreal;
	using idx = typename TRAITS::idx;
	using idx3d = typename TRAITS::idx3d;
	const idx XYZ = (idx) S * S * S;
	typename NSE::DATA SD;
	SD.indexer.s[0] = SD.indexer.s[1] = SD.indexer.s[2] = S;
	SD.XYZ = XYZ;
	SD.lbmViscosity = (dreal) 1e-3;
	SD.fx = (dreal) 1e-6;
	dreal* df[2] = {nullptr, nullptr};
	for (int i = 0; i < DFMAX; i++)
		CK(cudaMalloc(&df[i], sizeof(dreal) * 27 * XYZ));
	CK(cudaMalloc(&SD.dmacro, sizeof(dreal) * 4 * XYZ));
	CK(cudaMalloc(&SD.dmap, sizeof(short) * XYZ));
	{
		std::vector<short> hm((size_t) XYZ, (short) 7);	 // GEO_PERIODIC
		CK(cudaMemcpy(SD.dmap, hm.data(), sizeof(short) * XYZ, cudaMemcpyHostToDevice));
		// uniform equilibrium (rho = 1, small velocity) computed with the reference's EQ on the host for one cell, replicated
		dreal feq[27];
		struct V { dreal* p; dreal& operator()(int q, idx, idx, idx) { return p[q]; } } v{feq};
		NSE::COLL::setEquilibriumLat(v, 0, 0, 0, 1.0, 0.03, 0.01, -0.02);
		std::vector<dreal> plane((size_t) XYZ);
		for (int q = 0; q < 27; q++) {
			std::fill(plane.begin(), plane.end(), feq[q]);
			for (int i = 0; i < DFMAX; i++)
				CK(cudaMemcpy(df[i] + (size_t) q * XYZ, plane.data(), sizeof(dreal) * XYZ, cudaMemcpyHostToDevice));
		}
	}
	const int by = sizeof(dreal) == 8 ? 128 : 256;
	dim3 block(1, by, 1), grid(S, (S + by - 1) / by, S);
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	float total = 0;
	for (int it = -4; it < steps; it++) {
		SD.even_iter = ((it + 4) % 2) == 0;
		const int i = (it + 4) % DFMAX;
		for (int k = 0; k < DFMAX; k++) {
			int knew = (k - i) <= 0 ? (k - i + DFMAX) % DFMAX : k - i;
			SD.dfs[k] = df[knew];
		}
		CK(cudaEventRecord(e0));
		cudaLBMKernel<NSE><<<grid, block>>>(SD, (short) 1, idx3d(0, 0, 0), idx3d(S, S, S));
		CK(cudaEventRecord(e1));
		CK(cudaStreamSynchronize(0));
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (it >= 0)
			total += ms;
	}
	CK(cudaGetLastError());
	cudaFuncAttributes fa;
	cudaFuncGetAttributes(&fa, cudaLBMKernel<NSE>);
	const double mlups = (double) XYZ * steps / (total * 1e-3) / 1e6;
#ifdef AA_PATTERN
	const char* pat = "A-A";
#else
	const char* pat = "A-B";
#endif
	printf("{\"kernel\": \"reference cudaLBMKernel<D3Q27_CUM,EQ_INV_CUM> recompiled for sm_100a\", \"streaming\": \"%s\", \"real\": \"%s\", \"size\": %d, \"steps\": %d, "
		   "\"ms_per_step\": %.4f, \"MLUPS\": %.1f, \"GBs_algorithmic\": %.1f, \"registers\": %d, \"block\": [1, %d, 1]}\n",
		   pat, sizeof(dreal) == 8 ? "fp64" : "fp32", S, steps, total / steps, mlups, mlups * 27 * 2 * sizeof(dreal) / 1e3, fa.numRegs, by);
	return 0;
}

int main(int argc, char** argv)
{
	const int S = argc > 1 ? atoi(argv[1]) : 256;
	const int steps = argc > 2 ? atoi(argv[2]) : 20;
	const bool f32 = argc > 3 && strcmp(argv[3], "f32") == 0;
	return f32 ? run<TraitsSP>(S, steps) : run<TraitsDP>(S, steps);
}
