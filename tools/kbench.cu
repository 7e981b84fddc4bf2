// kbench.cu -- kernel-variant timing harness (development tool, not part of the product library).
// Compiled several times with different -D tuning macros (see tools/kbench.sh); runs the KB_LAT cumulant fp64 bulk kernel
// on an all-GEO_PERIODIC box and reports the average even / odd / A-B step time measured with CUDA events.
#include "../tnl_lbm_b200/csrc/kernels.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
using namespace lbmx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
#ifndef KB_REAL
#define KB_REAL double
#endif
#ifndef KB_LAT
#define KB_LAT D3Q27
#endif
#ifndef KB_KIND
#define KB_KIND K_CUM
#endif
int main(int argc, char** argv)
{
	using R = KB_REAL;
	const int S = argc > 1 ? atoi(argv[1]) : 256;
	const int iters = argc > 2 ? atoi(argv[2]) : 20;
	const int SZ = KB_LAT::NDIM == 3 ? S : 1;
	const long long XYZ = (long long) S * S * SZ;
	R *a, *b, *mac;
	int16_t* map;
	CK(cudaMalloc(&a, sizeof(R) * KB_LAT::Q * XYZ));
	CK(cudaMalloc(&b, sizeof(R) * KB_LAT::Q * XYZ));
	CK(cudaMalloc(&mac, sizeof(R) * 4 * XYZ));
	CK(cudaMalloc(&map, sizeof(int16_t) * XYZ));
	std::vector<int16_t> hm(XYZ, (int16_t) KB_LAT::PERIODIC);
	CK(cudaMemcpy(map, hm.data(), sizeof(int16_t) * XYZ, cudaMemcpyHostToDevice));
	k_set_equilibrium<KB_LAT, R><<<(unsigned) ((XYZ + 127) / 128), 128>>>(a, XYZ, XYZ, 0, 1, nullptr, nullptr, nullptr, nullptr, 1.0, 0.03, 0.01, -0.02);
	CK(cudaMemcpy(b, a, sizeof(R) * KB_LAT::Q * XYZ, cudaMemcpyDeviceToDevice));
	KParams<R> p{};
	p.cur = a; p.out = b; p.macro = mac; p.map = map; p.XYZ = XYZ; p.X = p.Y = S; p.Z = SZ; p.ox = 0; p.YZ = S * SZ; p.x_begin = 0;
	p.wrap = 1; p.eq = 1; p.out_mode = OUT_NONE; p.phys.nu = R(1e-3); p.phys.omega1 = R(1) / (R(3) * p.phys.nu + R(0.5));
	{ unsigned L = 0; while ((1u << L) < (unsigned) S) L++; p.ydiv_mul = (unsigned) ((((unsigned long long) 1 << (31 + L)) + S - 1) / S); p.ydiv_shift = L - 1; }
	auto set_bases = [&](bool aa) { for (int q = 0; q < KB_LAT::Q; q++) { p.rd[q] = p.cur + (size_t) q * XYZ; p.wr[q] = (aa ? p.cur : p.out) + (size_t) q * XYZ; } };
	set_bases(true); p.phys.fx = R(1e-6); p.phys.fy = p.phys.fz = 0;
	const int BS = LBMX_BULK_BLOCK;
	auto grid_for = [&](int cpt) { return dim3((unsigned) ((p.YZ + BS * cpt - 1) / (BS * cpt)), (unsigned) S); };
	const dim3 grid_e = grid_for(bulk_cpt<KB_LAT, R, S_AA_EVEN>()), grid_o = grid_for(bulk_cpt<KB_LAT, R, S_AA_ODD>()), grid_ab = grid_for(bulk_cpt<KB_LAT, R, S_AB>());
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float t_even = 0, t_odd = 0, t_ab = 0;
	for (int it = -4; it < 2 * iters; it++) {
		const bool even = (it & 1) == 0;
		CK(cudaEventRecord(e0));
		if (even) k_bulk<KB_LAT, KB_KIND, R, S_AA_EVEN><<<grid_e, BS>>>(p); else k_bulk<KB_LAT, KB_KIND, R, S_AA_ODD><<<grid_o, BS>>>(p);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
		if (it >= 0) (even ? t_even : t_odd) += ms;
	}
	for (int it = -2; it < iters; it++) {
		CK(cudaEventRecord(e0));
		set_bases(false);
		k_bulk<KB_LAT, KB_KIND, R, S_AB><<<grid_ab, BS>>>(p);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
		if (it >= 0) t_ab += ms;
		R* t = p.cur; p.cur = p.out; p.out = t;
	}
	CK(cudaGetLastError());
	const double bytes = (double) XYZ * KB_LAT::Q * 2 * sizeof(R);
	cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_bulk<KB_LAT, KB_KIND, R, S_AA_EVEN>);
	std::vector<R> h(8); CK(cudaMemcpy(h.data(), a, sizeof(R) * 8, cudaMemcpyDeviceToHost));
	printf("%-28s S=%d regs=%d  even %.4f ms %.0f GB/s | odd %.4f ms %.0f GB/s | AA avg %.0f MLUPS | AB %.4f ms %.0f GB/s  (f0=%.6f)\n", KB_NAME, S, fa.numRegs,
		   t_even / iters, bytes / (t_even / iters) / 1e6, t_odd / iters, bytes / (t_odd / iters) / 1e6, 2.0 * XYZ / ((t_even + t_odd) / iters) / 1e3, t_ab / iters, bytes / (t_ab / iters) / 1e6, (double) h[0]);
	return 0;
}
