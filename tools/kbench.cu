// kbench.cu -- kernel-variant timing harness (development tool, not part of the product library).
// Compiled several times with different -D tuning macros (see tools/kbench.sh); runs the KB_LAT / KB_KIND bulk kernels on an
// all-GEO_PERIODIC box and reports the average even / odd / A-B step time measured with CUDA events, for the plain kernels and for the
// TMA kernels (k_bulk_tma), with and without the macroscopic output, and checks that both kernel flavours leave identical bits.
//   kbench [S] [iters] [macro: 0 none | 1 rho,u every step | 2 MACRO_Mean] [field: 0 uniform | 1 the bench's sinusoidal field] [chain: 0 | 1]
//   chain = 1 adds a line for 2*iters A-A steps enqueued back to back (no host synchronisation between kernels), timed as one region --
//   what lbmx_step() does; the per-kernel figures above it are taken with a synchronisation after every kernel.
#include "../tnl_lbm_b200/csrc/kernels.cuh"
#include "../tnl_lbm_b200/csrc/tma_host.h"
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
#ifndef KB_NAME
#define KB_NAME "default"
#endif

template <typename R>
__global__ void k_count_diff(const R* a, const R* b, long long n, unsigned long long* cnt)
{
	unsigned long long local = 0;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
		local += (a[i] != b[i]) || (a[i] != a[i]);
	if (local)
		atomicAdd(cnt, local);
}

__global__ void k_bench_field(double* rho, double* vx, double* vy, double* vz, int S, int SZ)
{
	// bench.py initial_fields(): rho = 1 + 0.01 sin(2 pi x / X), vx = 0.05 sin(2 pi y / Y), vy = 0.02 cos(2 pi z / Z), vz = 0.01
	const long long n = (long long) S * S * SZ;
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		const int y = (int) (i % S), z = (int) ((i / S) % SZ), x = (int) (i / ((long long) S * SZ));
		rho[i] = 1.0 + 0.01 * sin(2.0 * 3.14159265358979323846 * x / S);
		vx[i] = 0.05 * sin(2.0 * 3.14159265358979323846 * y / S);
		vy[i] = 0.02 * cos(2.0 * 3.14159265358979323846 * z / SZ);
		vz[i] = 0.01;
	}
}

int main(int argc, char** argv)
{
	using R = KB_REAL;
	using L = KB_LAT;
	const int S = argc > 1 ? atoi(argv[1]) : 256;
	const int iters = argc > 2 ? atoi(argv[2]) : 20;
	const int macro = argc > 3 ? atoi(argv[3]) : 0;
	const int field = argc > 4 ? atoi(argv[4]) : 0;
	const int chain = argc > 5 ? atoi(argv[5]) : 0;
	const int SZ = L::NDIM == 3 ? S : 1;
	const long long XYZ = (long long) S * S * SZ;
	const int NM = macro == 2 ? (L::NDIM == 3 ? 13 : 8) : 4;
	R *a, *b, *mac, *ref;
	int16_t* map;
	unsigned long long* d_cnt;
	CK(cudaMalloc(&a, sizeof(R) * L::Q * XYZ));
	CK(cudaMalloc(&b, sizeof(R) * L::Q * XYZ));
	CK(cudaMalloc(&mac, sizeof(R) * NM * XYZ));
	CK(cudaMemset(mac, 0, sizeof(R) * NM * XYZ));
	CK(cudaMalloc(&map, sizeof(int16_t) * XYZ));
	CK(cudaMalloc(&d_cnt, 8));
	std::vector<int16_t> hm(XYZ, (int16_t) L::PERIODIC);
	CK(cudaMemcpy(map, hm.data(), sizeof(int16_t) * XYZ, cudaMemcpyHostToDevice));
	auto init = [&]() {
		if (field) {
			double* fld;
			CK(cudaMalloc(&fld, sizeof(double) * 4 * XYZ));
			k_bench_field<<<1184, 256>>>(fld, fld + XYZ, fld + 2 * XYZ, fld + 3 * XYZ, S, SZ);
			k_set_equilibrium<L, R><<<(unsigned) ((XYZ + 127) / 128), 128>>>(a, XYZ, XYZ, 0, L::Q == 27 ? 1 : 0, fld, fld + XYZ, fld + 2 * XYZ, fld + 3 * XYZ, 1.0, 0.0, 0.0, 0.0);
			CK(cudaDeviceSynchronize());
			CK(cudaFree(fld));
		}
		else
			k_set_equilibrium<L, R><<<(unsigned) ((XYZ + 127) / 128), 128>>>(a, XYZ, XYZ, 0, L::Q == 27 ? 1 : 0, nullptr, nullptr, nullptr, nullptr, 1.0, 0.03, 0.01, L::NDIM == 3 ? -0.02 : 0.0);
		CK(cudaDeviceSynchronize());
	};
	init();
	CK(cudaMemcpy(b, a, sizeof(R) * L::Q * XYZ, cudaMemcpyDeviceToDevice));
	KParams<R> p{};
	p.cur = a; p.out = b; p.macro = mac; p.map = map; p.XYZ = XYZ; p.X = p.Y = S; p.Z = SZ; p.ox = 0; p.YZ = S * SZ; p.x_begin = 0;
	p.wrap = 1; p.eq = L::Q == 27 ? 1 : 0; p.out_mode = macro == 0 ? OUT_NONE : (macro == 1 ? OUT_DEFAULT : OUT_MEAN); p.phys.nu = R(1e-3); set_rates(p.phys);
	{ unsigned Lg = 0; while ((1u << Lg) < (unsigned) S) Lg++; p.ydiv_mul = (unsigned) ((((unsigned long long) 1 << (31 + Lg)) + S - 1) / S); p.ydiv_shift = Lg - 1; }
	auto set_bases = [&](bool aa) { for (int q = 0; q < L::Q; q++) { p.rd[q] = p.cur + (size_t) q * XYZ; p.wr[q] = (aa ? p.cur : p.out) + (size_t) q * XYZ; } };
	set_bases(true); p.phys.fx = R(1e-6); p.phys.fy = p.phys.fz = 0;
	const int BS = LBMX_BULK_BLOCK;
	auto grid_for = [&](int cpt) { return dim3((unsigned) ((p.YZ + BS * cpt - 1) / (BS * cpt)), (unsigned) S); };
	const dim3 grid_e = grid_for(bulk_cpt<L, R, S_AA_EVEN>()), grid_o = grid_for(bulk_cpt<L, R, S_AA_ODD>()), grid_ab = grid_for(bulk_cpt<L, R, S_AB>());
	// TMA flavour
	const int ty = tma_tile_y(S, (int) sizeof(R));
	const bool have_tma = ty > 0;
	if (! have_tma)
		printf("no TMA flavour: tile geometry\n");
	p.tile_y = ty;
	p.tile_y_shift = 0;
	while ((1 << p.tile_y_shift) < ty) p.tile_y_shift++;
	const int TZ = ty ? 128 / ty : 1;
	const dim3 grid_t((unsigned) ((S / std::max(ty, 1)) * ((SZ + TZ - 1) / TZ)), (unsigned) S);
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float t_even[2] = {0, 0}, t_odd[2] = {0, 0}, t_ab = 0;
	const double bytes = (double) XYZ * L::Q * 2 * sizeof(R);
	cudaFuncAttributes fa, fo, ft; cudaFuncGetAttributes(&fa, k_bulk<L, KB_KIND, R, S_AA_EVEN>); cudaFuncGetAttributes(&fo, k_bulk<L, KB_KIND, R, S_AA_ODD>);
	cudaFuncGetAttributes(&ft, k_bulk_tma<L, KB_KIND, R, S_AA_ODD>);
	for (int tma_flavour = 0; tma_flavour < (have_tma ? 2 : 1); tma_flavour++) {
		if (tma_flavour == 1) {
			printf("%-22s S=%d macro=%d regs=%d/%d  even %.4f ms %.0f GB/s | odd %.4f ms %.0f GB/s | AA avg %.0f MLUPS\n", KB_NAME, S, macro, fa.numRegs, fo.numRegs,
				   t_even[0] / iters, bytes / (t_even[0] / iters) / 1e6, t_odd[0] / iters, bytes / (t_odd[0] / iters) / 1e6, 2.0 * XYZ / ((t_even[0] + t_odd[0]) / iters) / 1e3);
			fflush(stdout);
		}
		init();
		for (int it = -4; it < 2 * iters; it++) {
			const bool even = (it & 1) == 0;
			p.stat_counter = it + 4;
			CK(cudaEventRecord(e0));
			if (tma_flavour) {
				if (even) k_bulk_tma<L, KB_KIND, R, S_AA_EVEN><<<grid_t, 128>>>(p); else k_bulk_tma<L, KB_KIND, R, S_AA_ODD><<<grid_t, 128>>>(p);
			}
			else {
				if (even) k_bulk<L, KB_KIND, R, S_AA_EVEN><<<grid_e, BS>>>(p); else k_bulk<L, KB_KIND, R, S_AA_ODD><<<grid_o, BS>>>(p);
			}
			CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
			float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
			if (it >= 0) (even ? t_even : t_odd)[tma_flavour] += ms;
		}
		CK(cudaGetLastError());
		if (tma_flavour == 0 && have_tma) {
			CK(cudaMalloc(&ref, sizeof(R) * L::Q * XYZ));
			CK(cudaMemcpy(ref, a, sizeof(R) * L::Q * XYZ, cudaMemcpyDeviceToDevice));
		}
	}
	if (chain) {
		init();
		set_bases(true);
		float ms = 0;
		for (int rep = 0; rep < 2; rep++) {	 // the first repetition warms up
			CK(cudaEventRecord(e0));
			for (int it = 0; it < 2 * iters; it++) {
				p.stat_counter = it;
				if ((it & 1) == 0) k_bulk<L, KB_KIND, R, S_AA_EVEN><<<grid_e, BS>>>(p); else k_bulk<L, KB_KIND, R, S_AA_ODD><<<grid_o, BS>>>(p);
			}
			CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
			CK(cudaEventElapsedTime(&ms, e0, e1));
		}
		printf("%-22s S=%d macro=%d field=%d CHAINED %d steps: %.4f ms per step, %.0f GB/s, %.0f MLUPS\n", KB_NAME, S, macro, field, 2 * iters, ms / (2 * iters), bytes / (ms / (2 * iters)) / 1e6,
			   XYZ / (ms / (2 * iters)) / 1e3);
		fflush(stdout);
	}
	unsigned long long diff = 0;
	if (have_tma) {
		CK(cudaMemset(d_cnt, 0, 8));
		k_count_diff<R><<<1184, 256>>>(a, ref, (long long) L::Q * XYZ, d_cnt);
		CK(cudaMemcpy(&diff, d_cnt, 8, cudaMemcpyDeviceToHost));
		CK(cudaFree(ref));
	}
	std::vector<R> h(8); CK(cudaMemcpy(h.data(), a, sizeof(R) * 8, cudaMemcpyDeviceToHost));
	init();
	CK(cudaMemcpy(b, a, sizeof(R) * L::Q * XYZ, cudaMemcpyDeviceToDevice));
	for (int it = -2; it < iters; it++) {
		CK(cudaEventRecord(e0));
		set_bases(false);
		k_bulk<L, KB_KIND, R, S_AB><<<grid_ab, BS>>>(p);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
		if (it >= 0) t_ab += ms;
		R* t = p.cur; p.cur = p.out; p.out = t;
	}
	CK(cudaGetLastError());
	printf("%-22s S=%d macro=%d regs=%d/%d  even %.4f ms %.0f GB/s | odd %.4f ms %.0f GB/s | AA avg %.0f MLUPS | AB %.4f ms %.0f GB/s  (f0=%.6f)\n", KB_NAME, S, macro, fa.numRegs, fo.numRegs,
		   t_even[0] / iters, bytes / (t_even[0] / iters) / 1e6, t_odd[0] / iters, bytes / (t_odd[0] / iters) / 1e6, 2.0 * XYZ / ((t_even[0] + t_odd[0]) / iters) / 1e3, t_ab / iters, bytes / (t_ab / iters) / 1e6, (double) h[0]);
	if (have_tma)
		printf("%-22s   TMA tile_y=%d regs=%d  even %.4f ms %.0f GB/s | odd %.4f ms %.0f GB/s | AA avg %.0f MLUPS | plain-even+tma-odd %.0f MLUPS | differing elements vs plain: %llu\n", KB_NAME, ty, ft.numRegs,
			   t_even[1] / iters, bytes / (t_even[1] / iters) / 1e6, t_odd[1] / iters, bytes / (t_odd[1] / iters) / 1e6, 2.0 * XYZ / ((t_even[1] + t_odd[1]) / iters) / 1e3,
			   2.0 * XYZ / ((t_even[0] + t_odd[1]) / iters) / 1e3, diff);
	return 0;
}
