// Host-side check of the engine's per-cell operators (collide*.cuh) against the CPU restatement (test tool; g++ -ffp-contract=off;
// driven by tests/test_operators_on_host.py).
#define __host__
#define __device__
#define __forceinline__ inline
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using std::sqrt;
#ifndef LBMX_STRICT
#define LBMX_STRICT 1
#endif
#include "../../tnl_lbm_b200/csrc/collide.cuh"
#include "../../oracle/oracle_api.h"
using namespace lbmx;
template <typename R, int KIND>
int run(int coll, int eq, int prec)
{
	const int N = 6;
	oracle_desc d{};
	d.lattice = ORC_D3Q27; d.coll = coll; d.eq = eq; d.streaming = ORC_STREAM_AB; d.macro = ORC_MACRO_DEFAULT; d.inflow = ORC_INFLOW_CONST;
	d.precision = prec; d.nproc = 1; d.X = d.Y = d.Z = N; d.ox = 0;
	oracle_params p{};
	p.lbmViscosity = 1e-3; p.fx = 1e-6; p.fy = -2e-6; p.fz = 3e-6;
	const size_t XYZ = N * N * N;
	std::vector<R> a(27 * XYZ), b(27 * XYZ), mac(4 * XYZ), mine(27 * XYZ);
	std::vector<int16_t> map(XYZ, 7);  // GEO_PERIODIC
	srand(5);
	for (int q = 0; q < 27; q++) {
		int n = (D3Q27::cx(q) != 0) + (D3Q27::cy(q) != 0) + (D3Q27::cz(q) != 0);
		double w = n == 0 ? 8. / 27 : n == 1 ? 2. / 27 : n == 2 ? 1. / 54 : 1. / 216;
		for (size_t i = 0; i < XYZ; i++) a[q * XYZ + i] = (R) (w * (1 + 0.05 * (rand() / (double) RAND_MAX - 0.5)));
	}
	Phys<R> P; P.nu = (R) p.lbmViscosity; set_rates(P); P.fx = (R) p.fx; P.fy = (R) p.fy; P.fz = (R) p.fz;
	for (int x = 0; x < N; x++) for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) {
		R f[27];
		for (int q = 0; q < 27; q++) {
			int xs = (x - D3Q27::cx(q) + N) % N, ys = (y - D3Q27::cy(q) + N) % N, zs = (z - D3Q27::cz(q) + N) % N;
			f[q] = a[q * XYZ + ((size_t) xs * N + zs) * N + ys];
		}
		R rho, vx, vy, vz;
		density_velocity<KIND == K_CUM_HP_RHO>(f, P, rho, vx, vy, vz);
		collide<KIND>(f, P, eq, rho, vx, vy, vz);
		for (int q = 0; q < 27; q++) mine[q * XYZ + ((size_t) x * N + z) * N + y] = f[q];
	}
	oracle_step(&d, &p, a.data(), b.data(), mac.data(), map.data(), 0, 1, 1);
	size_t nd = 0; double mx = 0;
	for (size_t i = 0; i < 27 * XYZ; i++) if (mine[i] != b[i]) { nd++; mx = std::fmax(mx, std::fabs((double) mine[i] - (double) b[i])); }
	printf("coll %d eq %d prec %d: %zu of %zu differ, max abs %.3e\n", coll, eq, prec, nd, 27 * XYZ, mx);
	return nd != 0;
}
// D2Q9 (d2q9/col_srt.h, col_clbm.h): X x Y x 1 periodic box, storage index (x * Z + z) * Y + y with Z = 1
template <typename R, int KIND>
int run2d(int coll, int prec)
{
	const int NX = 7, NY = 6;
	oracle_desc d{};
	d.lattice = ORC_D2Q9; d.coll = coll; d.eq = ORC_EQ_STD; d.streaming = ORC_STREAM_AB; d.macro = ORC_MACRO_DEFAULT; d.inflow = ORC_INFLOW_CONST;
	d.precision = prec; d.nproc = 1; d.X = NX; d.Y = NY; d.Z = 1; d.ox = 0;
	oracle_params p{};
	p.lbmViscosity = 2e-2; p.fx = 2e-6; p.fy = -1e-6;
	const size_t XYZ = (size_t) NX * NY;
	std::vector<R> a(9 * XYZ), b(9 * XYZ), mac(3 * XYZ), mine(9 * XYZ);
	std::vector<int16_t> map(XYZ, 6);  // D2Q9 GEO_PERIODIC
	srand(7);
	for (int q = 0; q < 9; q++) {
		int n = (D2Q9::cx(q) != 0) + (D2Q9::cy(q) != 0);
		double w = n == 0 ? 4. / 9 : n == 1 ? 1. / 9 : 1. / 36;
		for (size_t i = 0; i < XYZ; i++) a[q * XYZ + i] = (R) (w * (1 + 0.05 * (rand() / (double) RAND_MAX - 0.5)));
	}
	Phys<R> P; P.nu = (R) p.lbmViscosity; set_rates(P); P.fx = (R) p.fx; P.fy = (R) p.fy; P.fz = 0;
	for (int x = 0; x < NX; x++) for (int y = 0; y < NY; y++) {
		R f[9];
		for (int q = 0; q < 9; q++) {
			int xs = (x - D2Q9::cx(q) + NX) % NX, ys = (y - D2Q9::cy(q) + NY) % NY;
			f[q] = a[q * XYZ + (size_t) xs * NY + ys];
		}
		R rho, vx, vy, vz;
		density_velocity(f, P, rho, vx, vy, vz);
		collide<KIND>(f, P, 0, rho, vx, vy, vz);
		for (int q = 0; q < 9; q++) mine[q * XYZ + (size_t) x * NY + y] = f[q];
	}
	oracle_step(&d, &p, a.data(), b.data(), mac.data(), map.data(), 0, 1, 1);
	size_t nd = 0; double mx = 0;
	for (size_t i = 0; i < 9 * XYZ; i++) if (mine[i] != b[i]) { nd++; mx = std::fmax(mx, std::fabs((double) mine[i] - (double) b[i])); }
	printf("coll %d eq %d prec %d: %zu of %zu differ, max abs %.3e (D2Q9)\n", coll, 0, prec, nd, 9 * XYZ, mx);
	return nd != 0;
}

int main()
{
	int r = 0;
	r |= run<float, K_CUM>(ORC_COLL_CUM, ORC_EQ_INV_CUM, ORC_F32);
	r |= run<double, K_CUM>(ORC_COLL_CUM, ORC_EQ_INV_CUM, ORC_F64);
	r |= run<float, K_SRT>(ORC_COLL_SRT, ORC_EQ_STD, ORC_F32);
	r |= run<float, K_BGK>(ORC_COLL_BGK, ORC_EQ_STD, ORC_F32);
	r |= run<float, K_BGK_GAL>(ORC_COLL_BGK_GALILEAN, ORC_EQ_STD, ORC_F32);
	r |= run<double, K_BGK_GAL>(ORC_COLL_BGK_GALILEAN, ORC_EQ_STD, ORC_F64);
	r |= run<float, K_CUM_HP_RHO>(ORC_COLL_CUM_HP_RHO, ORC_EQ_INV_CUM, ORC_F32);
	r |= run<double, K_CUM_HP_RHO>(ORC_COLL_CUM_HP_RHO, ORC_EQ_INV_CUM, ORC_F64);
	r |= run<float, K_MRT>(ORC_COLL_MRT_LES, ORC_EQ_STD, ORC_F32);
	r |= run<float, K_CLBM>(ORC_COLL_CLBM, ORC_EQ_STD, ORC_F32);
	r |= run<double, K_CLBM>(ORC_COLL_CLBM, ORC_EQ_STD, ORC_F64);
	r |= run<float, K_SRT_MF>(ORC_COLL_SRT_MODIF_FORCE, ORC_EQ_INV_CUM, ORC_F32);
	r |= run<double, K_SRT_MF>(ORC_COLL_SRT_MODIF_FORCE, ORC_EQ_STD, ORC_F64);
	r |= run<float, K_CUM_2017>(ORC_COLL_CUM_2017, ORC_EQ_INV_CUM, ORC_F32);
	r |= run<double, K_CUM_2017>(ORC_COLL_CUM_2017, ORC_EQ_INV_CUM, ORC_F64);
	r |= run<float, K_CUM_AALIAS>(ORC_COLL_CUM_ANTIALIAS, ORC_EQ_INV_CUM, ORC_F32);
	r |= run<double, K_CUM_AALIAS>(ORC_COLL_CUM_ANTIALIAS, ORC_EQ_INV_CUM, ORC_F64);
	r |= run<float, K_CUM_2017_AALIAS>(ORC_COLL_CUM_2017_ANTIALIAS, ORC_EQ_INV_CUM, ORC_F32);
	r |= run<double, K_CUM_2017_AALIAS>(ORC_COLL_CUM_2017_ANTIALIAS, ORC_EQ_INV_CUM, ORC_F64);
	r |= run<float, K_KBC_N1>(ORC_COLL_KBC_N1, ORC_EQ_STD, ORC_F32);
	r |= run<double, K_KBC_N2>(ORC_COLL_KBC_N2, ORC_EQ_STD, ORC_F64);
	r |= run<float, K_KBC_N3>(ORC_COLL_KBC_N3, ORC_EQ_STD, ORC_F32);
	r |= run<double, K_KBC_N4>(ORC_COLL_KBC_N4, ORC_EQ_STD, ORC_F64);
	r |= run<float, K_KBC_C1>(ORC_COLL_KBC_C1, ORC_EQ_STD, ORC_F32);
	r |= run<double, K_KBC_C2>(ORC_COLL_KBC_C2, ORC_EQ_STD, ORC_F64);
	r |= run<float, K_KBC_C3>(ORC_COLL_KBC_C3, ORC_EQ_STD, ORC_F32);
	r |= run<float, K_KBC_C4>(ORC_COLL_KBC_C4, ORC_EQ_STD, ORC_F32);
	r |= run<double, K_KBC_C4>(ORC_COLL_KBC_C4, ORC_EQ_STD, ORC_F64);
	r |= run2d<float, K_SRT>(ORC_COLL_SRT, ORC_F32);
	r |= run2d<double, K_SRT>(ORC_COLL_SRT, ORC_F64);
	r |= run2d<float, K_CLBM>(ORC_COLL_CLBM, ORC_F32);
	r |= run2d<double, K_CLBM>(ORC_COLL_CLBM, ORC_F64);
	return r;
}
