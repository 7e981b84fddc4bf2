// collide_ext.cuh -- further D3Q27 collision operators of the reference in the same trait slot.
//
// Written once, in the reference's floating-point association (cited per function, paths relative to the reference's
// include/lbm3d/): the default build compiles them with FMA contraction, the parity-arithmetic build (-fmad=false) reproduces the
// reference's strict CPU build bit for bit.  All of them are HBM-bound like the operators in collide.cuh; the one place where
// work is skipped in the default build is the CLBM forcing term when the body force is exactly zero (it is then an exact zero).
//
//   D3Q27_CLBM             d3q27/col_clbm.h:6-447
//   D3Q27_SRT_MODIF_FORCE  d3q27/col_srt_modif_force.h:9-120
//   D3Q27_KBC_N1..N4, C1..C4  d3q27/col_kbc_n.h:254-1272, col_kbc_c.h:283-1301
#pragma once
#include "lattice.cuh"

namespace lbmx {
namespace ext {

// S_row = (1/den) * sum_j coef(row, j) * m_j over ascending j, added to population dir(row) (col_clbm.h:342-443)
struct ClbmForce
{
	LBMX_HD static constexpr int coef(int row, int j)
	{
		constexpr signed char t[27][27] = {
			{  1,   0,   0,   0,   0,   0,   0,   0,   0,  -3,   0,   0,   0,   0,   0,   0,   0,   3,   0,   0,   0,   0,   0,   0,   0,   0,  -1},
			{  4,   6,   0,   0,   0,   0,   0,   9,   3,  -6,  -6,   0,   0,   0,   0,   0,   0,   0,  -6,   0,   0,   0,   0,   6,   0,   0,   2},
			{  4,  -6,   0,   0,   0,   0,   0,   9,   3,  -6,   6,   0,   0,   0,   0,   0,   0,   0,  -6,   0,   0,   0,   0,  -6,   0,   0,   2},
			{  4,   0,   6,   0,   0,   0,   0,  -9,   3,  -6,   0,  -6,   0,   0,   0,   0,   0,   0,   3,  -9,   0,   0,   0,   0,   6,   0,   2},
			{  4,   0,  -6,   0,   0,   0,   0,  -9,   3,  -6,   0,   6,   0,   0,   0,   0,   0,   0,   3,  -9,   0,   0,   0,   0,  -6,   0,   2},
			{  4,   0,   0,   6,   0,   0,   0,   0,  -6,  -6,   0,   0,  -6,   0,   0,   0,   0,   0,   3,   9,   0,   0,   0,   0,   0,   6,   2},
			{  4,   0,   0,  -6,   0,   0,   0,   0,  -6,  -6,   0,   0,   6,   0,   0,   0,   0,   0,   3,   9,   0,   0,   0,   0,   0,  -6,   2},
			{  8,  12,  12,   0,  18,   0,   0,   0,  12,   0,  -3,  -3,   0,  27,  27,   0,   0,  -6,   3,   9,   0,   0, -18,  -6,  -6,   0,  -2},
			{  8, -12,  12,   0, -18,   0,   0,   0,  12,   0,   3,  -3,   0, -27,  27,   0,   0,  -6,   3,   9,   0,   0,  18,   6,  -6,   0,  -2},
			{  8,  12, -12,   0, -18,   0,   0,   0,  12,   0,  -3,   3,   0,  27, -27,   0,   0,  -6,   3,   9,   0,   0,  18,  -6,   6,   0,  -2},
			{  8, -12, -12,   0,  18,   0,   0,   0,  12,   0,   3,   3,   0, -27, -27,   0,   0,  -6,   3,   9,   0,   0, -18,   6,   6,   0,  -2},
			{  8,  12,   0,  12,   0,  18,   0,  18,  -6,   0,  -3,   0,  -3, -27,   0,  27,   0,  -6,   3,  -9,   0, -18,   0,  -6,   0,  -6,  -2},
			{  8, -12,   0,  12,   0, -18,   0,  18,  -6,   0,   3,   0,  -3,  27,   0,  27,   0,  -6,   3,  -9,   0,  18,   0,   6,   0,  -6,  -2},
			{  8,  12,   0, -12,   0, -18,   0,  18,  -6,   0,  -3,   0,   3, -27,   0, -27,   0,  -6,   3,  -9,   0,  18,   0,  -6,   0,   6,  -2},
			{  8, -12,   0, -12,   0,  18,   0,  18,  -6,   0,   3,   0,   3,  27,   0, -27,   0,  -6,   3,  -9,   0, -18,   0,   6,   0,   6,  -2},
			{  8,   0,  12,  12,   0,   0,  18, -18,  -6,   0,   0,  -3,  -3,   0, -27, -27,   0,  -6,  -6,   0, -18,   0,   0,   0,  -6,  -6,  -2},
			{  8,   0, -12,  12,   0,   0, -18, -18,  -6,   0,   0,   3,  -3,   0,  27, -27,   0,  -6,  -6,   0,  18,   0,   0,   0,   6,  -6,  -2},
			{  8,   0,  12, -12,   0,   0, -18, -18,  -6,   0,   0,  -3,   3,   0, -27,  27,   0,  -6,  -6,   0,  18,   0,   0,   0,  -6,   6,  -2},
			{  8,   0, -12, -12,   0,   0,  18, -18,  -6,   0,   0,   3,   3,   0,  27,  27,   0,  -6,  -6,   0, -18,   0,   0,   0,   6,   6,  -2},
			{  8,  12,  12,  12,  18,  18,  18,   0,   0,  12,   6,   6,   6,   0,   0,   0,  27,   6,   0,   0,   9,   9,   9,   3,   3,   3,   1},
			{  8, -12,  12,  12, -18, -18,  18,   0,   0,  12,  -6,   6,   6,   0,   0,   0, -27,   6,   0,   0,   9,  -9,  -9,  -3,   3,   3,   1},
			{  8,  12, -12,  12, -18,  18, -18,   0,   0,  12,   6,  -6,   6,   0,   0,   0, -27,   6,   0,   0,  -9,   9,  -9,   3,  -3,   3,   1},
			{  8, -12, -12,  12,  18, -18, -18,   0,   0,  12,  -6,  -6,   6,   0,   0,   0,  27,   6,   0,   0,  -9,  -9,   9,  -3,  -3,   3,   1},
			{  8,  12,  12, -12,  18, -18, -18,   0,   0,  12,   6,   6,  -6,   0,   0,   0, -27,   6,   0,   0,  -9,  -9,   9,   3,   3,  -3,   1},
			{  8, -12,  12, -12, -18,  18, -18,   0,   0,  12,  -6,   6,  -6,   0,   0,   0,  27,   6,   0,   0,  -9,   9,  -9,  -3,   3,  -3,   1},
			{  8,  12, -12, -12, -18, -18,  18,   0,   0,  12,   6,  -6,  -6,   0,   0,   0,  27,   6,   0,   0,   9,  -9,  -9,   3,  -3,  -3,   1},
			{  8, -12, -12, -12,  18,  18,  18,   0,   0,  12,  -6,  -6,  -6,   0,   0,   0, -27,   6,   0,   0,   9,   9,   9,  -3,  -3,  -3,   1},
		};
		return t[row][j];
	}
	LBMX_HD static constexpr int den(int row)
	{
		constexpr int t[27] = {27, 108, 108, 108, 108, 108, 108, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216, 216};
		return t[row];
	}
	LBMX_HD static constexpr int dir(int row)
	{
		constexpr signed char t[27][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1}, {-1, 0, -1}, {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}, {1, 1, 1}, {-1, 1, 1}, {1, -1, 1}, {-1, -1, 1}, {1, 1, -1}, {-1, 1, -1}, {1, -1, -1}, {-1, -1, -1}};
		return D3Q27::find(t[row][0], t[row][1], t[row][2]);
	}
};

template <typename R>
LBMX_D void clbm_force_moments(R (&m)[27], R u, R v, R w, R Fx, R Fy, R Fz)  // col_clbm.h:303-340
{
	const R c2 = 2, c3 = 3, c4 = 4, c6 = 6, c8 = 8, c9 = 9, c12 = 12, c18 = 18;
	m[0] = 0;
	m[1] = Fx;
	m[2] = Fy;
	m[3] = Fz;
	m[4] = (Fx * v + Fy * u);
	m[5] = (Fx * w + Fz * u);
	m[6] = (Fy * w + Fz * v);
	m[7] = c2 * (Fx * u - Fy * v);
	m[8] = c2 * (Fx * u + Fy * v - c2 * Fz * w);
	m[9] = c2 * (Fx * u + Fy * v + Fz * w);
	m[10] = (c3 * v * v + c3 * w * w - c4) * Fx + c6 * u * v * Fy + c6 * u * w * Fz;
	m[11] = c6 * u * v * Fx + (c3 * u * u + c3 * w * w - c4) * Fy + c6 * w * v * Fz;
	m[12] = c6 * u * w * Fx + c6 * w * v * Fy + (c3 * u * u + c3 * v * v - c4) * Fz;
	m[13] = (v * v - w * w) * Fx + c2 * u * v * Fy - c2 * u * w * Fz;
	m[14] = c2 * u * v * Fx + (u * u - w * w) * Fy - c2 * w * v * Fz;
	m[15] = c2 * u * w * Fx - c2 * w * v * Fy + (u * u - v * v) * Fz;
	m[16] = Fx * v * w + Fy * u * w + Fz * u * v;
	m[17] = (c6 * v * v + c6 * w * w - c8) * u * Fx + (c6 * u * u * v + c6 * v * w * w - c8 * v) * Fy + (c6 * u * u * w + c6 * v * v * w - c8 * w) * Fz;
	m[18] = (c6 * v * v + c6 * w * w - c8) * u * Fx + (c6 * u * u * v - c12 * v * w * w + c4 * v) * Fy + (c6 * u * u * w - c12 * v * v * w + c4 * w) * Fz;
	m[19] = (c6 * v * v - c6 * w * w) * u * Fx + (c6 * u * u * v - c4 * v) * Fy + (-c6 * u * u * w + c4 * w) * Fz;
	m[20] = c6 * u * v * w * Fx + (c3 * u * u * w - c2 * w) * Fy + (c3 * u * u * v - c2 * v) * Fz;
	m[21] = (c3 * v * v * w - c2 * w) * Fx + c6 * u * v * w * Fy + (c3 * u * v * v - c2 * u) * Fz;
	m[22] = (c3 * v * w * w - c2 * v) * Fx + (c3 * u * w * w - c2 * u) * Fy + c6 * u * v * w * Fz;
	m[23] = ((c9 * w * w - c6) * v * v - c6 * w * w + c4) * Fx + (c18 * w * w - c12) * v * u * Fy + c6 * u * w * (c3 * v * v - c2) * Fz;
	m[24] = (c18 * w * w - c12) * v * u * Fx + ((c9 * w * w - c6) * u * u - c6 * w * w + c4) * Fy + c6 * w * v * (c3 * u * u - c2) * Fz;
	m[25] = c6 * u * w * (c3 * v * v - c2) * Fx + c6 * w * v * (c3 * u * u - c2) * Fy + ((c9 * v * v - c6) * u * u - c6 * v * v + c4) * Fz;
	m[26] = (c6 * (c3 * w * w - c2)) * (c3 * v * v - c2) * u * Fx + (c6 * (c3 * w * w - c2)) * (c3 * u * u - c2) * v * Fy
		  + c6 * w * (c3 * u * u - c2) * (c3 * v * v - c2) * Fz;
}

// to_central / from_central: strict::to_central, strict::from_central (collide_strict.cuh) -- Eq 6-14 and Eq 57-65 of
// col_clbm.h:18-117,202-300 are the cumulant operator's transforms
template <bool SKIP_ZERO_FORCE, typename R, typename PHYS>
LBMX_D void collide_clbm(R (&f)[27], const PHYS& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	const R one = R(1), two = R(2), three = R(3), half = R(0.5), third = R(1.0 / 3.0), n1o27 = R(1.0 / 27.0);
	R k[3][3][3];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
#pragma unroll
			for (int c = 0; c < 3; c++)
				k[a][b][c] = f[L::find(a - 1, b - 1, c - 1)];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
			strict::to_central(k[a][b][0], k[a][b][1], k[a][b][2], vz);
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int c = 0; c < 3; c++)
			strict::to_central(k[a][0][c], k[a][1][c], k[a][2][c], vy);
#pragma unroll
	for (int b = 0; b < 3; b++)
#pragma unroll
		for (int c = 0; c < 3; c++)
			strict::to_central(k[0][b][c], k[1][b][c], k[2][b][c], vx);

	// relaxation, default build of the reference: omega2..omega10 = 1, no antialias derivatives (col_clbm.h:119-200).  Terms that
	// are an exact +-0 for finite inputs ((1 - omega_n) * k, products with the zero derivatives) are left out.
	const R omega1 = one / (three * P.nu + half);
	const R keep = one - omega1;
	R s[3][3][3];
	const R d4 = keep * (k[2][0][0] - k[0][2][0]), d5 = keep * (k[2][0][0] - k[0][0][2]), d6 = rho;
	s[2][0][0] = third * (d4 + d5 + d6);
	s[0][2][0] = third * (-two * d4 + d5 + d6);
	s[0][0][2] = third * (d4 - two * d5 + d6);
	s[1][2][0] = (-k[1][0][2] - k[1][2][0]) * half + (k[1][0][2] - k[1][2][0]) * half + k[1][2][0];
	s[1][0][2] = (-k[1][0][2] - k[1][2][0]) * half + (-k[1][0][2] + k[1][2][0]) * half + k[1][0][2];
	s[2][1][0] = (-k[0][1][2] - k[2][1][0]) * half + (k[0][1][2] - k[2][1][0]) * half + k[2][1][0];
	s[0][1][2] = (-k[0][1][2] - k[2][1][0]) * half + (-k[0][1][2] + k[2][1][0]) * half + k[0][1][2];
	s[0][2][1] = (-k[0][2][1] - k[2][0][1]) * half + (-k[0][2][1] + k[2][0][1]) * half + k[0][2][1];
	s[2][0][1] = (-k[0][2][1] - k[2][0][1]) * half + (k[0][2][1] - k[2][0][1]) * half + k[2][0][1];
	s[1][1][1] = R(0);
	const R d16 = rho * third;
	s[2][2][0] = third * d16;  // third * (0 + 0 + d16)
	s[2][0][2] = third * d16;
	s[0][2][2] = third * d16;
	s[2][1][1] = s[1][2][1] = s[1][1][2] = R(0);
	s[2][2][1] = s[2][1][2] = s[1][2][2] = R(0);
	s[2][2][2] = rho * n1o27;
	s[0][0][0] = k[0][0][0];
	s[1][0][0] = k[1][0][0];
	s[0][1][0] = k[0][1][0];
	s[0][0][1] = k[0][0][1];
	s[1][0][1] = keep * k[1][0][1];
	s[0][1][1] = keep * k[0][1][1];
	s[1][1][0] = keep * k[1][1][0];
#pragma unroll
	for (int b = 0; b < 3; b++)
#pragma unroll
		for (int c = 0; c < 3; c++)
			strict::from_central(s[0][b][c], s[1][b][c], s[2][b][c], vx);
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int c = 0; c < 3; c++)
			strict::from_central(s[a][0][c], s[a][1][c], s[a][2][c], vy);
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
			strict::from_central(s[a][b][0], s[a][b][1], s[a][b][2], vz);
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
#pragma unroll
			for (int c = 0; c < 3; c++)
				f[L::find(a - 1, b - 1, c - 1)] = s[a][b][c];

	if (SKIP_ZERO_FORCE && P.fx == R(0) && P.fy == R(0) && P.fz == R(0))
		return;	 // every m_j is proportional to the force: the term is an exact zero (uniform branch)
	R m[27];
	clbm_force_moments(m, vx, vy, vz, P.fx, P.fy, P.fz);
	static_for<27>([&](auto rc) {
		constexpr int row = rc;
		R acc = R(0);
		bool first = true;
		static_for<26>([&](auto jc) {
			constexpr int j = jc + 1;  // m_0 = 0
			constexpr int c = ClbmForce::coef(row, j);
			if constexpr (c != 0) {
				const R t = R(c) * m[j];
				acc = first ? t : acc + t;
				first = false;
			}
		});
		f[ClbmForce::dir(row)] += R(1.0 / ClbmForce::den(row)) * acc;
	});
}

// col_srt_modif_force.h:17-118.  The reference writes the source term with double literals: sums that contain one are
// evaluated in double also for dreal = float (products of two dreal variables stay in dreal), rounded once when stored.
LBMX_HD constexpr int dir_comp(int q, int a)
{
	return a == 0 ? D3Q27::cx(q) : (a == 1 ? D3Q27::cy(q) : D3Q27::cz(q));
}
template <int q, typename R>
LBMX_D R modif_force_source(const R (&v)[3], const R (&F)[3])
{
	constexpr int n = (dir_comp(q, 0) != 0) + (dir_comp(q, 1) != 0) + (dir_comp(q, 2) != 0);
	if constexpr (n == 0)
		return R(-8.0 / 9.0 * double(v[0] * F[0] + F[1] * v[1] + v[2] * F[2]));
	else if constexpr (n == 1) {
		constexpr int a = dir_comp(q, 0) != 0 ? 0 : (dir_comp(q, 1) != 0 ? 1 : 2), o1 = a == 0 ? 1 : 0, o2 = a == 2 ? 1 : 2;
		const double own = (4.0 * double(v[a]) + (dir_comp(q, a) > 0 ? 2.0 : -2.0)) * double(F[a]) / 9.0;
		return R(own - 2.0 / 9.0 * double(v[o1] * F[o1] + v[o2] * F[o2]));
	}
	else {
		constexpr double den = n == 2 ? 18.0 : 72.0;
		double sum = 0;
		bool first = true;
		static_for<3>([&](auto ac) {
			constexpr int a = ac;
			if constexpr (dir_comp(q, a) != 0) {
				double A = 0;
				bool f1 = true;
				static_for<3>([&](auto bc) {
					constexpr int b = bc;
					if constexpr (dir_comp(q, b) != 0) {
						const double t = (b == a ? 2.0 : 3.0 * double(dir_comp(q, a) * dir_comp(q, b))) * double(v[b]);
						A = f1 ? t : A + t;
						f1 = false;
					}
				});
				A = A + double(dir_comp(q, a));
				const double term = A * double(F[a]) / den;
				sum = first ? term : sum + term;
				first = false;
			}
		});
		if constexpr (n == 2) {
			constexpr int z = dir_comp(q, 0) == 0 ? 0 : (dir_comp(q, 1) == 0 ? 1 : 2);
			sum = sum - double(F[z] * v[z]) / 18.0;
		}
		return R(sum);
	}
}

// the same source term for the default build: evaluated in R, one multiplication by 1/72, 1/18 or 1/9 per population instead of
// three double-precision divisions (81 fp64 divisions per cell made this operator compute-bound at 28 % of the HBM roofline)
template <int q, typename R>
LBMX_D R modif_force_source_fast(const R (&v)[3], const R (&F)[3])
{
	constexpr int n = (dir_comp(q, 0) != 0) + (dir_comp(q, 1) != 0) + (dir_comp(q, 2) != 0);
	if constexpr (n == 0)
		return R(-8.0 / 9.0) * ((v[0] * F[0] + v[1] * F[1]) + v[2] * F[2]);
	else if constexpr (n == 1) {
		constexpr int a = dir_comp(q, 0) != 0 ? 0 : (dir_comp(q, 1) != 0 ? 1 : 2), o1 = a == 0 ? 1 : 0, o2 = a == 2 ? 1 : 2;
		return (R(4.0 / 9.0) * v[a] + R(2.0 / 9.0 * dir_comp(q, a))) * F[a] - R(2.0 / 9.0) * (v[o1] * F[o1] + v[o2] * F[o2]);
	}
	else {
		R sum = R(0);
		static_for<3>([&](auto ac) {
			constexpr int a = ac;
			if constexpr (dir_comp(q, a) != 0) {
				R A = R(dir_comp(q, a));
				static_for<3>([&](auto bc) {
					constexpr int b = bc;
					if constexpr (dir_comp(q, b) != 0)
						A += R(b == a ? 2.0 : 3.0 * (dir_comp(q, a) * dir_comp(q, b))) * v[b];
				});
				sum += A * F[a];
			}
			else
				sum -= F[a] * v[a];	 // only an edge direction has a zero component
		});
		return sum * R(n == 2 ? 1.0 / 18.0 : 1.0 / 72.0);
	}
}

template <bool EXACT, typename R, typename PHYS>
LBMX_D void collide_srt_modif(R (&f)[27], const R (&feq)[27], const PHYS& P, R vx, R vy, R vz)
{
	const R one = R(1), half = R(0.5);
	const R tau = R(3) * P.nu + half;
	const R v[3] = {vx, vy, vz}, F[3] = {P.fx, P.fy, P.fz};
	if constexpr (EXACT) {
		static_for<27>([&](auto qc) {
			constexpr int q = qc;
			const R S = modif_force_source<q>(v, F);
			f[q] += (feq[q] - f[q]) / tau + (one - half / tau) * S;
		});
	}
	else {
		const R itau = one / tau, pre = one - half * itau;
		static_for<27>([&](auto qc) {
			constexpr int q = qc;
			f[q] += (feq[q] - f[q]) * itau + pre * modif_force_source_fast<q>(v, F);
		});
	}
}

// ---------------------------------------------------------------------------------------------------------------------------
// KBC family (d3q27/col_kbc_n.h:254-1272, col_kbc_c.h:283-1301): f = k + s + h, the shear part s relaxes with 2 beta, the
// higher-order part h with gamma * beta (gamma: entropic stabiliser).  One template covers the eight models:
//   CENTRAL = false: N1..N4, s from raw moments;  true: C1..C4, s from central moments
//   USE_T, USE_Q: s = D (+ T) (+ Q)   ->   N1/C1: D, N2/C2: D+T, N3/C3: D+Q, N4/C4: D+T+Q
// ---------------------------------------------------------------------------------------------------------------------------
struct KbcMoments
{
	enum { M200, M020, M002, M110, M101, M011, M111, M201, M102, M210, M120, M021, M012, COUNT };
	// populations of every raw moment in the order the reference adds them up (col_kbc_n.h:351-386); -1 = end of list
	LBMX_HD static constexpr int member(int moment, int i)
	{
		constexpr signed char t[COUNT][18] = {
			{20, 22,  8, 24, 26, 10, 12, 14,  2, 25, 23,  9, 21, 19,  7, 13, 11,  1},  // M200
			{20, 22,  8, 24, 26, 10, 16, 18,  4, 17, 15,  3, 25, 23,  9, 21, 19,  7},  // M020
			{20, 22, 12, 14, 24, 26, 16, 18,  6,  5, 17, 15, 25, 23, 13, 11, 21, 19},  // M002
			{20,  8, 22, 24, 10, 26, 25,  9, 23, 21,  7, 19, -1, -1, -1, -1, -1, -1},  // M110
			{20, 22, 12, 14, 24, 26, 25, 23, 13, 11, 21, 19, -1, -1, -1, -1, -1, -1},  // M101
			{20, 22, 24, 26, 16, 18, 17, 15, 25, 23, 21, 19, -1, -1, -1, -1, -1, -1},  // M011
			{19, 22, 24, 26, 25, 23, 21, 20, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1},  // M111
			{19, 22, 12, 14, 24, 26, 25, 23, 13, 11, 21, 20, -1, -1, -1, -1, -1, -1},  // M201
			{19, 22, 12, 14, 24, 26, 25, 23, 13, 11, 21, 20, -1, -1, -1, -1, -1, -1},  // M102
			{19,  8, 22, 24, 10, 26, 25,  9, 23, 21,  7, 20, -1, -1, -1, -1, -1, -1},  // M210
			{19,  8, 22, 24, 10, 26, 25,  9, 23, 21,  7, 20, -1, -1, -1, -1, -1, -1},  // M120
			{19, 22, 24, 26, 16, 18, 17, 15, 25, 23, 21, 20, -1, -1, -1, -1, -1, -1},  // M021
			{19, 22, 24, 26, 16, 18, 17, 15, 25, 23, 21, 20, -1, -1, -1, -1, -1, -1},  // M012
		};
		return t[moment][i];
	}
	LBMX_HD static constexpr int power(int moment, int axis)
	{
		constexpr signed char t[COUNT][3] = {{2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}, {2, 0, 1}, {1, 0, 2}, {2, 1, 0}, {1, 2, 0}, {0, 2, 1}, {0, 1, 2}};
		return t[moment][axis];
	}
	LBMX_HD static constexpr int sign(int moment, int q)
	{
		int s = 1;
		for (int a = 0; a < 3; a++)
			for (int i = 0; i < power(moment, a); i++)
				s *= dir_comp(q, a);
		return s;
	}
};

template <bool CENTRAL, typename R>
LBMX_D R kbc_scale(R x, int den)  // raw-moment models multiply by 1/6, 1/4, 1/2, 1/8; central-moment models divide by 6, 4, 2, 8
{
	if constexpr (CENTRAL)
		return x / R(den);
	else
		return x * (den == 6 ? R(1.0 / 6.0) : den == 4 ? R(0.25) : den == 2 ? R(0.5) : R(0.125));
}

// the shear-part tensors per direction (col_kbc_n.h:56-252, col_kbc_c.h:85-281)
template <bool CENTRAL, int q, typename R>
LBMX_D R kbc_tensor_d(R nxz, R nyz, R pxy, R pxz, R pyz)
{
	constexpr int cx = dir_comp(q, 0), cy = dir_comp(q, 1), cz = dir_comp(q, 2), n = (cx != 0) + (cy != 0) + (cz != 0);
	if constexpr (n == 1)
		return cx != 0 ? kbc_scale<CENTRAL>(R(2) * nxz - nyz, 6) : (cy != 0 ? kbc_scale<CENTRAL>(-nxz + R(2) * nyz, 6) : kbc_scale<CENTRAL>(-nxz - nyz, 6));
	else if constexpr (n == 2) {
		const R p = cz == 0 ? pxy : (cy == 0 ? pxz : pyz);
		constexpr int s = cz == 0 ? cx * cy : (cy == 0 ? cx * cz : cy * cz);
		return s > 0 ? kbc_scale<CENTRAL>(p, 4) : kbc_scale<CENTRAL>(-p, 4);
	}
	else
		return R(0);
}
template <bool CENTRAL, int q, typename R>
LBMX_D R kbc_tensor_q(R qxxy, R qxxz, R qxyy, R qyyz, R qxzz, R qyzz, R qxyz)
{
	constexpr int cx = dir_comp(q, 0), cy = dir_comp(q, 1), cz = dir_comp(q, 2), n = (cx != 0) + (cy != 0) + (cz != 0);
	if constexpr (n == 1) {
		const R s = cx != 0 ? (qxyy + qxzz) : (cy != 0 ? (qxxy + qyzz) : (qxxz + qyyz));
		return (cx + cy + cz) > 0 ? kbc_scale<CENTRAL>(-s, 2) : kbc_scale<CENTRAL>(s, 2);
	}
	else if constexpr (n == 2) {
		const R a = cz == 0 ? qxyy : (cy == 0 ? qxzz : qyzz), b = cz == 0 ? qxxy : (cy == 0 ? qxxz : qyyz);
		constexpr int sa = cz == 0 ? cx : (cy == 0 ? cx : cy), sb = cz == 0 ? cy : cz;
		const R ta = sa > 0 ? a : -a;
		return kbc_scale<CENTRAL>(sb > 0 ? ta + b : ta - b, 4);
	}
	else if constexpr (n == 3)
		return cx * cy * cz > 0 ? kbc_scale<CENTRAL>(qxyz, 8) : kbc_scale<CENTRAL>(-qxyz, 8);
	else
		return R(0);
}

// EXACT = the reference's arithmetic as written (parity build).  The default build uses 1/feq_i = (-1/rho) (1/g_x)(1/g_y)(1/g_z)
// and S_i * (1/rho): 10 reciprocals per cell instead of 54 divisions, which is what bounds the exact form.
template <bool CENTRAL, bool USE_T, bool USE_Q, bool EXACT, typename R, typename PHYS>
LBMX_D void collide_kbc(R (&f)[27], const PHYS& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	using KM = KbcMoments;
	const R one = R(1), two = R(2), three = R(3), six = R(6), half = R(0.5), third = R(1.0 / 3.0);
	const R v[3] = {vx, vy, vz};
	R g[3][3];	// product-form equilibrium factors (col_kbc_n.h:293-321), as in col_bgk.h
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const R z = third - one + v[a] * v[a];
		const R p = -half * (z + one + v[a]);
		g[a][1] = z;
		g[a][2] = p;
		g[a][0] = p + v[a];
	}
	R M[KM::COUNT];
	static_for<KM::COUNT>([&](auto kc) {
		constexpr int k = kc;
		R acc = R(0);
		static_for<18>([&](auto ic) {
			constexpr int i = ic;
			constexpr int q = KM::member(k, i);
			if constexpr (q >= 0) {
				if constexpr (i == 0)
					acc = KM::sign(k, q) > 0 ? f[q] : -f[q];
				else
					acc = KM::sign(k, q) > 0 ? acc + f[q] : acc - f[q];
			}
		});
		M[k] = acc;
	});
	R T = (M[KM::M200] + M[KM::M020] + M[KM::M002]), Nxz = (M[KM::M200] - M[KM::M002]), Nyz = (M[KM::M020] - M[KM::M002]);
	R Pxy = M[KM::M110], Pxz = M[KM::M101], Pyz = M[KM::M011];
	R Qxxy = M[KM::M210], Qxxz = M[KM::M201], Qxyy = M[KM::M120], Qyyz = M[KM::M021], Qxzz = M[KM::M102], Qyzz = M[KM::M012], Qxyz = M[KM::M111];
	R eT, eNxz = 0, eNyz = 0, ePxy = 0, ePxz = 0, ePyz = 0, eQxxy = 0, eQxxz = 0, eQxyy = 0, eQyyz = 0, eQxzz = 0, eQyzz = 0, eQxyz = 0;
	if constexpr (! CENTRAL) {	// col_kbc_n.h:28-54
		eT = (rho * (three * third + vx * vx + vy * vy + vz * vz));
		eNxz = (rho * (vx * vx - vz * vz));
		eNyz = (rho * (vy * vy - vz * vz));
		ePxy = (rho * vx * vy);
		ePxz = (rho * vx * vz);
		ePyz = (rho * vy * vz);
		eQxxy = (rho * vy * (third + vx * vx));
		eQxxz = (rho * vz * (third + vx * vx));
		eQxyy = (rho * vx * (third + vy * vy));
		eQyyz = (rho * vz * (third + vy * vy));
		eQxzz = (rho * vx * (third + vz * vz));
		eQyzz = (rho * vy * (third + vz * vz));
		eQxyz = (rho * vx * vy * vz);
	}
	else {	// col_kbc_c.h:56-83: central moments; their equilibria are 0 except the trace
		const R rT = T, rNxz = Nxz, rNyz = Nyz, rPxy = Pxy, rPxz = Pxz, rPyz = Pyz;
		const R rQxxy = Qxxy, rQxxz = Qxxz, rQxyy = Qxyy, rQyyz = Qyyz, rQxzz = Qxzz, rQyzz = Qyzz, rQxyz = Qxyz;
		T = (rT - rho * (vx * vx + vy * vy + vz * vz));
		Nxz = (rNxz + rho * (vz * vz - vx * vx));
		Nyz = (rNyz + rho * (vz * vz - vy * vy));
		Pxy = (rPxy - rho * vx * vy);
		Pxz = (rPxz - rho * vx * vz);
		Pyz = (rPyz - rho * vy * vz);
		Qxxy = (rQxxy - third * (six * vx * Pxy + vy * (three * vx * vx + two * Nxz - Nyz + T)));
		Qxxz = (rQxxz - third * (six * vx * Pxz + vz * (three * vx * vx + two * Nxz - Nyz + T)));
		Qxyy = (rQxyy - third * (six * vy * Pxy + vx * (three * vy * vy + two * Nyz - Nxz + T)));
		Qyyz = (rQyyz - third * (six * vy * Pyz + vz * (three * vy * vy + two * Nyz - Nxz + T)));
		Qxzz = (rQxzz - third * (six * vz * Pxz + vx * (three * vz * vz - Nyz - Nxz + T)));
		Qyzz = (rQyzz - third * (six * vz * Pyz + vy * (three * vz * vz - Nyz - Nxz + T)));
		Qxyz = (rQxyz - vx * Pyz - vy * Pxz - vz * Pxy - vx * vy * vz);
		eT = (rho * three * third);
	}
	// Default arithmetic, raw-moment models: the shear tensors are linear in the moments, so the equilibrium moments are subtracted once
	// here (13 values live instead of 26) instead of per direction as the reference does (col_kbc_n.h:56-252); parity arithmetic keeps
	// the reference's association.
	constexpr bool FOLD = ! EXACT && ! CENTRAL;
	if constexpr (FOLD) {
		T -= eT;
		Nxz -= eNxz;
		Nyz -= eNyz;
		Pxy -= ePxy;
		Pxz -= ePxz;
		Pyz -= ePyz;
		Qxxy -= eQxxy;
		Qxxz -= eQxxz;
		Qxyy -= eQxyy;
		Qyyz -= eQyyz;
		Qxzz -= eQxzz;
		Qyzz -= eQyzz;
		Qxyz -= eQxyz;
	}
	// feq, delta-s and delta-h of a population are cheap functions of 9 equilibrium factors and 13 moments: they are recomputed
	// in the two passes below instead of being stored (4 x 27 live values cost more registers than the kernel can keep resident;
	// the recomputation is the same expression, so the result is unchanged)
	auto feq_of = [&](auto qc) -> R {
		constexpr int q = qc;
		return -rho * g[0][L::cx(q) + 1] * g[1][L::cy(q) + 1] * g[2][L::cz(q) + 1];
	};
	auto ds_of = [&](auto qc) -> R {
		constexpr int q = qc;
		constexpr int n = (L::cx(q) != 0) + (L::cy(q) != 0) + (L::cz(q) != 0);
		R acc = R(0);
		if constexpr (n == 1 || n == 2) {
			acc = kbc_tensor_d<CENTRAL, q>(Nxz, Nyz, Pxy, Pxz, Pyz);
			if constexpr (! CENTRAL && ! FOLD)
				acc = acc - kbc_tensor_d<CENTRAL, q>(eNxz, eNyz, ePxy, ePxz, ePyz);
		}
		if constexpr (USE_T && n <= 1) {
			const R t = n == 0 ? -T : kbc_scale<CENTRAL>(T, 6);
			if constexpr (FOLD)
				acc = acc + t;
			else {
				const R et = n == 0 ? -eT : kbc_scale<CENTRAL>(eT, 6);
				acc = (acc + t) - et;
			}
		}
		if constexpr (USE_Q && n >= 1) {
			acc = acc + kbc_tensor_q<CENTRAL, q>(Qxxy, Qxxz, Qxyy, Qyyz, Qxzz, Qyzz, Qxyz);
			if constexpr (! CENTRAL && ! FOLD)
				acc = acc - kbc_tensor_q<CENTRAL, q>(eQxxy, eQxxz, eQxyy, eQyyz, eQxzz, eQyzz, eQxyz);
		}
		return acc;
	};
	const R beta = (one / (two * P.nu / third + one));
	const R irho = one / rho, nirho = -irho;
	R ig[3][3];
	if constexpr (! EXACT) {
#pragma unroll
		for (int a = 0; a < 3; a++)
#pragma unroll
			for (int c = 0; c < 3; c++)
				ig[a][c] = one / g[a][c];
	}
	// <Ds|Dh> and <Dh|Dh> (weights 1/feq), summed in the order mmm, mmz, mmp, mzm, ... ppp (col_kbc_n.h:233-252)
	R sd = R(0), hh = R(0);
	static_for<27>([&](auto ic) {
		constexpr int i = ic;
		constexpr int q = L::find(i / 9 - 1, (i / 3) % 3 - 1, i % 3 - 1);
		const R fe = feq_of(std::integral_constant<int, q>{}), ds = ds_of(std::integral_constant<int, q>{});
		const R dh = f[q] - fe - ds;
		R ifeq;
		if constexpr (EXACT)
			ifeq = one / fe;
		else
			ifeq = (nirho * ig[0][L::cx(q) + 1]) * (ig[1][L::cy(q) + 1] * ig[2][L::cz(q) + 1]);
		const R t1 = ds * dh * ifeq, t2 = dh * dh * ifeq;
		if constexpr (i == 0) {
			sd = t1;
			hh = t2;
		}
		else {
			sd = sd + t1;
			hh = hh + t2;
		}
	});
	const R gamma = (one / beta - (two - one / beta) * sd / hh);
#if defined(__CUDA_ARCH__) && ! defined(LBMX_KBC_NO_PASS_BARRIER)
	// keep the compiler from carrying the 27 delta-s / feq values of the first pass over to the second (common subexpressions): the point of
	// recomputing them is that they do NOT occupy registers in between
	if constexpr (! EXACT) {
		if constexpr (sizeof(R) == 8)
			asm volatile("" : "+d"(T), "+d"(Nxz), "+d"(Nyz), "+d"(Pxy), "+d"(Pxz), "+d"(Pyz), "+d"(rho));
		else
			asm volatile("" : "+f"(T), "+f"(Nxz), "+f"(Nyz), "+f"(Pxy), "+f"(Pxz), "+f"(Pyz), "+f"(rho));
	}
#endif
	static_for<27>([&](auto qc) {
		constexpr int q = qc;
		const R fe = feq_of(qc), ds = ds_of(qc);
		const R dh = f[q] - fe - ds;
		R S = strict::force_projection(L::cx(q), L::cy(q), L::cz(q), vx, vy, vz, P);
		if constexpr (EXACT)
			S = S / rho;
		else
			S = S * irho;
		f[q] -= beta * (two * ds + gamma * dh) - (one - beta) * S * fe;
	});
}

// 1 / x for weights and scale factors of the default-arithmetic operators: fp32 on the device takes the hardware approximation (MUFU.RCP,
// 1 ulp) instead of the IEEE sequence (~8 instructions with a slow path; the fp32 KBC kernels are bound by instruction issue).
// Measured on the KBC_N4 fp32 kernels (profiles/kbench_r2_kbc_variants.txt): A-A even 5050 -> 5209 GB/s, odd 4138 -> 4465, but A-B
// 4841 -> 3741 (the same source, only this function differs) -- so the A-B bulk kernel keeps the IEEE division (HW = false).
template <bool HW = true, typename R>
LBMX_D R rcp_fast(R x)
{
#if defined(__CUDA_ARCH__) && ! defined(LBMX_NO_RCP_FAST)
	if constexpr (HW && sizeof(R) == 4)
		return __fdividef(R(1), x);
#endif
	return R(1) / x;
}

// The same eight models in default arithmetic, organised for the fp64 / fp32 pipes instead of the reference's statement order (the form
// above needs ~1080 floating-point instructions per cell, which bounds the fp64 kernels on the FP64 pipe before HBM does):
//   * the 13 raw moments come from column sums (z, then y, then x: 72 additions instead of 190);
//   * fp64: pass 1 leaves delta-h in the place of f, so pass 2 is  f' = (1 - beta gamma) dh + (1 - 2 beta) ds + (1 + (1 - beta) S) feq
//     (the reference's  f - beta (2 ds + gamma dh) + (1 - beta) S feq  with f = dh + ds + feq substituted); fp32 keeps the incremental form;
//   * feq_q and 1/feq_q are one multiplication each: the (x,y) factor pairs are formed once per three populations;
//   * (1 - beta) S_q = sx[cx] + sy[cy] + sz[cz] with three values per axis (S is linear in the lattice velocity, col_bgk.h:62-88).
// Same quantities as collide_kbc<..., EXACT = false>, agreement with the reference to rounding.
template <bool CENTRAL, bool USE_T, bool USE_Q, bool HW_RCP = true, typename R, typename PHYS>
LBMX_D void collide_kbc_fast(R (&f)[27], const PHYS& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	constexpr bool INCREMENTAL = sizeof(R) == 4;
	const R one = R(1), two = R(2), three = R(3), six = R(6), half = R(0.5), third = R(1.0 / 3.0);
	const R v[3] = {vx, vy, vz};
	R g[3][3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const R z = third - one + v[a] * v[a];
		const R p = -half * (z + one + v[a]);
		g[a][1] = z;
		g[a][2] = p;
		g[a][0] = p + v[a];
	}
	// ---- raw moments (col_kbc_n.h:351-386) from column sums
	R T, Nxz, Nyz, Pxy, Pxz, Pyz, Qxxy, Qxxz, Qxyy, Qyyz, Qxzz, Qyzz, Qxyz;
	{
		R y00[3], y10[3], y20[3], y01[3], y11[3], y21[3], y02[3], y12[3];
#pragma unroll
		for (int a = 0; a < 3; a++) {
			R z0[3], z1[3], z2[3];
#pragma unroll
			for (int b = 0; b < 3; b++) {
				const R fm = f[L::find(a - 1, b - 1, -1)], f0 = f[L::find(a - 1, b - 1, 0)], fp = f[L::find(a - 1, b - 1, 1)];
				z2[b] = fp + fm;
				z1[b] = fp - fm;
				z0[b] = z2[b] + f0;
			}
			y20[a] = z0[2] + z0[0];
			y10[a] = z0[2] - z0[0];
			y00[a] = y20[a] + z0[1];
			y21[a] = z1[2] + z1[0];
			y11[a] = z1[2] - z1[0];
			y01[a] = y21[a] + z1[1];
			y02[a] = (z2[2] + z2[0]) + z2[1];
			y12[a] = z2[2] - z2[0];
		}
		const R m200 = y00[2] + y00[0], m020 = (y20[2] + y20[0]) + y20[1], m002 = (y02[2] + y02[0]) + y02[1];
		T = (m200 + m020) + m002;
		Nxz = m200 - m002;
		Nyz = m020 - m002;
		Pxy = y10[2] - y10[0];
		Pxz = y01[2] - y01[0];
		Pyz = (y11[2] + y11[0]) + y11[1];
		Qxyz = y11[2] - y11[0];
		Qxxz = y01[2] + y01[0];
		Qxzz = y02[2] - y02[0];
		Qxxy = y10[2] + y10[0];
		Qxyy = y20[2] - y20[0];
		Qyyz = (y21[2] + y21[0]) + y21[1];
		Qyzz = (y12[2] + y12[0]) + y12[1];
	}
	if constexpr (! CENTRAL) {	// minus the equilibrium moments (col_kbc_n.h:28-54): the shear tensors are linear in them
		const R xx = vx * vx, yy = vy * vy, zz = vz * vz;
		const R rx = rho * vx, ry = rho * vy, rz = rho * vz;
		T -= rho * (((one + xx) + yy) + zz);
		Nxz -= rho * (xx - zz);
		Nyz -= rho * (yy - zz);
		Pxy -= rx * vy;
		Pxz -= rx * vz;
		Pyz -= ry * vz;
		Qxxy -= ry * (third + xx);
		Qxxz -= rz * (third + xx);
		Qxyy -= rx * (third + yy);
		Qyyz -= rz * (third + yy);
		Qxzz -= rx * (third + zz);
		Qyzz -= ry * (third + zz);
		Qxyz -= rx * vy * vz;
	}
	else {	// col_kbc_c.h:56-83: central moments; their equilibria are 0 except the trace (rho)
		T = T - rho * (vx * vx + vy * vy + vz * vz);
		Nxz = Nxz + rho * (vz * vz - vx * vx);
		Nyz = Nyz + rho * (vz * vz - vy * vy);
		Pxy = Pxy - rho * vx * vy;
		Pxz = Pxz - rho * vx * vz;
		Pyz = Pyz - rho * vy * vz;
		const R cxx = (three * vx * vx + two * Nxz - Nyz) + T, cyy = (three * vy * vy + two * Nyz - Nxz) + T, czz = (three * vz * vz - Nyz - Nxz) + T;
		Qxxy = Qxxy - third * (six * vx * Pxy + vy * cxx);
		Qxxz = Qxxz - third * (six * vx * Pxz + vz * cxx);
		Qxyy = Qxyy - third * (six * vy * Pxy + vx * cyy);
		Qyyz = Qyyz - third * (six * vy * Pyz + vz * cyy);
		Qxzz = Qxzz - third * (six * vz * Pxz + vx * czz);
		Qyzz = Qyzz - third * (six * vz * Pyz + vy * czz);
		Qxyz = Qxyz - vx * Pyz - vy * Pxz - vz * Pxy - vx * vy * vz;
		T -= rho;
	}
	auto ds_of = [&](auto qc) -> R {
		constexpr int q = qc;
		constexpr int n = (L::cx(q) != 0) + (L::cy(q) != 0) + (L::cz(q) != 0);
		R acc = R(0);
		if constexpr (n == 1 || n == 2)
			acc = kbc_tensor_d<false, q>(Nxz, Nyz, Pxy, Pxz, Pyz);
		if constexpr (USE_T && n <= 1)
			acc = acc + (n == 0 ? -T : T * R(1.0 / 6.0));
		if constexpr (USE_Q && n >= 1)
			acc = acc + kbc_tensor_q<false, q>(Qxxy, Qxxz, Qxyy, Qyyz, Qxzz, Qyzz, Qxyz);
		return acc;
	};
	const R beta = (one / (two * P.nu / third + one));
	const R irho = rcp_fast<HW_RCP>(rho);
	R ig[3][3], nrg0[3], nig0[3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
#pragma unroll
		for (int c = 0; c < 3; c++)
			ig[a][c] = rcp_fast<HW_RCP>(g[a][c]);
		nrg0[a] = -rho * g[0][a];
		nig0[a] = -irho * ig[0][a];
	}
	// ---- pass 1: <Ds|Dh> and <Dh|Dh> (weights 1/feq); delta-h replaces f
	R sd = R(0), hh = R(0);
	static_for<9>([&](auto abc) {
		constexpr int a = abc / 3, b = abc % 3;
		const R gxy = nrg0[a] * g[1][b], igxy = nig0[a] * ig[1][b];
		static_for<3>([&](auto cc) {
			constexpr int c = cc;
			constexpr int q = L::find(a - 1, b - 1, c - 1);
			const R fe = gxy * g[2][c], ds = ds_of(std::integral_constant<int, q>{});
			const R dh = f[q] - fe - ds;
			const R t = dh * (igxy * ig[2][c]);
			sd = sd + ds * t;
			hh = hh + dh * t;
			if constexpr (! INCREMENTAL)
				f[q] = dh;
		});
	});
	const R gamma = (one / beta - (two - one / beta) * sd / hh);
#if defined(__CUDA_ARCH__) && ! defined(LBMX_KBC_NO_PASS_BARRIER)
	// keep the compiler from carrying the 27 delta-s / feq values of the first pass over to the second (common subexpressions): the point of
	// recomputing them is that they do NOT occupy registers in between
	if constexpr (sizeof(R) == 8)
		asm volatile("" : "+d"(T), "+d"(Nxz), "+d"(Nyz), "+d"(Pxy), "+d"(Pxz), "+d"(Pyz), "+d"(rho));
	else
		asm volatile("" : "+f"(T), "+f"(Nxz), "+f"(Nyz), "+f"(Pxy), "+f"(Pxz), "+f"(Pyz), "+f"(rho));
#endif
	// ---- pass 2.  fp64: the compact form on the stored delta-h.  fp32: the reference's incremental form f - (beta (2 ds + gamma dh) - (1 - beta) S feq)
	// on the untouched f (delta-h recomputed): one rounding at the magnitude of f per step instead of three -- what 1000-step fp32 runs need to stay
	// within the reference's own noise (cf. srt_update in collide.cuh); it costs the fp32 kernels ~70 of their ~700 floating-point instructions
	const R c1 = one - beta * gamma, c2 = one - two * beta;
	const R ks = three * (one - beta) * irho;
	const R F[3] = {P.fx, P.fy, P.fz};
	R sv[3][3];	 // (1 - beta) S_q = sv[0][cx+1] + sv[1][cy+1] + sv[2][cz+1]
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const R kf = ks * F[a];
		sv[a][1] = -v[a] * kf;
		sv[a][0] = sv[a][1] - kf;
		sv[a][2] = sv[a][1] + kf;
	}
#pragma unroll
	for (int a = 0; a < 3; a++)
		nrg0[a] = -rho * g[0][a];
	static_for<9>([&](auto abc) {
		constexpr int a = abc / 3, b = abc % 3;
		const R gxy = nrg0[a] * g[1][b], exy = (INCREMENTAL ? sv[0][a] : one + sv[0][a]) + sv[1][b];
		static_for<3>([&](auto cc) {
			constexpr int c = cc;
			constexpr int q = L::find(a - 1, b - 1, c - 1);
			const R fe = gxy * g[2][c], ds = ds_of(std::integral_constant<int, q>{});
			if constexpr (INCREMENTAL) {
				const R dh = f[q] - fe - ds;
				f[q] = f[q] - (((beta * gamma) * dh + (two * beta) * ds) - (exy + sv[2][c]) * fe);
			}
			else
				f[q] = c1 * f[q] + (c2 * ds + (exy + sv[2][c]) * fe);
		});
	});
}

}  // namespace ext
}  // namespace lbmx
