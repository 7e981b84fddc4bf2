// collide_ext.cuh -- further D3Q27 collision operators of the reference in the same trait slot.
//
// Written once, in the reference's floating-point association (cited per function, paths relative to the reference's
// include/lbm3d/): the default build compiles them with FMA contraction, the parity-arithmetic build (-fmad=false) reproduces the
// reference's strict CPU build bit for bit.  All of them are HBM-bound like the operators in collide.cuh; the one place where
// work is skipped in the default build is the CLBM forcing term when the body force is exactly zero (it is then an exact zero).
//
//   D3Q27_CLBM             d3q27/col_clbm.h:6-447
//   D3Q27_SRT_MODIF_FORCE  d3q27/col_srt_modif_force.h:9-120
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

}  // namespace ext
}  // namespace lbmx
