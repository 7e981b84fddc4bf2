// collide.cuh -- per-cell arithmetic of the hot path, written for registers on sm_100a.
//
// Everything here works on one cell's populations held in registers (f[Q], fully unrolled, compile-time direction
// tables).  The operators compute what the reference's trait classes compute (cited per function, paths relative to
// the reference's include/lbm3d/), but are organised for the GPU: the moment transforms are pruned to the moments the
// default cumulant build actually relaxes, all divisions by rho collapse into one reciprocal, products with the
// compile-time zeros/ones of the reference are dropped, and FMA contraction is allowed.  Agreement with the reference
// is therefore to rounding (<= 1e-12 relative in fp64 over 1000 steps, tests/test_gpu_parity.py), not bit-for-bit.
#pragma once
#include "lattice.cuh"

namespace lbmx {

template <typename R>
struct Phys
{
	R nu;		  // lbmViscosity (KS.lbmViscosity)
	R omega1;	  // 1 / (3 nu + 1/2), evaluated once on the host in precision R (col_cum.h:175, col_srt.h:19, col_bgk.h:19)
	R fx, fy, fz; // homogeneous body force (NSE_Data, lbm_data.h:87-96)
	R g17[5];	  // omega3, omega4, omega5, A, B of the 2017 cumulant parametrisation (col_cum.h:177-208): functions of nu alone, evaluated
				  // once on the host (set_rates) for the default-arithmetic kernels; the parity-arithmetic kernels evaluate them per cell as the reference does
};

// every rate that depends on the viscosity alone; call after setting P.nu
template <typename R>
inline __host__ __device__ void set_rates(Phys<R>& P)
{
	const R one = 1, two = 2, three = 3, four = 4, five = 5, six = 6, seven = 7, eight = 8, nine = 9, n10 = 10, n11 = 11, n13 = 13, n15 = 15, n16 = 16, n18 = 18, n24 = 24, n26 = 26,
			n28 = 28, n42 = 42, n46 = 46, n48 = 48, n56 = 56, n216 = 216;
	P.omega1 = one / (three * P.nu + R(0.5));  // IEEE division in precision R: the same bits the reference computes per cell
	const R omega1 = P.omega1, omega2 = one;
	P.g17[0] = eight * (omega1 - two) * (omega2 * (three * omega1 - one) - five * omega1) / (eight * (five - two * omega1) * omega1 + omega2 * (eight + omega1 * (nine * omega1 - n26)));
	P.g17[1] = eight * (omega1 - two) * (omega1 + omega2 * (three * omega1 - seven)) / (omega2 * (n56 - n42 * omega1 + nine * omega1 * omega1) - eight * omega1);
	P.g17[2] = n24 * (omega1 - two) * (four * omega1 * omega1 + omega1 * omega2 * (n18 - n13 * omega1) + omega2 * omega2 * (two + omega1 * (six * omega1 - n11)))
			 / (n16 * omega1 * omega1 * (omega1 - six) - two * omega1 * omega2 * (n216 + five * omega1 * (nine * omega1 - n46))
				+ omega2 * omega2 * (omega1 * (three * omega1 - n10) * (n15 * omega1 - n28) - n48));
	P.g17[3] = (four * omega1 * omega1 + two * omega1 * omega2 * (omega1 - six) + omega2 * omega2 * (omega1 * (n10 - three * omega1) - four)) / (omega1 - omega2)
			 / (omega2 * (two + three * omega1) - eight * omega1);
	P.g17[4] = (four * omega1 * omega2 * (nine * omega1 - n16) - four * omega1 * omega1 - two * omega2 * omega2 * (two + nine * omega1 * (omega1 - two))) / three / (omega1 - omega2)
			 / (omega2 * (two + three * omega1) - eight * omega1);
}

}  // namespace lbmx

// LBMX_STRICT=1 (with nvcc -fmad=false) selects the reference-association variants of collide_strict.cuh for this object file
#ifndef LBMX_STRICT
	#define LBMX_STRICT 0
#endif
#include "collide_strict.cuh"
#include "collide_ext.cuh"

namespace lbmx {
constexpr bool kStrict = LBMX_STRICT != 0;

// --------------------------------------------------------------------------------------------------------------------
// density / velocity: COMMON::computeDensityAndVelocity (d3q27/common.h:16-50, d2q9/common.h:16-36).
// The sums are organised as the z-column sums the cumulant transform needs anyway (the compiler merges them); the
// association differs from the reference's tree, i.e. agreement to a few ulp.
// --------------------------------------------------------------------------------------------------------------------
// KAHAN: the build with USE_HIGH_PRECISION_RHO (defs.h:252, d3q27/common.h:19-29; D3Q27 only here) -- strict::kahan_sum for the density
template <bool KAHAN = false, typename R>
LBMX_D void density_velocity(const R (&f)[27], const Phys<R>& P, R& rho, R& vx, R& vy, R& vz)
{
	using L = D3Q27;
	if constexpr (kStrict) {
		strict::density_velocity<KAHAN>(f, P, rho, vx, vy, vz);
		return;
	}
	R k0[3][3], d[3][3];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++) {
			const R fm = f[L::find(a - 1, b - 1, -1)], f0 = f[L::find(a - 1, b - 1, 0)], fp = f[L::find(a - 1, b - 1, 1)];
			k0[a][b] = (fp + fm) + f0;
			d[a][b] = fp - fm;
		}
	R A[3], By[3], Dz[3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
		A[a] = (k0[a][2] + k0[a][0]) + k0[a][1];
		By[a] = k0[a][2] - k0[a][0];
		Dz[a] = (d[a][2] + d[a][0]) + d[a][1];
	}
	rho = KAHAN ? strict::kahan_sum(f) : (A[2] + A[0]) + A[1];
	const R jx = A[2] - A[0];
	const R jy = (By[2] + By[0]) + By[1];
	const R jz = (Dz[2] + Dz[0]) + Dz[1];
	const R ir = R(1) / rho;
	vx = (jx + R(0.5) * P.fx) * ir;
	vy = (jy + R(0.5) * P.fy) * ir;
	vz = (jz + R(0.5) * P.fz) * ir;
}

template <bool KAHAN = false, typename R>
LBMX_D void density_velocity(const R (&f)[9], const Phys<R>& P, R& rho, R& vx, R& vy, R& vz)
{
	using L = D2Q9;
	if constexpr (kStrict) {
		strict::density_velocity(f, P, rho, vx, vy, vz);
		return;
	}
	R k0[3], dy[3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const R fm = f[L::find(a - 1, -1)], f0 = f[L::find(a - 1, 0)], fp = f[L::find(a - 1, 1)];
		k0[a] = (fp + fm) + f0;
		dy[a] = fp - fm;
	}
	rho = (k0[2] + k0[0]) + k0[1];
	const R ir = R(1) / rho;
	vx = ((k0[2] - k0[0]) + R(0.5) * P.fx) * ir;
	vy = (((dy[2] + dy[0]) + dy[1]) + R(0.5) * P.fy) * ir;
	vz = R(0);
}

// --------------------------------------------------------------------------------------------------------------------
// equilibria: D3Q27_EQ (d3q27/eq.h:13-130), D3Q27_EQ_INV_CUM (d3q27/eq_inv_cum.h:24-136), D2Q9_EQ (d2q9/eq.h:13-61)
// --------------------------------------------------------------------------------------------------------------------
template <typename R>
LBMX_HD constexpr R w27(int q)
{
	const int n = (D3Q27::cx(q) != 0) + (D3Q27::cy(q) != 0) + (D3Q27::cz(q) != 0);
	return n == 0 ? R(8.0 / 27.0) : n == 1 ? R(2.0 / 27.0) : n == 2 ? R(1.0 / 54.0) : R(1.0 / 216.0);
}
template <typename R>
LBMX_HD constexpr R w9(int q)
{
	const int n = (D2Q9::cx(q) != 0) + (D2Q9::cy(q) != 0);
	return n == 0 ? R(4.0 / 9.0) : n == 1 ? R(1.0 / 9.0) : R(1.0 / 36.0);
}

// all Q equilibrium populations at once (the per-axis factors are shared between directions)
template <typename R>
LBMX_D void equilibrium(R (&feq)[27], int eqkind, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	if (eqkind == 1) {	// product form +-w' rho gx gy gz with g(0) = 3v^2-2, g(+-1) = 3v^2 +- 3v + 1
		const R v[3] = {vx, vy, vz};
		R g[3][3];
#pragma unroll
		for (int a = 0; a < 3; a++) {
			const R t = R(3) * v[a] * v[a];
			g[a][1] = t - R(2);
			g[a][2] = (t + R(3) * v[a]) + R(1);
			g[a][0] = (t - R(3) * v[a]) + R(1);
		}
		static_for<27>([&](auto qc) {
			constexpr int q = qc;
			constexpr int n = (L::cx(q) != 0) + (L::cy(q) != 0) + (L::cz(q) != 0);
			constexpr R w = n == 0 ? -R(1.0 / 27.0) : n == 1 ? R(1.0 / 54.0) : n == 2 ? -R(1.0 / 108.0) : R(1.0 / 216.0);
			feq[q] = (w * rho) * ((g[0][L::cx(q) + 1] * g[1][L::cy(q) + 1]) * g[2][L::cz(q) + 1]);
		});
	}
	else {	// second-order polynomial w rho (1 - 3/2 u.u + 3 c.u + 9/2 (c.u)^2)
		const R base = R(1) - R(1.5) * ((vx * vx + vy * vy) + vz * vz);
		static_for<27>([&](auto qc) {
			constexpr int q = qc;
			const R cu = (R(L::cx(q)) * vx + R(L::cy(q)) * vy) + R(L::cz(q)) * vz;
			feq[q] = (w27<R>(q) * rho) * ((base + R(3) * cu) + (R(4.5) * cu) * cu);
		});
	}
}

// d3q27/eq_entropic.h:11-211 (default equilibrium of D3Q27_KBC_N2..N4, C2..C4): rho * W(cx) W(cy) W(cz) * prod_a (2 - s_a) *
// prod_a B_a^{c_a}, s_a = sqrt(1 + 3 v_a^2), B_a = (2 v_a + s_a) / (1 - v_a), evaluated left to right as the reference writes it.
// Its unqualified sqrt() is ::sqrt(double) in the host build: for dreal = float the chain is in double from the first (2 - s_a)
// on and rounded once at the end.  Only initialisation and boundary cells use an equilibrium of this family, so the double
// arithmetic is off the hot path.
template <typename R>
LBMX_D void equilibrium_entropic(R (&feq)[27], R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	const R v[3] = {vx, vy, vz};
	double s[3], B[3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
		s[a] = sqrt(double(R(1) + R(3) * v[a] * v[a]));
		B[a] = (double(R(2) * v[a]) + s[a]) / double(R(1) - v[a]);
	}
	static_for<27>([&](auto qc) {
		constexpr int q = qc;
		constexpr R w6 = R(1.0 / 6.0), w23 = R(2.0 / 3.0);
		constexpr R w = (L::cx(q) ? w6 : w23) * (L::cy(q) ? w6 : w23) * (L::cz(q) ? w6 : w23);
		double chain = double(w) * (2.0 - s[0]) * (2.0 - s[1]) * (2.0 - s[2]);
		static_for<3>([&](auto ac) {
			constexpr int a = ac;
			constexpr int c = a == 0 ? L::cx(q) : (a == 1 ? L::cy(q) : L::cz(q));
			if constexpr (c < 0)
				chain = chain * 1.0 / B[a];
			else if constexpr (c > 0)
				chain = chain * B[a];
		});
		feq[q] = R(double(rho) * chain);
	});
}

template <typename R>
LBMX_D void equilibrium(R (&feq)[9], int, R rho, R vx, R vy, R)
{
	using L = D2Q9;
	const R base = R(1) - R(1.5) * (vx * vx + vy * vy);
	static_for<9>([&](auto qc) {
		constexpr int q = qc;
		const R cu = R(L::cx(q)) * vx + R(L::cy(q)) * vy;
		feq[q] = (w9<R>(q) * rho) * ((base + R(3) * cu) + (R(4.5) * cu) * cu);
	});
}

// --------------------------------------------------------------------------------------------------------------------
// D3Q27 cumulant collision, default build of the reference (d3q27/col_cum.h:14-485 with omega2..omega10 = 1, A = B = 0,
// no antialias terms, col_cum.h:208-247).
//
// With every rate but omega1 equal to one, all cumulants of order >= 3 relax to zero, so of the forward transform only
// the ten moments of order <= 2 are needed (col_cum.h:52-148 pruned): 102 instead of 162 operations.  The backward
// chain (col_cum.h:312-445) is kept complete, with the identically-zero inputs removed from the first (x) level.
// `rho`, `v*` are the pre-collision density and (half-force shifted) velocity the caller computed or imposed.
// --------------------------------------------------------------------------------------------------------------------
template <bool Z0, bool Z1, bool Z2, typename R>
LBMX_D void back3(R k0, R k1, R k2, const R (&c)[6], R& fm, R& f0, R& fp)
{
	// c = {1-v^2, -2v, (v^2-v)/2, v-1/2, (v^2+v)/2, v+1/2}; Eq G2015(88)-(96), zero inputs elided at compile time
	R z = R(0), m = R(0), p = R(0);
	if constexpr (! Z2) {
		z = -k2;
		m = R(0.5) * k2;
		p = m;
	}
	if constexpr (! Z1) {
		z = c[1] * k1 + z;
		m = c[3] * k1 + m;
		p = c[5] * k1 + p;
	}
	if constexpr (! Z0) {
		z = c[0] * k0 + z;
		m = c[2] * k0 + m;
		p = c[4] * k0 + p;
	}
	f0 = z;
	fm = m;
	fp = p;
}

template <typename R>
LBMX_D void back_coeffs(R v, R (&c)[6])
{
	const R v2 = v * v;
	c[0] = R(1) - v2;
	c[1] = R(-2) * v;
	c[2] = R(0.5) * (v2 - v);
	c[3] = v - R(0.5);
	c[4] = R(0.5) * (v2 + v);
	c[5] = v + R(0.5);
}

// forward transform to the central moments of order <= 2 (col_cum.h:52-148 pruned; the same equations open col_clbm.h:18-117): all the
// default builds of D3Q27_CUM and D3Q27_CLBM relax
template <typename R>
LBMX_D void central_moments_to_second_order(const R (&f)[27], R vx, R vy, R vz, R& k000, R& k100, R& k200, R& k010, R& k110, R& k001, R& k101, R& k020, R& k002, R& k011)
{
	using L = D3Q27;
	// ---- forward, z: per (x,y) column, orders 0..2 (Eq 6-8)
	R z0[3][3], z1[3][3], z2[3][3];
	{
		const R vv = vz * vz, m2v = R(-2) * vz;
#pragma unroll
		for (int a = 0; a < 3; a++)
#pragma unroll
			for (int b = 0; b < 3; b++) {
				const R fm = f[L::find(a - 1, b - 1, -1)], f0 = f[L::find(a - 1, b - 1, 0)], fp = f[L::find(a - 1, b - 1, 1)];
				const R s = fp + fm, d = fp - fm;
				const R k0 = s + f0;
				z0[a][b] = k0;
				z1[a][b] = d - vz * k0;
				z2[a][b] = (s + m2v * d) + vv * k0;
			}
	}
	// ---- forward, y: only beta + gamma <= 2 (Eq 9-11)
	R y00[3], y10[3], y20[3], y01[3], y11[3], y02[3];
	{
		const R vv = vy * vy, m2v = R(-2) * vy;
#pragma unroll
		for (int a = 0; a < 3; a++) {
			{
				const R s = z0[a][2] + z0[a][0], d = z0[a][2] - z0[a][0];
				const R k0 = s + z0[a][1];
				y00[a] = k0;
				y10[a] = d - vy * k0;
				y20[a] = (s + m2v * d) + vv * k0;
			}
			{
				const R s = z1[a][2] + z1[a][0], d = z1[a][2] - z1[a][0];
				const R k0 = s + z1[a][1];
				y01[a] = k0;
				y11[a] = d - vy * k0;
			}
			y02[a] = (z2[a][2] + z2[a][0]) + z2[a][1];
		}
	}
	// ---- forward, x: alpha + beta + gamma <= 2 (Eq 12-14)
	{
		const R vv = vx * vx, m2v = R(-2) * vx;
		{
			const R s = y00[2] + y00[0], d = y00[2] - y00[0];
			k000 = s + y00[1];
			k100 = d - vx * k000;
			k200 = (s + m2v * d) + vv * k000;
		}
		{
			const R s = y10[2] + y10[0], d = y10[2] - y10[0];
			k010 = s + y10[1];
			k110 = d - vx * k010;
		}
		{
			const R s = y01[2] + y01[0], d = y01[2] - y01[0];
			k001 = s + y01[1];
			k101 = d - vx * k001;
		}
		k020 = (y20[2] + y20[0]) + y20[1];
		k002 = (y02[2] + y02[0]) + y02[1];
		k011 = (y11[2] + y11[0]) + y11[1];
	}
}

template <typename R>
LBMX_D void collide_cum(R (&f)[27], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	R k000, k100, k200, k010, k110, k001, k101, k020, k002, k011;
	central_moments_to_second_order(f, vx, vy, vz, k000, k100, k200, k010, k110, k001, k101, k020, k002, k011);
	// ---- relaxation (col_cum.h:175,223-256): second order with omega1, trace with omega2 = 1, first order sign flip
	const R omega1 = P.omega1;
	const R keep = R(1) - omega1;
	const R s110 = keep * k110, s101 = keep * k101, s011 = keep * k011;
	const R r33 = keep * (k200 - k020), r34 = keep * (k200 - k002), r35 = k000;
	const R third = R(1.0 / 3.0);
	const R s200 = third * ((r33 + r34) + r35);
	const R s020 = third * ((r34 - R(2) * r33) + r35);
	const R s002 = third * ((r33 - R(2) * r34) + r35);
	const R s100 = -k100, s010 = -k010, s001 = -k001;  // col_cum.h:341-345
	// ---- cumulants -> central moments, orders 4..6 (Eq G2015(81)-(84), col_cum.h:312-338); all third-order and
	//      fifth-order post-collision cumulants vanish, which removes S_12x-type moments altogether
	const R ir = R(1) / rho;
	const R s211 = (s200 * s011 + R(2) * s101 * s110) * ir;
	const R s121 = (s020 * s101 + R(2) * s110 * s011) * ir;
	const R s112 = (s002 * s110 + R(2) * s011 * s101) * ir;
	const R s220 = (s020 * s200 + R(2) * s110 * s110) * ir;
	const R s022 = (s002 * s020 + R(2) * s011 * s011) * ir;
	const R s202 = (s200 * s002 + R(2) * s101 * s101) * ir;
	const R s222 = (((s200 * s022 + s020 * s202) + s002 * s220) + R(4) * ((s011 * s211 + s101 * s121) + s110 * s112)) * ir
				 - ((R(16) * s110 * s101 * s011 + R(4) * ((s101 * s101 * s020 + s011 * s011 * s200) + s110 * s110 * s002)) + R(2) * s200 * s020 * s002)
					   * (ir * ir);
	// ---- backward, x (Eq 88-90): x[a][beta][gamma]
	R cx_[6], cy_[6], cz_[6];
	back_coeffs(vx, cx_);
	back_coeffs(vy, cy_);
	back_coeffs(vz, cz_);
	R x[3][3][3];
	back3<false, false, false>(k000, s100, s200, cx_, x[0][0][0], x[1][0][0], x[2][0][0]);
	back3<false, false, true>(s001, s101, R(0), cx_, x[0][0][1], x[1][0][1], x[2][0][1]);
	back3<false, true, false>(s002, R(0), s202, cx_, x[0][0][2], x[1][0][2], x[2][0][2]);
	back3<false, false, true>(s010, s110, R(0), cx_, x[0][1][0], x[1][1][0], x[2][1][0]);
	back3<false, true, false>(s011, R(0), s211, cx_, x[0][1][1], x[1][1][1], x[2][1][1]);
	back3<true, false, true>(R(0), s112, R(0), cx_, x[0][1][2], x[1][1][2], x[2][1][2]);
	back3<false, true, false>(s020, R(0), s220, cx_, x[0][2][0], x[1][2][0], x[2][2][0]);
	back3<true, false, true>(R(0), s121, R(0), cx_, x[0][2][1], x[1][2][1], x[2][2][1]);
	back3<false, true, false>(s022, R(0), s222, cx_, x[0][2][2], x[1][2][2], x[2][2][2]);
	// ---- backward, y (Eq 91-93) then z (Eq 94-96)
#pragma unroll
	for (int a = 0; a < 3; a++) {
		R yb[3][3];	 // [b][gamma]
#pragma unroll
		for (int g = 0; g < 3; g++)
			back3<false, false, false>(x[a][0][g], x[a][1][g], x[a][2][g], cy_, yb[0][g], yb[1][g], yb[2][g]);
#pragma unroll
		for (int b = 0; b < 3; b++) {
			R fm, f0, fp;
			back3<false, false, false>(yb[b][0], yb[b][1], yb[b][2], cz_, fm, f0, fp);
			f[L::find(a - 1, b - 1, -1)] = fm;
			f[L::find(a - 1, b - 1, 0)] = f0;
			f[L::find(a - 1, b - 1, 1)] = fp;
		}
	}
}

// --------------------------------------------------------------------------------------------------------------------
// D3Q27_CUM built with USE_GEIER_CUM_2017 and / or USE_GEIER_CUM_ANTIALIAS (defs.h:254-255; col_cum.h:177-229, 258-276), default
// arithmetic.  Same organisation as collide_cum: one reciprocal of rho, coefficient form of the backward chain, forward transform
// pruned by the compiler to the moments that are read (order <= 3 with the 2017 limiter, <= 2 without); omega3..5, A and B come from
// Phys::g17 (they depend on the viscosity alone: ~100 operations and 8 divisions per cell in the reference's form), and the limiter is
// evaluated in R (the reference's ::fabs(double) promotes it to double for dreal = float: kept in the parity-arithmetic build).
// --------------------------------------------------------------------------------------------------------------------
template <typename R>
LBMX_D void fwd3(R& lo, R& mid, R& hi, R v, R vv, R m2v)
{
	const R s = hi + lo, d = hi - lo;
	const R k0 = s + mid;
	lo = k0;
	mid = d - v * k0;
	hi = (s + m2v * d) + vv * k0;
}

template <bool G2017, bool ANTIALIAS, typename R>
LBMX_D void collide_cum_switches(R (&f)[27], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	const R one = R(1), two = R(2), three = R(3), four = R(4), half = R(0.5), third = R(1.0 / 3.0);
	R m[3][3][3];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
#pragma unroll
			for (int c = 0; c < 3; c++)
				m[a][b][c] = f[L::find(a - 1, b - 1, c - 1)];
	{
		const R vv = vz * vz, m2v = R(-2) * vz;
#pragma unroll
		for (int a = 0; a < 3; a++)
#pragma unroll
			for (int b = 0; b < 3; b++)
				fwd3(m[a][b][0], m[a][b][1], m[a][b][2], vz, vv, m2v);
	}
	{
		const R vv = vy * vy, m2v = R(-2) * vy;
#pragma unroll
		for (int a = 0; a < 3; a++)
#pragma unroll
			for (int c = 0; c < 3; c++)
				fwd3(m[a][0][c], m[a][1][c], m[a][2][c], vy, vv, m2v);
	}
	{
		const R vv = vx * vx, m2v = R(-2) * vx;
#pragma unroll
		for (int b = 0; b < 3; b++)
#pragma unroll
			for (int c = 0; c < 3; c++)
				fwd3(m[0][b][c], m[1][b][c], m[2][b][c], vx, vv, m2v);
	}
	const R omega1 = P.omega1, keep = one - omega1;
	const R ir = one / rho;
	const R s000 = m[0][0][0];
	const R s100 = -m[1][0][0], s010 = -m[0][1][0], s001 = -m[0][0][1];	 // col_cum.h:341-345
	const R s110 = keep * m[1][1][0], s101 = keep * m[1][0][1], s011 = keep * m[0][1][1];
	R s120 = 0, s102 = 0, s210 = 0, s012 = 0, s021 = 0, s201 = 0, s111 = 0;
	if constexpr (G2017) {	// col_cum.h:183-197, 258-276: limited rates on the sums and differences of the third-order cumulants
		const R lam = rho * R(0.01);
		auto relax = [&](R w, R x) -> R {
			const R ax = x < R(0) ? -x : x;
			return (one - (w + (one - w) * ax * ext::rcp_fast(lam + ax))) * x;
		};
		const R e117 = relax(P.g17[0], m[1][2][0] + m[1][0][2]), e118 = relax(P.g17[0], m[2][1][0] + m[0][1][2]), e119 = relax(P.g17[0], m[2][0][1] + m[0][2][1]);
		const R e120 = relax(P.g17[1], m[1][2][0] - m[1][0][2]), e121 = relax(P.g17[1], m[2][1][0] - m[0][1][2]), e122 = relax(P.g17[1], m[2][0][1] - m[0][2][1]);
		s120 = half * (e120 + e117);
		s102 = half * (e117 - e120);
		s210 = half * (e121 + e118);
		s012 = half * (e118 - e121);
		s021 = half * (e119 - e122);
		s201 = half * (e122 + e119);
		s111 = relax(P.g17[2], m[1][1][1]);
	}
	R r33 = keep * (m[2][0][0] - m[0][2][0]), r34 = keep * (m[2][0][0] - m[0][0][2]), r35 = m[0][0][0];
	R c220 = 0, c202 = 0, c022 = 0, c211 = 0, c121 = 0, c112 = 0;
	if constexpr (ANTIALIAS) {	// col_cum.h:215-229 and the derivative terms of Eq 33-35, 43-48 (omega2 = 1)
		const R ho = half * omega1 * ir;
		const R Dxu = -ho * ((two * m[2][0][0] - m[0][2][0]) - m[0][0][2]) - half * ir * (((m[2][0][0] + m[0][2][0]) + m[0][0][2]) - (rho - one));
		const R Dyv = Dxu + three * ho * (m[2][0][0] - m[0][2][0]);
		const R Dzw = Dxu + three * ho * (m[2][0][0] - m[0][0][2]);
		const R xx = vx * vx * Dxu, yy = vy * vy * Dyv, zz = vz * vz * Dzw;
		const R k1 = three * rho * (one - omega1 * half);
		r33 = r33 - k1 * (xx - yy);
		r34 = r34 - k1 * (xx - zz);
		r35 = r35 - (R(1.5) * rho) * ((xx + yy) + zz);
		if constexpr (G2017) {
			const R w = (one / omega1 - half) * rho;
			const R ka = R(2.0 / 3.0) * w * P.g17[3];
			const R e43 = ka * ((Dxu - two * Dyv) + Dzw), e44 = ka * ((Dxu + Dyv) - two * Dzw), e45 = -two * ka * ((Dxu + Dyv) + Dzw);
			c220 = third * ((e43 + e44) + e45);
			c202 = third * (e45 - e43);
			c022 = third * (e45 - e44);
			const R kb = (w * P.g17[4]) * (omega1 * ir);  // -1/3 (1/omega1 - 1/2) B rho * (-3 omega1 / rho * C)
			c211 = kb * m[0][1][1];
			c121 = kb * m[1][0][1];
			c112 = kb * m[1][1][0];
		}
	}
	const R s200 = third * ((r33 + r34) + r35);
	const R s020 = third * ((r34 - two * r33) + r35);
	const R s002 = third * ((r33 - two * r34) + r35);
	// cumulants -> central moments, Eq G2015(81)-(84), col_cum.h:312-338; post-collision cumulants of order 5 and 6 are zero
	const R s211 = c211 + (s200 * s011 + two * s101 * s110) * ir;
	const R s121 = c121 + (s020 * s101 + two * s110 * s011) * ir;
	const R s112 = c112 + (s002 * s110 + two * s011 * s101) * ir;
	const R s220 = c220 + (s020 * s200 + two * s110 * s110) * ir;
	const R s022 = c022 + (s002 * s020 + two * s011 * s011) * ir;
	const R s202 = c202 + (s200 * s002 + two * s101 * s101) * ir;
	R s122 = 0, s212 = 0, s221 = 0;
	R s222 = (((s200 * s022 + s020 * s202) + s002 * s220) + four * ((s011 * s211 + s101 * s121) + s110 * s112)) * ir
		   - ((R(16) * s110 * s101 * s011 + four * ((s101 * s101 * s020 + s011 * s011 * s200) + s110 * s110 * s002)) + two * s200 * s020 * s002) * (ir * ir);
	if constexpr (G2017) {
		s122 = (((s020 * s102 + s002 * s120) + four * s011 * s111) + two * (s110 * s012 + s101 * s021)) * ir;
		s212 = (((s002 * s210 + s200 * s012) + four * s101 * s111) + two * (s011 * s201 + s110 * s102)) * ir;
		s221 = (((s200 * s021 + s020 * s201) + four * s110 * s111) + two * (s101 * s120 + s011 * s210)) * ir;
		s222 = s222 + (four * s111 * s111 + two * ((s120 * s102 + s210 * s012) + s201 * s021)) * ir;
	}
	R cx_[6], cy_[6], cz_[6];
	back_coeffs(vx, cx_);
	back_coeffs(vy, cy_);
	back_coeffs(vz, cz_);
	R x[3][3][3];  // x[a][beta][gamma]
	back3<false, false, false>(s000, s100, s200, cx_, x[0][0][0], x[1][0][0], x[2][0][0]);
	back3<false, false, ! G2017>(s001, s101, s201, cx_, x[0][0][1], x[1][0][1], x[2][0][1]);
	back3<false, ! G2017, false>(s002, s102, s202, cx_, x[0][0][2], x[1][0][2], x[2][0][2]);
	back3<false, false, ! G2017>(s010, s110, s210, cx_, x[0][1][0], x[1][1][0], x[2][1][0]);
	back3<false, ! G2017, false>(s011, s111, s211, cx_, x[0][1][1], x[1][1][1], x[2][1][1]);
	back3<! G2017, false, ! G2017>(s012, s112, s212, cx_, x[0][1][2], x[1][1][2], x[2][1][2]);
	back3<false, ! G2017, false>(s020, s120, s220, cx_, x[0][2][0], x[1][2][0], x[2][2][0]);
	back3<! G2017, false, ! G2017>(s021, s121, s221, cx_, x[0][2][1], x[1][2][1], x[2][2][1]);
	back3<false, ! G2017, false>(s022, s122, s222, cx_, x[0][2][2], x[1][2][2], x[2][2][2]);
#pragma unroll
	for (int a = 0; a < 3; a++) {
		R yb[3][3];	 // [b][gamma]
#pragma unroll
		for (int g = 0; g < 3; g++)
			back3<false, false, false>(x[a][0][g], x[a][1][g], x[a][2][g], cy_, yb[0][g], yb[1][g], yb[2][g]);
#pragma unroll
		for (int b = 0; b < 3; b++) {
			R fm, f0, fp;
			back3<false, false, false>(yb[b][0], yb[b][1], yb[b][2], cz_, fm, f0, fp);
			f[L::find(a - 1, b - 1, -1)] = fm;
			f[L::find(a - 1, b - 1, 0)] = f0;
			f[L::find(a - 1, b - 1, 1)] = fp;
		}
	}
}

// --------------------------------------------------------------------------------------------------------------------
// D3Q27 cascaded operator, default build of the reference (d3q27/col_clbm.h:6-447 with omega2..omega10 = 1 and no antialias
// derivatives, col_clbm.h:119-200), default arithmetic.
//
// With those rates the post-collision central moments of order >= 3 are constants of the cell: 0, rho/9 (k220, k202, k022) and rho/27
// (k222); the third-order expressions (-a - b)/2 + (a - b)/2 + b of col_clbm.h:157-162 are zero up to the rounding of a third-order
// moment.  So the forward transform stops at order 2 like the cumulant operator's (102 instead of 162 operations) and the backward chain
// (col_clbm.h:202-300 = Eq G2015(88)-(96)) loses its zero inputs at compile time.
// The forcing term (col_clbm.h:303-443) is the population set whose raw moments are M_abc = F . grad_u (u^a v^b w^c), a, b, c <= 2, which the
// reference builds as 27 source moments times a dense 27 x 27 matrix (~600 operations).  Its CENTRAL moments about u are F . grad_u' of
// (u' - u)^a (v' - v)^b (w' - w)^c at u' = u: (Fx, Fy, Fz) at first order and zero everywhere else -- so here the force is added to the three
// first-order central moments before the backward transform, which is the same linear map applied to the same input.
// The parity-arithmetic build keeps the reference's form (ext::collide_clbm).
// --------------------------------------------------------------------------------------------------------------------
template <typename R>
LBMX_D void collide_clbm_fast(R (&f)[27], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	R k000, k100, k200, k010, k110, k001, k101, k020, k002, k011;
	central_moments_to_second_order(f, vx, vy, vz, k000, k100, k200, k010, k110, k001, k101, k020, k002, k011);
	const R keep = R(1) - P.omega1;
	const R third = R(1.0 / 3.0);
	const R d4 = keep * (k200 - k020), d5 = keep * (k200 - k002), d6 = rho;	 // col_clbm.h:147-154
	const R s200 = third * ((d4 + d5) + d6);
	const R s020 = third * ((d5 - R(2) * d4) + d6);
	const R s002 = third * ((d4 - R(2) * d5) + d6);
	const R s110 = keep * k110, s101 = keep * k101, s011 = keep * k011;	 // col_clbm.h:196-198
	const R s100 = k100 + P.fx, s010 = k010 + P.fy, s001 = k001 + P.fz;	 // col_clbm.h:191-193 and the forcing term
	const R s22 = rho * R(1.0 / 9.0), s222 = rho * R(1.0 / 27.0);			 // col_clbm.h:172-186
	R cx_[6], cy_[6], cz_[6];
	back_coeffs(vx, cx_);
	back_coeffs(vy, cy_);
	back_coeffs(vz, cz_);
	// ---- backward, x: x[a][beta][gamma]; the rows (beta, gamma) = (1,2) and (2,1) are zero
	R x[3][3][3];
	back3<false, false, false>(k000, s100, s200, cx_, x[0][0][0], x[1][0][0], x[2][0][0]);
	back3<false, false, true>(s001, s101, R(0), cx_, x[0][0][1], x[1][0][1], x[2][0][1]);
	back3<false, true, false>(s002, R(0), s22, cx_, x[0][0][2], x[1][0][2], x[2][0][2]);
	back3<false, false, true>(s010, s110, R(0), cx_, x[0][1][0], x[1][1][0], x[2][1][0]);
	back3<false, true, true>(s011, R(0), R(0), cx_, x[0][1][1], x[1][1][1], x[2][1][1]);
	back3<false, true, false>(s020, R(0), s22, cx_, x[0][2][0], x[1][2][0], x[2][2][0]);
	back3<false, true, false>(s22, R(0), s222, cx_, x[0][2][2], x[1][2][2], x[2][2][2]);
	// ---- backward, y then z
#pragma unroll
	for (int a = 0; a < 3; a++) {
		R yb[3][3];	 // [b][gamma]
		back3<false, false, false>(x[a][0][0], x[a][1][0], x[a][2][0], cy_, yb[0][0], yb[1][0], yb[2][0]);
		back3<false, false, true>(x[a][0][1], x[a][1][1], R(0), cy_, yb[0][1], yb[1][1], yb[2][1]);
		back3<false, true, false>(x[a][0][2], R(0), x[a][2][2], cy_, yb[0][2], yb[1][2], yb[2][2]);
#pragma unroll
		for (int b = 0; b < 3; b++) {
			R fm, f0, fp;
			back3<false, false, false>(yb[b][0], yb[b][1], yb[b][2], cz_, fm, f0, fp);
			f[L::find(a - 1, b - 1, -1)] = fm;
			f[L::find(a - 1, b - 1, 0)] = f0;
			f[L::find(a - 1, b - 1, 1)] = fp;
		}
	}
}

// --------------------------------------------------------------------------------------------------------------------
// D3Q27 SRT (col_srt.h:16-108), BGK (col_bgk.h:16-145, no Galilean correction), MRT_LES (col_mrt.h:13-141)
// --------------------------------------------------------------------------------------------------------------------
// SRT in default arithmetic, one routine for D3Q27 and D3Q19 (col_srt.h:16-108):
//   f' = f + ((feq - f) / tau + ((1 - 1/(2 tau)) S) feq),   S = 3 (c - u).F / rho.
// The equilibrium is evaluated population by population and never stored (f[Q] and feq[Q] together pushed the fp64 kernels to 160
// registers = 3 CTAs per SM), (1 - 1/(2 tau)) S is a sum of three per-axis values, and the equilibrium family is a compile-time argument
// (EQ: 0 = second-order polynomial, eq.h:13-130; 1 = product form, eq_inv_cum.h:24-136).  The update keeps the reference's incremental
// form: feq - f is exact where it matters (Sterbenz), so a step rounds once at the magnitude of f.  The algebraically equal
// (1 - 1/tau) f + feq (1/tau + ...) saves one operation per population and rounds three times there -- after 1000 fp32 steps the
// populations were 4e-5 off the reference instead of < 1e-5 (tests/test_gpu_parity.py::test_1000_steps_fp32_srt_and_d2q9): fp32 keeps
// the incremental form, fp64 (where three roundings at 1e-16 do not show) takes the shorter one (A-A odd kernel 5.75 -> 6.49 TB/s).
template <typename L, int EQ, typename R>
LBMX_D void srt_update(R (&f)[L::Q], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	const R itau = P.omega1;
	const R iRho = R(1) / (rho == R(0) ? R(1) : rho);
	const R pre = (R(1) - R(0.5) * itau) * (R(3) * iRho);
	const R v[3] = {vx, vy, vz}, F[3] = {P.fx, P.fy, P.fz};
	R s[3][3];	// pre (c - u).F = s[0][cx+1] + s[1][cy+1] + s[2][cz+1]
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const R kf = pre * F[a];
		const R mid = -v[a] * kf;
		s[a][1] = mid;
		s[a][0] = mid - kf;
		s[a][2] = mid + kf;
	}
	R g[3][3];	// EQ 1: per-axis factors g(0) = 3v^2 - 2, g(+-1) = 3v^2 +- 3v + 1; EQ 0: +-3v and 0
	R base = R(0);
#pragma unroll
	for (int a = 0; a < 3; a++) {
		if constexpr (EQ == 1) {
			const R t = R(3) * v[a] * v[a];
			g[a][1] = t - R(2);
			g[a][2] = (t + R(3) * v[a]) + R(1);
			g[a][0] = (t - R(3) * v[a]) + R(1);
		}
		else {
			g[a][1] = R(0);
			g[a][2] = R(3) * v[a];
			g[a][0] = R(-3) * v[a];
		}
	}
	if constexpr (EQ == 0)
		base = R(1) - R(1.5) * ((vx * vx + vy * vy) + vz * vz);
	static_for<L::Q>([&](auto qc) {
		constexpr int q = qc;
		constexpr int cx = L::cx(q), cy = L::cy(q), cz = L::cz(q), n = (cx != 0) + (cy != 0) + (cz != 0);
		R feq;
		if constexpr (EQ == 1) {
			constexpr R w = n == 0 ? -R(1.0 / 27.0) : n == 1 ? R(1.0 / 54.0) : n == 2 ? -R(1.0 / 108.0) : R(1.0 / 216.0);
			feq = (w * rho) * ((g[0][cx + 1] * g[1][cy + 1]) * g[2][cz + 1]);
		}
		else {
			constexpr R w = L::Q == 27 ? (n == 0 ? R(8.0 / 27.0) : n == 1 ? R(2.0 / 27.0) : n == 2 ? R(1.0 / 54.0) : R(1.0 / 216.0))
									   : (n == 0 ? R(1.0 / 3.0) : n == 1 ? R(1.0 / 18.0) : R(1.0 / 36.0));
			R cu3 = R(0);  // 3 c.u from the components that are not zero
			if constexpr (n > 0) {
				if constexpr (cx != 0 && cy != 0 && cz != 0)
					cu3 = (g[0][cx + 1] + g[1][cy + 1]) + g[2][cz + 1];
				else if constexpr (cx != 0 && cy != 0)
					cu3 = g[0][cx + 1] + g[1][cy + 1];
				else if constexpr (cx != 0 && cz != 0)
					cu3 = g[0][cx + 1] + g[2][cz + 1];
				else if constexpr (cy != 0 && cz != 0)
					cu3 = g[1][cy + 1] + g[2][cz + 1];
				else
					cu3 = cx != 0 ? g[0][cx + 1] : (cy != 0 ? g[1][cy + 1] : g[2][cz + 1]);
				feq = (w * rho) * ((base + cu3) + (R(0.5) * cu3) * cu3);
			}
			else
				feq = (w * rho) * base;
		}
		const R ps = (s[0][cx + 1] + s[1][cy + 1]) + s[2][cz + 1];
		if constexpr (sizeof(R) == 8)
			f[q] = (R(1) - itau) * f[q] + feq * (itau + ps);  // fp64: three roundings at 1e-16 instead of one, 1000 steps stay within 1e-12
		else
			f[q] = f[q] + ((feq - f[q]) * itau + ps * feq);
	});
}

// D3Q27_SRT_MODIF_FORCE in default arithmetic (col_srt_modif_force.h:9-120): SRT with the source term
//   (1 - 1/(2 tau)) w_q [3 (c - u).F + 9 (c.u)(c.F)]  =  w_q [(cg)(1 + cu3) - B],   cg = (1 - 1/(2 tau)) 3 c.F,  cu3 = 3 c.u,  B = (1 - 1/(2 tau)) 3 u.F
// (the reference writes the 27 brackets out with three double-precision divisions each).  Same organisation as srt_update: nothing stored
// per population, cu3 shared with the polynomial equilibrium, fp32 incremental / fp64 compact update.
template <int EQ, typename R>
LBMX_D void srt_modif_update(R (&f)[27], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	const R itau = P.omega1;
	const R pre = R(1) - R(0.5) * itau;
	const R v[3] = {vx, vy, vz}, F[3] = {P.fx, P.fy, P.fz};
	R u3[3][3], g3[3][3], g[3][3];
	R B = R(0);
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const R gf = (R(3) * pre) * F[a];
		u3[a][0] = R(-3) * v[a];
		u3[a][1] = R(0);
		u3[a][2] = R(3) * v[a];
		g3[a][0] = -gf;
		g3[a][1] = R(0);
		g3[a][2] = gf;
		B = a == 0 ? v[a] * gf : B + v[a] * gf;
		if constexpr (EQ == 1) {
			const R t = R(3) * v[a] * v[a];
			g[a][1] = t - R(2);
			g[a][2] = (t + R(3) * v[a]) + R(1);
			g[a][0] = (t - R(3) * v[a]) + R(1);
		}
	}
	const R base = R(1) - R(1.5) * ((vx * vx + vy * vy) + vz * vz);
	static_for<27>([&](auto qc) {
		constexpr int q = qc;
		constexpr int cx = L::cx(q), cy = L::cy(q), cz = L::cz(q), n = (cx != 0) + (cy != 0) + (cz != 0);
		constexpr R w = n == 0 ? R(8.0 / 27.0) : n == 1 ? R(2.0 / 27.0) : n == 2 ? R(1.0 / 54.0) : R(1.0 / 216.0);
		// sums over the components that are not zero, associated (x + y) + z so that pairs are shared between populations
		R cu3 = R(0), cg = R(0);
		if constexpr (cx != 0 && cy != 0 && cz != 0) {
			cu3 = (u3[0][cx + 1] + u3[1][cy + 1]) + u3[2][cz + 1];
			cg = (g3[0][cx + 1] + g3[1][cy + 1]) + g3[2][cz + 1];
		}
		else if constexpr (cx != 0 && cy != 0) {
			cu3 = u3[0][cx + 1] + u3[1][cy + 1];
			cg = g3[0][cx + 1] + g3[1][cy + 1];
		}
		else if constexpr (cx != 0 && cz != 0) {
			cu3 = u3[0][cx + 1] + u3[2][cz + 1];
			cg = g3[0][cx + 1] + g3[2][cz + 1];
		}
		else if constexpr (cy != 0 && cz != 0) {
			cu3 = u3[1][cy + 1] + u3[2][cz + 1];
			cg = g3[1][cy + 1] + g3[2][cz + 1];
		}
		else if constexpr (n == 1) {
			cu3 = cx != 0 ? u3[0][cx + 1] : (cy != 0 ? u3[1][cy + 1] : u3[2][cz + 1]);
			cg = cx != 0 ? g3[0][cx + 1] : (cy != 0 ? g3[1][cy + 1] : g3[2][cz + 1]);
		}
		const R src = n == 0 ? -(w * B) : w * (cg * cu3 + (cg - B));
		R feq;
		if constexpr (EQ == 1) {
			constexpr R wp = n == 0 ? -R(1.0 / 27.0) : n == 1 ? R(1.0 / 54.0) : n == 2 ? -R(1.0 / 108.0) : R(1.0 / 216.0);
			feq = (wp * rho) * ((g[0][cx + 1] * g[1][cy + 1]) * g[2][cz + 1]);
		}
		else
			feq = n == 0 ? (w * rho) * base : (w * rho) * ((base + cu3) + (R(0.5) * cu3) * cu3);
		if constexpr (sizeof(R) == 8)
			f[q] = (R(1) - itau) * f[q] + (itau * feq + src);
		else
			f[q] = f[q] + ((feq - f[q]) * itau + src);
	});
}

template <typename R>
LBMX_D void collide_srt(R (&f)[27], const Phys<R>& P, int eqkind, R rho, R vx, R vy, R vz)
{
	if (eqkind == 1)
		srt_update<D3Q27, 1>(f, P, rho, vx, vy, vz);
	else
		srt_update<D3Q27, 0>(f, P, rho, vx, vy, vz);
}

// GALILEAN: the build with USE_GALILEAN_CORRECTION (defs.h:253, col_bgk.h:20-45) -- the diagonal second moments (here from the
// z-column sums, any order: default arithmetic) correct the zero-velocity factor of each axis
template <bool GALILEAN = false, typename R>
LBMX_D void collide_bgk(R (&f)[27], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	const R omega1 = P.omega1;
	const R irho3 = R(3) / rho;
	const R pre = (R(1) - R(0.5) * omega1) * irho3;
	const R v[3] = {vx, vy, vz};
	R G[3] = {R(0), R(0), R(0)};
	if constexpr (GALILEAN) {
		R m[3] = {R(0), R(0), R(0)};
		static_for<27>([&](auto qc) {
			constexpr int q = qc;
			if constexpr (L::cx(q) != 0)
				m[0] += f[q];
			if constexpr (L::cy(q) != 0)
				m[1] += f[q];
			if constexpr (L::cz(q) != 0)
				m[2] += f[q];
		});
		const R k = (R(1) / omega1 - R(0.5)) * (R(1.5) * omega1);  // -3 v^2 D (1/omega - 1/2) with D = -omega/2 (3 m/rho - 1 - 3 v^2)
#pragma unroll
		for (int a = 0; a < 3; a++)
			G[a] = (v[a] * v[a]) * k * ((m[a] * irho3 - R(1)) - R(3) * (v[a] * v[a]));
	}
	R g[3][3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const R z = ((R(1.0 / 3.0) - R(1)) + v[a] * v[a]) + G[a];
		const R p = R(-0.5) * ((z + R(1)) + v[a]);
		g[a][1] = z;
		g[a][2] = p;
		g[a][0] = p + v[a];
	}
	// pre (c - u).F as a sum of three per-axis values; the update in the reference's incremental form (see srt_update)
	const R F[3] = {P.fx, P.fy, P.fz};
	R s[3][3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const R kf = pre * F[a];
		const R mid = -v[a] * kf;
		s[a][1] = mid;
		s[a][0] = mid - kf;
		s[a][2] = mid + kf;
	}
	static_for<27>([&](auto qc) {
		constexpr int q = qc;
		const R feq = ((-rho * g[0][L::cx(q) + 1]) * g[1][L::cy(q) + 1]) * g[2][L::cz(q) + 1];
		const R ps = (s[0][L::cx(q) + 1] + s[1][L::cy(q) + 1]) + s[2][L::cz(q) + 1];
		if constexpr (sizeof(R) == 8)
			f[q] = (R(1) - omega1) * f[q] + feq * (omega1 + ps);
		else
			f[q] = f[q] + ((feq - f[q]) * omega1 + ps * feq);
	});
}

template <typename R>
LBMX_D void collide_mrt(R (&f)[27], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	// raw second moments from the z-column sums
	R Pxx = 0, Pyy = 0, Pzz = 0, Pxy = 0, Pxz = 0, Pyz = 0;
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++) {
			const R fm = f[L::find(a - 1, b - 1, -1)], f0 = f[L::find(a - 1, b - 1, 0)], fp = f[L::find(a - 1, b - 1, 1)];
			const R s = fp + fm, d = fp - fm, k0 = s + f0;
			Pzz += s;
			if (a != 1)
				Pxx += k0;
			if (b != 1)
				Pyy += k0;
			if (a != 1 && b != 1)
				Pxy += R((a - 1) * (b - 1)) * k0;
			if (a != 1)
				Pxz += R(a - 1) * d;
			if (b != 1)
				Pyz += R(b - 1) * d;
		}
	const R Nxx = Pxx - rho * (R(1.0 / 3.0) + vx * vx);
	const R Nyy = Pyy - rho * (R(1.0 / 3.0) + vy * vy);
	const R Nzz = Pzz - rho * (R(1.0 / 3.0) + vz * vz);
	const R Nxy = Pxy - rho * vx * vy;
	const R Nxz = Pxz - rho * vx * vz;
	const R Nyz = Pyz - rho * vy * vz;
	const R Qn = R(2) * (((Nxx * Nxx + Nyy * Nyy) + Nzz * Nzz) + R(2) * ((Nxy * Nxy + Nxz * Nxz) + Nyz * Nyz));
	const R tau = R(3) * P.nu + R(0.5);
	const R Csm = R(0.0342);
	const R omega = R(2) / (sqrt(tau * tau + (R(2) * Csm * R(9)) * sqrt(Qn) / rho) + tau);	 // Smagorinsky rate
	Pxx -= omega * Nxx;
	Pyy -= omega * Nyy;
	Pzz -= omega * Nzz;
	Pxy -= omega * Nxy;
	Pxz -= omega * Nxz;
	Pyz -= omega * Nyz;
	const R tr = R(1.5) * ((Pxx + Pyy) + Pzz);
	static_for<27>([&](auto qc) {
		constexpr int q = qc;
		constexpr int a = L::cx(q), b = L::cy(q), c = L::cz(q);
		constexpr int n = (a != 0) + (b != 0) + (c != 0);
		const R lin = rho * ((R(2.5) - R(1.5) * R(n)) + R(3) * ((vx * R(a) + vy * R(b)) + vz * R(c)));
		const R quad = ((Pxx * R(a * a) + Pyy * R(b * b)) + Pzz * R(c * c)) + R(2) * ((Pxy * R(a * b) + Pxz * R(a * c)) + Pyz * R(b * c));
		f[q] = w27<R>(q) * ((lin + R(4.5) * quad) - tr);
	});
}

// --------------------------------------------------------------------------------------------------------------------
// D2Q9 SRT (d2q9/col_srt.h:16-44) and cascaded CLBM (d2q9/col_clbm.h:13-89)
// --------------------------------------------------------------------------------------------------------------------
// d2q9/col_srt.h:16-44 in default arithmetic: f' = f + (feq - f) / tau + F_q with the source term
//   F_q = (1 - 1/(2 tau)) w_q [3 (c - u).F + 9 (c.u)(c.F)] = (1 - 1/(2 tau)) w_q [(3 c.F)(1 + 3 c.u) - 3 u.F],
// so with cu3 = 3 c.u shared between the equilibrium polynomial and the source:
//   f' = f + ( (w_q rho (base + cu3 + cu3^2 / 2) - f) / tau + w_q pre ((3 c.F)(1 + cu3) - 3 u.F) )
// (the reference divides by 9 and 36 per population and evaluates all nine brackets in full: ~225 fp64 instructions per cell, which is
// most of the time of a step on an L2-resident lattice -- BASELINE configs[1]; this form takes ~130.  Incremental like the reference:
// see srt_update)
template <typename R>
LBMX_D void collide_srt(R (&f)[9], const Phys<R>& P, int, R rho, R vx, R vy, R)
{
	using L = D2Q9;
	const R itau = P.omega1;
	const R pre = R(1) - R(0.5) * itau;
	const R base = R(1) - R(1.5) * (vx * vx + vy * vy);
	const R u3[2][3] = {{R(-3) * vx, R(0), R(3) * vx}, {R(-3) * vy, R(0), R(3) * vy}};
	const R gx = (R(3) * pre) * P.fx, gy = (R(3) * pre) * P.fy;	 // pre * 3 F
	const R g3[2][3] = {{-gx, R(0), gx}, {-gy, R(0), gy}};
	const R B = vx * gx + vy * gy;	// pre * 3 u.F
	static_for<9>([&](auto qc) {
		constexpr int q = qc;
		constexpr int cx = L::cx(q), cy = L::cy(q), n = (cx != 0) + (cy != 0);
		constexpr R w = n == 0 ? R(4.0 / 9.0) : n == 1 ? R(1.0 / 9.0) : R(1.0 / 36.0);
		constexpr bool compact = sizeof(R) == 8;  // fp64: (1 - 1/tau) f + w (...), fp32: the incremental form (see srt_update)
		if constexpr (n == 0) {
			if constexpr (compact)
				f[q] = (R(1) - itau) * f[q] + w * ((itau * rho) * base - B);
			else
				f[q] = f[q] + (((w * rho) * base - f[q]) * itau - w * B);
		}
		else {
			const R cu3 = n == 2 ? u3[0][cx + 1] + u3[1][cy + 1] : (cx != 0 ? u3[0][cx + 1] : u3[1][cy + 1]);
			const R cg = n == 2 ? g3[0][cx + 1] + g3[1][cy + 1] : (cx != 0 ? g3[0][cx + 1] : g3[1][cy + 1]);
			const R poly = (base + cu3) + (R(0.5) * cu3) * cu3;
			if constexpr (compact)
				f[q] = (R(1) - itau) * f[q] + w * ((itau * rho) * poly + (cg * cu3 + (cg - B)));
			else
				f[q] = f[q] + (((w * rho) * poly - f[q]) * itau + w * (cg * cu3 + (cg - B)));
		}
	});
}

// d2q9/col_clbm.h:13-89 in default arithmetic.  Same quantities as the reference's statements, organised around the six raw moments of
// order >= 2 (the reference spells every sum out population by population: ~315 floating-point instructions per cell, this form ~170):
//   m20 = sum cx^2 f, m02, m11, m21 = sum cx^2 cy f, m12, m22   (rho and u are the caller's: imposed on inflow / outflow cells, so no
//   identity that needs rho = sum f is used -- rho - zp - zm - zz stays what it is).
// The cascade increments leave through four shared partial sums, and the Premnath-Banerjee source -- the population set with raw
// moments F . grad_u (u^a v^b), a, b <= 2 -- is separable:  S_ij = Fx a'_i b_j + Fy a_i b'_j  with a = ((u^2-u)/2, 1-u^2, (u^2+u)/2) and
// a' = (u - 1/2, -2u, u + 1/2) per axis (the inverse raw-moment transform applied to (1, u, u^2) and to its derivative), instead of eight
// source moments times a 9 x 8 table.  The update stays incremental: f + (cascade + source).
template <typename R>
LBMX_D void collide_clbm(R (&f)[9], const Phys<R>& P, R rho, R vx, R vy)
{
	using L = D2Q9;
	const R tau = R(3) * P.nu + R(0.5);
	const R fx = P.fx, fy = P.fy;
	const R zz = f[L::find(0, 0)], pz = f[L::find(1, 0)], mz = f[L::find(-1, 0)], zp = f[L::find(0, 1)], zm = f[L::find(0, -1)];
	const R pp = f[L::find(1, 1)], mm = f[L::find(-1, -1)], pm = f[L::find(1, -1)], mp = f[L::find(-1, 1)];
	const R q4t = R(0.25) / tau;
	const R dpos = pp + mm, dneg = pm + mp;			 // diagonal pairs
	const R cup = pp + mp, cdn = pm + mm;			 // cy = +1 / -1 corners
	const R crt = pp + pm, clf = mp + mm;			 // cx = +1 / -1 corners
	const R m22 = dpos + dneg;
	const R sx = pz + mz, sy = zp + zm;
	const R m20 = sx + m22, m02 = sy + m22;
	const R m11 = dpos - dneg, m21 = cup - cdn, m12 = crt - clf;
	const R vx2 = vx * vx, vy2 = vy * vy, vxy = vx * vy;
	const R Fv = fx * vx + fy * vy, Fd = fx * vx - fy * vy, Fc = fx * vy + fy * vx;
	// cascaded relaxation of the second moments (trace Pm, normal difference NE, shear V) and the higher central moments
	const R Pm = R(1.0 / 12.0) * (((rho * (vx2 + vy2) - (m20 + m02)) + R(2.0 / 3.0) * rho) - Fv);
	const R NE = q4t * (((m02 - m20) + rho * (vx2 - vy2)) - Fd);
	const R V = q4t * ((m11 - vxy * rho) + R(0.5) * Fc);
	const R P6 = R(6) * Pm, N2 = R(2) * NE;
	const R kxxyy = (((m20 - vx2 * rho) + N2) + P6) * (((m02 - vy2 * rho) - N2) + P6);
	const R UP = -(R(0.25) * ((((vy * (rho - (sy + zz)) - m21) - R(2) * vx2 * vy * rho) + R(0.5) * vx2 * fy) + fx * vxy) + (vy * R(0.5)) * (R(3) * Pm + NE) + vx * (m11 * R(0.5) - R(2) * V));
	const R RIGHT = -(R(0.25) * ((((vx * (rho - (sx + zz)) - m12) - R(2) * vy2 * vx * rho) + R(0.5) * vy2 * fx) + fy * vxy) + (vx * R(0.5)) * (R(3) * Pm - NE) + vy * (m11 * R(0.5) - R(2) * V));
	const R NP = R(0.25)
			   * ((((kxxyy - m22) - R(8) * Pm) + R(2) * (vx * (m12 - R(4) * RIGHT) + vy * (m21 - R(4) * UP))) + R(4) * vxy * (R(4) * V - m11)
				  + (vx2 * ((N2 - m02) - P6) + vy2 * (((-m20 - N2) - P6) + R(3) * vx2 * rho)) - vxy * Fc);
	// source populations, separable
	const R ax[3] = {R(0.5) * (vx2 - vx), R(1) - vx2, R(0.5) * (vx2 + vx)}, ay[3] = {R(0.5) * (vy2 - vy), R(1) - vy2, R(0.5) * (vy2 + vy)};
	const R dx[3] = {fx * (vx - R(0.5)), fx * (R(-2) * vx), fx * (vx + R(0.5))}, dy[3] = {fy * (vy - R(0.5)), fy * (R(-2) * vy), fy * (vy + R(0.5))};
	// cascade increments
	const R a = R(2) * Pm + NP, b = -Pm - R(2) * NP;
	const R aV = a + V, amV = a - V, ru = RIGHT - UP, rs = RIGHT + UP;
	const R bN = b + NE, bmN = b - NE;
	f[L::find(-1, 1)] = mp + ((aV + ru) + (dx[0] * ay[2] + ax[0] * dy[2]));
	f[L::find(1, -1)] = pm + ((aV - ru) + (dx[2] * ay[0] + ax[2] * dy[0]));
	f[L::find(-1, -1)] = mm + ((amV + rs) + (dx[0] * ay[0] + ax[0] * dy[0]));
	f[L::find(1, 1)] = pp + ((amV - rs) + (dx[2] * ay[2] + ax[2] * dy[2]));
	f[L::find(-1, 0)] = mz + ((bN - R(2) * RIGHT) + (dx[0] * ay[1] + ax[0] * dy[1]));
	f[L::find(1, 0)] = pz + ((bN + R(2) * RIGHT) + (dx[2] * ay[1] + ax[2] * dy[1]));
	f[L::find(0, -1)] = zm + ((bmN - R(2) * UP) + (dx[1] * ay[0] + ax[1] * dy[0]));
	f[L::find(0, 1)] = zp + ((bmN + R(2) * UP) + (dx[1] * ay[2] + ax[1] * dy[2]));
	f[L::find(0, 0)] = zz + (R(4) * (NP - Pm) + (dx[1] * ay[1] + ax[1] * dy[1]));
}

// --------------------------------------------------------------------------------------------------------------------
// D3Q19 (no reference implementation -- parity unpinned): the D3Q27 operators restricted to the 19-velocity set with the
// standard weights 1/3, 1/18, 1/36.  SRT with the reference's source-term form (col_srt.h) and the polynomial equilibrium
// (eq.h); "MRT_LES" = the regularised second-moment operator with Smagorinsky rate of col_mrt.h, which is lattice-agnostic.
// --------------------------------------------------------------------------------------------------------------------
template <typename R>
LBMX_HD constexpr R w19(int q)
{
	const int n = (D3Q19::cx(q) != 0) + (D3Q19::cy(q) != 0) + (D3Q19::cz(q) != 0);
	return n == 0 ? R(1.0 / 3.0) : n == 1 ? R(1.0 / 18.0) : R(1.0 / 36.0);
}

template <bool KAHAN = false, typename R>
LBMX_D void density_velocity(const R (&f)[19], const Phys<R>& P, R& rho, R& vx, R& vy, R& vz)
{
	using L = D3Q19;
	R r = f[0], jx = R(0), jy = R(0), jz = R(0);
	static_for<9>([&](auto ic) {  // opposite pairs are adjacent: (1,2), (3,4), ...
		constexpr int q = 1 + 2 * ic;
		const R s = f[q] + f[q + 1], d = f[q] - f[q + 1];
		r += s;
		if constexpr (L::cx(q) != 0)
			jx += R(L::cx(q)) * d;
		if constexpr (L::cy(q) != 0)
			jy += R(L::cy(q)) * d;
		if constexpr (L::cz(q) != 0)
			jz += R(L::cz(q)) * d;
	});
	rho = r;
	const R ir = R(1) / rho;
	vx = (jx + R(0.5) * P.fx) * ir;
	vy = (jy + R(0.5) * P.fy) * ir;
	vz = (jz + R(0.5) * P.fz) * ir;
}

template <typename R>
LBMX_D void equilibrium(R (&feq)[19], int, R rho, R vx, R vy, R vz)
{
	using L = D3Q19;
	const R base = R(1) - R(1.5) * ((vx * vx + vy * vy) + vz * vz);
	static_for<19>([&](auto qc) {
		constexpr int q = qc;
		const R cu = (R(L::cx(q)) * vx + R(L::cy(q)) * vy) + R(L::cz(q)) * vz;
		feq[q] = (w19<R>(q) * rho) * ((base + R(3) * cu) + (R(4.5) * cu) * cu);
	});
}

template <typename R>
LBMX_D void collide_srt(R (&f)[19], const Phys<R>& P, int, R rho, R vx, R vy, R vz)
{
	srt_update<D3Q19, 0>(f, P, rho, vx, vy, vz);
}

template <typename R>
LBMX_D void collide_mrt(R (&f)[19], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q19;
	R Pxx = 0, Pyy = 0, Pzz = 0, Pxy = 0, Pxz = 0, Pyz = 0;
	static_for<19>([&](auto qc) {
		constexpr int q = qc;
		constexpr int a = L::cx(q), b = L::cy(q), c = L::cz(q);
		if constexpr (a != 0)
			Pxx += f[q];
		if constexpr (b != 0)
			Pyy += f[q];
		if constexpr (c != 0)
			Pzz += f[q];
		if constexpr (a * b != 0)
			Pxy += R(a * b) * f[q];
		if constexpr (a * c != 0)
			Pxz += R(a * c) * f[q];
		if constexpr (b * c != 0)
			Pyz += R(b * c) * f[q];
	});
	const R Nxx = Pxx - rho * (R(1.0 / 3.0) + vx * vx);
	const R Nyy = Pyy - rho * (R(1.0 / 3.0) + vy * vy);
	const R Nzz = Pzz - rho * (R(1.0 / 3.0) + vz * vz);
	const R Nxy = Pxy - rho * vx * vy;
	const R Nxz = Pxz - rho * vx * vz;
	const R Nyz = Pyz - rho * vy * vz;
	const R Qn = R(2) * (((Nxx * Nxx + Nyy * Nyy) + Nzz * Nzz) + R(2) * ((Nxy * Nxy + Nxz * Nxz) + Nyz * Nyz));
	const R tau = R(3) * P.nu + R(0.5);
	const R omega = R(2) / (sqrt(tau * tau + (R(2) * R(0.0342) * R(9)) * sqrt(Qn) / rho) + tau);
	Pxx -= omega * Nxx;
	Pyy -= omega * Nyy;
	Pzz -= omega * Nzz;
	Pxy -= omega * Nxy;
	Pxz -= omega * Nxz;
	Pyz -= omega * Nyz;
	const R tr = R(1.5) * ((Pxx + Pyy) + Pzz);
	static_for<19>([&](auto qc) {
		constexpr int q = qc;
		constexpr int a = L::cx(q), b = L::cy(q), c = L::cz(q);
		constexpr int n = (a != 0) + (b != 0) + (c != 0);
		const R lin = rho * ((R(2.5) - R(1.5) * R(n)) + R(3) * ((vx * R(a) + vy * R(b)) + vz * R(c)));
		const R quad = ((Pxx * R(a * a) + Pyy * R(b * b)) + Pzz * R(c * c)) + R(2) * ((Pxy * R(a * b) + Pxz * R(a * c)) + Pyz * R(b * c));
		f[q] = w19<R>(q) * ((lin + R(4.5) * quad) - tr);
	});
}

// equilibrium for initialisation and boundary cells: every family the solver may name as EQ (eqkind 3 = entropic, D3Q27 only)
template <typename R>
LBMX_D void equilibrium_any(R (&feq)[27], int eqkind, R rho, R vx, R vy, R vz)
{
	if (eqkind == 3)
		equilibrium_entropic(feq, rho, vx, vy, vz);
	else
		equilibrium(feq, eqkind, rho, vx, vy, vz);
}
template <typename R, int Q>
LBMX_D void equilibrium_any(R (&feq)[Q], int eqkind, R rho, R vx, R vy, R vz)
{
	equilibrium(feq, eqkind, rho, vx, vy, vz);
}

// --------------------------------------------------------------------------------------------------------------------
// operator tags: what COLL means for a kernel instantiation
// HW_RCP: fp32 KBC weights from the hardware reciprocal (ext::rcp_fast); the A-B bulk kernel turns it off (see there)
// --------------------------------------------------------------------------------------------------------------------
enum CollKind : int { K_CUM = 0, K_SRT = 1, K_BGK = 2, K_MRT = 3, K_CLBM = 4 /* D2Q9_CLBM or D3Q27_CLBM, by lattice */, K_SRT_MF = 5, K_CUM_2017 = 10, K_CUM_AALIAS = 11, K_CUM_2017_AALIAS = 12 /* D3Q27_CUM built with the switches of defs.h:254-255 */,
						K_KBC_N1 = 13, K_KBC_N2, K_KBC_N3, K_KBC_N4, K_KBC_C1, K_KBC_C2, K_KBC_C3, K_KBC_C4,
						K_BGK_GAL = 21 /* D3Q27_BGK built with USE_GALILEAN_CORRECTION (defs.h:253) */,
						K_CUM_HP_RHO = 22 /* D3Q27_CUM built with USE_HIGH_PRECISION_RHO (defs.h:252): density_velocity<true>, same collision */ };

template <int KIND, bool HW_RCP = true, typename R>
LBMX_D void collide(R (&f)[27], const Phys<R>& P, int eqkind, R rho, R vx, R vy, R vz)
{
	if constexpr (KIND == K_CLBM) {
		if constexpr (kStrict)
			ext::collide_clbm<false>(f, P, rho, vx, vy, vz);
		else
			collide_clbm_fast(f, P, rho, vx, vy, vz);
	}
	else if constexpr (KIND >= K_KBC_N1 && KIND <= K_KBC_C4) {
		constexpr bool central = KIND >= K_KBC_C1, use_t = (KIND - K_KBC_N1) % 4 == 1 || (KIND - K_KBC_N1) % 4 == 3, use_q = (KIND - K_KBC_N1) % 4 >= 2;
#ifdef LBMX_KBC_REFERENCE_ORDER	 // the reference's statement order in default arithmetic (tools/kbench comparison)
		ext::collide_kbc<central, use_t, use_q, kStrict>(f, P, rho, vx, vy, vz);
#else
		if constexpr (kStrict)
			ext::collide_kbc<central, use_t, use_q, true>(f, P, rho, vx, vy, vz);
		else
			ext::collide_kbc_fast<central, use_t, use_q, HW_RCP>(f, P, rho, vx, vy, vz);
#endif
	}
	else if constexpr (KIND == K_CUM_2017 || KIND == K_CUM_AALIAS || KIND == K_CUM_2017_AALIAS) {
		constexpr bool g2017 = KIND != K_CUM_AALIAS, aalias = KIND != K_CUM_2017;
#ifdef LBMX_CUM_REFERENCE_ORDER	 // the reference's statement order in default arithmetic (tools/kbench comparison)
		strict::collide_cum<g2017, aalias>(f, P, rho, vx, vy, vz);
#else
		if constexpr (kStrict)
			strict::collide_cum<g2017, aalias>(f, P, rho, vx, vy, vz);
		else
			collide_cum_switches<g2017, aalias>(f, P, rho, vx, vy, vz);
#endif
	}
	else if constexpr (KIND == K_SRT_MF) {
		if constexpr (kStrict) {
			R feq[27];
			equilibrium(feq, eqkind, rho, vx, vy, vz);
			ext::collide_srt_modif<true>(f, feq, P, vx, vy, vz);
		}
		else if (eqkind == 1)
			srt_modif_update<1>(f, P, rho, vx, vy, vz);
		else
			srt_modif_update<0>(f, P, rho, vx, vy, vz);
	}
	else if constexpr (kStrict) {
		if constexpr (KIND == K_CUM || KIND == K_CUM_HP_RHO)
			strict::collide_cum(f, P, rho, vx, vy, vz);
		else if constexpr (KIND == K_SRT) {
			R feq[27];
			equilibrium(feq, eqkind, rho, vx, vy, vz);
			strict::collide_srt(f, feq, P, rho, vx, vy, vz);
		}
		else if constexpr (KIND == K_BGK)
			strict::collide_bgk(f, P, rho, vx, vy, vz);
		else if constexpr (KIND == K_BGK_GAL)
			strict::collide_bgk<true>(f, P, rho, vx, vy, vz);
		else
			strict::collide_mrt(f, P, rho, vx, vy, vz);
	}
	else if constexpr (KIND == K_CUM || KIND == K_CUM_HP_RHO)
		collide_cum(f, P, rho, vx, vy, vz);
	else if constexpr (KIND == K_SRT)
		collide_srt(f, P, eqkind, rho, vx, vy, vz);
	else if constexpr (KIND == K_BGK)
		collide_bgk(f, P, rho, vx, vy, vz);
	else if constexpr (KIND == K_BGK_GAL)
		collide_bgk<true>(f, P, rho, vx, vy, vz);
	else
		collide_mrt(f, P, rho, vx, vy, vz);
}
template <int KIND, bool HW_RCP = true, typename R>
LBMX_D void collide(R (&f)[19], const Phys<R>& P, int eqkind, R rho, R vx, R vy, R vz)
{
	if constexpr (KIND == K_SRT)
		collide_srt(f, P, eqkind, rho, vx, vy, vz);
	else
		collide_mrt(f, P, rho, vx, vy, vz);
}
template <int KIND, bool HW_RCP = true, typename R>
LBMX_D void collide(R (&f)[9], const Phys<R>& P, int eqkind, R rho, R vx, R vy, R vz)
{
	if constexpr (kStrict) {
		if constexpr (KIND == K_SRT) {
			R feq[9];
			equilibrium(feq, 0, rho, vx, vy, R(0));
			strict::collide_srt(f, feq, P, vx, vy);
		}
		else
			strict::collide_clbm(f, P, rho, vx, vy);
	}
	else if constexpr (KIND == K_SRT)
		collide_srt(f, P, eqkind, rho, vx, vy, vz);
	else
		collide_clbm(f, P, rho, vx, vy);
}

}  // namespace lbmx
