// collide_strict.cuh -- "parity arithmetic": the per-cell operators with the reference's floating-point association.
//
// Selected per object file with -DLBMX_STRICT=1 together with nvcc -fmad=false (tnl_lbm_b200/build.py builds the fp32 kernel
// families this way; lbmx_desc.flags & LBMX_FLAG_STRICT_ARITH picks them).  Why it exists: in fp32 the reference is not
// reproducible to the 1e-5 tolerance against ITSELF once the compiler may contract a*b+c into FMAs or the sums are re-associated
// (its strict and its contracted CPU builds differ by 2e-5 in velocity after 1000 steps).  These variants evaluate every
// expression in the order the reference writes it (cited per function, paths relative to the reference's include/lbm3d/), with
// true divisions, so that the fp32 engine agrees with the reference's strict build to rounding of the few terms that differ:
// the identically-zero products that are dropped, and the <= 1e-9-relative residues the reference carries in its third-order
// cumulants (col_cum.h:278-285 with omega3 = omega4 = 1).  The fast variants in collide.cuh remain the default.
#pragma once
#include "lattice.cuh"

namespace lbmx {

template <typename R>
struct Phys;

namespace strict {

// the density of a build with USE_HIGH_PRECISION_RHO (defs.h:252; d3q27/common.h:19-29): compensated (Kahan) sum over the populations in
// index order.  No multiplications: FMA contraction cannot touch it, and nvcc does not reassociate without -use_fast_math.
template <typename R, int Q>
LBMX_D R kahan_sum(const R (&f)[Q])
{
	R s = R(0), c = R(0);
	static_for<Q>([&](auto qc) {
		const R y = f[qc] - c;
		const R t = s + y;
		c = (t - s) - y;
		s = t;
	});
	return s;
}

// d3q27/common.h:16-50 -- s(q) = f[q] + f[opp q], d(q) = f[q] - f[opp q]
template <bool KAHAN = false, typename R>
LBMX_D void density_velocity(const R (&f)[27], const Phys<R>& P, R& rho, R& vx, R& vy, R& vz)
{
	using L = D3Q27;
#define SQ(a, b, c) (f[L::find(a, b, c)] + f[L::find(-(a), -(b), -(c))])
#define DQ(a, b, c) (f[L::find(a, b, c)] - f[L::find(-(a), -(b), -(c))])
	const R corners = (SQ(1, 1, 1) + SQ(1, -1, 1)) + (SQ(1, 1, -1) + SQ(-1, 1, 1));
	const R edges = ((SQ(0, 1, 1) + SQ(0, 1, -1)) + (SQ(1, 0, 1) + SQ(1, 0, -1))) + (SQ(1, 1, 0) + SQ(1, -1, 0));
	const R axes = (SQ(1, 0, 0) + SQ(0, 1, 0)) + SQ(0, 0, 1);
	rho = KAHAN ? kahan_sum(f) : ((corners + edges) + axes) + f[L::find(0, 0, 0)];
	const R half = R(0.5);
	const R cz = (DQ(1, 1, 1) + DQ(-1, 1, 1)) + (DQ(1, -1, 1) + DQ(-1, -1, 1));
	const R ez = (DQ(0, 1, 1) + DQ(0, -1, 1)) + (DQ(1, 0, 1) + DQ(-1, 0, 1));
	vz = (((cz + ez) + DQ(0, 0, 1)) + P.fz * half) / rho;
	const R cx = (DQ(1, 1, 1) + DQ(1, -1, 1)) + (DQ(1, 1, -1) + DQ(1, -1, -1));
	const R ex = (DQ(1, 0, 1) + DQ(1, 0, -1)) + (DQ(1, 1, 0) + DQ(1, -1, 0));
	vx = (((cx + ex) + DQ(1, 0, 0)) + P.fx * half) / rho;
	const R cy = (DQ(1, 1, 1) + DQ(1, 1, -1)) + (DQ(-1, 1, 1) + DQ(-1, 1, -1));
	const R ey = (DQ(1, 1, 0) + DQ(-1, 1, 0)) + (DQ(0, 1, 1) + DQ(0, 1, -1));
	vy = (((cy + ey) + DQ(0, 1, 0)) + P.fy * half) / rho;
#undef SQ
#undef DQ
}

// d2q9/common.h:16-36
template <bool KAHAN = false, typename R>
LBMX_D void density_velocity(const R (&f)[9], const Phys<R>& P, R& rho, R& vx, R& vy, R& vz)
{
	using L = D2Q9;
#define F2(a, b) f[L::find(a, b)]
	const R half = R(0.5);
	rho = F2(0, 0) + (((F2(1, 0) + F2(-1, 0)) + (F2(0, -1) + F2(0, 1))) + ((F2(1, 1) + F2(-1, -1)) + (F2(-1, 1) + F2(1, -1))));
	vx = (((F2(1, 0) - F2(-1, 0)) + ((F2(1, -1) - F2(-1, 1)) + (F2(1, 1) - F2(-1, -1)))) + half * P.fx) / rho;
	vy = (((F2(0, 1) - F2(0, -1)) + ((F2(-1, 1) - F2(1, -1)) + (F2(1, 1) - F2(-1, -1)))) + half * P.fy) / rho;
	vz = R(0);
#undef F2
}

// Eq 6-8 / 9-11 / 12-14 (col_cum.h:52-148): (f-, f0, f+) -> (k0, k1, k2) along one axis
template <typename R>
LBMX_D void to_central(R& lo, R& mid, R& hi, R v)
{
	const R fm = lo, fz = mid, fp = hi;
	const R k0 = (fp + fm) + fz;
	lo = k0;
	mid = (fp - fm) - v * k0;
	hi = (fp + fm) - R(2) * v * (fp - fm) + v * v * k0;
}

// Eq G2015(88)-(96), col_cum.h:349-445: (k0, k1, k2) -> (f-, f0, f+) along one axis
template <typename R>
LBMX_D void from_central(R& lo, R& mid, R& hi, R v)
{
	const R k0 = lo, k1 = mid, k2 = hi;
	const R one = R(1), two = R(2), half = R(0.5);
	mid = k0 * (one - v * v) - two * v * k1 - k2;
	lo = (k0 * (v * v - v) + k1 * (two * v - one) + k2) * half;
	hi = (k0 * (v * v + v) + k1 * (two * v + one) + k2) * half;
}

// rate limiter of the 2017 parametrisation (col_cum.h:183-197): w + (1 - w) * fabs(x) / (rho * lambda + fabs(x)).  The reference's
// unqualified fabs() is ::fabs(double) in the host build: for dreal = float everything downstream of it is evaluated in double and
// rounded once when stored (like sqrt() in col_mrt.h).
template <typename R>
LBMX_D R limited_rate(R w, R x, R rho, R lambda)
{
	const double ax = fabs(double(x));
	return R(double(w) + double(R(1) - w) * ax / (double(rho * lambda) + ax));
}

// d3q27/col_cum.h:14-485, evaluated the way the reference evaluates it: all 27 central moments, and the third-order cumulants
// keep the rounding residue that (-a-b)/2 + (a-b)/2 + b leaves behind (col_cum.h:278-285) -- dropping it changes single bits, and a
// single bit is enough for two fp32 runs to drift apart like two different roundings.  Only terms that are exactly +-0 for finite
// inputs are left out: (1 - omega_n) * C with omega_n = 1, and products with A, B or a velocity derivative that is a constant 0.
//   G2017     = the reference compiled with -DUSE_GEIER_CUM_2017 (defs.h:254): parametrised omega3..5 with limiter, A and B
//   ANTIALIAS = ... with -DUSE_GEIER_CUM_ANTIALIAS (defs.h:255): velocity-derivative terms of Eq 33-35, 43-48
// Default build of the reference: both off (omega2..omega10 = 1, A = B = 0).
template <bool G2017 = false, bool ANTIALIAS = false, typename R, typename PHYS>
LBMX_D void collide_cum(R (&f)[27], const PHYS& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	const R one = R(1), two = R(2), three = R(3), four = R(4), half = R(0.5), third = R(1.0 / 3.0);
	R m[3][3][3];  // index 0,1,2 = velocity sign -,0,+ before an axis is transformed, moment order 0,1,2 after
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
#pragma unroll
			for (int c = 0; c < 3; c++)
				m[a][b][c] = f[L::find(a - 1, b - 1, c - 1)];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
			to_central(m[a][b][0], m[a][b][1], m[a][b][2], vz);
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int c = 0; c < 3; c++)
			to_central(m[a][0][c], m[a][1][c], m[a][2][c], vy);
#pragma unroll
	for (int b = 0; b < 3; b++)
#pragma unroll
		for (int c = 0; c < 3; c++)
			to_central(m[0][b][c], m[1][b][c], m[2][b][c], vx);

	const R omega1 = one / (three * P.nu + half);
	const R omega2 = one;
	const R keep = one - omega1;
	R A = R(0), B = R(0);
	R S[3][3][3];
	S[0][0][0] = m[0][0][0];
	S[1][0][0] = -m[1][0][0];  // col_cum.h:341-345
	S[0][1][0] = -m[0][1][0];
	S[0][0][1] = -m[0][0][1];
	S[1][1][0] = keep * m[1][1][0];
	S[1][0][1] = keep * m[1][0][1];
	S[0][1][1] = keep * m[0][1][1];
	// cumulants of order <= 3 equal the central moments (col_cum.h:151-171 only changes orders 4-6, which relax to 0 here)
	if constexpr (G2017) {	// col_cum.h:177-208, 258-276
		const R five = 5, six = 6, seven = 7, eight = 8, nine = 9, n10 = 10, n11 = 11, n13 = 13, n15 = 15, n16 = 16, n18 = 18, n24 = 24, n26 = 26, n28 = 28,
				n42 = 42, n46 = 46, n48 = 48, n56 = 56, n216 = 216;
		const R lambda3 = R(0.01), lambda4 = R(0.01), lambda5 = R(0.01);
		const R omega3 = eight * (omega1 - two) * (omega2 * (three * omega1 - one) - five * omega1)
					   / (eight * (five - two * omega1) * omega1 + omega2 * (eight + omega1 * (nine * omega1 - n26)));
		const R omega4 = eight * (omega1 - two) * (omega1 + omega2 * (three * omega1 - seven)) / (omega2 * (n56 - n42 * omega1 + nine * omega1 * omega1) - eight * omega1);
		const R omega5 = n24 * (omega1 - two)
					   * (four * omega1 * omega1 + omega1 * omega2 * (n18 - n13 * omega1) + omega2 * omega2 * (two + omega1 * (six * omega1 - n11)))
					   / (n16 * omega1 * omega1 * (omega1 - six) - two * omega1 * omega2 * (n216 + five * omega1 * (nine * omega1 - n46))
						  + omega2 * omega2 * (omega1 * (three * omega1 - n10) * (n15 * omega1 - n28) - n48));
		A = (four * omega1 * omega1 + two * omega1 * omega2 * (omega1 - six) + omega2 * omega2 * (omega1 * (n10 - three * omega1) - four)) / (omega1 - omega2)
		  / (omega2 * (two + three * omega1) - eight * omega1);
		B = (four * omega1 * omega2 * (nine * omega1 - n16) - four * omega1 * omega1 - two * omega2 * omega2 * (two + nine * omega1 * (omega1 - two))) / three
		  / (omega1 - omega2) / (omega2 * (two + three * omega1) - eight * omega1);
		const R e117 = (one - limited_rate(omega3, m[1][2][0] + m[1][0][2], rho, lambda3)) * (m[1][2][0] + m[1][0][2]);
		const R e118 = (one - limited_rate(omega3, m[2][1][0] + m[0][1][2], rho, lambda3)) * (m[2][1][0] + m[0][1][2]);
		const R e119 = (one - limited_rate(omega3, m[2][0][1] + m[0][2][1], rho, lambda3)) * (m[2][0][1] + m[0][2][1]);
		const R e120 = (one - limited_rate(omega4, m[1][2][0] - m[1][0][2], rho, lambda4)) * (m[1][2][0] - m[1][0][2]);
		const R e121 = (one - limited_rate(omega4, m[2][1][0] - m[0][1][2], rho, lambda4)) * (m[2][1][0] - m[0][1][2]);
		const R e122 = (one - limited_rate(omega4, m[2][0][1] - m[0][2][1], rho, lambda4)) * (m[2][0][1] - m[0][2][1]);
		S[1][2][0] = half * (e120 + e117);
		S[1][0][2] = half * (-e120 + e117);
		S[2][1][0] = half * (e121 + e118);
		S[0][1][2] = half * (-e121 + e118);
		S[0][2][1] = half * (-e122 + e119);
		S[2][0][1] = half * (e122 + e119);
		S[1][1][1] = (one - limited_rate(omega5, m[1][1][1], rho, lambda5)) * m[1][1][1];
	}
	else {	// Eq 36-41 with omega3 = omega4 = 1 (x * 1 is exact), Eq 42 with omega5 = 1
		S[1][2][0] = (-m[1][0][2] - m[1][2][0]) * half + (m[1][0][2] - m[1][2][0]) * half + m[1][2][0];
		S[1][0][2] = (-m[1][0][2] - m[1][2][0]) * half + (-m[1][0][2] + m[1][2][0]) * half + m[1][0][2];
		S[2][1][0] = (-m[0][1][2] - m[2][1][0]) * half + (m[0][1][2] - m[2][1][0]) * half + m[2][1][0];
		S[0][1][2] = (-m[0][1][2] - m[2][1][0]) * half + (-m[0][1][2] + m[2][1][0]) * half + m[0][1][2];
		S[0][2][1] = (-m[0][2][1] - m[2][0][1]) * half + (-m[0][2][1] + m[2][0][1]) * half + m[0][2][1];
		S[2][0][1] = (-m[0][2][1] - m[2][0][1]) * half + (m[0][2][1] - m[2][0][1]) * half + m[2][0][1];
		S[1][1][1] = R(0);
	}
	R r33 = keep * (m[2][0][0] - m[0][2][0]), r34 = keep * (m[2][0][0] - m[0][0][2]), r35 = m[0][0][0];
	R c220 = R(0), c202 = R(0), c022 = R(0), c211 = R(0), c121 = R(0), c112 = R(0);	 // post-collision cumulants of order 4
	if constexpr (ANTIALIAS) {	// col_cum.h:215-229 and the derivative terms of Eq 33-35, 43-48
		const R n3o2 = R(1.5), n2o3 = R(2.0 / 3.0), n4o3 = R(4.0 / 3.0);
		const R Dxu = -omega1 / two / rho * (two * m[2][0][0] - m[0][2][0] - m[0][0][2]) - omega2 / two / rho * (m[2][0][0] + m[0][2][0] + m[0][0][2] - (-one + rho));
		const R Dyv = Dxu + n3o2 * omega1 / rho * (m[2][0][0] - m[0][2][0]);
		const R Dzw = Dxu + n3o2 * omega1 / rho * (m[2][0][0] - m[0][0][2]);
		r33 = r33 - three * rho * (one - omega1 * half) * (vx * vx * Dxu - vy * vy * Dyv);
		r34 = r34 - three * rho * (one - omega1 * half) * (vx * vx * Dxu - vz * vz * Dzw);
		r35 = r35 - three * rho * (one - omega2 / two) * (vx * vx * Dxu + vy * vy * Dyv + vz * vz * Dzw);
		if constexpr (G2017) {	// A, B != 0 only in the 2017 parametrisation
			const R DxvDyu = -three * omega1 / rho * m[1][1][0];
			const R DxwDzu = -three * omega1 / rho * m[1][0][1];
			const R DywDzv = -three * omega1 / rho * m[0][1][1];
			const R e43 = n2o3 * (one / omega1 - half) * A * rho * (Dxu - two * Dyv + Dzw);
			const R e44 = n2o3 * (one / omega1 - half) * A * rho * (Dxu + Dyv - two * Dzw);
			const R e45 = -n4o3 * (one / omega1 - half) * A * rho * (Dxu + Dyv + Dzw);
			c220 = third * (e43 + e44 + e45);
			c202 = third * (-e43 + e45);
			c022 = third * (-e44 + e45);
			c211 = -third * (one / omega1 - half) * B * rho * DywDzv;
			c121 = -third * (one / omega1 - half) * B * rho * DxwDzu;
			c112 = -third * (one / omega1 - half) * B * rho * DxvDyu;
		}
	}
	S[2][0][0] = third * (r33 + r34 + r35);
	S[0][2][0] = third * (-two * r33 + r34 + r35);
	S[0][0][2] = third * (r33 - two * r34 + r35);
	// Eq G2015(81)-(84), col_cum.h:312-338; the post-collision cumulants of order 5 and 6 are +-0
#define s(a, b, c) S[a][b][c]
	S[2][1][1] = c211 + (s(2, 0, 0) * s(0, 1, 1) + two * s(1, 0, 1) * s(1, 1, 0)) / rho;
	S[1][2][1] = c121 + (s(0, 2, 0) * s(1, 0, 1) + two * s(1, 1, 0) * s(0, 1, 1)) / rho;
	S[1][1][2] = c112 + (s(0, 0, 2) * s(1, 1, 0) + two * s(0, 1, 1) * s(1, 0, 1)) / rho;
	S[2][2][0] = c220 + (s(0, 2, 0) * s(2, 0, 0) + two * s(1, 1, 0) * s(1, 1, 0)) / rho;
	S[0][2][2] = c022 + (s(0, 0, 2) * s(0, 2, 0) + two * s(0, 1, 1) * s(0, 1, 1)) / rho;
	S[2][0][2] = c202 + (s(2, 0, 0) * s(0, 0, 2) + two * s(1, 0, 1) * s(1, 0, 1)) / rho;
	S[1][2][2] = (s(0, 2, 0) * s(1, 0, 2) + s(0, 0, 2) * s(1, 2, 0) + four * s(0, 1, 1) * s(1, 1, 1) + two * (s(1, 1, 0) * s(0, 1, 2) + s(1, 0, 1) * s(0, 2, 1))) / rho;
	S[2][1][2] = (s(0, 0, 2) * s(2, 1, 0) + s(2, 0, 0) * s(0, 1, 2) + four * s(1, 0, 1) * s(1, 1, 1) + two * (s(0, 1, 1) * s(2, 0, 1) + s(1, 1, 0) * s(1, 0, 2))) / rho;
	S[2][2][1] = (s(2, 0, 0) * s(0, 2, 1) + s(0, 2, 0) * s(2, 0, 1) + four * s(1, 1, 0) * s(1, 1, 1) + two * (s(1, 0, 1) * s(1, 2, 0) + s(0, 1, 1) * s(2, 1, 0))) / rho;
	S[2][2][2] = (four * s(1, 1, 1) * s(1, 1, 1) + s(2, 0, 0) * s(0, 2, 2) + s(0, 2, 0) * s(2, 0, 2) + s(0, 0, 2) * s(2, 2, 0)
				  + four * (s(0, 1, 1) * s(2, 1, 1) + s(1, 0, 1) * s(1, 2, 1) + s(1, 1, 0) * s(1, 1, 2))
				  + two * (s(1, 2, 0) * s(1, 0, 2) + s(2, 1, 0) * s(0, 1, 2) + s(2, 0, 1) * s(0, 2, 1)))
					 / rho
			   - (R(16) * s(1, 1, 0) * s(1, 0, 1) * s(0, 1, 1)
				  + four * (s(1, 0, 1) * s(1, 0, 1) * s(0, 2, 0) + s(0, 1, 1) * s(0, 1, 1) * s(2, 0, 0) + s(1, 1, 0) * s(1, 1, 0) * s(0, 0, 2))
				  + two * s(2, 0, 0) * s(0, 2, 0) * s(0, 0, 2))
					 / rho / rho;
#undef s
#pragma unroll
	for (int b = 0; b < 3; b++)
#pragma unroll
		for (int c = 0; c < 3; c++)
			from_central(S[0][b][c], S[1][b][c], S[2][b][c], vx);
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int c = 0; c < 3; c++)
			from_central(S[a][0][c], S[a][1][c], S[a][2][c], vy);
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
			from_central(S[a][b][0], S[a][b][1], S[a][b][2], vz);
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = 0; b < 3; b++)
#pragma unroll
			for (int c = 0; c < 3; c++)
				f[L::find(a - 1, b - 1, c - 1)] = S[a][b][c];
}

// source factor 3 ((c-u).F), col_srt.h:25-51 / col_bgk.h:62-88
template <typename R>
LBMX_D R force_projection(int cx, int cy, int cz, R vx, R vy, R vz, const Phys<R>& P)
{
	const R one = R(1);
	const R tx = cx < 0 ? (-vx - one) * P.fx : (cx > 0 ? (-vx + one) * P.fx : -vx * P.fx);
	const R ty = cy < 0 ? (-vy - one) * P.fy : (cy > 0 ? (-vy + one) * P.fy : -vy * P.fy);
	const R tz = cz < 0 ? (-vz - one) * P.fz : (cz > 0 ? (-vz + one) * P.fz : -vz * P.fz);
	return R(3) * (tx + ty + tz);
}

// col_srt.h:16-108 -- feq[] supplied by the caller (equilibrium() in collide.cuh is already in the reference's association)
template <typename R>
LBMX_D void collide_srt(R (&f)[27], const R (&feq)[27], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	const R one = R(1), half = R(0.5);
	const R tau = R(3) * P.nu + half;
	const R iRho = one / (rho == R(0) ? one : rho);
	static_for<27>([&](auto qc) {
		constexpr int q = qc;
		const R S = force_projection(L::cx(q), L::cy(q), L::cz(q), vx, vy, vz, P) * iRho;
		f[q] += (feq[q] - f[q]) / tau + (one - half / tau) * S * feq[q];
	});
}

// col_bgk.h:16-145; GALILEAN: built with USE_GALILEAN_CORRECTION (defs.h:253, col_bgk.h:20-45), second moments summed in the
// reference's order: m_200 over x in (-,+) with (y,z) = mm mp mz pm pp pz zm zp zz; m_020 over x in (-,0,+), y in (-,+), z in (-,+,0);
// m_002 over x, y in (-,0,+), z in (-,+)
template <bool GALILEAN = false, typename R>
LBMX_D void collide_bgk(R (&f)[27], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	const R one = R(1), half = R(0.5), third = R(1.0 / 3.0), three = R(3);
	const R omega1 = one / (R(3) * P.nu + half);
	const R v[3] = {vx, vy, vz};
	R G[3] = {R(0), R(0), R(0)};
	if constexpr (GALILEAN) {
		R m[3];
		static_for<18>([&](auto ic) {
			constexpr int i = ic;
			constexpr int s3[3] = {-1, 1, 0}, x3[3] = {-1, 0, 1};
			constexpr int o9[9][2] = {{-1, -1}, {-1, 1}, {-1, 0}, {1, -1}, {1, 1}, {1, 0}, {0, -1}, {0, 1}, {0, 0}};
			constexpr int q0 = L::find(i < 9 ? -1 : 1, o9[i % 9][0], o9[i % 9][1]);
			constexpr int q1 = L::find(x3[i / 6], (i / 3) % 2 == 0 ? -1 : 1, s3[i % 3]);
			constexpr int q2 = L::find(x3[i / 6], x3[(i / 2) % 3], i % 2 == 0 ? -1 : 1);
			if constexpr (i == 0) {
				m[0] = f[q0];
				m[1] = f[q1];
				m[2] = f[q2];
			}
			else {
				m[0] = m[0] + f[q0];
				m[1] = m[1] + f[q1];
				m[2] = m[2] + f[q2];
			}
		});
#pragma unroll
		for (int a = 0; a < 3; a++) {
			const R D = -omega1 * half * (three * m[a] / rho - one - three * v[a] * v[a]);
			G[a] = -three * v[a] * v[a] * D * (one / omega1 - half);
		}
	}
	R g[3][3];
#pragma unroll
	for (int a = 0; a < 3; a++) {
		const R z = GALILEAN ? third - one + v[a] * v[a] + G[a] : third - one + v[a] * v[a];
		const R p = -half * (z + one + v[a]);
		g[a][1] = z;
		g[a][2] = p;
		g[a][0] = p + v[a];
	}
	static_for<27>([&](auto qc) {
		constexpr int q = qc;
		const R S = force_projection(L::cx(q), L::cy(q), L::cz(q), vx, vy, vz, P) / rho;
		const R feq = -rho * g[0][L::cx(q) + 1] * g[1][L::cy(q) + 1] * g[2][L::cz(q) + 1];
		f[q] += (feq - f[q]) * omega1 + (one - half * omega1) * S * feq;
	});
}

// col_mrt.h:13-141 ("MRT_LES")
template <typename R>
LBMX_D void collide_mrt(R (&f)[27], const Phys<R>& P, R rho, R vx, R vy, R vz)
{
	using L = D3Q27;
	const R two = R(2), three = R(3), third = R(1.0 / 3.0);
	// running sums in lexicographic (x,y,z) order over the signs -,0,+ (col_mrt.h:18-31)
	R Pxx = 0, Pyy = 0, Pzz = 0, Pxy = 0, Pxz = 0, Pyz = 0;
	static_for<27>([&](auto ic) {
		constexpr int i = ic;
		constexpr int a = i / 9 - 1, b = (i / 3) % 3 - 1, c = i % 3 - 1;
		const R v = f[L::find(a, b, c)];
		if constexpr (a != 0)
			Pxx = Pxx + v;
		if constexpr (b != 0)
			Pyy = Pyy + v;
		if constexpr (c != 0)
			Pzz = Pzz + v;
		if constexpr (a * b > 0)
			Pxy = Pxy + v;
		if constexpr (a * b < 0)
			Pxy = Pxy - v;
		if constexpr (a * c > 0)
			Pxz = Pxz + v;
		if constexpr (a * c < 0)
			Pxz = Pxz - v;
		if constexpr (b * c > 0)
			Pyz = Pyz + v;
		if constexpr (b * c < 0)
			Pyz = Pyz - v;
	});
	const R Nxx = Pxx - rho * (third + vx * vx);
	const R Nyy = Pyy - rho * (third + vy * vy);
	const R Nzz = Pzz - rho * (third + vz * vz);
	const R Nxz = Pxz - rho * vx * vz;
	const R Nxy = Pxy - rho * vx * vy;
	const R Nyz = Pyz - rho * vy * vz;
	const R Qn = two * (Nxx * Nxx + Nyy * Nyy + Nzz * Nzz + two * (Nxy * Nxy + Nxz * Nxz + Nyz * Nyz));
	const R tau = three * P.nu + R(0.5);
	const R Csm = R(0.0342);
	// the reference's unqualified sqrt() is ::sqrt(double) in the host build this mode is pinned to, also for dreal = float:
	// the rate is evaluated in double and rounded once (col_mrt.h:63-66)
	const double inner = (double) (tau * tau) + (double) (two * Csm * three * three) * ::sqrt((double) Qn) / (double) rho;
	const R omega = (R) ((double) two / (::sqrt(inner) + (double) tau));
	Pxx -= omega * Nxx;
	Pyy -= omega * Nyy;
	Pzz -= omega * Nzz;
	Pxy -= omega * Nxy;
	Pxz -= omega * Nxz;
	Pyz -= omega * Nyz;
	static_for<27>([&](auto qc) {
		constexpr int q = qc;
		constexpr int a = L::cx(q), b = L::cy(q), c = L::cz(q);
		constexpr int n = (a != 0) + (b != 0) + (c != 0);
		constexpr R w = n == 0 ? R(8.0 / 27.0) : n == 1 ? R(2.0 / 27.0) : n == 2 ? R(1.0 / 54.0) : R(1.0 / 216.0);
		f[q] = w
			 * (rho * (R(2.5) - R(1.5) * R(n) + three * (vx * R(a) + vy * R(b) + vz * R(c)))
				+ R(4.5) * (Pxx * R(a * a) + Pyy * R(b * b) + Pzz * R(c * c) + two * (Pxy * R(a * b) + Pxz * R(a * c) + Pyz * R(b * c))) - R(1.5) * (Pxx + Pyy + Pzz));
	});
}

// d2q9/col_srt.h:16-44
template <typename R>
LBMX_D void collide_srt(R (&f)[9], const R (&feq)[9], const Phys<R>& P, R vx, R vy)
{
	using L = D2Q9;
	const R one = R(1), half = R(0.5), three = R(3), four = R(4), nine = R(9), n36 = R(36);
	const R tau = three * P.nu + half;
	const R fx = P.fx, fy = P.fy;
	const R pre = one - half / tau;
	R F[9];
	F[L::find(0, 0)] = pre * four / nine * (three * (-vx * fx - vy * fy));
	F[L::find(1, 0)] = pre / nine * (three * ((one - vx) * fx - vy * fy) + nine * vx * fx);
	F[L::find(-1, 0)] = pre / nine * (three * ((-one - vx) * fx - vy * fy) + nine * vx * fx);
	F[L::find(0, 1)] = pre / nine * (three * (-vx * fx + (one - vy) * fy) + nine * vy * fy);
	F[L::find(0, -1)] = pre / nine * (three * (-vx * fx + (-one - vy) * fy) + nine * vy * fy);
	F[L::find(1, 1)] = pre / n36 * (three * ((one - vx) * fx + (one - vy) * fy) + nine * (vx + vy) * (fx + fy));
	F[L::find(-1, -1)] = pre / n36 * (three * ((-one - vx) * fx + (-one - vy) * fy) + nine * (vx + vy) * (fx + fy));
	F[L::find(1, -1)] = pre / n36 * (three * ((one - vx) * fx + (-one - vy) * fy) + nine * (vx - vy) * (fx - fy));
	F[L::find(-1, 1)] = pre / n36 * (three * ((-one - vx) * fx + (one - vy) * fy) + nine * (vx - vy) * (fx - fy));
	static_for<9>([&](auto qc) {
		constexpr int q = qc;
		f[q] += (feq[q] - f[q]) / tau + F[q];
	});
}

// d2q9/col_clbm.h:13-89
template <typename R>
LBMX_D void collide_clbm(R (&f)[9], const Phys<R>& P, R rho, R vx, R vy)
{
	using L = D2Q9;
	const R tau = R(3) * P.nu + R(0.5);
	const R fx = P.fx, fy = P.fy;
	const R zz = f[L::find(0, 0)], pz = f[L::find(1, 0)], mz = f[L::find(-1, 0)], zp = f[L::find(0, 1)], zm = f[L::find(0, -1)];
	const R pp = f[L::find(1, 1)], mm = f[L::find(-1, -1)], pm = f[L::find(1, -1)], mp = f[L::find(-1, 1)];
	const R c2 = 2, c3 = 3, c4 = 4, c6 = 6, c8 = 8, c9 = 9, c36 = 36, q25 = R(.25), h5 = R(.5);
	const R Pm = R(1.) / R(12.) * (rho * (vx * vx + vy * vy) - pz - zp - zm - mz - c2 * (pm + mm + pp + mp - R(1.) / R(3.) * rho) - (fx * vx + fy * vy));
	const R NE = q25 / tau * (zp + zm - pz - mz + rho * (vx * vx - vy * vy) - (fx * vx - fy * vy));
	const R V = q25 / tau * ((pp + mm - mp - pm) - vx * vy * rho + h5 * (fx * vy + fy * vx));
	const R kxxyy = (pz + pp + mp + pm + mm + mz - vx * vx * rho + c2 * NE + c6 * Pm) * (zp + pp + mp + zm + pm + mm - vy * vy * rho - c2 * NE + c6 * Pm);
	const R UP = (-(q25 * (pm + mm - pp - mp - c2 * vx * vx * vy * rho + vy * (rho - zp - zm - zz) - h5 * (-vx * vx) * fy + fx * vx * vy)
					- vy * h5 * (-c3 * Pm - NE) + vx * ((pp - mp - pm + mm) * h5 - c2 * V)));
	const R RIGHT = (-(q25 * (mm + mp - pm - pp - c2 * vy * vy * vx * rho + vx * (rho - zz - mz - pz) - h5 * (-vy * vy) * fx + fy * vy * vx)
					   - vx * h5 * (-c3 * Pm + NE) + vy * ((pp + mm - pm - mp) * h5 - c2 * V)));
	const R NP = (q25
				  * (kxxyy - pp - mp - pm - mm - c8 * Pm + c2 * (vx * (pp - mp + pm - mm - c4 * RIGHT) + vy * (pp + mp - pm - mm - c4 * UP))
					 + c4 * vx * vy * (-pp + mp + pm - mm + c4 * V) + vx * vx * (-zp - pp - mp - zm - pm - mm + c2 * NE - c6 * Pm)
					 + vy * vy * ((-pz - pp - mp - pm - mm - mz - c2 * NE - c6 * Pm) + c3 * vx * vx * rho) - (fx * vx * vy * vy + fy * vy * vx * vx)));
	f[L::find(-1, 1)] += c2 * Pm + NP + V - UP + RIGHT;
	f[L::find(-1, 0)] += -Pm - c2 * NP + NE - c2 * RIGHT;
	f[L::find(-1, -1)] += c2 * Pm + NP - V + UP + RIGHT;
	f[L::find(0, -1)] += -Pm - c2 * NP - NE - c2 * UP;
	f[L::find(1, -1)] += c2 * Pm + NP + V + UP - RIGHT;
	f[L::find(1, 0)] += -Pm - c2 * NP + NE + c2 * RIGHT;
	f[L::find(1, 1)] += c2 * Pm + NP - V - UP - RIGHT;
	f[L::find(0, 1)] += -Pm - c2 * NP - NE + c2 * UP;
	f[L::find(0, 0)] += (c4 * (-Pm + NP));
	const R m1 = fx, m2 = fy;
	const R m3 = c6 * (fx * vx + fy * vy);
	const R m4 = c2 * (fx * vx - fy * vy);
	const R m5 = fx * vy + fy * vx;
	const R m6 = (c2 - c3 * vx * vx) * fy - c6 * fx * vx * vy;
	const R m7 = (c2 - c3 * vy * vy) * fx - c6 * fy * vx * vy;
	const R m8 = c6 * ((c3 * vy * vy - c2) * fx * vx + (c3 * vx * vx - c2) * fy * vy);
	f[L::find(0, 0)] += (-m3 + m8) / c9;
	f[L::find(1, 0)] += (c6 * m1 - m3 + c9 * m4 + c6 * m7 - c2 * m8) / c36;
	f[L::find(0, 1)] += (c6 * m2 - m3 - c9 * m4 + c6 * m6 - c2 * m8) / c36;
	f[L::find(-1, 0)] += (-c6 * m1 - m3 + c9 * m4 - c6 * m7 - c2 * m8) / c36;
	f[L::find(0, -1)] += (-c6 * m2 - m3 - c9 * m4 - c6 * m6 - c2 * m8) / c36;
	f[L::find(1, 1)] += (c6 * m1 + c6 * m2 + c2 * m3 + c9 * m5 - c3 * m6 - c3 * m7 + m8) / c36;
	f[L::find(-1, 1)] += (-c6 * m1 + c6 * m2 + c2 * m3 - c9 * m5 - c3 * m6 + c3 * m7 + m8) / c36;
	f[L::find(-1, -1)] += (-c6 * m1 - c6 * m2 + c2 * m3 + c9 * m5 + c3 * m6 + c3 * m7 + m8) / c36;
	f[L::find(1, -1)] += (c6 * m1 - c6 * m2 + c2 * m3 - c9 * m5 + c3 * m6 - c3 * m7 + m8) / c36;
}

}  // namespace strict
}  // namespace lbmx
