// lbm_oracle_ext.h -- CPU restatement of the reference's further D3Q27 collision operators (TEST INFRASTRUCTURE ONLY, part of
// oracle/lbm_oracle.cpp).  Pinned bit for bit against the reference's own headers by tests/test_oracle_vs_reference.py.
//
//   D3Q27_CLBM             include/lbm3d/d3q27/col_clbm.h:6-447
//   D3Q27_SRT_MODIF_FORCE  include/lbm3d/d3q27/col_srt_modif_force.h:9-120
//   D3Q27_KBC_N1..N4       include/lbm3d/d3q27/col_kbc_n.h:254-1272
//   D3Q27_KBC_C1..C4       include/lbm3d/d3q27/col_kbc_c.h:283-1301
#pragma once

// ---------------------------------------------------------------------------------------------
// D3Q27_CLBM: central-moment relaxation (col_clbm.h:16-200) + the 27-moment forcing term (col_clbm.h:303-416)
// ---------------------------------------------------------------------------------------------
// S_q = pre_q * sum_j coef[q][j] * m_j, terms in ascending j (col_clbm.h:342-415); dir = the population the row is added to
// (col_clbm.h:417-443); den: pre_q = (dreal)(1.0/den)
struct ClbmForceRow
{
	int dir[3];
	int den;
	int coef[27];
};
static const ClbmForceRow CLBM_FORCE[27] = {
	{{ 0,  0,  0},  27, {  1,   0,   0,   0,   0,   0,   0,   0,   0,  -3,   0,   0,   0,   0,   0,   0,   0,   3,   0,   0,   0,   0,   0,   0,   0,   0,  -1}},
	{{ 1,  0,  0}, 108, {  4,   6,   0,   0,   0,   0,   0,   9,   3,  -6,  -6,   0,   0,   0,   0,   0,   0,   0,  -6,   0,   0,   0,   0,   6,   0,   0,   2}},
	{{-1,  0,  0}, 108, {  4,  -6,   0,   0,   0,   0,   0,   9,   3,  -6,   6,   0,   0,   0,   0,   0,   0,   0,  -6,   0,   0,   0,   0,  -6,   0,   0,   2}},
	{{ 0,  1,  0}, 108, {  4,   0,   6,   0,   0,   0,   0,  -9,   3,  -6,   0,  -6,   0,   0,   0,   0,   0,   0,   3,  -9,   0,   0,   0,   0,   6,   0,   2}},
	{{ 0, -1,  0}, 108, {  4,   0,  -6,   0,   0,   0,   0,  -9,   3,  -6,   0,   6,   0,   0,   0,   0,   0,   0,   3,  -9,   0,   0,   0,   0,  -6,   0,   2}},
	{{ 0,  0,  1}, 108, {  4,   0,   0,   6,   0,   0,   0,   0,  -6,  -6,   0,   0,  -6,   0,   0,   0,   0,   0,   3,   9,   0,   0,   0,   0,   0,   6,   2}},
	{{ 0,  0, -1}, 108, {  4,   0,   0,  -6,   0,   0,   0,   0,  -6,  -6,   0,   0,   6,   0,   0,   0,   0,   0,   3,   9,   0,   0,   0,   0,   0,  -6,   2}},
	{{ 1,  1,  0}, 216, {  8,  12,  12,   0,  18,   0,   0,   0,  12,   0,  -3,  -3,   0,  27,  27,   0,   0,  -6,   3,   9,   0,   0, -18,  -6,  -6,   0,  -2}},
	{{-1,  1,  0}, 216, {  8, -12,  12,   0, -18,   0,   0,   0,  12,   0,   3,  -3,   0, -27,  27,   0,   0,  -6,   3,   9,   0,   0,  18,   6,  -6,   0,  -2}},
	{{ 1, -1,  0}, 216, {  8,  12, -12,   0, -18,   0,   0,   0,  12,   0,  -3,   3,   0,  27, -27,   0,   0,  -6,   3,   9,   0,   0,  18,  -6,   6,   0,  -2}},
	{{-1, -1,  0}, 216, {  8, -12, -12,   0,  18,   0,   0,   0,  12,   0,   3,   3,   0, -27, -27,   0,   0,  -6,   3,   9,   0,   0, -18,   6,   6,   0,  -2}},
	{{ 1,  0,  1}, 216, {  8,  12,   0,  12,   0,  18,   0,  18,  -6,   0,  -3,   0,  -3, -27,   0,  27,   0,  -6,   3,  -9,   0, -18,   0,  -6,   0,  -6,  -2}},
	{{-1,  0,  1}, 216, {  8, -12,   0,  12,   0, -18,   0,  18,  -6,   0,   3,   0,  -3,  27,   0,  27,   0,  -6,   3,  -9,   0,  18,   0,   6,   0,  -6,  -2}},
	{{ 1,  0, -1}, 216, {  8,  12,   0, -12,   0, -18,   0,  18,  -6,   0,  -3,   0,   3, -27,   0, -27,   0,  -6,   3,  -9,   0,  18,   0,  -6,   0,   6,  -2}},
	{{-1,  0, -1}, 216, {  8, -12,   0, -12,   0,  18,   0,  18,  -6,   0,   3,   0,   3,  27,   0, -27,   0,  -6,   3,  -9,   0, -18,   0,   6,   0,   6,  -2}},
	{{ 0,  1,  1}, 216, {  8,   0,  12,  12,   0,   0,  18, -18,  -6,   0,   0,  -3,  -3,   0, -27, -27,   0,  -6,  -6,   0, -18,   0,   0,   0,  -6,  -6,  -2}},
	{{ 0, -1,  1}, 216, {  8,   0, -12,  12,   0,   0, -18, -18,  -6,   0,   0,   3,  -3,   0,  27, -27,   0,  -6,  -6,   0,  18,   0,   0,   0,   6,  -6,  -2}},
	{{ 0,  1, -1}, 216, {  8,   0,  12, -12,   0,   0, -18, -18,  -6,   0,   0,  -3,   3,   0, -27,  27,   0,  -6,  -6,   0,  18,   0,   0,   0,  -6,   6,  -2}},
	{{ 0, -1, -1}, 216, {  8,   0, -12, -12,   0,   0,  18, -18,  -6,   0,   0,   3,   3,   0,  27,  27,   0,  -6,  -6,   0, -18,   0,   0,   0,   6,   6,  -2}},
	{{ 1,  1,  1}, 216, {  8,  12,  12,  12,  18,  18,  18,   0,   0,  12,   6,   6,   6,   0,   0,   0,  27,   6,   0,   0,   9,   9,   9,   3,   3,   3,   1}},
	{{-1,  1,  1}, 216, {  8, -12,  12,  12, -18, -18,  18,   0,   0,  12,  -6,   6,   6,   0,   0,   0, -27,   6,   0,   0,   9,  -9,  -9,  -3,   3,   3,   1}},
	{{ 1, -1,  1}, 216, {  8,  12, -12,  12, -18,  18, -18,   0,   0,  12,   6,  -6,   6,   0,   0,   0, -27,   6,   0,   0,  -9,   9,  -9,   3,  -3,   3,   1}},
	{{-1, -1,  1}, 216, {  8, -12, -12,  12,  18, -18, -18,   0,   0,  12,  -6,  -6,   6,   0,   0,   0,  27,   6,   0,   0,  -9,  -9,   9,  -3,  -3,   3,   1}},
	{{ 1,  1, -1}, 216, {  8,  12,  12, -12,  18, -18, -18,   0,   0,  12,   6,   6,  -6,   0,   0,   0, -27,   6,   0,   0,  -9,  -9,   9,   3,   3,  -3,   1}},
	{{-1,  1, -1}, 216, {  8, -12,  12, -12, -18,  18, -18,   0,   0,  12,  -6,   6,  -6,   0,   0,   0,  27,   6,   0,   0,  -9,   9,  -9,  -3,   3,  -3,   1}},
	{{ 1, -1, -1}, 216, {  8,  12, -12, -12, -18, -18,  18,   0,   0,  12,   6,  -6,  -6,   0,   0,   0,  27,   6,   0,   0,   9,  -9,  -9,   3,  -3,  -3,   1}},
	{{-1, -1, -1}, 216, {  8, -12, -12, -12,  18,  18,  18,   0,   0,  12,  -6,  -6,  -6,   0,   0,   0, -27,   6,   0,   0,   9,   9,   9,  -3,  -3,  -3,   1}},
};

template <typename R>
void clbm_force_moments(R (&m)[27], R u, R v, R w, R Fx, R Fy, R Fz)  // col_clbm.h:303-340
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

template <typename R>
void collide_clbm27(Cell<R, 27>& K)
{
	const R one = 1, two = 2, three = 3, half = (R) 0.5, third = (R) (1.0 / 3.0), n1o27 = (R) (1.0 / 27.0);
	const R rho = K.rho, vx = K.vx, vy = K.vy, vz = K.vz;
	R k[3][3][3];
	for (int q = 0; q < 27; q++)
		k[C27[q][0] + 1][C27[q][1] + 1][C27[q][2] + 1] = K.f[q];
	// forward central-moment transform z, y, x (col_clbm.h:18-117): the cumulant operator's Eq 6-14
	for (int a = 0; a < 3; a++)
		for (int b = 0; b < 3; b++)
			to_central(k[a][b][0], k[a][b][1], k[a][b][2], vz);
	for (int a = 0; a < 3; a++)
		for (int c = 0; c < 3; c++)
			to_central(k[a][0][c], k[a][1][c], k[a][2][c], vy);
	for (int b = 0; b < 3; b++)
		for (int c = 0; c < 3; c++)
			to_central(k[0][b][c], k[1][b][c], k[2][b][c], vx);

	// relaxation (col_clbm.h:119-200); default build: omega2..10 = 1, no antialias derivatives
	const R omega1 = one / (three * K.nu + half);
	const R omega2 = one, omega3 = one, omega4 = one, omega5 = one, omega6 = one, omega7 = one, omega8 = one, omega9 = one, omega10 = one;
	const R Dxu = 0, Dyv = 0, Dzw = 0;
	R s[3][3][3];
	const R d4 = (one - omega1) * (k[2][0][0] - k[0][2][0]) - three * rho * (one - omega1 * half) * (vx * vx * Dxu - vy * vy * Dyv);
	const R d5 = (one - omega1) * (k[2][0][0] - k[0][0][2]) - three * rho * (one - omega1 * half) * (vx * vx * Dxu - vz * vz * Dzw);
	const R d6 = rho * omega2 + (one - omega2) * (k[2][0][0] + k[0][2][0] + k[0][0][2])
			   - three * rho * (one - omega2 / two) * (vx * vx * Dxu + vy * vy * Dyv + vz * vz * Dzw);
	s[2][0][0] = third * (d4 + d5 + d6);
	s[0][2][0] = third * (-two * d4 + d5 + d6);
	s[0][0][2] = third * (d4 - two * d5 + d6);
	s[1][2][0] = (-k[1][0][2] - k[1][2][0]) * omega3 * half + (k[1][0][2] - k[1][2][0]) * omega4 * half + k[1][2][0];
	s[1][0][2] = (-k[1][0][2] - k[1][2][0]) * omega3 * half + (-k[1][0][2] + k[1][2][0]) * omega4 * half + k[1][0][2];
	s[2][1][0] = (-k[0][1][2] - k[2][1][0]) * omega3 * half + (k[0][1][2] - k[2][1][0]) * omega4 * half + k[2][1][0];
	s[0][1][2] = (-k[0][1][2] - k[2][1][0]) * omega3 * half + (-k[0][1][2] + k[2][1][0]) * omega4 * half + k[0][1][2];
	s[0][2][1] = (-k[0][2][1] - k[2][0][1]) * omega3 * half + (-k[0][2][1] + k[2][0][1]) * omega4 * half + k[0][2][1];
	s[2][0][1] = (-k[0][2][1] - k[2][0][1]) * omega3 * half + (k[0][2][1] - k[2][0][1]) * omega4 * half + k[2][0][1];
	s[1][1][1] = (one - omega5) * k[1][1][1];
	const R d14 = (one - omega6) * (k[2][2][0] - two * k[2][0][2] + k[0][2][2]);
	const R d15 = (one - omega6) * (k[2][2][0] + k[2][0][2] - two * k[0][2][2]);
	const R d16 = (one - omega7) * (k[2][2][0] + k[2][0][2] + k[0][2][2]) + omega7 * rho * third;
	s[2][2][0] = third * (d14 + d15 + d16);
	s[2][0][2] = third * (-d14 + d16);
	s[0][2][2] = third * (-d15 + d16);
	s[2][1][1] = (one - omega8) * k[2][1][1];
	s[1][2][1] = (one - omega8) * k[1][2][1];
	s[1][1][2] = (one - omega8) * k[1][1][2];
	s[2][2][1] = (one - omega9) * k[2][2][1];
	s[2][1][2] = (one - omega9) * k[2][1][2];
	s[1][2][2] = (one - omega9) * k[1][2][2];
	s[2][2][2] = (one - omega10) * k[2][2][2] + omega10 * rho * n1o27;
	s[0][0][0] = k[0][0][0];  // no sign change of the first moments here (col_clbm.h:188-192), unlike the cumulant operator
	s[1][0][0] = k[1][0][0];
	s[0][1][0] = k[0][1][0];
	s[0][0][1] = k[0][0][1];
	s[1][0][1] = (one - omega1) * k[1][0][1];
	s[0][1][1] = (one - omega1) * k[0][1][1];
	s[1][1][0] = (one - omega1) * k[1][1][0];

	// backward transform x, y, z (col_clbm.h:202-300)
	for (int b = 0; b < 3; b++)
		for (int c = 0; c < 3; c++)
			from_central(s[0][b][c], s[1][b][c], s[2][b][c], vx);
	for (int a = 0; a < 3; a++)
		for (int c = 0; c < 3; c++)
			from_central(s[a][0][c], s[a][1][c], s[a][2][c], vy);
	for (int a = 0; a < 3; a++)
		for (int b = 0; b < 3; b++)
			from_central(s[a][b][0], s[a][b][1], s[a][b][2], vz);
	for (int q = 0; q < 27; q++)
		K.f[q] = s[C27[q][0] + 1][C27[q][1] + 1][C27[q][2] + 1];

	// forcing (col_clbm.h:303-443)
	R m[27];
	clbm_force_moments(m, vx, vy, vz, K.fx, K.fy, K.fz);
	for (const ClbmForceRow& row : CLBM_FORCE) {
		R acc = 0;
		bool first = true;
		for (int j = 1; j < 27; j++) {	// m_0 = 0: its term is an exact zero at the head of every sum
			if (row.coef[j] == 0)
				continue;
			const R t = (R) row.coef[j] * m[j];
			acc = first ? t : acc + t;
			first = false;
		}
		K.f[find27(row.dir[0], row.dir[1], row.dir[2])] += (R) (1.0 / row.den) * acc;
	}
}

// ---------------------------------------------------------------------------------------------
// D3Q27_SRT_MODIF_FORCE (col_srt_modif_force.h:17-118): SRT with a first-order-in-u source term.  The reference writes the
// source with double literals, so for dreal = float those sums are evaluated in double (products of two dreal variables stay
// in dreal) and rounded once when stored.
// ---------------------------------------------------------------------------------------------
template <typename R>
R modif_force_source(int q, R vx, R vy, R vz, R fx, R fy, R fz)
{
	const int c[3] = {C27[q][0], C27[q][1], C27[q][2]};
	const R v[3] = {vx, vy, vz}, F[3] = {fx, fy, fz};
	const int n = (c[0] != 0) + (c[1] != 0) + (c[2] != 0);
	if (n == 0)
		return (R) (-8.0 / 9.0 * (double) (v[0] * F[0] + F[1] * v[1] + v[2] * F[2]));
	if (n == 1) {
		int a = c[0] != 0 ? 0 : (c[1] != 0 ? 1 : 2), o1 = a == 0 ? 1 : 0, o2 = a == 2 ? 1 : 2;
		const double own = (4.0 * (double) v[a] + (c[a] > 0 ? 2.0 : -2.0)) * (double) F[a] / 9.0;
		return (R) (own - 2.0 / 9.0 * (double) (v[o1] * F[o1] + v[o2] * F[o2]));
	}
	const double den = n == 2 ? 18.0 : 72.0;
	double sum = 0;
	bool first = true;
	for (int a = 0; a < 3; a++) {
		if (c[a] == 0)
			continue;
		double A = 0;
		bool f1 = true;
		for (int b = 0; b < 3; b++) {
			if (c[b] == 0)
				continue;
			const double t = (b == a ? 2.0 : 3.0 * (double) (c[a] * c[b])) * (double) v[b];
			A = f1 ? t : A + t;
			f1 = false;
		}
		A = A + (double) c[a];
		const double term = A * (double) F[a] / den;
		sum = first ? term : sum + term;
		first = false;
	}
	if (n == 2) {
		const int z = c[0] == 0 ? 0 : (c[1] == 0 ? 1 : 2);
		sum = sum - (double) (F[z] * v[z]) / 18.0;
	}
	return (R) sum;
}

template <typename R>
void collide_srt_modif27(Cell<R, 27>& K, int eqkind)
{
	const R one = 1, half = (R) 0.5;
	const R tau = (R) 3.0 * K.nu + half;
	R S[27], feq[27];
	for (int q = 0; q < 27; q++) {
		S[q] = modif_force_source(q, K.vx, K.vy, K.vz, K.fx, K.fy, K.fz);
		feq[q] = equilibrium(K, eqkind, q, K.rho, K.vx, K.vy, K.vz);
	}
	for (int q = 0; q < 27; q++)
		K.f[q] += (feq[q] - K.f[q]) / tau + (one - half / tau) * S[q];
}

// ---------------------------------------------------------------------------------------------
// KBC family (col_kbc_n.h:254-1272, col_kbc_c.h:283-1301): f = k + s + h; the shear part s relaxes with 2 beta, the higher-order
// part h with gamma * beta, gamma = the entropic stabiliser.  The eight models differ in the shear part only:
//   N1..N4: s built from raw moments,     D | D+T | D+Q | D+T+Q     (col_kbc_n.h:28-252)
//   C1..C4: s built from central moments, D~| D~+T~ | D~+Q~ | D~+T~+Q~ (col_kbc_c.h:56-281)
// ---------------------------------------------------------------------------------------------
// raw moments M_abc = sum_i cx^a cy^b cz^c f_i, summed left to right in the order the reference lists the populations
// (col_kbc_n.h:351-386); direction codes m, z, p = -1, 0, +1 for (x, y, z)
struct KbcMoment
{
	int a, b, c;
	const char* order;
};
static const KbcMoment KBC_MOMENTS[13] = {
	{2, 0, 0, "mmm mmp mmz mpm mpp mpz mzm mzp mzz pmm pmp pmz ppm ppp ppz pzm pzp pzz"},
	{0, 2, 0, "mmm mmp mmz mpm mpp mpz zmm zmp zmz zpm zpp zpz pmm pmp pmz ppm ppp ppz"},
	{0, 0, 2, "mmm mmp mzm mzp mpm mpp zmm zmp zzm zzp zpm zpp pmm pmp pzm pzp ppm ppp"},
	{1, 1, 0, "mmm mmz mmp mpm mpz mpp pmm pmz pmp ppm ppz ppp"},
	{1, 0, 1, "mmm mmp mzm mzp mpm mpp pmm pmp pzm pzp ppm ppp"},
	{0, 1, 1, "mmm mmp mpm mpp zmm zmp zpm zpp pmm pmp ppm ppp"},
	{1, 1, 1, "ppp mmp mpm mpp pmm pmp ppm mmm"},
	{2, 0, 1, "ppp mmp mzm mzp mpm mpp pmm pmp pzm pzp ppm mmm"},
	{1, 0, 2, "ppp mmp mzm mzp mpm mpp pmm pmp pzm pzp ppm mmm"},
	{2, 1, 0, "ppp mmz mmp mpm mpz mpp pmm pmz pmp ppm ppz mmm"},
	{1, 2, 0, "ppp mmz mmp mpm mpz mpp pmm pmz pmp ppm ppz mmm"},
	{0, 2, 1, "ppp mmp mpm mpp zmm zmp zpm zpp pmm pmp ppm mmm"},
	{0, 1, 2, "ppp mmp mpm mpp zmm zmp zpm zpp pmm pmp ppm mmm"},
};
enum { KM200, KM020, KM002, KM110, KM101, KM011, KM111, KM201, KM102, KM210, KM120, KM021, KM012 };

template <typename R>
void collide_kbc27(Cell<R, 27>& K, bool central, bool useT, bool useQ)
{
	const R zero = 0, one = 1, two = 2, three = 3, four = 4, six = 6, eight = 8, half = (R) 0.5, third = (R) (1.0 / 3.0);
	const R n1o4 = (R) 0.25, n1o6 = (R) (1.0 / 6.0), n1o8 = (R) 0.125;
	const R rho = K.rho, vx = K.vx, vy = K.vy, vz = K.vz;
	const R v[3] = {vx, vy, vz};
	// product-form equilibrium (col_kbc_n.h:293-321), as in col_bgk.h
	R g[3][3];
	for (int a = 0; a < 3; a++) {
		const R z = third - one + v[a] * v[a];
		const R p = -half * (z + one + v[a]);
		g[a][1] = z;
		g[a][2] = p;
		g[a][0] = p + v[a];
	}
	R feq[27], ifeq[27];
	for (int q = 0; q < 27; q++) {
		feq[q] = -rho * g[0][C27[q][0] + 1] * g[1][C27[q][1] + 1] * g[2][C27[q][2] + 1];
		ifeq[q] = one / feq[q];
	}
	R M[13];
	for (int k = 0; k < 13; k++) {
		const KbcMoment& mo = KBC_MOMENTS[k];
		R acc = 0;
		bool first = true;
		for (const char* s = mo.order; *s; s += (s[3] ? 4 : 3)) {
			auto comp = [](char ch) { return ch == 'm' ? -1 : (ch == 'p' ? 1 : 0); };
			const int cx = comp(s[0]), cy = comp(s[1]), cz = comp(s[2]);
			int sign = 1;
			for (int i = 0; i < mo.a; i++) sign *= cx;
			for (int i = 0; i < mo.b; i++) sign *= cy;
			for (int i = 0; i < mo.c; i++) sign *= cz;
			const R t = K.f[find27(cx, cy, cz)];
			acc = first ? (sign > 0 ? t : -t) : (sign > 0 ? acc + t : acc - t);
			first = false;
		}
		M[k] = acc;
	}
	// the "special moments" and their equilibria (col_kbc_n.h:28-54, col_kbc_c.h:56-83)
	R T = (M[KM200] + M[KM020] + M[KM002]), Nxz = (M[KM200] - M[KM002]), Nyz = (M[KM020] - M[KM002]);
	R Pxy = M[KM110], Pxz = M[KM101], Pyz = M[KM011];
	R Qxxy = M[KM210], Qxxz = M[KM201], Qxyy = M[KM120], Qyyz = M[KM021], Qxzz = M[KM102], Qyzz = M[KM012], Qxyz = M[KM111];
	R eT, eNxz = 0, eNyz = 0, ePxy = 0, ePxz = 0, ePyz = 0, eQxxy = 0, eQxxz = 0, eQxyy = 0, eQyyz = 0, eQxzz = 0, eQyzz = 0, eQxyz = 0;
	if (! central) {
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
	else {
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
	// scaling of a tensor entry: the raw-moment models multiply by 1/6, 1/4, 1/2, 1/8, the central-moment models divide by 6, 4, 2, 8
	auto sc = [&](R x, int den) -> R {
		if (central)
			return x / (den == 6 ? six : den == 4 ? four : den == 2 ? two : eight);
		return x * (den == 6 ? n1o6 : den == 4 ? n1o4 : den == 2 ? half : n1o8);
	};
	// shear-part tensors per direction (col_kbc_n.h:56-252, col_kbc_c.h:85-281); has = the entry is not the constant 0
	auto tensorD = [&](int cx, int cy, int cz, R nxz, R nyz, R pxy, R pxz, R pyz, bool& has) -> R {
		const int n = (cx != 0) + (cy != 0) + (cz != 0);
		has = n == 1 || n == 2;
		if (n == 1)
			return cx != 0 ? sc(two * nxz - nyz, 6) : (cy != 0 ? sc(-nxz + two * nyz, 6) : sc(-nxz - nyz, 6));
		if (n == 2) {
			const R p = cz == 0 ? pxy : (cy == 0 ? pxz : pyz);
			const int s = cz == 0 ? cx * cy : (cy == 0 ? cx * cz : cy * cz);
			return s > 0 ? sc(p, 4) : sc(-p, 4);
		}
		return zero;
	};
	auto tensorT = [&](int cx, int cy, int cz, R t, bool& has) -> R {
		const int n = (cx != 0) + (cy != 0) + (cz != 0);
		has = n <= 1;
		return n == 0 ? -t : (n == 1 ? sc(t, 6) : zero);
	};
	auto tensorQ = [&](int cx, int cy, int cz, R qxxy, R qxxz, R qxyy, R qyyz, R qxzz, R qyzz, R qxyz, bool& has) -> R {
		const int n = (cx != 0) + (cy != 0) + (cz != 0);
		has = n >= 1;
		if (n == 1) {
			const R s = cx != 0 ? (qxyy + qxzz) : (cy != 0 ? (qxxy + qyzz) : (qxxz + qyyz));
			return (cx + cy + cz) > 0 ? sc(-s, 2) : sc(s, 2);
		}
		if (n == 2) {
			const R a = cz == 0 ? qxyy : (cy == 0 ? qxzz : qyzz), b = cz == 0 ? qxxy : (cy == 0 ? qxxz : qyyz);
			const int sa = cz == 0 ? cx : (cy == 0 ? cx : cy), sb = cz == 0 ? cy : cz;
			const R ta = sa > 0 ? a : -a;
			return sc(sb > 0 ? ta + b : ta - b, 4);
		}
		if (n == 3)
			return cx * cy * cz > 0 ? sc(qxyz, 8) : sc(-qxyz, 8);
		return zero;
	};
	R Ds[27], Dh[27];
	for (int q = 0; q < 27; q++) {
		const int cx = C27[q][0], cy = C27[q][1], cz = C27[q][2];
		R acc = zero;
		bool has;
		const R d = tensorD(cx, cy, cz, Nxz, Nyz, Pxy, Pxz, Pyz, has);
		if (has) {
			acc = d;
			if (! central)	// the central-moment equilibria of D~ are the constant 0
				acc = acc - tensorD(cx, cy, cz, eNxz, eNyz, ePxy, ePxz, ePyz, has);
		}
		if (useT) {
			const R t = tensorT(cx, cy, cz, T, has);
			if (has)
				acc = (acc + t) - tensorT(cx, cy, cz, eT, has);
		}
		if (useQ) {
			const R qq = tensorQ(cx, cy, cz, Qxxy, Qxxz, Qxyy, Qyyz, Qxzz, Qyzz, Qxyz, has);
			if (has) {
				acc = acc + qq;
				if (! central)
					acc = acc - tensorQ(cx, cy, cz, eQxxy, eQxxz, eQxyy, eQyyz, eQxzz, eQyzz, eQxyz, has);
			}
		}
		Ds[q] = acc;
		Dh[q] = K.f[q] - feq[q] - Ds[q];
	}
	const R beta = (one / (two * K.nu / third + one));
	// <Ds|Dh> and <Dh|Dh>, summed in the order mmm, mmz, mmp, mzm, ... ppp (col_kbc_n.h:233-252)
	R sd = 0, hh = 0;
	bool first = true;
	for (int a = -1; a <= 1; a++)
		for (int b = -1; b <= 1; b++)
			for (int c = -1; c <= 1; c++) {
				const int q = find27(a, b, c);
				const R t1 = Ds[q] * Dh[q] * ifeq[q], t2 = Dh[q] * Dh[q] * ifeq[q];
				sd = first ? t1 : sd + t1;
				hh = first ? t2 : hh + t2;
				first = false;
			}
	const R gamma = (one / beta - (two - one / beta) * sd / hh);
	for (int q = 0; q < 27; q++) {
		const R S = force_projection(q, vx, vy, vz, K.fx, K.fy, K.fz) / rho;
		K.f[q] -= beta * (two * Ds[q] + gamma * Dh[q]) - (one - beta) * S * feq[q];
	}
}
