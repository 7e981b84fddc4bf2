// lbm_oracle.cpp -- CPU restatement ("port") of the reference's fused collide-and-stream path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA engine: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the
// library built from it; the product never routes through it (see oracle_api.h).
//
// PARITY PIN: every function below is pinned bit-for-bit (strict IEEE build, -ffp-contract=off)
// against the reference's own code compiled into oracle/_ref/ by oracle/Makefile
// (tests/test_oracle_vs_reference.py, run wherever /root/reference was available to build
// oracle/_ref), and against the golden vectors under tests/golden/ generated from that
// reference build by tests/golden/make_golden.py.  The reference has no tests or golden vectors
// of its own.
//
// It is written from the algorithm, not from the text, of the reference: populations live in a
// 3x3x3 array indexed by velocity sign, the moment transforms are one triplet routine applied
// along each axis, summation trees are tables.  Floating-point association follows the reference
// exactly (cited per function) so that the strict build reproduces it to the last bit.
//
// Paths cited as file:line are relative to /root/reference/include/lbm3d/ unless stated.

#include "oracle_api.h"

#include <cmath>
#include <cstdint>
#include <cstring>

namespace {

typedef int64_t idx;

// ---------------------------------------------------------------------------------------------
// lattices: direction numbering of defs.h:257-305
// ---------------------------------------------------------------------------------------------
constexpr int C27[27][3] = {
	{0, 0, 0},	  {1, 0, 0},   {-1, 0, 0},	{0, 1, 0},	{0, -1, 0},	 {0, 0, 1},	  {0, 0, -1},  {1, 1, 0},	{-1, -1, 0},
	{1, -1, 0},	  {-1, 1, 0},  {1, 0, 1},	{-1, 0, -1}, {1, 0, -1}, {-1, 0, 1},  {0, 1, 1},   {0, -1, -1}, {0, 1, -1},
	{0, -1, 1},	  {1, 1, 1},   {-1, -1, -1}, {1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, -1}, {1, -1, -1}, {-1, 1, 1},
};
constexpr int C9[9][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}};

constexpr int find27(int cx, int cy, int cz)
{
	for (int q = 0; q < 27; q++)
		if (C27[q][0] == cx && C27[q][1] == cy && C27[q][2] == cz)
			return q;
	return -1;
}
constexpr int find9(int cx, int cy)
{
	for (int q = 0; q < 9; q++)
		if (C9[q][0] == cx && C9[q][1] == cy)
			return q;
	return -1;
}
constexpr int opp27(int q)
{
	return find27(-C27[q][0], -C27[q][1], -C27[q][2]);
}
constexpr int opp9(int q)
{
	return find9(-C9[q][0], -C9[q][1]);
}

// short names: D(cx,cy,cz)
#define D(a, b, c) find27(a, b, c)
#define E(a, b) find9(a, b)

// ---------------------------------------------------------------------------------------------
// per-cell state (KernelStruct, defs.h:122-167) and per-block data (LBM_Data / NSE_Data, lbm_data.h:7-131)
// ---------------------------------------------------------------------------------------------
template <typename R, int Q_>
struct Cell
{
	static constexpr int Q = Q_;
	R f[Q_];
	R fx = 0, fy = 0, fz = 0;
	R vx = 0, vy = 0, vz = 0;
	R rho = 1, nu = 1;	// defaults matter: MACRO_Void never overwrites them (d3q27/macro.h:174-188)
	bool kahan_rho = false;	 // the reference compiled with -DUSE_HIGH_PRECISION_RHO (defs.h:252)
};

template <typename R>
struct Block
{
	idx X, Y, Z, ox, XYZ;
	bool even;
	int nproc;
	R* cur;	 // dfs[df_cur]
	R* out;	 // dfs[df_out] (A-B) -- equal to cur for A-A
	R* macro;
	const int16_t* map;
	R nu, fx, fy, fz;
	R in_vx, in_vy, in_vz;
	const R* profile;
	idx profile_sy;
	int stat_counter;
	int macro_gates = 0;
	int inflow_kind;
	const R* bouzidi;  // [8][XYZ] or nullptr

	idx cell(idx x, idx y, idx z) const	 // lbm_data.h:49-57 with permutation (x,z,y), overlap in x only
	{
		return ((x + ox) * Z + z) * Y + y;
	}
	idx at(int q, idx x, idx y, idx z) const
	{
		return q * XYZ + cell(x, y, z);
	}
};

// ---------------------------------------------------------------------------------------------
// neighbour coordinates: kernelInitIndices, kernels.h:6-58 (HAVE_MPI form; identical to the plain
// form when every overlap is zero)
// ---------------------------------------------------------------------------------------------
struct Nbr
{
	idx xm, x, xp, ym, y, yp, zm, z, zp;
};

template <typename R>
Nbr neighbours(const Block<R>& B, bool periodic_cell, bool aa, idx x, idx y, idx z)
{
	Nbr n;
	n.x = x;
	n.y = y;
	n.z = z;
	if (periodic_cell) {
		const bool wrap = B.nproc == 1;
		n.xp = (wrap && x == B.X - 1) ? 0 : x + 1;
		n.xm = (wrap && x == 0) ? B.X - 1 : x - 1;
		n.yp = (wrap && y == B.Y - 1) ? 0 : y + 1;
		n.ym = (wrap && y == 0) ? B.Y - 1 : y - 1;
		n.zp = (wrap && z == B.Z - 1) ? 0 : z + 1;
		n.zm = (wrap && z == 0) ? B.Z - 1 : z - 1;
	}
	else if (aa) {
		n.xp = x + 1;
		n.xm = x - 1;
		n.yp = y + 1;
		n.ym = y - 1;
		n.zp = z + 1;
		n.zm = z - 1;
	}
	else {
		n.xp = x + 1 < B.X - 1 + B.ox ? x + 1 : B.X - 1 + B.ox;
		n.xm = x - 1 > -B.ox ? x - 1 : -B.ox;
		n.yp = y + 1 < B.Y - 1 ? y + 1 : B.Y - 1;
		n.ym = y - 1 > 0 ? y - 1 : 0;
		n.zp = z + 1 < B.Z - 1 ? z + 1 : B.Z - 1;
		n.zm = z - 1 > 0 ? z - 1 : 0;
	}
	return n;
}

inline idx pick(int c, idx m, idx z, idx p)
{
	return c < 0 ? m : (c > 0 ? p : z);
}

// ---------------------------------------------------------------------------------------------
// D3Q27 lattice policy
// ---------------------------------------------------------------------------------------------
struct L27
{
	static constexpr int Q = 27;
	static constexpr int NDIM = 3;
	static int c(int q, int a) { return C27[q][a]; }
	static int opp(int q) { return opp27(q); }
	static int find(int cx, int cy, int cz) { return find27(cx, cy, cz); }
	enum { FLUID, WALL, INFLOW, INFLOW_LEFT, OUTFLOW_EQ, OUTFLOW_RIGHT, OUTFLOW_RIGHT_INTERP, PERIODIC, NOTHING, SYM_TOP, SYM_BOTTOM, SYM_LEFT, SYM_RIGHT, SYM_BACK, SYM_FRONT };	// d3q27/bc.h:17-34
	static bool collides(int m) { return m == FLUID || m == PERIODIC || m == OUTFLOW_RIGHT || m == OUTFLOW_RIGHT_INTERP || m == INFLOW_LEFT; }	// d3q27/bc.h:243-248
};

// D3Q19 = the first 19 directions of the D3Q27 numbering.  PARITY UNPINNED: the reference has no such lattice,
// so this part of the file is a restatement of nothing -- it applies the reference's SRT (col_srt.h) and MRT_LES (col_mrt.h)
// formulas with the standard D3Q19 weights and exists only so that the engine's D3Q19 kernels have an independent CPU check.
struct L19
{
	static constexpr int Q = 19;
	static constexpr int NDIM = 3;
	static int c(int q, int a) { return C27[q][a]; }
	static int opp(int q) { return opp27(q); }
	static int find(int cx, int cy, int cz) { return find27(cx, cy, cz); }
	enum { FLUID, WALL, INFLOW, INFLOW_LEFT, OUTFLOW_EQ, OUTFLOW_RIGHT, OUTFLOW_RIGHT_INTERP, PERIODIC, NOTHING, SYM_TOP, SYM_BOTTOM, SYM_LEFT, SYM_RIGHT, SYM_BACK, SYM_FRONT };
	static bool collides(int m) { return L27::collides(m); }
};

struct L9
{
	static constexpr int Q = 9;
	static constexpr int NDIM = 2;
	static int c(int q, int a) { return C9[q][a]; }
	static int opp(int q) { return opp9(q); }
	static int find(int cx, int cy, int) { return find9(cx, cy); }
	enum { FLUID, WALL, INFLOW, OUTFLOW_EQ, OUTFLOW_RIGHT, OUTFLOW_RIGHT_INTERP, PERIODIC, NOTHING, SYM_TOP, SYM_BOTTOM, SYM_LEFT, SYM_RIGHT, FLUID_NEAR_WALL, INFLOW_LEFT = -100 };	 // d2q9/bc.h:16-34
	static bool collides(int m) { return m == FLUID || m == FLUID_NEAR_WALL || m == PERIODIC || m == OUTFLOW_RIGHT || m == OUTFLOW_RIGHT_INTERP; }	// d2q9/bc.h:198-203
};

// ---------------------------------------------------------------------------------------------
// density and velocity
// ---------------------------------------------------------------------------------------------
// d3q27/common.h:16-50.  s(q) = f[q] + f[opp q], d(q) = f[q] - f[opp q]; the trees are
//   rho = ((corners + edges) + axes) + rest
//   v_a = (((corner_a + edge_a) + axis_a) + F_a/2) / rho
// with the member order given by the tables below (Geier 2015 App. J order as used by the reference).
template <typename R>
void density_velocity(Cell<R, 27>& K)
{
	const R* f = K.f;
	auto s = [&](int q) { return f[q] + f[opp27(q)]; };
	auto d = [&](int q) { return f[q] - f[opp27(q)]; };
	const R corners = (s(D(1, 1, 1)) + s(D(1, -1, 1))) + (s(D(1, 1, -1)) + s(D(-1, 1, 1)));
	const R edges = ((s(D(0, 1, 1)) + s(D(0, 1, -1))) + (s(D(1, 0, 1)) + s(D(1, 0, -1)))) + (s(D(1, 1, 0)) + s(D(1, -1, 0)));
	const R axes = (s(D(1, 0, 0)) + s(D(0, 1, 0))) + s(D(0, 0, 1));
	K.rho = ((corners + edges) + axes) + f[D(0, 0, 0)];
	if (K.kahan_rho) {	// d3q27/common.h:19-29: Kahan summation over the populations in index order
		K.rho = 0;
		R c = 0;
		for (int i = 0; i < 27; i++) {
			const R y = f[i] - c;
			const R t = K.rho + y;
			c = (t - K.rho) - y;
			K.rho = t;
		}
	}

	const R half = (R) 0.5;
	const R cz = (d(D(1, 1, 1)) + d(D(-1, 1, 1))) + (d(D(1, -1, 1)) + d(D(-1, -1, 1)));
	const R ez = (d(D(0, 1, 1)) + d(D(0, -1, 1))) + (d(D(1, 0, 1)) + d(D(-1, 0, 1)));
	K.vz = (((cz + ez) + d(D(0, 0, 1))) + K.fz * half) / K.rho;
	const R cx = (d(D(1, 1, 1)) + d(D(1, -1, 1))) + (d(D(1, 1, -1)) + d(D(1, -1, -1)));
	const R ex = (d(D(1, 0, 1)) + d(D(1, 0, -1))) + (d(D(1, 1, 0)) + d(D(1, -1, 0)));
	K.vx = (((cx + ex) + d(D(1, 0, 0))) + K.fx * half) / K.rho;
	const R cy = (d(D(1, 1, 1)) + d(D(1, 1, -1))) + (d(D(-1, 1, 1)) + d(D(-1, 1, -1)));
	const R ey = (d(D(1, 1, 0)) + d(D(-1, 1, 0))) + (d(D(0, 1, 1)) + d(D(0, 1, -1)));
	K.vy = (((cy + ey) + d(D(0, 1, 0))) + K.fy * half) / K.rho;
}

// D3Q19 (unpinned): plain pairwise sums over the 9 opposite pairs
template <typename R>
void density_velocity(Cell<R, 19>& K)
{
	R rho = K.f[0], j[3] = {0, 0, 0};
	for (int q = 1; q < 19; q += 2) {
		rho += K.f[q] + K.f[q + 1];
		for (int a = 0; a < 3; a++)
			j[a] += (R) C27[q][a] * (K.f[q] - K.f[q + 1]);
	}
	K.rho = rho;
	K.vx = (j[0] + (R) 0.5 * K.fx) / rho;
	K.vy = (j[1] + (R) 0.5 * K.fy) / rho;
	K.vz = (j[2] + (R) 0.5 * K.fz) / rho;
}

// d2q9/common.h:16-36
template <typename R>
void density_velocity(Cell<R, 9>& K)
{
	const R* f = K.f;
	const R half = (R) 0.5;
	K.rho = f[E(0, 0)] + (((f[E(1, 0)] + f[E(-1, 0)]) + (f[E(0, -1)] + f[E(0, 1)])) + ((f[E(1, 1)] + f[E(-1, -1)]) + (f[E(-1, 1)] + f[E(1, -1)])));
	K.vx = (((f[E(1, 0)] - f[E(-1, 0)]) + ((f[E(1, -1)] - f[E(-1, 1)]) + (f[E(1, 1)] - f[E(-1, -1)]))) + half * K.fx) / K.rho;
	K.vy = (((f[E(0, 1)] - f[E(0, -1)]) + ((f[E(-1, 1)] - f[E(1, -1)]) + (f[E(1, 1)] - f[E(-1, -1)]))) + half * K.fy) / K.rho;
}

// ---------------------------------------------------------------------------------------------
// equilibria
// ---------------------------------------------------------------------------------------------
// d3q27/eq.h:13-130: w*rho*(1 - 3/2 u.u + 3 c.u + 9/2 (c.u)^2); weights as (dreal)(a/b) literals (ciselnik.h)
template <typename R>
R eq27_std(int q, R rho, R vx, R vy, R vz)
{
	const R qx = (R) C27[q][0], qy = (R) C27[q][1], qz = (R) C27[q][2];
	const int n = (C27[q][0] != 0) + (C27[q][1] != 0) + (C27[q][2] != 0);
	const R w = n == 0 ? (R) (8.0 / 27.0) : n == 1 ? (R) (2.0 / 27.0) : n == 2 ? (R) (1.0 / 54.0) : (R) (1.0 / 216.0);
	const R cu = qx * vx + qy * vy + qz * vz;
	const R poly = (R) 1.0 - (R) 1.5 * (vx * vx + vy * vy + vz * vz) + (R) 3.0 * cu + (R) 4.5 * cu * cu;
	return w * rho * poly;
}

// d3q27/eq_inv_cum.h:24-136: +-w'*rho*(gx*gy*gz), g(0)=3v^2-2, g(+-1)=3v^2+-3v+1
template <typename R>
R eq27_inv_cum(int q, R rho, R vx, R vy, R vz)
{
	auto g = [](int c, R v) -> R {
		const R three = (R) 3.0;
		if (c == 0)
			return three * v * v - (R) 2.0;
		if (c > 0)
			return three * v * v + three * v + (R) 1.0;
		return three * v * v - three * v + (R) 1.0;
	};
	const int n = (C27[q][0] != 0) + (C27[q][1] != 0) + (C27[q][2] != 0);
	const R w = n == 0 ? -(R) (1.0 / 27.0) : n == 1 ? (R) (1.0 / 54.0) : n == 2 ? -(R) (1.0 / 108.0) : (R) (1.0 / 216.0);
	return w * rho * (g(C27[q][0], vx) * g(C27[q][1], vy) * g(C27[q][2], vz));
}

// d2q9/eq.h:13-61
template <typename R>
R eq9_std(int q, R rho, R vx, R vy)
{
	const R qx = (R) C9[q][0], qy = (R) C9[q][1];
	const int n = (C9[q][0] != 0) + (C9[q][1] != 0);
	const R w = n == 0 ? (R) (4.0 / 9.0) : n == 1 ? (R) (1.0 / 9.0) : (R) (1.0 / 36.0);
	const R cu = qx * vx + qy * vy;
	const R poly = (R) 1.0 - (R) 1.5 * (vx * vx + vy * vy) + (R) 3.0 * cu + (R) 4.5 * cu * cu;
	return w * rho * poly;
}

// d3q27/eq_entropic.h:11-211: rho * W(cx) W(cy) W(cz) * prod_a (2 - s_a) * prod_a B_a^{c_a}, s_a = sqrt(1 + 3 v_a^2),
// B_a = (2 v_a + s_a) / (1 - v_a), evaluated left to right: weights, the three (2 - s_a), then per axis "* 1 / B_a" (c_a = -1) or
// "* B_a" (c_a = +1).  The reference's unqualified sqrt() is ::sqrt(double) in the host build, so for dreal = float the chain is
// in double from the first (2 - s_a) on and rounded once when returned.
template <typename R>
R eq27_entropic(int q, R rho, R vx, R vy, R vz)
{
	const R v[3] = {vx, vy, vz};
	const int c[3] = {C27[q][0], C27[q][1], C27[q][2]};
	const R w6 = (R) (1.0 / 6.0), w23 = (R) (2.0 / 3.0);
	const R w = (c[0] ? w6 : w23) * (c[1] ? w6 : w23) * (c[2] ? w6 : w23);
	double s[3], B[3];
	for (int a = 0; a < 3; a++) {
		s[a] = ::sqrt((double) ((R) 1.0 + (R) 3.0 * v[a] * v[a]));
		B[a] = ((double) ((R) 2.0 * v[a]) + s[a]) / (double) ((R) 1.0 - v[a]);
	}
	double chain = (double) w * (2.0 - s[0]) * (2.0 - s[1]) * (2.0 - s[2]);
	for (int a = 0; a < 3; a++) {
		if (c[a] < 0)
			chain = chain * 1.0 / B[a];
		else if (c[a] > 0)
			chain = chain * B[a];
	}
	return (R) ((double) rho * chain);
}

template <typename R>
R equilibrium(const Cell<R, 27>&, int eqkind, int q, R rho, R vx, R vy, R vz)
{
	if (eqkind == ORC_EQ_ENTROPIC)
		return eq27_entropic(q, rho, vx, vy, vz);
	return eqkind == ORC_EQ_INV_CUM ? eq27_inv_cum(q, rho, vx, vy, vz) : eq27_std(q, rho, vx, vy, vz);
}
template <typename R>
R equilibrium(const Cell<R, 19>&, int, int q, R rho, R vx, R vy, R vz)  // D3Q19 (unpinned): eq.h polynomial, weights 1/3, 1/18, 1/36
{
	const R qx = (R) C27[q][0], qy = (R) C27[q][1], qz = (R) C27[q][2];
	const int n = (C27[q][0] != 0) + (C27[q][1] != 0) + (C27[q][2] != 0);
	const R w = n == 0 ? (R) (1.0 / 3.0) : n == 1 ? (R) (1.0 / 18.0) : (R) (1.0 / 36.0);
	const R cu = qx * vx + qy * vy + qz * vz;
	return w * rho * ((R) 1.0 - (R) 1.5 * (vx * vx + vy * vy + vz * vz) + (R) 3.0 * cu + (R) 4.5 * cu * cu);
}
template <typename R>
R equilibrium(const Cell<R, 9>&, int, int q, R rho, R vx, R vy, R)
{
	return eq9_std(q, rho, vx, vy);
}

// setEquilibrium / setEquilibriumDecomposition: d3q27/common.h:61-124, d2q9/common.h:47-72
template <typename R, int Q>
void set_equilibrium(Cell<R, Q>& K, int eqkind)
{
	for (int q = 0; q < Q; q++)
		K.f[q] = equilibrium(K, eqkind, q, K.rho, K.vx, K.vy, K.vz);
}
template <typename R, int Q>
void add_equilibrium_difference(Cell<R, Q>& K, int eqkind, R rho_out)
{
	for (int q = 0; q < Q; q++)
		K.f[q] += equilibrium(K, eqkind, q, rho_out, K.vx, K.vy, K.vz) - equilibrium(K, eqkind, q, K.rho, K.vx, K.vy, K.vz);
}

// ---------------------------------------------------------------------------------------------
// D3Q27 cumulant collision: d3q27/col_cum.h:14-485 (default build: no USE_GEIER_CUM_2017 / _ANTIALIAS)
// ---------------------------------------------------------------------------------------------
// m[a][b][c]: index 0,1,2 = velocity sign -,0,+ before an axis is transformed and moment order 0,1,2 after.
template <typename R>
inline void to_central(R& lo, R& mid, R& hi, R v)  // Eq 6-8 / 9-11 / 12-14 (col_cum.h:52-148): in (f-,f0,f+), out (k0,k1,k2)
{
	const R fm = lo, fz = mid, fp = hi;
	const R k0 = (fp + fm) + fz;
	const R k1 = (fp - fm) - v * k0;
	const R k2 = (fp + fm) - (R) 2.0 * v * (fp - fm) + v * v * k0;
	lo = k0;
	mid = k1;
	hi = k2;
}
template <typename R>
inline void from_central(R& lo, R& mid, R& hi, R v)	 // Eq G2015(88)-(96) (col_cum.h:349-445): in (k0,k1,k2), out (f-,f0,f+)
{
	const R k0 = lo, k1 = mid, k2 = hi;
	const R one = (R) 1.0, two = (R) 2.0, half = (R) 0.5;
	const R f0 = k0 * (one - v * v) - two * v * k1 - k2;
	const R fm = (k0 * (v * v - v) + k1 * (two * v - one) + k2) * half;
	const R fp = (k0 * (v * v + v) + k1 * (two * v + one) + k2) * half;
	lo = fm;
	mid = f0;
	hi = fp;
}

// rate limiter of the 2017 parametrisation (col_cum.h:183-197): w + (1 - w) * fabs(x) / (rho * lambda + fabs(x)).  The
// reference calls an unqualified fabs(): in its host build that is ::fabs(double), so for dreal = float everything downstream
// of it is evaluated in double and rounded once when stored (same situation as sqrt() in col_mrt.h).
template <typename R>
inline R limited_rate(R w, R x, R rho, R lambda)
{
	const double ax = std::fabs((double) x);
	return (R) ((double) w + (double) ((R) 1.0 - w) * ax / ((double) (rho * lambda) + ax));
}

// G2017: -DUSE_GEIER_CUM_2017 (parametrised rates, limiter, A and B terms); ANTIALIAS: -DUSE_GEIER_CUM_ANTIALIAS (velocity derivatives)
template <typename R, bool G2017 = false, bool ANTIALIAS = false>
void collide_cum(Cell<R, 27>& K)
{
	const R one = 1, two = 2, three = 3, four = 4, sixteen = 16, half = (R) 0.5;
	const R third = (R) (1.0 / 3.0), n2o3 = (R) (2.0 / 3.0), n4o3 = (R) (4.0 / 3.0);
	const R rho = K.rho, vx = K.vx, vy = K.vy, vz = K.vz;
	R m[3][3][3];
	for (int q = 0; q < 27; q++)
		m[C27[q][0] + 1][C27[q][1] + 1][C27[q][2] + 1] = K.f[q];

	// forward central-moment transform: z, then y, then x
	for (int a = 0; a < 3; a++)
		for (int b = 0; b < 3; b++)
			to_central(m[a][b][0], m[a][b][1], m[a][b][2], vz);
	for (int a = 0; a < 3; a++)
		for (int c = 0; c < 3; c++)
			to_central(m[a][0][c], m[a][1][c], m[a][2][c], vy);
	for (int b = 0; b < 3; b++)
		for (int c = 0; c < 3; c++)
			to_central(m[0][b][c], m[1][b][c], m[2][b][c], vx);
#define k(a, b, c) m[a][b][c]

	// cumulants of order 4-6, Eq G2015(51)-(54) (col_cum.h:151-171); lower orders equal the central moments
	R C[3][3][3];
	std::memcpy(C, m, sizeof(C));
	C[2][1][1] = k(2, 1, 1) - (k(2, 0, 0) * k(0, 1, 1) + two * k(1, 0, 1) * k(1, 1, 0)) / rho;
	C[1][2][1] = k(1, 2, 1) - (k(0, 2, 0) * k(1, 0, 1) + two * k(1, 1, 0) * k(0, 1, 1)) / rho;
	C[1][1][2] = k(1, 1, 2) - (k(0, 0, 2) * k(1, 1, 0) + two * k(0, 1, 1) * k(1, 0, 1)) / rho;
	C[2][2][0] = k(2, 2, 0) - (k(0, 2, 0) * k(2, 0, 0) + two * k(1, 1, 0) * k(1, 1, 0)) / rho;
	C[0][2][2] = k(0, 2, 2) - (k(0, 0, 2) * k(0, 2, 0) + two * k(0, 1, 1) * k(0, 1, 1)) / rho;
	C[2][0][2] = k(2, 0, 2) - (k(2, 0, 0) * k(0, 0, 2) + two * k(1, 0, 1) * k(1, 0, 1)) / rho;
	C[1][2][2] = k(1, 2, 2) - (k(0, 2, 0) * k(1, 0, 2) + k(0, 0, 2) * k(1, 2, 0) + four * k(0, 1, 1) * k(1, 1, 1) + two * (k(1, 1, 0) * k(0, 1, 2) + k(1, 0, 1) * k(0, 2, 1))) / rho;
	C[2][1][2] = k(2, 1, 2) - (k(0, 0, 2) * k(2, 1, 0) + k(2, 0, 0) * k(0, 1, 2) + four * k(1, 0, 1) * k(1, 1, 1) + two * (k(0, 1, 1) * k(2, 0, 1) + k(1, 1, 0) * k(1, 0, 2))) / rho;
	C[2][2][1] = k(2, 2, 1) - (k(2, 0, 0) * k(0, 2, 1) + k(0, 2, 0) * k(2, 0, 1) + four * k(1, 1, 0) * k(1, 1, 1) + two * (k(1, 0, 1) * k(1, 2, 0) + k(0, 1, 1) * k(2, 1, 0))) / rho;
	C[2][2][2] = k(2, 2, 2)
			   - (four * k(1, 1, 1) * k(1, 1, 1) + k(2, 0, 0) * k(0, 2, 2) + k(0, 2, 0) * k(2, 0, 2) + k(0, 0, 2) * k(2, 2, 0)
				  + four * (k(0, 1, 1) * k(2, 1, 1) + k(1, 0, 1) * k(1, 2, 1) + k(1, 1, 0) * k(1, 1, 2))
				  + two * (k(1, 2, 0) * k(1, 0, 2) + k(2, 1, 0) * k(0, 1, 2) + k(2, 0, 1) * k(0, 2, 1)))
					 / rho
			   + (sixteen * k(1, 1, 0) * k(1, 0, 1) * k(0, 1, 1)
				  + four * (k(1, 0, 1) * k(1, 0, 1) * k(0, 2, 0) + k(0, 1, 1) * k(0, 1, 1) * k(2, 0, 0) + k(1, 1, 0) * k(1, 1, 0) * k(0, 0, 2))
				  + two * k(2, 0, 0) * k(0, 2, 0) * k(0, 0, 2))
					 / rho / rho;

	// relaxation rates (col_cum.h:175-220): only omega1 depends on the viscosity; everything else is 1, A = B = 0,
	// and the velocity-derivative (antialias) terms are 0.  They stay in the formulas so that the evaluation,
	// including products with 0 and 1, is the reference's.
	const R omega1 = one / (three * K.nu + half);
	const R omega2 = one;
	R omega3 = one, omega4 = one, omega5 = one, A = 0, B = 0;
	const R omega6 = one, omega7 = one, omega8 = one, omega9 = one, omega10 = one;
	R w120p102 = 0, w210p012 = 0, w201p021 = 0, w120m102 = 0, w210m012 = 0, w201m021 = 0, w111 = 0;
	if (G2017) {  // col_cum.h:177-208
		const R five = 5, seven = 7, eight = 8, nine = 9, n10 = 10, n11 = 11, n13 = 13, n15 = 15, n18 = 18, n24 = 24, n26 = 26, n28 = 28, n42 = 42, n46 = 46,
				n48 = 48, n56 = 56, n216 = 216, six = 6;
		const R lambda3 = (R) 0.01, lambda4 = (R) 0.01, lambda5 = (R) 0.01;
		omega3 = eight * (omega1 - two) * (omega2 * (three * omega1 - one) - five * omega1)
			   / (eight * (five - two * omega1) * omega1 + omega2 * (eight + omega1 * (nine * omega1 - n26)));
		w120p102 = limited_rate(omega3, C[1][2][0] + C[1][0][2], rho, lambda3);
		w210p012 = limited_rate(omega3, C[2][1][0] + C[0][1][2], rho, lambda3);
		w201p021 = limited_rate(omega3, C[2][0][1] + C[0][2][1], rho, lambda3);
		omega4 = eight * (omega1 - two) * (omega1 + omega2 * (three * omega1 - seven)) / (omega2 * (n56 - n42 * omega1 + nine * omega1 * omega1) - eight * omega1);
		w120m102 = limited_rate(omega4, C[1][2][0] - C[1][0][2], rho, lambda4);
		w210m012 = limited_rate(omega4, C[2][1][0] - C[0][1][2], rho, lambda4);
		w201m021 = limited_rate(omega4, C[2][0][1] - C[0][2][1], rho, lambda4);
		omega5 = n24 * (omega1 - two)
			   * (four * omega1 * omega1 + omega1 * omega2 * (n18 - n13 * omega1) + omega2 * omega2 * (two + omega1 * (six * omega1 - n11)))
			   / (sixteen * omega1 * omega1 * (omega1 - six) - two * omega1 * omega2 * (n216 + five * omega1 * (nine * omega1 - n46))
				  + omega2 * omega2 * (omega1 * (three * omega1 - n10) * (n15 * omega1 - n28) - n48));
		w111 = limited_rate(omega5, C[1][1][1], rho, lambda5);
		A = (four * omega1 * omega1 + two * omega1 * omega2 * (omega1 - six) + omega2 * omega2 * (omega1 * (n10 - three * omega1) - four)) / (omega1 - omega2)
		  / (omega2 * (two + three * omega1) - eight * omega1);
		B = (four * omega1 * omega2 * (nine * omega1 - sixteen) - four * omega1 * omega1 - two * omega2 * omega2 * (two + nine * omega1 * (omega1 - two))) / three
		  / (omega1 - omega2) / (omega2 * (two + three * omega1) - eight * omega1);
	}
	R Dxu = 0, Dyv = 0, Dzw = 0, DxvDyu = 0, DxwDzu = 0, DywDzv = 0;
	if (ANTIALIAS) {  // col_cum.h:215-229
		const R n3o2 = (R) 1.5;
		Dxu = -omega1 / two / rho * (two * C[2][0][0] - C[0][2][0] - C[0][0][2]) - omega2 / two / rho * (C[2][0][0] + C[0][2][0] + C[0][0][2] - (-one + rho));
		Dyv = Dxu + n3o2 * omega1 / rho * (C[2][0][0] - C[0][2][0]);
		Dzw = Dxu + n3o2 * omega1 / rho * (C[2][0][0] - C[0][0][2]);
		DxvDyu = -three * omega1 / rho * C[1][1][0];
		DxwDzu = -three * omega1 / rho * C[1][0][1];
		DywDzv = -three * omega1 / rho * C[0][1][1];
	}

	R S[3][3][3];  // post-collision cumulants (Cs_*), then central moments (ks_*)
	S[1][1][0] = (one - omega1) * C[1][1][0];
	S[1][0][1] = (one - omega1) * C[1][0][1];
	S[0][1][1] = (one - omega1) * C[0][1][1];
	// Eq 33-35 (col_cum.h:249-256)
	const R r33 = (one - omega1) * (C[2][0][0] - C[0][2][0]) - three * rho * (one - omega1 * half) * (vx * vx * Dxu - vy * vy * Dyv);
	const R r34 = (one - omega1) * (C[2][0][0] - C[0][0][2]) - three * rho * (one - omega1 * half) * (vx * vx * Dxu - vz * vz * Dzw);
	const R r35 = k(0, 0, 0) * omega2 + (one - omega2) * (C[2][0][0] + C[0][2][0] + C[0][0][2])
				- three * rho * (one - omega2 / two) * (vx * vx * Dxu + vy * vy * Dyv + vz * vz * Dzw);
	S[2][0][0] = third * (r33 + r34 + r35);
	S[0][2][0] = third * (-two * r33 + r34 + r35);
	S[0][0][2] = third * (r33 - two * r34 + r35);
	if (G2017) {  // limited rates, col_cum.h:258-276
		const R e117 = (one - w120p102) * (C[1][2][0] + C[1][0][2]);
		const R e118 = (one - w210p012) * (C[2][1][0] + C[0][1][2]);
		const R e119 = (one - w201p021) * (C[2][0][1] + C[0][2][1]);
		const R e120 = (one - w120m102) * (C[1][2][0] - C[1][0][2]);
		const R e121 = (one - w210m012) * (C[2][1][0] - C[0][1][2]);
		const R e122 = (one - w201m021) * (C[2][0][1] - C[0][2][1]);
		S[1][2][0] = half * (e120 + e117);
		S[1][0][2] = half * (-e120 + e117);
		S[2][1][0] = half * (e121 + e118);
		S[0][1][2] = half * (-e121 + e118);
		S[0][2][1] = half * (-e122 + e119);
		S[2][0][1] = half * (e122 + e119);
		S[1][1][1] = (one - w111) * C[1][1][1];
	}
	else {
		// Eq 36-41 (col_cum.h:278-285)
		S[1][2][0] = (-C[1][0][2] - C[1][2][0]) * omega3 * half + (C[1][0][2] - C[1][2][0]) * omega4 * half + C[1][2][0];
		S[1][0][2] = (-C[1][0][2] - C[1][2][0]) * omega3 * half + (-C[1][0][2] + C[1][2][0]) * omega4 * half + C[1][0][2];
		S[2][1][0] = (-C[0][1][2] - C[2][1][0]) * omega3 * half + (C[0][1][2] - C[2][1][0]) * omega4 * half + C[2][1][0];
		S[0][1][2] = (-C[0][1][2] - C[2][1][0]) * omega3 * half + (-C[0][1][2] + C[2][1][0]) * omega4 * half + C[0][1][2];
		S[0][2][1] = (-C[0][2][1] - C[2][0][1]) * omega3 * half + (-C[0][2][1] + C[2][0][1]) * omega4 * half + C[0][2][1];
		S[2][0][1] = (-C[0][2][1] - C[2][0][1]) * omega3 * half + (C[0][2][1] - C[2][0][1]) * omega4 * half + C[2][0][1];
		S[1][1][1] = (one - omega5) * C[1][1][1];  // Eq 42
	}
	// Eq 43-45 (col_cum.h:288-297)
	const R r43 = n2o3 * (one / omega1 - half) * omega6 * A * rho * (Dxu - two * Dyv + Dzw) + (one - omega6) * (C[2][2][0] - two * C[2][0][2] + C[0][2][2]);
	const R r44 = n2o3 * (one / omega1 - half) * omega6 * A * rho * (Dxu + Dyv - two * Dzw) + (one - omega6) * (C[2][2][0] + C[2][0][2] - two * C[0][2][2]);
	const R r45 = -n4o3 * (one / omega1 - half) * omega7 * A * rho * (Dxu + Dyv + Dzw) + (one - omega7) * (C[2][2][0] + C[2][0][2] + C[0][2][2]);
	S[2][2][0] = third * (r43 + r44 + r45);
	S[2][0][2] = third * (-r43 + r45);
	S[0][2][2] = third * (-r44 + r45);
	// Eq 46-52 (col_cum.h:299-306)
	S[2][1][1] = -third * (one / omega1 - half) * omega8 * B * rho * DywDzv + (one - omega8) * C[2][1][1];
	S[1][2][1] = -third * (one / omega1 - half) * omega8 * B * rho * DxwDzu + (one - omega8) * C[1][2][1];
	S[1][1][2] = -third * (one / omega1 - half) * omega8 * B * rho * DxvDyu + (one - omega8) * C[1][1][2];
	S[2][2][1] = (one - omega9) * C[2][2][1];
	S[2][1][2] = (one - omega9) * C[2][1][2];
	S[1][2][2] = (one - omega9) * C[1][2][2];
	S[2][2][2] = (one - omega10) * C[2][2][2];
#undef k
#define s(a, b, c) S[a][b][c]
	// cumulants -> central moments, Eq G2015(81)-(84) (col_cum.h:312-338); evaluation order of the reference:
	// the three 211-type, the three 220-type, the three 122-type, then 222
	const R c211 = s(2, 1, 1), c121 = s(1, 2, 1), c112 = s(1, 1, 2), c220 = s(2, 2, 0), c022 = s(0, 2, 2), c202 = s(2, 0, 2);
	const R c122 = s(1, 2, 2), c212 = s(2, 1, 2), c221 = s(2, 2, 1), c222 = s(2, 2, 2);
	S[2][1][1] = c211 + (s(2, 0, 0) * s(0, 1, 1) + two * s(1, 0, 1) * s(1, 1, 0)) / rho;
	S[1][2][1] = c121 + (s(0, 2, 0) * s(1, 0, 1) + two * s(1, 1, 0) * s(0, 1, 1)) / rho;
	S[1][1][2] = c112 + (s(0, 0, 2) * s(1, 1, 0) + two * s(0, 1, 1) * s(1, 0, 1)) / rho;
	S[2][2][0] = c220 + (s(0, 2, 0) * s(2, 0, 0) + two * s(1, 1, 0) * s(1, 1, 0)) / rho;
	S[0][2][2] = c022 + (s(0, 0, 2) * s(0, 2, 0) + two * s(0, 1, 1) * s(0, 1, 1)) / rho;
	S[2][0][2] = c202 + (s(2, 0, 0) * s(0, 0, 2) + two * s(1, 0, 1) * s(1, 0, 1)) / rho;
	S[1][2][2] = c122 + (s(0, 2, 0) * s(1, 0, 2) + s(0, 0, 2) * s(1, 2, 0) + four * s(0, 1, 1) * s(1, 1, 1) + two * (s(1, 1, 0) * s(0, 1, 2) + s(1, 0, 1) * s(0, 2, 1))) / rho;
	S[2][1][2] = c212 + (s(0, 0, 2) * s(2, 1, 0) + s(2, 0, 0) * s(0, 1, 2) + four * s(1, 0, 1) * s(1, 1, 1) + two * (s(0, 1, 1) * s(2, 0, 1) + s(1, 1, 0) * s(1, 0, 2))) / rho;
	S[2][2][1] = c221 + (s(2, 0, 0) * s(0, 2, 1) + s(0, 2, 0) * s(2, 0, 1) + four * s(1, 1, 0) * s(1, 1, 1) + two * (s(1, 0, 1) * s(1, 2, 0) + s(0, 1, 1) * s(2, 1, 0))) / rho;
	S[2][2][2] = c222
			   + (four * s(1, 1, 1) * s(1, 1, 1) + s(2, 0, 0) * s(0, 2, 2) + s(0, 2, 0) * s(2, 0, 2) + s(0, 0, 2) * s(2, 2, 0)
				  + four * (s(0, 1, 1) * s(2, 1, 1) + s(1, 0, 1) * s(1, 2, 1) + s(1, 1, 0) * s(1, 1, 2))
				  + two * (s(1, 2, 0) * s(1, 0, 2) + s(2, 1, 0) * s(0, 1, 2) + s(2, 0, 1) * s(0, 2, 1)))
					 / rho
			   - (sixteen * s(1, 1, 0) * s(1, 0, 1) * s(0, 1, 1)
				  + four * (s(1, 0, 1) * s(1, 0, 1) * s(0, 2, 0) + s(0, 1, 1) * s(0, 1, 1) * s(2, 0, 0) + s(1, 1, 0) * s(1, 1, 0) * s(0, 0, 2))
				  + two * s(2, 0, 0) * s(0, 2, 0) * s(0, 0, 2))
					 / rho / rho;
#undef s
	// zeroth moment kept, first central moments change sign (the reference's forcing convention, col_cum.h:341-345)
	S[0][0][0] = m[0][0][0];
	S[1][0][0] = -m[1][0][0];
	S[0][1][0] = -m[0][1][0];
	S[0][0][1] = -m[0][0][1];

	// backward transform: x, then y, then z
	for (int b = 0; b < 3; b++)
		for (int c = 0; c < 3; c++)
			from_central(S[0][b][c], S[1][b][c], S[2][b][c], vx);
	for (int a = 0; a < 3; a++)
		for (int c = 0; c < 3; c++)
			from_central(S[a][0][c], S[a][1][c], S[a][2][c], vy);
	for (int a = 0; a < 3; a++)
		for (int b = 0; b < 3; b++)
			from_central(S[a][b][0], S[a][b][1], S[a][b][2], vz);
	for (int q = 0; q < 27; q++)
		K.f[q] = S[C27[q][0] + 1][C27[q][1] + 1][C27[q][2] + 1];
}

// ---------------------------------------------------------------------------------------------
// D3Q27 SRT / BGK / MRT_LES
// ---------------------------------------------------------------------------------------------
// source factor shared by SRT and BGK (col_srt.h:25-51, col_bgk.h:62-88): 3*((c-u).F), each axis term written
// (-v-1)*F, -v*F, (-v+1)*F for c = -1, 0, +1
template <typename R>
inline R force_projection(int q, R vx, R vy, R vz, R fx, R fy, R fz)
{
	auto t = [](int c, R v, R F) -> R { return c < 0 ? (-v - (R) 1.0) * F : (c > 0 ? (-v + (R) 1.0) * F : -v * F); };
	return (R) 3.0 * (t(C27[q][0], vx, fx) + t(C27[q][1], vy, fy) + t(C27[q][2], vz, fz));
}

template <typename R>
void collide_srt27(Cell<R, 27>& K, int eqkind)	// col_srt.h:16-108
{
	const R one = 1, half = (R) 0.5;
	const R tau = (R) 3.0 * K.nu + half;
	const R iRho = one / (K.rho == 0 ? one : K.rho);
	R S[27], feq[27];
	for (int q = 0; q < 27; q++) {
		S[q] = force_projection(q, K.vx, K.vy, K.vz, K.fx, K.fy, K.fz) * iRho;
		feq[q] = equilibrium(K, eqkind, q, K.rho, K.vx, K.vy, K.vz);
	}
	for (int q = 0; q < 27; q++)
		K.f[q] += (feq[q] - K.f[q]) / tau + (one - half / tau) * S[q] * feq[q];
}

// col_bgk.h:16-145; product-form equilibrium, EQ argument unused.  GALILEAN: the build with -DUSE_GALILEAN_CORRECTION (defs.h:253,
// col_bgk.h:20-45): the diagonal second moments, summed in the reference's order, correct the zero-velocity factor of each axis.
template <typename R, bool GALILEAN = false>
void collide_bgk27(Cell<R, 27>& K)
{
	const R one = 1, half = (R) 0.5, third = (R) (1.0 / 3.0), three = 3;
	const R omega1 = one / ((R) 3.0 * K.nu + half);
	R g[3][3];	// g[axis][sign+1]
	const R v[3] = {K.vx, K.vy, K.vz};
	R G[3] = {0, 0, 0};
	if (GALILEAN) {
		// m_200: x in (-,+), (y,z) in the order mm mp mz pm pp pz zm zp zz; m_020: x in (-,0,+), y in (-,+), z in (-,+,0);
		// m_002: x in (-,0,+), y in (-,0,+), z in (-,+)   (col_bgk.h:21-26)
		static const int o9[9][2] = {{-1, -1}, {-1, 1}, {-1, 0}, {1, -1}, {1, 1}, {1, 0}, {0, -1}, {0, 1}, {0, 0}};
		static const int x3[3] = {-1, 0, 1}, z3[3] = {-1, 1, 0};
		R m[3] = {0, 0, 0};
		int n = 0;
		for (int a = -1; a <= 1; a += 2)
			for (int i = 0; i < 9; i++, n++)
				m[0] = n == 0 ? K.f[find27(a, o9[i][0], o9[i][1])] : m[0] + K.f[find27(a, o9[i][0], o9[i][1])];
		n = 0;
		for (int a = 0; a < 3; a++)
			for (int b = -1; b <= 1; b += 2)
				for (int c = 0; c < 3; c++, n++)
					m[1] = n == 0 ? K.f[find27(x3[a], b, z3[c])] : m[1] + K.f[find27(x3[a], b, z3[c])];
		n = 0;
		for (int a = 0; a < 3; a++)
			for (int b = 0; b < 3; b++)
				for (int c = -1; c <= 1; c += 2, n++)
					m[2] = n == 0 ? K.f[find27(x3[a], x3[b], c)] : m[2] + K.f[find27(x3[a], x3[b], c)];
		for (int a = 0; a < 3; a++) {
			const R D = -omega1 * half * (three * m[a] / K.rho - one - three * v[a] * v[a]);
			G[a] = -three * v[a] * v[a] * D * (one / omega1 - half);
		}
	}
	for (int a = 0; a < 3; a++) {
		const R z = GALILEAN ? third - one + v[a] * v[a] + G[a] : third - one + v[a] * v[a];
		const R p = -half * (z + one + v[a]);
		g[a][1] = z;
		g[a][2] = p;
		g[a][0] = p + v[a];
	}
	for (int q = 0; q < 27; q++) {
		const R S = force_projection(q, K.vx, K.vy, K.vz, K.fx, K.fy, K.fz) / K.rho;
		const R feq = -K.rho * g[0][C27[q][0] + 1] * g[1][C27[q][1] + 1] * g[2][C27[q][2] + 1];
		K.f[q] += (feq - K.f[q]) * omega1 + (one - half * omega1) * S * feq;
	}
}

template <typename R>
void collide_mrt27(Cell<R, 27>& K)	// col_mrt.h:13-141 ("MRT_LES": relax the non-equilibrium stress with a Smagorinsky rate, rebuild f; no force)
{
	const R one = 1, two = 2, three = 3, third = (R) (1.0 / 3.0);
	// second moments: running sums in lexicographic (x,y,z) order over sign -,0,+ (col_mrt.h:18-31)
	R P[6] = {0, 0, 0, 0, 0, 0};  // xx, yy, zz, xy, xz, yz
	for (int a = -1; a <= 1; a++)
		for (int b = -1; b <= 1; b++)
			for (int c = -1; c <= 1; c++) {
				const R f = K.f[find27(a, b, c)];
				const int w[6] = {a * a, b * b, c * c, a * b, a * c, b * c};
				for (int i = 0; i < 6; i++) {
					if (w[i] > 0)
						P[i] = P[i] + f;
					else if (w[i] < 0)
						P[i] = P[i] - f;
				}
			}
	R &Pxx = P[0], &Pyy = P[1], &Pzz = P[2], &Pxy = P[3], &Pxz = P[4], &Pyz = P[5];
	const R Nxx = Pxx - K.rho * (third + K.vx * K.vx);
	const R Nyy = Pyy - K.rho * (third + K.vy * K.vy);
	const R Nzz = Pzz - K.rho * (third + K.vz * K.vz);
	const R Nxz = Pxz - K.rho * K.vx * K.vz;
	const R Nxy = Pxy - K.rho * K.vx * K.vy;
	const R Nyz = Pyz - K.rho * K.vy * K.vz;
	const R Qn = two * (Nxx * Nxx + Nyy * Nyy + Nzz * Nzz + two * (Nxy * Nxy + Nxz * Nxz + Nyz * Nyz));
	const R tau = three * K.nu + (R) 0.5;
	const R Csm = (R) 0.0342;
	// The reference writes an unqualified sqrt(): on the host g++ resolves it to ::sqrt(double) for either dreal, so for
	// dreal=float the rate is evaluated in double and rounded once (device code would pick sqrtf: a 1-ulp-level difference).
	const double inner = (double) (tau * tau) + (double) (two * Csm * three * three) * ::sqrt((double) Qn) / (double) K.rho;
	const R omega = (R) ((double) two / (::sqrt(inner) + (double) tau));
	Pxx -= omega * Nxx;
	Pyy -= omega * Nyy;
	Pzz -= omega * Nzz;
	Pxy -= omega * Nxy;
	Pxz -= omega * Nxz;
	Pyz -= omega * Nyz;
	for (int q = 0; q < 27; q++) {
		const int a = C27[q][0], b = C27[q][1], c = C27[q][2];
		const int n = (a != 0) + (b != 0) + (c != 0);
		const R w = n == 0 ? (R) (8.0 / 27.0) : n == 1 ? (R) (2.0 / 27.0) : n == 2 ? (R) (1.0 / 54.0) : (R) (1.0 / 216.0);
		K.f[q] = w
			   * (K.rho * ((R) 2.5 - (R) 1.5 * (R) n + three * (K.vx * (R) a + K.vy * (R) b + K.vz * (R) c))
				  + (R) 4.5 * (Pxx * (R) (a * a) + Pyy * (R) (b * b) + Pzz * (R) (c * c) + two * (Pxy * (R) (a * b) + Pxz * (R) (a * c) + Pyz * (R) (b * c)))
				  - (R) 1.5 * (Pxx + Pyy + Pzz));
	}
}

// ---------------------------------------------------------------------------------------------
// D2Q9 SRT and cascaded (CLBM)
// ---------------------------------------------------------------------------------------------
template <typename R>
void collide_srt9(Cell<R, 9>& K)  // d2q9/col_srt.h:16-44
{
	const R one = 1, half = (R) 0.5, three = 3, four = 4, nine = 9, n36 = 36;
	const R tau = three * K.nu + half;
	const R vx = K.vx, vy = K.vy, fx = K.fx, fy = K.fy;
	const R pre = one - half / tau;
	R F[9];
	F[E(0, 0)] = pre * four / nine * (three * (-vx * fx - vy * fy));
	F[E(1, 0)] = pre / nine * (three * ((one - vx) * fx - vy * fy) + nine * vx * fx);
	F[E(-1, 0)] = pre / nine * (three * ((-one - vx) * fx - vy * fy) + nine * vx * fx);
	F[E(0, 1)] = pre / nine * (three * (-vx * fx + (one - vy) * fy) + nine * vy * fy);
	F[E(0, -1)] = pre / nine * (three * (-vx * fx + (-one - vy) * fy) + nine * vy * fy);
	F[E(1, 1)] = pre / n36 * (three * ((one - vx) * fx + (one - vy) * fy) + nine * (vx + vy) * (fx + fy));
	F[E(-1, -1)] = pre / n36 * (three * ((-one - vx) * fx + (-one - vy) * fy) + nine * (vx + vy) * (fx + fy));
	F[E(1, -1)] = pre / n36 * (three * ((one - vx) * fx + (-one - vy) * fy) + nine * (vx - vy) * (fx - fy));
	F[E(-1, 1)] = pre / n36 * (three * ((-one - vx) * fx + (one - vy) * fy) + nine * (vx - vy) * (fx - fy));
	for (int q = 0; q < 9; q++)
		K.f[q] += (eq9_std(q, K.rho, vx, vy) - K.f[q]) / tau + F[q];
}

template <typename R>
void collide_clbm9(Cell<R, 9>& K)  // d2q9/col_clbm.h:13-89 (cascaded operator + Premnath-Banerjee central-moment forcing)
{
	const R tau = (R) 3.0 * K.nu + (R) 0.5;
	const R rho = K.rho, vx = K.vx, vy = K.vy, fx = K.fx, fy = K.fy;
	const R zz = K.f[E(0, 0)], pz = K.f[E(1, 0)], mz = K.f[E(-1, 0)], zp = K.f[E(0, 1)], zm = K.f[E(0, -1)];
	const R pp = K.f[E(1, 1)], mm = K.f[E(-1, -1)], pm = K.f[E(1, -1)], mp = K.f[E(-1, 1)];
	const R c2 = 2, c3 = 3, c4 = 4, c6 = 6, c8 = 8, c9 = 9, c36 = 36, q25 = (R) .25, h5 = (R) .5;

	const R P = (R) 1. / (R) 12. * (rho * (vx * vx + vy * vy) - pz - zp - zm - mz - c2 * (pm + mm + pp + mp - (R) 1. / (R) 3. * rho) - (fx * vx + fy * vy));
	const R NE = q25 / tau * (zp + zm - pz - mz + rho * (vx * vx - vy * vy) - (fx * vx - fy * vy));
	const R V = q25 / tau * ((pp + mm - mp - pm) - vx * vy * rho + h5 * (fx * vy + fy * vx));
	const R kxxyy = (pz + pp + mp + pm + mm + mz - vx * vx * rho + c2 * NE + c6 * P) * (zp + pp + mp + zm + pm + mm - vy * vy * rho - c2 * NE + c6 * P);
	const R UP = (-(q25 * (pm + mm - pp - mp - c2 * vx * vx * vy * rho + vy * (rho - zp - zm - zz) - h5 * (-vx * vx) * fy + fx * vx * vy)
					- vy * h5 * (-c3 * P - NE) + vx * ((pp - mp - pm + mm) * h5 - c2 * V)));
	const R RIGHT = (-(q25 * (mm + mp - pm - pp - c2 * vy * vy * vx * rho + vx * (rho - zz - mz - pz) - h5 * (-vy * vy) * fx + fy * vy * vx)
					   - vx * h5 * (-c3 * P + NE) + vy * ((pp + mm - pm - mp) * h5 - c2 * V)));
	const R NP = (q25
				  * (kxxyy - pp - mp - pm - mm - c8 * P + c2 * (vx * (pp - mp + pm - mm - c4 * RIGHT) + vy * (pp + mp - pm - mm - c4 * UP))
					 + c4 * vx * vy * (-pp + mp + pm - mm + c4 * V) + vx * vx * (-zp - pp - mp - zm - pm - mm + c2 * NE - c6 * P)
					 + vy * vy * ((-pz - pp - mp - pm - mm - mz - c2 * NE - c6 * P) + c3 * vx * vx * rho) - (fx * vx * vy * vy + fy * vy * vx * vx)));

	K.f[E(-1, 1)] += c2 * P + NP + V - UP + RIGHT;
	K.f[E(-1, 0)] += -P - c2 * NP + NE - c2 * RIGHT;
	K.f[E(-1, -1)] += c2 * P + NP - V + UP + RIGHT;
	K.f[E(0, -1)] += -P - c2 * NP - NE - c2 * UP;
	K.f[E(1, -1)] += c2 * P + NP + V + UP - RIGHT;
	K.f[E(1, 0)] += -P - c2 * NP + NE + c2 * RIGHT;
	K.f[E(1, 1)] += c2 * P + NP - V - UP - RIGHT;
	K.f[E(0, 1)] += -P - c2 * NP - NE + c2 * UP;
	K.f[E(0, 0)] += (c4 * (-P + NP));

	const R m1 = fx, m2 = fy;
	const R m3 = c6 * (fx * vx + fy * vy);
	const R m4 = c2 * (fx * vx - fy * vy);
	const R m5 = fx * vy + fy * vx;
	const R m6 = (c2 - c3 * vx * vx) * fy - c6 * fx * vx * vy;
	const R m7 = (c2 - c3 * vy * vy) * fx - c6 * fy * vx * vy;
	const R m8 = c6 * ((c3 * vy * vy - c2) * fx * vx + (c3 * vx * vx - c2) * fy * vy);
	K.f[E(0, 0)] += (-m3 + m8) / c9;
	K.f[E(1, 0)] += (c6 * m1 - m3 + c9 * m4 + c6 * m7 - c2 * m8) / c36;
	K.f[E(0, 1)] += (c6 * m2 - m3 - c9 * m4 + c6 * m6 - c2 * m8) / c36;
	K.f[E(-1, 0)] += (-c6 * m1 - m3 + c9 * m4 - c6 * m7 - c2 * m8) / c36;
	K.f[E(0, -1)] += (-c6 * m2 - m3 - c9 * m4 - c6 * m6 - c2 * m8) / c36;
	K.f[E(1, 1)] += (c6 * m1 + c6 * m2 + c2 * m3 + c9 * m5 - c3 * m6 - c3 * m7 + m8) / c36;
	K.f[E(-1, 1)] += (-c6 * m1 + c6 * m2 + c2 * m3 - c9 * m5 - c3 * m6 + c3 * m7 + m8) / c36;
	K.f[E(-1, -1)] += (-c6 * m1 - c6 * m2 + c2 * m3 + c9 * m5 + c3 * m6 + c3 * m7 + m8) / c36;
	K.f[E(1, -1)] += (c6 * m1 - c6 * m2 + c2 * m3 - c9 * m5 + c3 * m6 - c3 * m7 + m8) / c36;
}

#include "lbm_oracle_ext.h"

template <typename R>
void collide(Cell<R, 27>& K, const oracle_desc& d)
{
	switch (d.coll) {
		case ORC_COLL_CUM: case ORC_COLL_CUM_HP_RHO: collide_cum(K); break;
		case ORC_COLL_SRT: collide_srt27(K, d.eq); break;
		case ORC_COLL_BGK: collide_bgk27(K); break;
		case ORC_COLL_BGK_GALILEAN: collide_bgk27<R, true>(K); break;
		case ORC_COLL_MRT_LES: collide_mrt27(K); break;
		case ORC_COLL_CLBM: collide_clbm27(K); break;
		case ORC_COLL_SRT_MODIF_FORCE: collide_srt_modif27(K, d.eq); break;
		case ORC_COLL_KBC_N1: case ORC_COLL_KBC_N2: case ORC_COLL_KBC_N3: case ORC_COLL_KBC_N4:
		case ORC_COLL_KBC_C1: case ORC_COLL_KBC_C2: case ORC_COLL_KBC_C3: case ORC_COLL_KBC_C4: {
			const int k = (d.coll - ORC_COLL_KBC_N1) % 4;  // 0: D, 1: D+T, 2: D+Q, 3: D+T+Q
			collide_kbc27(K, d.coll >= ORC_COLL_KBC_C1, k == 1 || k == 3, k >= 2);
			break;
		}
		case ORC_COLL_CUM_2017: collide_cum<R, true, false>(K); break;
		case ORC_COLL_CUM_ANTIALIAS: collide_cum<R, false, true>(K); break;
		case ORC_COLL_CUM_2017_ANTIALIAS: collide_cum<R, true, true>(K); break;
	}
}
// D3Q19 (unpinned): SRT and MRT_LES with the formulas of col_srt.h / col_mrt.h over 19 velocities
template <typename R>
void collide(Cell<R, 19>& K, const oracle_desc& d)
{
	const R one = 1, half = (R) 0.5, two = 2, three = 3, third = (R) (1.0 / 3.0);
	if (d.coll == ORC_COLL_SRT) {
		const R tau = three * K.nu + half;
		const R iRho = one / (K.rho == 0 ? one : K.rho);
		R S[19], feq[19];
		for (int q = 0; q < 19; q++) {
			S[q] = force_projection(q, K.vx, K.vy, K.vz, K.fx, K.fy, K.fz) * iRho;
			feq[q] = equilibrium(K, 0, q, K.rho, K.vx, K.vy, K.vz);
		}
		for (int q = 0; q < 19; q++)
			K.f[q] += (feq[q] - K.f[q]) / tau + (one - half / tau) * S[q] * feq[q];
		return;
	}
	R P[6] = {0, 0, 0, 0, 0, 0};
	for (int q = 0; q < 19; q++) {
		const int a = C27[q][0], b = C27[q][1], c = C27[q][2];
		const int w[6] = {a * a, b * b, c * c, a * b, a * c, b * c};
		for (int i = 0; i < 6; i++)
			P[i] += (R) w[i] * K.f[q];
	}
	const R N[6] = {P[0] - K.rho * (third + K.vx * K.vx), P[1] - K.rho * (third + K.vy * K.vy), P[2] - K.rho * (third + K.vz * K.vz),
					P[3] - K.rho * K.vx * K.vy, P[4] - K.rho * K.vx * K.vz, P[5] - K.rho * K.vy * K.vz};
	const R Qn = two * (N[0] * N[0] + N[1] * N[1] + N[2] * N[2] + two * (N[3] * N[3] + N[4] * N[4] + N[5] * N[5]));
	const R tau = three * K.nu + half;
	const R omega = two / (std::sqrt(tau * tau + two * (R) 0.0342 * three * three * std::sqrt(Qn) / K.rho) + tau);
	for (int i = 0; i < 6; i++)
		P[i] -= omega * N[i];
	for (int q = 0; q < 19; q++) {
		const int a = C27[q][0], b = C27[q][1], c = C27[q][2];
		const int n = (a != 0) + (b != 0) + (c != 0);
		const R w = n == 0 ? third : n == 1 ? (R) (1.0 / 18.0) : (R) (1.0 / 36.0);
		K.f[q] = w
			   * (K.rho * ((R) 2.5 - (R) 1.5 * (R) n + three * (K.vx * (R) a + K.vy * (R) b + K.vz * (R) c))
				  + (R) 4.5 * (P[0] * (R) (a * a) + P[1] * (R) (b * b) + P[2] * (R) (c * c) + two * (P[3] * (R) (a * b) + P[4] * (R) (a * c) + P[5] * (R) (b * c)))
				  - (R) 1.5 * (P[0] + P[1] + P[2]));
	}
}

template <typename R>
void collide(Cell<R, 9>& K, const oracle_desc& d)
{
	if (d.coll == ORC_COLL_SRT)
		collide_srt9(K);
	else
		collide_clbm9(K);
}

// ---------------------------------------------------------------------------------------------
// streaming
// ---------------------------------------------------------------------------------------------
// A-B pull (d3q27/streaming_AB.h:21-58, d2q9/streaming_AB.h:22-36): f[q](x) <- df_cur[q](x - c_q)
// A-A (d3q27/streaming_AA.h:78-116): even: f[q] <- df[q](x);  odd: f[opp q] <- df[q](x + c_q)
template <typename L, typename R, int Q>
void stream_in(const Block<R>& B, Cell<R, Q>& K, const Nbr& n, bool aa)
{
	for (int q = 0; q < Q; q++) {
		const int cx = L::c(q, 0), cy = L::c(q, 1), cz = L::c(q, 2);
		if (! aa)
			K.f[q] = B.cur[B.at(q, pick(-cx, n.xm, n.x, n.xp), pick(-cy, n.ym, n.y, n.yp), pick(-cz, n.zm, n.z, n.zp))];
		else if (B.even)
			K.f[q] = B.cur[B.at(q, n.x, n.y, n.z)];
		else
			K.f[L::opp(q)] = B.cur[B.at(q, pick(cx, n.xm, n.x, n.xp), pick(cy, n.ym, n.y, n.yp), pick(cz, n.zm, n.z, n.zp))];
	}
}

// A-B: df_out[q](x) <- f[q] (streaming_AB.h:12-19).  A-A (streaming_AA.h:12-76): even: df[opp q](x) <- f[q]; odd: df[q](x + c_q) <- f[q]
template <typename L, typename R, int Q>
void stream_out(const Block<R>& B, const Cell<R, Q>& K, const Nbr& n, bool aa)
{
	for (int q = 0; q < Q; q++) {
		const int cx = L::c(q, 0), cy = L::c(q, 1), cz = L::c(q, 2);
		if (! aa)
			B.out[B.at(q, n.x, n.y, n.z)] = K.f[q];
		else if (B.even)
			B.cur[B.at(L::opp(q), n.x, n.y, n.z)] = K.f[q];
		else
			B.cur[B.at(q, pick(cx, n.xm, n.x, n.xp), pick(cy, n.ym, n.y, n.yp), pick(cz, n.zm, n.z, n.zp))] = K.f[q];
	}
}

// c_s-weighted interpolation for the populations that enter through the right face (A-B only;
// d3q27/streaming_AB.h:209-242, d2q9/streaming_AB.h:60-74)
template <typename L, typename R, int Q>
void stream_in_interp_right(const Block<R>& B, Cell<R, Q>& K, const Nbr& n)
{
	constexpr R cs = (R) 0.5773502691896257;
	for (int q = 0; q < Q; q++) {
		const int cx = L::c(q, 0), cy = L::c(q, 1), cz = L::c(q, 2);
		const idx yy = pick(-cy, n.ym, n.y, n.yp), zz = pick(-cz, n.zm, n.z, n.zp);
		if (cx < 0)
			K.f[q] = cs * B.cur[B.at(q, n.xm, yy, zz)] + (1 - cs) * B.cur[B.at(q, n.x, yy, zz)];
		else
			K.f[q] = B.cur[B.at(q, pick(-cx, n.xm, n.x, n.xp), yy, zz)];
	}
}

// ---------------------------------------------------------------------------------------------
// inflow velocity (lbm_data.h:98-131, sim_NSE/sim_2.cu:16-33, sim_2D/sim2d_1.cu:20-35)
// ---------------------------------------------------------------------------------------------
template <typename R, int Q>
void inflow(const Block<R>& B, Cell<R, Q>& K, idx, idx y, idx z)
{
	switch (B.inflow_kind) {
		case ORC_INFLOW_CONST:
			K.vx = B.in_vx;
			K.vy = B.in_vy;
			if (Q != 9)
				K.vz = B.in_vz;
			break;
		case ORC_INFLOW_PROFILE_YZ:
			K.vx = B.profile[y + z * B.profile_sy];
			K.vy = 0;
			K.vz = 0;
			break;
		case ORC_INFLOW_PARABOLIC_Y: {	// sim_2D/sim2d_3.cu:46-53: the double literals make the profile a double expression
			R s = (R) (y - (idx) B.in_vy) * B.in_vz;  // in_vy carries y0, in_vz carries 1 / (y1 - y0)
			if (s < 0)
				s = 0;
			else if (s > 1)
				s = 1;
			K.vx = (R) ((double) B.in_vx * (4.0 * (double) s * (1.0 - (double) s)));
			K.vy = 0;
			break;
		}
		default:
			K.rho = 1;
			K.vx = 0;
			K.vy = 0;
			K.vz = 0;
	}
}

// ---------------------------------------------------------------------------------------------
// cell rules before the collision: D3Q27_BC_All::preCollision (d3q27/bc.h:51-241), D2Q9_BC_All::preCollision (d2q9/bc.h:89-196)
// ---------------------------------------------------------------------------------------------
// moment inflow condition on the left face (Eichler 2024), d3q27/bc.h:82-136
template <typename R>
void inflow_left_moments(Cell<R, 27>& K)
{
	R* f = K.f;
	const R third = (R) (1.0 / 3.0), ninth = (R) (1.0 / 9.0);
	const R vx = K.vx, vy = K.vy, vz = K.vz;
	const R zero_plane = f[D(0, 0, 0)]
					   + (+((f[D(0, 1, 1)] + f[D(0, -1, -1)]) + (f[D(0, 1, -1)] + f[D(0, -1, 1)])) + ((f[D(0, 1, 0)] + f[D(0, -1, 0)]) + (f[D(0, 0, 1)] + f[D(0, 0, -1)])));
	const R minus_corner_edge = +((f[D(-1, 1, 1)] + f[D(-1, -1, -1)]) + (f[D(-1, 1, -1)] + f[D(-1, -1, 1)])) + ((f[D(-1, 1, 0)] + f[D(-1, -1, 0)]) + (f[D(-1, 0, 1)] + f[D(-1, 0, -1)]));
	K.rho = (R) 1.0 / (1 - vx) * ((zero_plane) + 2 * (f[D(-1, 0, 0)] + (minus_corner_edge)));
	const R rho = K.rho;
	const R m100 = rho * vx;
	const R m010 = rho * vy;
	const R m001 = rho * vz;
	const R m011 = rho * (vy * vz);
	const R m020 = third * rho + rho * (vy * vy);
	const R m002 = third * rho + rho * (vz * vz);
	const R m021 = third * rho * vz + rho * ((vy * vy) * vz);
	const R m012 = third * rho * vy + rho * (vy * (vz * vz));
	const R m022 = ninth * rho + third * rho * (vy * vy + vz * vz) + rho * (vy * vy) * (vz * vz);
	const R zero_ring = +((f[D(0, 1, 1)] + f[D(0, -1, -1)]) + (f[D(0, 1, -1)] + f[D(0, -1, 1)])) + ((f[D(0, 0, 1)] + f[D(0, 0, -1)]) + (f[D(0, 1, 0)] + f[D(0, -1, 0)]));
	f[D(1, 0, 0)] = m100 + (m022 - (m020 + m002)) + f[D(-1, 0, 0)] + (zero_ring) + 2 * (minus_corner_edge);
	const R h = (R) 0.5, qt = (R) 0.25;
	f[D(1, 1, 0)] = h * ((m020 - m022) + (-m012 + m010)) - (f[D(-1, 1, 0)] + f[D(0, 1, 0)]);
	f[D(1, -1, 0)] = h * ((m020 - m022) + (m012 - m010)) - (f[D(-1, -1, 0)] + f[D(0, -1, 0)]);
	f[D(1, 0, 1)] = h * ((m002 - m022) + (-m021 + m001)) - (f[D(-1, 0, 1)] + f[D(0, 0, 1)]);
	f[D(1, 0, -1)] = h * ((m002 - m022) + (m021 - m001)) - (f[D(-1, 0, -1)] + f[D(0, 0, -1)]);
	f[D(1, 1, 1)] = qt * ((m022 + m011) + (m021 + m012)) - (f[D(-1, 1, 1)] + f[D(0, 1, 1)]);
	f[D(1, 1, -1)] = qt * ((m022 - m011) + (-m021 + m012)) - (f[D(-1, 1, -1)] + f[D(0, 1, -1)]);
	f[D(1, -1, 1)] = qt * ((m022 - m011) + (m021 - m012)) - (f[D(-1, -1, 1)] + f[D(0, -1, 1)]);
	f[D(1, -1, -1)] = qt * ((m022 + m011) + (-m021 - m012)) - (f[D(-1, -1, -1)] + f[D(0, -1, -1)]);
}
template <typename R>
void inflow_left_moments(Cell<R, 9>&)
{}
template <typename R>
void inflow_left_moments(Cell<R, 19>& K)  // D3Q19 (unpinned): no moment condition exists for 19 velocities; impose the equilibrium of (1, u_in)
{
	K.rho = 1;
	for (int q = 0; q < 19; q++)
		K.f[q] = equilibrium(K, 0, q, K.rho, K.vx, K.vy, K.vz);
}

// REFERENCE QUIRK, reproduced on purpose (D2Q9 only).  D2Q9_BC_All::preCollision takes coordinate parameters named
// `zm` and `zp` (d2q9/bc.h:90) which shadow the direction enumerators zp=3 / zm=4 (defs.h:262-263).  Inside that function
// KS.f[zm] and KS.f[zp] therefore index the populations with the neighbour z-COORDINATES: the wall rule (bc.h:135),
// SYM_TOP (bc.h:190) and SYM_BOTTOM (bc.h:184) do not touch the straight +-y populations but f[z-1] / f[z+1] instead.
// On the X x Y x 1 lattice under A-B both coordinates clamp to 0, so the statement degenerates to a no-op on f[0];
// under A-A a non-periodic cell has zm=-1, zp=1 and the reference indexes f[-1] (undefined behaviour): skipped here.
// Given the direction (cx,cy) of the nominal pair, rewrites (a,b) to what the reference really indexes; false = skip.
inline bool shadow_pair(int cx, int cy, const Nbr& n, int& a, int& b)
{
	if (cx != 0 || cy == 0)
		return true;  // only the (0,+1)/(0,-1) pair is written with the shadowed names
	const int ia = (int) (cy > 0 ? n.zp : n.zm), ib = (int) (cy > 0 ? n.zm : n.zp);
	if (ia < 0 || ia >= 9 || ib < 0 || ib >= 9)
		return false;
	a = ia;
	b = ib;
	return true;
}

// D2Q9 GEO_FLUID_NEAR_WALL: Bouzidi interpolated bounce-back on the links that hit a wall (d2q9/bc.h:61-87,140-167), A-B only.
// theta < 0: the link does not hit a wall, plain pull.  Called after the ordinary streaming has filled K.f.
// Same shadowing quirk as the wall rule: the two statements for the straight +-y links are written `KS.f[zp] = ...` /
// `KS.f[zm] = ...` with the shadowed names (bc.h:153,155), so they assign f[0] (overwritten by the rest-particle line right
// after) and the +-y populations keep their plainly streamed values.
template <typename R>
void bouzidi_near_wall(const Block<R>& B, Cell<R, 9>& K, const Nbr& n)
{
	auto theta = [&](int dir) -> R { return B.bouzidi ? B.bouzidi[dir * B.XYZ + B.cell(n.x, n.y, n.z)] : (R) -1; };
	auto fb = [&](R th, int k, int kbar, idx xB, idx yB, idx xS, idx yS) -> R {
		if (th < (R) 0)
			return B.cur[B.at(kbar, xS, yS, n.z)];
		const R fA = B.cur[B.at(k, n.x, n.y, n.z)], fOppA = B.cur[B.at(kbar, n.x, n.y, n.z)], fB = B.cur[B.at(k, xB, yB, n.z)];
		if (th <= (R) 0.5)
			return (R) 2.0 * th * fA + ((R) 1.0 - (R) 2.0 * th) * fB;
		const R w = (R) 0.5 / th;
		return ((R) 1.0 - w) * fOppA + w * fA;
	};
	const R th_e = theta(0), th_w = theta(2), th_ne = theta(4), th_nw = theta(5), th_sw = theta(6), th_se = theta(7);
	K.f[E(1, 0)] = fb(th_w, E(-1, 0), E(1, 0), n.xp, n.y, n.xm, n.y);
	K.f[E(-1, 0)] = fb(th_e, E(1, 0), E(-1, 0), n.xm, n.y, n.xp, n.y);
	K.f[E(1, 1)] = fb(th_sw, E(-1, -1), E(1, 1), n.xp, n.yp, n.xm, n.ym);
	K.f[E(-1, 1)] = fb(th_se, E(1, -1), E(-1, 1), n.xm, n.yp, n.xp, n.ym);
	K.f[E(-1, -1)] = fb(th_ne, E(1, 1), E(-1, -1), n.xm, n.ym, n.xp, n.yp);
	K.f[E(1, -1)] = fb(th_nw, E(-1, 1), E(1, -1), n.xp, n.ym, n.xm, n.yp);
	K.f[E(0, 0)] = B.cur[B.at(E(0, 0), n.x, n.y, n.z)];
}
template <typename R, int Q>
void bouzidi_near_wall(const Block<R>&, Cell<R, Q>&, const Nbr&)
{}

// returns false for GEO_NOTHING (cell neither reads nor writes distributions)
template <typename L, typename R, int Q>
bool pre_collision(const Block<R>& B, Cell<R, Q>& K, const oracle_desc& d, int m, Nbr n, bool aa)
{
	if (m == L::NOTHING) {
		K.rho = 1;
		K.vx = K.vy = K.vz = 0;
		return false;
	}
	if (m == L::OUTFLOW_RIGHT)	// pull from the cell to the left for all three x positions (bc.h:63-65)
		n.xp = n.x = n.xm;
	if (m != L::OUTFLOW_RIGHT_INTERP)
		stream_in<L>(B, K, n, aa);

	if (m == L::INFLOW) {
		inflow(B, K, n.x, n.y, n.z);
		K.rho = 1;
		set_equilibrium(K, d.eq);
	}
	else if (m == L::INFLOW_LEFT) {
		inflow(B, K, n.x, n.y, n.z);
		inflow_left_moments(K);
	}
	else if (m == L::OUTFLOW_EQ) {
		density_velocity(K);
		K.rho = 1;
		set_equilibrium(K, d.eq);
	}
	else if (m == L::OUTFLOW_RIGHT) {
		density_velocity(K);
		K.rho = 1;
	}
	else if (m == L::OUTFLOW_RIGHT_INTERP) {
		stream_in_interp_right<L>(B, K, n);
		density_velocity(K);
		add_equilibrium_difference(K, d.eq, (R) 1);
		K.rho = 1;
	}
	else if (m == L::WALL) {  // full-way bounce-back: swap every population with its opposite; reported rho=1, u=0
		K.rho = 1;
		K.vx = K.vy = K.vz = 0;
		for (int q = 1; q < Q; q++) {
			int o = L::opp(q), a = q;
			if (o < q)
				continue;
			if (L::NDIM == 2 && ! shadow_pair(L::c(q, 0), L::c(q, 1), n, a, o))
				continue;
			const R t = K.f[a];
			K.f[a] = K.f[o];
			K.f[o] = t;
		}
	}
	else if (m == L::SYM_TOP || m == L::SYM_BOTTOM || m == L::SYM_LEFT || m == L::SYM_RIGHT || (L::NDIM == 3 && (m == 13 || m == 14))) {
		// mirror the populations that point away from the symmetry plane (d3q27/bc.h:172-237, d2q9/bc.h:168-191):
		//   3-D: TOP  f[..-] <- f[..+]   BOTTOM f[..+] <- f[..-]   LEFT f[+..] <- f[-..]   RIGHT f[-..] <- f[+..]
		//        BACK f[.+.] <- f[.-.]   FRONT  f[.-.] <- f[.+.]
		//   2-D: TOP  f[.-]  <- f[.+]    BOTTOM f[.+]  <- f[.-]    LEFT f[+.]  <- f[-.]    RIGHT f[-.]  <- f[+.]
		int axis, dst;
		if (m == L::SYM_TOP) {
			axis = L::NDIM - 1;
			dst = -1;
		}
		else if (m == L::SYM_BOTTOM) {
			axis = L::NDIM - 1;
			dst = +1;
		}
		else if (m == L::SYM_LEFT) {
			axis = 0;
			dst = +1;
		}
		else if (m == L::SYM_RIGHT) {
			axis = 0;
			dst = -1;
		}
		else if (m == 13) {	 // SYM_BACK
			axis = 1;
			dst = +1;
		}
		else {	// SYM_FRONT
			axis = 1;
			dst = -1;
		}
		for (int q = 0; q < Q; q++)
			if (L::c(q, axis) == dst) {
				int cc[3] = {L::c(q, 0), L::c(q, 1), L::c(q, 2)};
				cc[axis] = -dst;
				int to = q, from = L::find(cc[0], cc[1], cc[2]);
				if (L::NDIM == 2 && ! shadow_pair(L::c(q, 0), L::c(q, 1), n, to, from))
					continue;
				K.f[to] = K.f[from];
			}
		density_velocity(K);
	}
	else {
		if (L::NDIM == 2 && m == 12 && ! aa)  // D2Q9 GEO_FLUID_NEAR_WALL (d2q9/bc.h:29); under A-A the reference's df_cur reads are meaningless
			bouzidi_near_wall(B, K, n);
		density_velocity(K);
	}
	return true;
}

// ---------------------------------------------------------------------------------------------
// macroscopic output: d3q27/macro.h:50-171, d2q9/macro.h:49-140
// ---------------------------------------------------------------------------------------------
template <typename R, int Q>
void output_macro(const Block<R>& B, const Cell<R, Q>& K, const oracle_desc& d, idx x, idx y, idx z)
{
	if (d.macro == ORC_MACRO_VOID)
		return;
	const idx c = B.cell(x, y, z);
	R* M = B.macro;
	const int nd = Q == 9 ? 2 : 3;
	const R v[3] = {K.vx, K.vy, K.vz};
	M[0 * B.XYZ + c] = K.rho;
	for (int a = 0; a < nd; a++)
		M[(1 + a) * B.XYZ + c] = v[a];
	if (d.macro == ORC_MACRO_WITH_MEAN_2D) {  // D2Q9_MACRO_WithMean::outputMacro, sim_2D/sim2d_2.cu:75-95
		if (B.macro_gates & ORC_GATE_MEANS) {
			M[3 * B.XYZ + c] += K.vx;
			M[4 * B.XYZ + c] += K.vy;
		}
		if (B.macro_gates & ORC_GATE_FLUCS) {
			const R dux = K.vx - M[5 * B.XYZ + c];
			const R duy = K.vy - M[6 * B.XYZ + c];
			const R mag = (R) std::sqrt(dux * dux + duy * duy);
			M[7 * B.XYZ + c] += mag;
			M[8 * B.XYZ + c] += dux * dux;
			M[9 * B.XYZ + c] += duy * duy;
		}
		return;
	}
	if (d.macro != ORC_MACRO_MEAN)
		return;
	// running mean and Welford co-moments; component order: means, then xx,yy,zz,xy,xz,yz (3-D) / xx,yy,xy (2-D)
	const R denom = R(1) / R(B.stat_counter + 1);
	const int base_mean = 1 + nd, base_cov = 1 + 2 * nd;
	R delta[3], delta_new[3];
	for (int a = 0; a < nd; a++) {
		const R old = M[(base_mean + a) * B.XYZ + c];
		delta[a] = v[a] - old;
		const R now = old + delta[a] * denom;
		delta_new[a] = v[a] - now;
		M[(base_mean + a) * B.XYZ + c] = now;
	}
	const int pairs3[6][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 1}, {0, 2}, {1, 2}};
	const int pairs2[3][2] = {{0, 0}, {1, 1}, {0, 1}};
	const int np = nd == 3 ? 6 : 3;
	for (int i = 0; i < np; i++) {
		const int a = nd == 3 ? pairs3[i][0] : pairs2[i][0];
		const int b = nd == 3 ? pairs3[i][1] : pairs2[i][1];
		const R old = M[(base_cov + i) * B.XYZ + c];
		M[(base_cov + i) * B.XYZ + c] = old + delta_new[a] * delta[b];
	}
}

// ---------------------------------------------------------------------------------------------
// one cell update: LBMKernel<NSE>, kernels.h:60-100
// ---------------------------------------------------------------------------------------------
template <typename L, typename R>
void cell_update(const Block<R>& B, const oracle_desc& d, idx x, idx y, idx z)
{
	constexpr int Q = L::Q;
	const bool aa = d.streaming == ORC_STREAM_AA;
	const int m = B.map[B.cell(x, y, z)];
	const Nbr n = neighbours(B, m == L::PERIODIC, aa, x, y, z);
	Cell<R, Q> K;
	K.kahan_rho = d.coll == ORC_COLL_CUM_HP_RHO;
	if (d.macro != ORC_MACRO_VOID) {  // copyQuantities (macro.h:73-80); MACRO_Void leaves the KernelStruct defaults
		K.nu = B.nu;
		K.fx = B.fx;
		K.fy = B.fy;
		K.fz = Q != 9 ? B.fz : 0;
	}
	const bool active = pre_collision<L>(B, K, d, m, n, aa);
	if (L::collides(m))
		collide(K, d);
	if (active)
		stream_out<L>(B, K, n, aa);	 // written at the true (x,y,z) also for OUTFLOW_RIGHT (kernels.h:97 passes the unmodified indices)
	output_macro(B, K, d, x, y, z);
}

template <typename R>
Block<R> make_block(const oracle_desc* d, const oracle_params* p)
{
	Block<R> B{};
	B.X = d->X;
	B.Y = d->Y;
	B.Z = d->Z;
	B.ox = d->ox;
	B.XYZ = (d->X + 2 * d->ox) * d->Y * d->Z;
	B.nproc = d->nproc;
	B.inflow_kind = d->inflow;
	if (p) {
		B.nu = (R) p->lbmViscosity;
		B.fx = (R) p->fx;
		B.fy = (R) p->fy;
		B.fz = (R) p->fz;
		B.in_vx = (R) p->inflow_vx;
		B.in_vy = (R) p->inflow_vy;
		B.in_vz = (R) p->inflow_vz;
		B.profile = (const R*) p->vx_profile;
		B.profile_sy = p->profile_size_y;
		B.stat_counter = p->stat_counter;
		B.macro_gates = p->macro_gates;
		B.bouzidi = (const R*) p->bouzidi_coeff;
	}
	return B;
}

template <typename L, typename R>
int run_steps(const oracle_desc* d, const oracle_params* p, void* df_a, void* df_b, void* macro, const int16_t* map, int64_t iteration, int nsteps,
			  int nthreads)
{
	Block<R> B = make_block<R>(d, p);
	B.macro = (R*) macro;
	B.map = map;
	const bool aa = d->streaming == ORC_STREAM_AA;
	for (int64_t it = iteration; it < iteration + nsteps; it++) {
		B.even = (it % 2) == 0;	 // lbm.hpp:318
		if (aa)
			B.cur = B.out = (R*) df_a;
		else {	// A-B ping-pong, lbm.hpp:320-327
			B.cur = (R*) (B.even ? df_a : df_b);
			B.out = (R*) (B.even ? df_b : df_a);
		}
		// visiting order and threading of the reference's host loop, state.hpp:1114-1123
#pragma omp parallel for schedule(static) collapse(2) num_threads(nthreads)
		for (idx x = 0; x < B.X; x++)
			for (idx z = 0; z < B.Z; z++)
				for (idx y = 0; y < B.Y; y++)
					cell_update<L, R>(B, *d, x, y, z);
	}
	return 0;
}

// LBM_BLOCK::setEquilibrium (lbm_block.hpp:219-250): all sites including ghost planes.  The reference passes
// `real` (double) arguments that are narrowed to dreal at the call of EQ::eq_* (common.h:126-158).
template <typename L, typename R>
int fill_equilibrium(const oracle_desc* d, void* df, const double* rho, const double* vx, const double* vy, const double* vz, double crho, double cvx,
					 double cvy, double cvz)
{
	Block<R> B = make_block<R>(d, nullptr);
	R* f = (R*) df;
	Cell<R, L::Q> K;
	for (idx x = -B.ox; x < B.X + B.ox; x++)
		for (idx z = 0; z < B.Z; z++)
			for (idx y = 0; y < B.Y; y++) {
				const idx c = B.cell(x, y, z);
				const R r = (R) (rho ? rho[c] : crho), ux = (R) (rho ? vx[c] : cvx), uy = (R) (rho ? vy[c] : cvy), uz = (R) (rho ? (vz ? vz[c] : 0.0) : cvz);
				for (int q = 0; q < L::Q; q++)
					f[q * B.XYZ + c] = equilibrium(K, d->eq, q, r, ux, uy, uz);
			}
	return 0;
}

// LBM_BLOCK::computeInitialMacro (lbm_block.hpp:252-277): local read, force zeroed, rho/u, output
template <typename L, typename R>
int initial_macro(const oracle_desc* d, const oracle_params* p, void* df, void* macro)
{
	Block<R> B = make_block<R>(d, p);
	B.macro = (R*) macro;
	const R* f = (const R*) df;
	for (idx x = 0; x < B.X; x++)
		for (idx z = 0; z < B.Z; z++)
			for (idx y = 0; y < B.Y; y++) {
				Cell<R, L::Q> K;
				K.kahan_rho = d->coll == ORC_COLL_CUM_HP_RHO;
				for (int q = 0; q < L::Q; q++)
					K.f[q] = f[B.at(q, x, y, z)];
				if (d->macro != ORC_MACRO_VOID)
					K.nu = B.nu;
				K.fx = K.fy = K.fz = 0;
				density_velocity(K);
				output_macro(B, K, *d, x, y, z);
			}
	return 0;
}

bool supported(const oracle_desc* d)
{
	if (d->precision != ORC_F32 && d->precision != ORC_F64)
		return false;
	if (d->streaming != ORC_STREAM_AB && d->streaming != ORC_STREAM_AA)
		return false;
	if (d->macro < ORC_MACRO_VOID || d->macro > ORC_MACRO_WITH_MEAN_2D || (d->macro == ORC_MACRO_WITH_MEAN_2D && d->lattice != ORC_D2Q9))
		return false;
	if (d->lattice == ORC_D3Q27)
		return ((d->coll >= ORC_COLL_CUM && d->coll <= ORC_COLL_SRT_MODIF_FORCE) || (d->coll >= ORC_COLL_CUM_2017 && d->coll <= ORC_COLL_CUM_HP_RHO))
			&& (d->eq == ORC_EQ_STD || d->eq == ORC_EQ_INV_CUM || d->eq == ORC_EQ_ENTROPIC);
	if (d->lattice == ORC_D2Q9)
		return (d->coll == ORC_COLL_SRT || d->coll == ORC_COLL_CLBM) && d->eq == ORC_EQ_STD && d->Z == 1;
	if (d->lattice == ORC_D3Q19)  // no reference implementation: PARITY UNPINNED (see L19)
		return (d->coll == ORC_COLL_SRT || d->coll == ORC_COLL_MRT_LES) && d->eq == ORC_EQ_STD;
	return false;
}

}  // namespace

#define DISPATCH(fn, ...)                                                                 \
	if (! supported(d))                                                                   \
		return -1;                                                                        \
	if (d->lattice == ORC_D3Q27)                                                          \
		return d->precision == ORC_F64 ? fn<L27, double>(__VA_ARGS__) : fn<L27, float>(__VA_ARGS__); \
	if (d->lattice == ORC_D3Q19)                                                          \
		return d->precision == ORC_F64 ? fn<L19, double>(__VA_ARGS__) : fn<L19, float>(__VA_ARGS__); \
	return d->precision == ORC_F64 ? fn<L9, double>(__VA_ARGS__) : fn<L9, float>(__VA_ARGS__);

extern "C" {

const char* oracle_kind(void)
{
	return "port";
}

int oracle_supported(const oracle_desc* d)
{
	return supported(d) ? 0 : -1;
}

int oracle_step(const oracle_desc* d, const oracle_params* p, void* df_a, void* df_b, void* macro, const int16_t* map, int64_t iteration,
				int32_t nsteps, int32_t nthreads)
{
	if (nthreads < 1)
		nthreads = 1;
	DISPATCH(run_steps, d, p, df_a, df_b, macro, map, iteration, nsteps, nthreads)
}

int oracle_set_equilibrium(const oracle_desc* d, void* df, double rho, double vx, double vy, double vz)
{
	DISPATCH(fill_equilibrium, d, df, nullptr, nullptr, nullptr, nullptr, rho, vx, vy, vz)
}

int oracle_set_equilibrium_field(const oracle_desc* d, void* df, const double* rho, const double* vx, const double* vy, const double* vz)
{
	DISPATCH(fill_equilibrium, d, df, rho, vx, vy, vz, 0, 0, 0, 0)
}

int oracle_initial_macro(const oracle_desc* d, const oracle_params* p, void* df, void* macro)
{
	DISPATCH(initial_macro, d, p, df, macro)
}
}
