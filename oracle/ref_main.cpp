// ref_main.cpp -- oracle_api.h implemented on top of the reference's own code (TEST INFRASTRUCTURE ONLY)
#include "ref_common.h"

static int ref_dispatch(const RefCall& c)
{
	int r;
	if ((r = ref_dispatch_d3q27_cum(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_srt(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_bgk(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_bgk_gal(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_cum_hprho(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_mrt(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_clbm(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_srtmf(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_cum2017(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_cumaa(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_cum2017aa(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_kbc_n(c)) != -1)
		return r;
	if ((r = ref_dispatch_d3q27_kbc_c(c)) != -1)
		return r;
	if ((r = ref_dispatch_d2q9(c)) != -1)
		return r;
	return -1;
}

extern "C" {

const char* oracle_kind(void)
{
	return "reference";
}

int oracle_supported(const oracle_desc* d)
{
	RefCall c{};
	c.op = 0;
	c.d = d;
	return ref_dispatch(c);
}

int oracle_step(const oracle_desc* d, const oracle_params* p, void* df_a, void* df_b, void* macro, const int16_t* map, int64_t iteration,
				int32_t nsteps, int32_t nthreads)
{
	RefCall c{};
	c.op = 1;
	c.d = d;
	c.p = p;
	c.df_a = df_a;
	c.df_b = df_b;
	c.macro = macro;
	c.map = map;
	c.iteration = iteration;
	c.nsteps = nsteps;
	c.nthreads = nthreads < 1 ? 1 : nthreads;
	return ref_dispatch(c);
}

int oracle_set_equilibrium(const oracle_desc* d, void* df, double rho, double vx, double vy, double vz)
{
	RefCall c{};
	c.op = 2;
	c.d = d;
	c.df_a = df;
	c.crho = rho;
	c.cvx = vx;
	c.cvy = vy;
	c.cvz = vz;
	return ref_dispatch(c);
}

int oracle_set_equilibrium_field(const oracle_desc* d, void* df, const double* rho, const double* vx, const double* vy, const double* vz)
{
	RefCall c{};
	c.op = 2;
	c.d = d;
	c.df_a = df;
	c.rho = rho;
	c.vx = vx;
	c.vy = vy;
	c.vz = vz;
	return ref_dispatch(c);
}

int oracle_initial_macro(const oracle_desc* d, const oracle_params* p, void* df, void* macro)
{
	RefCall c{};
	c.op = 3;
	c.d = d;
	c.p = p;
	c.df_a = df;
	c.macro = macro;
	return ref_dispatch(c);
}
}
