/* abi_minimal.c -- the C ABI of include/lbmx.h from plain C (C99): a periodic D3Q27 cumulant box with a body force, stepped on the
 * GPU, with the mass and momentum budget checked on the host.  No C++ mirror, no Python: this is the call sequence a binding in any
 * language reproduces.
 *
 *   gcc -std=c99 -Iinclude examples/abi_minimal.c -Ltnl_lbm_b200 -llbmx -lm -o abi_minimal && ./abi_minimal [N] [steps]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "lbmx.h"

#define CHECK(call)                                                                   \
	do {                                                                              \
		int rc_ = (call);                                                             \
		if (rc_ != LBMX_OK) {                                                         \
			fprintf(stderr, "%s -> %d: %s\n", #call, rc_, lbmx_last_error());         \
			return 1;                                                                 \
		}                                                                             \
	} while (0)

int main(int argc, char** argv)
{
	const int64_t N = argc > 1 ? atoll(argv[1]) : 32;
	const int64_t steps = argc > 2 ? atoll(argv[2]) : 100;
	const double fx = 1e-6;

	lbmx_desc d = {0};
	d.lattice = LBMX_D3Q27;
	d.coll = LBMX_COLL_CUM;
	d.eq = LBMX_EQ_INV_CUM;
	d.streaming = LBMX_STREAM_AA;
	d.macro = LBMX_MACRO_DEFAULT;
	d.inflow = LBMX_INFLOW_NONE;
	d.precision = LBMX_F64;
	d.macro_policy = LBMX_MACRO_LAST_STEP;
	d.X = d.Y = d.Z = N;
	d.rank = 0;
	d.nranks = 1;
	d.device = -1;

	lbmx_engine* e = NULL;
	CHECK(lbmx_create(&d, &e));

	const size_t cells = (size_t) (N * N * N);
	int16_t* map = (int16_t*) malloc(cells * sizeof(int16_t));
	double* macro = (double*) malloc(4 * cells * sizeof(double));
	if (! map || ! macro)
		return 1;
	for (size_t i = 0; i < cells; i++)
		map[i] = 7; /* D3Q27_BC_All::GEO_PERIODIC */
	CHECK(lbmx_map_upload(e, map, 0));
	CHECK(lbmx_df_set_equilibrium(e, 1.0, 0.0, 0.0, 0.0));

	lbmx_params p = {0};
	p.lbmViscosity = 1e-3;
	p.fx = fx;
	CHECK(lbmx_set_params(e, &p));
	CHECK(lbmx_macro_init(e));
	CHECK(lbmx_step(e, steps));
	CHECK(lbmx_sync(e));
	int32_t nan = 0;
	CHECK(lbmx_has_nan(e, &nan));
	CHECK(lbmx_macro_download(e, macro, 0));

	/* macro layout: [component][x][z][y]; rho and u are the pre-collision values of the last step, u carries the half-force shift */
	double mass = 0, jx = 0;
	for (size_t i = 0; i < cells; i++) {
		mass += macro[i];
		jx += macro[i] * macro[cells + i];
	}
	const double mass_err = fabs(mass / (double) cells - 1.0);
	const double jx_expected = ((double) steps - 0.5) * fx; /* F per cell per step (col_cum.h:341-345 forcing convention) */
	const double jx_err = fabs(jx / (double) cells - jx_expected);
	lbmx_stats st;
	CHECK(lbmx_get_stats(e, &st));
	printf("lbmx %d: %lld^3 cells, %lld steps, %lld kernel launches, nan=%d, mass error %.2e, momentum error %.2e\n", lbmx_version(), (long long) N,
		   (long long) steps, (long long) st.kernel_launches, (int) nan, mass_err, jx_err);
	CHECK(lbmx_destroy(e));
	free(map);
	free(macro);
	return (nan == 0 && mass_err < 1e-12 && jx_err < 1e-12) ? 0 : 2;
}
