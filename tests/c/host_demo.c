/*
 * tests/c/host_demo.c -- the C ABI driven from plain C, the way the reference's host (algo.c / cuts.c / subprob.c) would:
 * observations and dual vertices arrive one per iteration, tables grow by find-or-append, a cut is formed at x every
 * iteration.  Built twice by tests/test_c_host.py: against libsdgpu.so (-DSD_PREFIX=sdgpu_) and against the CPU oracle
 * (-DSD_PREFIX=sdo_); the two transcripts must agree (indices and iStar exactly, coefficients to 1e-9).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sdgpu.h"

#define SD_CAT2(a, b) a##b
#define SD_CAT(a, b) SD_CAT2(a, b)
#define FN(name) SD_CAT(SD_PREFIX, name)

/* the oracle library exports the same entry points under another prefix */
int  FN(create)(const sdgpu_problem *, const sdgpu_caps *, int, sdgpu_ctx **);
void FN(destroy)(sdgpu_ctx *);
const char *FN(last_error)(void);
int  FN(calc_omega)(sdgpu_ctx *, const double *, double, int *);
int  FN(calc_delta)(sdgpu_ctx *, int, int);
int  FN(update_dual)(sdgpu_ctx *, const double *, double, int, double, int *, int *, int *, int *);
int  FN(basis_find_or_append)(sdgpu_ctx *, int, int, int, int, int, const int32_t *, const int32_t *, int *);
int  FN(sd_cut)(sdgpu_ctx *, const double *, int, int, double, sdgpu_cut *);
int  FN(get_counts)(sdgpu_ctx *, sdgpu_counts *);

static unsigned long long rng_state = 88172645463325252ULL;
static double urand(void) {                       /* xorshift64*, uniform in (-1, 1) */
	rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
	return (double) ((rng_state * 2685821657736338717ULL) >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

int main(int argc, char **argv) {
	enum { ROWS = 12, COLS = 20, N1 = 6, R = 5, Q = 2, K = 60 };
	int iters = argc > 1 ? atoi(argv[1]) : K;
	int32_t CCols[N1 + 1] = {0, 1, 2, 3, 4, 5, 6}, rvRows[R + 1] = {0, 2, 4, 7, 9, 11}, rvbOmRows[R + 1] = {0, 2, 4, 7, 9, 11};
	int32_t rvCOmCols[Q + 1] = {0, 2, 5}, rvCOmRows[Q + 1] = {0, 4, 9}, rvCols[Q + 1] = {0, 2, 5};
	int32_t bcol[ROWS + 1], ccol[3 * ROWS + 1], crow[3 * ROWS + 1];
	double bval[ROWS + 1], cval[3 * ROWS + 1];
	sdgpu_problem p;
	sdgpu_caps caps = { 2 * K + 2, 2 * K + 2, 2 * K + 2, K + 1, 1 };
	sdgpu_ctx *ctx = NULL;
	double observ[R + Q + 1], pi[ROWS + 1], x[N1 + 1], beta[N1 + 1];
	double pool[8][ROWS + 1];
	int32_t istar[K + 1];
	int k, i;

	for (i = 1; i <= ROWS; i++) { bcol[i] = i; bval[i] = urand(); }
	for (i = 1; i <= 3 * ROWS; i++) { crow[i] = (i - 1) % ROWS + 1; ccol[i] = (i * 5) % N1 + 1; cval[i] = urand(); }
	for (k = 0; k < 8; k++) for (i = 0; i <= ROWS; i++) pool[k][i] = i ? urand() : 0.0;
	memset(&p, 0, sizeof p);
	p.num.rows = ROWS; p.num.cols = COLS; p.num.prevCols = N1; p.num.cntCcols = N1; p.num.rvRowCnt = R; p.num.rvbOmCnt = R;
	p.num.rvCOmCnt = Q; p.num.rvdOmCnt = 0; p.num.numRV = R + Q;
	p.coord.CCols = CCols; p.coord.rvRows = rvRows; p.coord.rvbOmRows = rvbOmRows; p.coord.rvCOmCols = rvCOmCols;
	p.coord.rvCOmRows = rvCOmRows; p.coord.rvCols = rvCols; p.coord.rvOffset[1] = R; p.coord.rvOffset[2] = R + Q;
	p.bBar.cnt = ROWS; p.bBar.col = bcol; p.bBar.val = bval;
	p.Cbar.cnt = 3 * ROWS; p.Cbar.col = ccol; p.Cbar.row = crow; p.Cbar.val = cval;
	if (FN(create)(&p, &caps, 0, &ctx) != 0) { fprintf(stderr, "create failed: %s\n", FN(last_error)()); return 2; }

	for (k = 1; k <= iters; k++) {
		int newObs = 0, li, nl, si, ns, newBasis = 0, o, b, st;
		sdgpu_cut cut;
		for (i = 1; i <= R + Q; i++) observ[i] = (k % 7 == 0) ? 0.5 : 3.0 * urand();       /* every 7th observation repeats */
		observ[0] = 0.0;
		{ int pick = (int) ((urand() + 1.0) * 4.0) % 8; for (i = 0; i <= ROWS; i++) pi[i] = pool[pick][i] + (i && k % 3 == 0 ? 2e-4 * urand() : 0.0); }
		for (i = 1; i <= N1; i++) x[i] = 0.5 * (urand() + 1.0);
		x[0] = 0.0;
		o = FN(calc_omega)(ctx, observ, 1e-3, &newObs);                                     /* algo.c:152 */
		if (newObs) FN(calc_delta)(ctx, 1, o);                                              /* stocUpdate.c:25 */
		if (FN(update_dual)(ctx, pi, 0.0, k, 1e-3, &li, &nl, &si, &ns) != 0) { fprintf(stderr, "%s\n", FN(last_error)()); return 3; }
		b = FN(basis_find_or_append)(ctx, ns, o, k, 1, 0, &si, NULL, &newBasis);            /* stocUpdate.c:101-131 */
		cut.beta = beta; cut.iStar = istar;
		st = FN(sd_cut)(ctx, x, k, 1, 0.0, &cut);                                           /* cuts.c:56 */
		if (st != 0) { printf("%d cut NULL (%d)\n", k, st); continue; }
		{
			long long chk = 0;
			for (i = 0; i < cut.omegaCnt; i++) chk = chk * 31 + istar[i] + 1;
			printf("%d o=%d%c l=%d%c s=%d%c b=%d%c N=%d istar#=%lld alpha=%.17g", k, o, newObs ? '*' : ' ', li, nl ? '*' : ' ', si, ns ? '*' : ' ',
					b, newBasis ? '*' : ' ', cut.omegaCnt, chk, cut.alpha);
			for (i = 1; i <= N1; i++) printf(" %.17g", beta[i]);
			printf(" | %.17g %.17g\n", cut.cummOld, cut.cummAll);
		}
	}
	{ sdgpu_counts c; FN(get_counts)(ctx, &c); printf("counts %lld %lld %lld %lld\n", (long long) c.omega, (long long) c.lambda, (long long) c.sigma, (long long) c.basis); }
	FN(destroy)(ctx);
	return 0;
}
