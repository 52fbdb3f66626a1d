/*
 * oracle/sd_oracle.c  --  TEST INFRASTRUCTURE ONLY.  CPU restatement ("port") of the reference's SD
 * cut-formation hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs may load this; the product library (libsdgpu.so) never does.
 *
 * Parity status: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
 * section 8c), and its algebra helpers live in the un-vendored, unpinned SMU-SODA/spAlgorithms.  This
 * restatement is therefore pinned the only way available: tests/test_oracle_vs_ref.py drives it side by
 * side with oracle/_ref/libsdref.so -- the reference's own stocUpdate.c / cuts.c / optimal.c compiled
 * from /root/reference against the header shim in oracle/shim/ -- and requires bit-identical tables,
 * iStar and cut coefficients; tests/golden/ holds vectors generated from that reference build.  What
 * stays unpinned is the ten shim helpers (vXv, vXvSparse, ...), whose semantics are inferred from call
 * sites.
 *
 * Same entry points as include/sdgpu.h with the prefix sdo_ instead of sdgpu_, so a test is "same calls,
 * compare outputs".  Layout is dense row-major (not the reference's pointer-per-cell AoS): the
 * per-element arithmetic, its left-to-right order and every tie-break are the reference's, the memory
 * walk is not.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off; never -march=native / -ffast-math).
 *
 * Citations are file:line under /root/reference/twoSD_src.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "sdgpu.h"

#define ABSV(x) ((x) > 0.0 ? (x) : -(x))     /* DBL_ABS as the reference's comparisons use it */

typedef struct {
	sdgpu_num  num;
	sdgpu_caps caps;
	int32_t *CCols, *rvRows, *rvbOmRows, *rvCOmCols, *rvCOmRows, *rvCols;
	int32_t  rvOffset[3];
	int32_t  bBarCnt, *bBarCol;  double *bBarVal;
	int32_t  CbarCnt, *CbarCol, *CbarRow;  double *CbarVal;

	/* omega: stoc.h:33-39 */
	int64_t  omegaCnt;  double *omegaVals;  int32_t *omegaW;          /* [maxOmega][numRV+1] */
	/* lambda: stoc.h:45-48 */
	int64_t  lambdaCnt; double *lambdaVals;                          /* [maxLambda][R+1]     */
	/* sigma: stoc.h:55-60 */
	int64_t  sigmaCnt;  double *sigmaPib, *sigmaPiC; int32_t *sigmaLambda, *sigmaCk;   /* piC [maxSigma][n1c+1] */
	/* delta: stoc.h:68-70; row l allocated when lambda l appears (stocUpdate.c:232) */
	double **deltaPib;  double **deltaPiC;                           /* [l] -> [maxOmega], [l] -> [maxOmega][Q+1] */
	/* basis records: stoc.h:72-97 (fields the argmax reads) */
	int64_t  basisCnt;  int32_t *bCk, *bFeas, *bPhiLen, *bWeight;  int64_t *bTermStart;
	int64_t  termCnt, termCap;  int32_t *tSigma, *tOmega;
	uint8_t **obsFeasible;                                           /* [b] -> [maxOmega] or NULL */
	/* checkBasisFeasibility inputs (randCost.c:202-258): per basis piDet, phi, gBar, psi values, cstat */
	int32_t *rvdOmCols; char *senx;
	double **fPiDet, **fPhi, **fGBar, **fPsi; int32_t **fCstat;
	/* cell->fcutsPool (twoSD.h:129): feasibility cuts kept by addCut2Pool */
	int64_t  fpCnt, fpCap;  double *fpAlpha, *fpBeta;
} oracleCtx;

static char g_err[512];
const char *sdo_last_error(void) { return g_err; }
int sdo_abi_version(void) { return SDGPU_ABI_VERSION; }
static int fail(const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); return SDGPU_ERR; }

static int32_t *copyI(const int32_t *src, int n) {
	int32_t *d = (int32_t *) calloc((size_t) n + 1, sizeof(int32_t));
	if (src) memcpy(d, src, ((size_t) n + 1) * sizeof(int32_t));
	return d;
}
static double *copyD(const double *src, int n) {
	double *d = (double *) calloc((size_t) n + 1, sizeof(double));
	if (src) memcpy(d, src, ((size_t) n + 1) * sizeof(double));
	return d;
}

/* ---- the un-vendored helpers, restated (same inferred semantics as oracle/shim/shim.c) ------------- */
/* vXv: cuts.c:106, stocUpdate.c:175 */
static double dotIdx(const double *a, const double *b, const int32_t *idx, int len) {
	double s = 0.0;
	for (int c = 1; c <= len; c++) s += a[c] * b[idx ? idx[c] : c];
	return s;
}

int sdo_create(const sdgpu_problem *p, const sdgpu_caps *caps, int device, oracleCtx **out) {
	oracleCtx *c = (oracleCtx *) calloc(1, sizeof *c);
	(void) device;
	if (!c) return fail("out of memory");
	c->num = p->num; c->caps = *caps;
	if (c->caps.maxTerms < 1) c->caps.maxTerms = 1;
	c->CCols = copyI(p->coord.CCols, p->num.cntCcols);       c->rvRows = copyI(p->coord.rvRows, p->num.rvRowCnt);
	c->rvbOmRows = copyI(p->coord.rvbOmRows, p->num.rvbOmCnt); c->rvCOmCols = copyI(p->coord.rvCOmCols, p->num.rvCOmCnt);
	c->rvCOmRows = copyI(p->coord.rvCOmRows, p->num.rvCOmCnt); c->rvCols = copyI(p->coord.rvCols, p->num.rvCOmCnt);
	memcpy(c->rvOffset, p->coord.rvOffset, sizeof c->rvOffset);
	c->bBarCnt = p->bBar.cnt; c->bBarCol = copyI(p->bBar.col, p->bBar.cnt); c->bBarVal = copyD(p->bBar.val, p->bBar.cnt);
	c->CbarCnt = p->Cbar.cnt; c->CbarCol = copyI(p->Cbar.col, p->Cbar.cnt); c->CbarRow = copyI(p->Cbar.row, p->Cbar.cnt);
	c->CbarVal = copyD(p->Cbar.val, p->Cbar.cnt);
	c->omegaVals = (double *) calloc((size_t) caps->maxOmega * (p->num.numRV + 1), sizeof(double));
	c->omegaW = (int32_t *) calloc((size_t) caps->maxOmega, sizeof(int32_t));
	c->lambdaVals = (double *) calloc((size_t) caps->maxLambda * (p->num.rvRowCnt + 1), sizeof(double));
	c->sigmaPib = (double *) calloc((size_t) caps->maxSigma, sizeof(double));
	c->sigmaPiC = (double *) calloc((size_t) caps->maxSigma * (p->num.cntCcols + 1), sizeof(double));
	c->sigmaLambda = (int32_t *) calloc((size_t) caps->maxSigma, sizeof(int32_t));
	c->sigmaCk = (int32_t *) calloc((size_t) caps->maxSigma, sizeof(int32_t));
	c->deltaPib = (double **) calloc((size_t) caps->maxLambda, sizeof(double *));
	c->deltaPiC = (double **) calloc((size_t) caps->maxLambda, sizeof(double *));
	c->bCk = (int32_t *) calloc((size_t) caps->maxBasis, sizeof(int32_t));
	c->bFeas = (int32_t *) calloc((size_t) caps->maxBasis, sizeof(int32_t));
	c->bPhiLen = (int32_t *) calloc((size_t) caps->maxBasis, sizeof(int32_t));
	c->bWeight = (int32_t *) calloc((size_t) caps->maxBasis, sizeof(int32_t));
	c->bTermStart = (int64_t *) calloc((size_t) caps->maxBasis + 1, sizeof(int64_t));
	c->termCap = caps->maxBasis * (int64_t) c->caps.maxTerms;
	c->tSigma = (int32_t *) calloc((size_t) c->termCap, sizeof(int32_t));
	c->tOmega = (int32_t *) calloc((size_t) c->termCap, sizeof(int32_t));
	c->obsFeasible = (uint8_t **) calloc((size_t) caps->maxBasis, sizeof(uint8_t *));
	c->fPiDet = (double **) calloc((size_t) caps->maxBasis, sizeof(double *)); c->fPhi = (double **) calloc((size_t) caps->maxBasis, sizeof(double *));
	c->fGBar = (double **) calloc((size_t) caps->maxBasis, sizeof(double *)); c->fPsi = (double **) calloc((size_t) caps->maxBasis, sizeof(double *));
	c->fCstat = (int32_t **) calloc((size_t) caps->maxBasis, sizeof(int32_t *));
	*out = c;
	return 0;
}

/* free*Type(..., partial = true): setup.c:242-246, stocUpdate.c:472-545 */
int sdo_reset(oracleCtx *c) {
	for (int64_t l = 0; l < c->lambdaCnt; l++) {
		free(c->deltaPib[l]); c->deltaPib[l] = NULL;
		free(c->deltaPiC[l]); c->deltaPiC[l] = NULL;
	}
	for (int64_t b = 0; b < c->basisCnt; b++) {
		free(c->obsFeasible[b]); c->obsFeasible[b] = NULL;
		free(c->fPiDet[b]); free(c->fPhi[b]); free(c->fGBar[b]); free(c->fPsi[b]); free(c->fCstat[b]);
		c->fPiDet[b] = c->fPhi[b] = c->fGBar[b] = c->fPsi[b] = NULL; c->fCstat[b] = NULL;
	}
	c->omegaCnt = c->lambdaCnt = c->sigmaCnt = c->basisCnt = c->termCnt = 0;
	c->fpCnt = 0;                                          /* freeCutsType(cell->fcutsPool, true) setup.c:236 */
	return 0;
}

void sdo_destroy(oracleCtx *c) {
	if (!c) return;
	sdo_reset(c);
	free(c->CCols); free(c->rvRows); free(c->rvbOmRows); free(c->rvCOmCols); free(c->rvCOmRows); free(c->rvCols);
	free(c->bBarCol); free(c->bBarVal); free(c->CbarCol); free(c->CbarRow); free(c->CbarVal);
	free(c->omegaVals); free(c->omegaW); free(c->lambdaVals); free(c->sigmaPib); free(c->sigmaPiC);
	free(c->sigmaLambda); free(c->sigmaCk); free(c->deltaPib); free(c->deltaPiC);
	free(c->bCk); free(c->bFeas); free(c->bPhiLen); free(c->bWeight); free(c->bTermStart);
	free(c->tSigma); free(c->tOmega); free(c->obsFeasible); free(c->fpAlpha); free(c->fpBeta);
	free(c);
}

int sdo_get_counts(oracleCtx *c, sdgpu_counts *out) {
	out->omega = c->omegaCnt; out->lambda = c->lambdaCnt; out->sigma = c->sigmaCnt; out->basis = c->basisCnt;
	return 0;
}

/* ---- omega: calcOmega stocUpdate.c:326-348 ---------------------------------------------------------- */
static double *omegaRow(oracleCtx *c, int64_t o) { return c->omegaVals + (size_t) o * (c->num.numRV + 1); }

int sdo_omega_find(oracleCtx *c, const double *observ, double tol) {
	/* equalVector over positions 1..numRV, first match wins (stocUpdate.c:330-335) */
	for (int64_t o = 0; o < c->omegaCnt; o++) {
		const double *v = omegaRow(c, o);
		int same = 1;
		for (int j = 1; j <= c->num.numRV; j++)
			if (ABSV(observ[j] - v[j]) > tol) { same = 0; break; }
		if (same) return (int) o;
	}
	return SDGPU_NONE;
}

int sdo_omega_append(oracleCtx *c, const double *observ, int weight) {
	if (c->omegaCnt >= c->caps.maxOmega) return fail("omega capacity exceeded");
	memcpy(omegaRow(c, c->omegaCnt), observ, ((size_t) c->num.numRV + 1) * sizeof(double));   /* duplicVector :338 */
	c->omegaW[c->omegaCnt] = weight;
	return (int) c->omegaCnt++;
}

int sdo_omega_bump(oracleCtx *c, int idx, int by) {
	if (idx < 0 || idx >= c->omegaCnt) return fail("omega index out of range");
	c->omegaW[idx] += by;
	return 0;
}

int sdo_omega_append_bulk(oracleCtx *c, int64_t n, const double *vals, const int32_t *weights) {
	for (int64_t i = 0; i < n; i++) {
		int r = sdo_omega_append(c, vals + (size_t) i * (c->num.numRV + 1), weights ? weights[i] : 1);
		if (r < 0) return r;
	}
	return 0;
}

int sdo_calc_omega(oracleCtx *c, const double *observ, double tol, int *newOmegaFlag) {
	int idx = sdo_omega_find(c, observ, tol);
	if (idx >= 0) {
		c->omegaW[idx]++;                                /* :333 */
		if (newOmegaFlag) *newOmegaFlag = 0;
		return idx;
	}
	if (newOmegaFlag) *newOmegaFlag = 1;
	return sdo_omega_append(c, observ, 1);               /* :338-340,347 */
}

/* ---- lambda: calcLambda stocUpdate.c:264-284 --------------------------------------------------------- */
static double *lambdaRow(oracleCtx *c, int64_t l) { return c->lambdaVals + (size_t) l * (c->num.rvRowCnt + 1); }

int sdo_calc_lambda(oracleCtx *c, const double *Pi, double tol, int *newLambdaFlag) {
	int R = c->num.rvRowCnt;
	double *cand = (double *) calloc((size_t) R + 1, sizeof(double));
	for (int i = 1; i <= R; i++) cand[i] = Pi[c->rvRows[i]];         /* reduceVector :269 */
	for (int64_t l = 0; l < c->lambdaCnt; l++) {                      /* :272-277 */
		const double *v = lambdaRow(c, l);
		int same = 1;
		for (int i = 1; i <= R; i++)
			if (ABSV(cand[i] - v[i]) > tol) { same = 0; break; }
		if (same) { free(cand); if (newLambdaFlag) *newLambdaFlag = 0; return (int) l; }
	}
	if (c->lambdaCnt >= c->caps.maxLambda) { free(cand); return fail("lambda capacity exceeded"); }
	memcpy(lambdaRow(c, c->lambdaCnt), cand, ((size_t) R + 1) * sizeof(double));   /* :280 */
	free(cand);
	if (newLambdaFlag) *newLambdaFlag = 1;
	return (int) c->lambdaCnt++;
}

/* ---- sigma: calcSigma stocUpdate.c:286-320 ----------------------------------------------------------- */
static double *sigmaPiCRow(oracleCtx *c, int64_t s) { return c->sigmaPiC + (size_t) s * (c->num.cntCcols + 1); }

int sdo_calc_sigma(oracleCtx *c, const double *pi, double mubBar, int idxLambda, int newLambdaFlag, int currentIter,
		double tol, int *newSigmaFlag) {
	int n1c = c->num.cntCcols;
	double pibBar = 0.0;
	for (int e = 1; e <= c->bBarCnt; e++) pibBar += c->bBarVal[e] * pi[c->bBarCol[e]];   /* vXvSparse :293 */
	pibBar = pibBar + mubBar;
	double *full = (double *) calloc((size_t) c->num.prevCols + 1, sizeof(double));       /* vxMSparse :295 */
	for (int e = 1; e <= c->CbarCnt; e++) full[c->CbarCol[e]] += pi[c->CbarRow[e]] * c->CbarVal[e];
	double *piCBar = (double *) calloc((size_t) n1c + 1, sizeof(double));
	for (int k = 1; k <= n1c; k++) piCBar[k] = full[c->CCols[k]];                         /* reduceVector :296 */
	free(full);

	if (!newLambdaFlag) {                                                                  /* :299-310 */
		for (int64_t s = 0; s < c->sigmaCnt; s++) {
			if (ABSV(pibBar - c->sigmaPib[s]) <= tol) {
				const double *v = sigmaPiCRow(c, s);
				int same = 1;
				for (int k = 1; k <= n1c; k++)
					if (ABSV(piCBar[k] - v[k]) > tol) { same = 0; break; }
				if (same && c->sigmaLambda[s] == idxLambda) {
					free(piCBar);
					if (newSigmaFlag) *newSigmaFlag = 0;
					return (int) s;
				}
			}
		}
	}
	if (c->sigmaCnt >= c->caps.maxSigma) { free(piCBar); return fail("sigma capacity exceeded"); }
	if (newSigmaFlag) *newSigmaFlag = 1;                                                   /* :312-318 */
	c->sigmaPib[c->sigmaCnt] = pibBar;
	memcpy(sigmaPiCRow(c, c->sigmaCnt), piCBar, ((size_t) n1c + 1) * sizeof(double));
	c->sigmaLambda[c->sigmaCnt] = idxLambda;
	c->sigmaCk[c->sigmaCnt] = currentIter;
	free(piCBar);
	return (int) c->sigmaCnt++;
}

/* ---- delta: calcDelta stocUpdate.c:196-257 ------------------------------------------------------------ */
static void deltaCell(oracleCtx *c, const double *lamFull, int64_t o, double *pib, double *piC, double *scratch) {
	const double *w = omegaRow(c, o);
	int Rb = c->num.rvbOmCnt, Q = c->num.rvCOmCnt;
	double s = 0.0;
	for (int j = 1; j <= Rb; j++) s += w[j] * lamFull[c->rvbOmRows[j]];                   /* vXvSparse :218,244 */
	*pib = s;
	if (Q > 0) {
		memset(scratch, 0, ((size_t) c->num.prevCols + 1) * sizeof(double));               /* vxMSparse :220,246 */
		for (int e = 1; e <= Q; e++) scratch[c->rvCOmCols[e]] += lamFull[c->rvCOmRows[e]] * w[Rb + e];
		for (int k = 1; k <= Q; k++) piC[k] = scratch[c->rvCOmCols[k]];                    /* reduceVector :221,247 */
	}
}

static double *expandLambda(oracleCtx *c, int64_t l) {                                     /* expandVector :214,236 */
	double *full = (double *) calloc((size_t) c->num.rows + 1, sizeof(double));
	const double *v = lambdaRow(c, l);
	for (int i = 1; i <= c->num.rvRowCnt; i++) full[c->rvRows[i]] = v[i];
	return full;
}

int sdo_calc_delta(oracleCtx *c, int newOmegaFlag, int elemIdx) {
	int Q = c->num.rvCOmCnt;
	double *scratch = (double *) calloc((size_t) c->num.prevCols + 1, sizeof(double));
	if (newOmegaFlag) {                                                                    /* case I :206-229 */
		if (elemIdx < 0 || elemIdx >= c->omegaCnt) { free(scratch); return fail("calc_delta: observation out of range"); }
		for (int64_t l = 0; l < c->lambdaCnt; l++) {
			double *full = expandLambda(c, l);
			deltaCell(c, full, elemIdx, &c->deltaPib[l][elemIdx], Q ? c->deltaPiC[l] + (size_t) elemIdx * (Q + 1) : NULL, scratch);
			free(full);
		}
	}
	else {                                                                                 /* case II :230-254 */
		if (elemIdx < 0 || elemIdx >= c->lambdaCnt) { free(scratch); return fail("calc_delta: lambda out of range"); }
		free(c->deltaPib[elemIdx]); free(c->deltaPiC[elemIdx]);
		c->deltaPib[elemIdx] = (double *) calloc((size_t) c->caps.maxOmega, sizeof(double));
		c->deltaPiC[elemIdx] = Q ? (double *) calloc((size_t) c->caps.maxOmega * (Q + 1), sizeof(double)) : NULL;
		double *full = expandLambda(c, elemIdx);
		for (int64_t o = 0; o < c->omegaCnt; o++)
			deltaCell(c, full, o, &c->deltaPib[elemIdx][o], Q ? c->deltaPiC[elemIdx] + (size_t) o * (Q + 1) : NULL, scratch);
		free(full);
	}
	free(scratch);
	return 0;
}

int sdo_calc_delta_block(oracleCtx *c, int64_t l0, int64_t l1, int64_t o0, int64_t o1) {
	int Q = c->num.rvCOmCnt;
	if (l0 < 0 || l1 > c->lambdaCnt || o0 < 0 || o1 > c->omegaCnt || l0 > l1 || o0 > o1) return fail("calc_delta_block: block out of range");
	/* rows are independent (stocUpdate.c:230-254): spread them over threads, arithmetic per cell unchanged */
#pragma omp parallel for schedule(dynamic, 4)
	for (int64_t l = l0; l < l1; l++) {
		double *scratch = (double *) calloc((size_t) c->num.prevCols + 1, sizeof(double));
		if (!c->deltaPib[l]) {
			c->deltaPib[l] = (double *) calloc((size_t) c->caps.maxOmega, sizeof(double));
			c->deltaPiC[l] = Q ? (double *) calloc((size_t) c->caps.maxOmega * (Q + 1), sizeof(double)) : NULL;
		}
		double *full = expandLambda(c, l);
		for (int64_t o = o0; o < o1; o++)
			deltaCell(c, full, o, &c->deltaPib[l][o], Q ? c->deltaPiC[l] + (size_t) o * (Q + 1) : NULL, scratch);
		free(full);
		free(scratch);
	}
	return 0;
}

/* stocUpdate.c:78-85 (and :90-97 for a phi column) */
int sdo_update_dual(oracleCtx *c, const double *pi, double mubBar, int currentIter, double tol,
		int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag) {
	int nl = 0, ns = 0;
	int li = sdo_calc_lambda(c, pi, tol, &nl);
	if (li < 0) return li;
	int si = sdo_calc_sigma(c, pi, mubBar, li, nl, currentIter, tol, &ns);
	if (si < 0) return si;
	if (nl) { int r = sdo_calc_delta(c, 0, li); if (r < 0) return r; }
	if (lambdaIdx) *lambdaIdx = li;
	if (newLambdaFlag) *newLambdaFlag = nl;
	if (sigmaIdx) *sigmaIdx = si;
	if (newSigmaFlag) *newSigmaFlag = ns;
	return 0;
}

/* stocUpdate.c:24-25 followed by :78-85 (the entry point the CUDA library serves with one launch) */
int sdo_update_dual_col(oracleCtx *c, int newOmegaIdx, const double *pi, double mubBar, int currentIter, double tol,
		int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag) {
	if (newOmegaIdx >= 0) { int r = sdo_calc_delta(c, 1, newOmegaIdx); if (r != 0) return r; }
	return sdo_update_dual(c, pi, mubBar, currentIter, tol, lambdaIdx, newLambdaFlag, sigmaIdx, newSigmaFlag);
}

int sdo_update_dual_bulk(oracleCtx *c, int64_t n, const double *pis, const double *mubBar, const int32_t *iters,
		double tol, int32_t *lambdaIdx, int32_t *sigmaIdx) {
	for (int64_t i = 0; i < n; i++) {
		int li, si, r = 0;
		const double *pi = pis + (size_t) i * (c->num.rows + 1);
		if (tol < 0.0) {        /* synthetic loader: append without the dedup scans and without the delta row */
			int nl = 1, ns = 1;
			if (c->lambdaCnt >= c->caps.maxLambda) return fail("lambda capacity exceeded");
			double *row = lambdaRow(c, c->lambdaCnt);
			for (int k = 1; k <= c->num.rvRowCnt; k++) row[k] = pi[c->rvRows[k]];
			li = (int) c->lambdaCnt++;
			si = sdo_calc_sigma(c, pi, mubBar ? mubBar[i] : 0.0, li, nl, iters ? iters[i] : (int) i + 1, 0.0, &ns);
			if (si < 0) return si;
		}
		else
			r = sdo_update_dual(c, pi, mubBar ? mubBar[i] : 0.0, iters ? iters[i] : (int) i + 1, tol, &li, NULL, &si, NULL);
		if (r < 0) return r;
		if (lambdaIdx) lambdaIdx[i] = li;
		if (sigmaIdx) sigmaIdx[i] = si;
	}
	return 0;
}

/* ---- basis records: stocUpdate.c:101-131 --------------------------------------------------------------- */
int sdo_basis_append(oracleCtx *c, int ck, int feasFlag, int phiLength, const int32_t *sigmaIdx, const int32_t *omegaIdx) {
	if (c->basisCnt >= c->caps.maxBasis) return fail("basis capacity exceeded");
	if (phiLength + 1 > c->caps.maxTerms || c->termCnt + phiLength + 1 > c->termCap) return fail("basis term capacity exceeded");
	int64_t b = c->basisCnt;
	c->bCk[b] = ck; c->bFeas[b] = feasFlag != 0; c->bPhiLen[b] = phiLength; c->bWeight[b] = 1;
	c->bTermStart[b] = c->termCnt;
	for (int t = 0; t <= phiLength; t++) {
		c->tSigma[c->termCnt] = sigmaIdx[t];
		c->tOmega[c->termCnt] = (t > 0 && omegaIdx) ? omegaIdx[t] : 0;
		c->termCnt++;
	}
	c->bTermStart[b + 1] = c->termCnt;
	if (feasFlag) {                                            /* :119-127, checkBasisFeasibility -> true by default */
		c->obsFeasible[b] = (uint8_t *) malloc((size_t) c->caps.maxOmega);
		memset(c->obsFeasible[b], 1, (size_t) c->caps.maxOmega);
	}
	else
		c->obsFeasible[b] = NULL;                              /* :129 */
	return (int) c->basisCnt++;
}

int sdo_basis_append_bulk(oracleCtx *c, int64_t n, const int32_t *ck, const int32_t *feas, const int32_t *sigmaIdx) {
	int first = (int) c->basisCnt;
	for (int64_t i = 0; i < n; i++) {
		int r = sdo_basis_append(c, ck[i], feas ? feas[i] : 1, 0, sigmaIdx + i, NULL);
		if (r < 0) return r;
	}
	return first;
}

int sdo_basis_find_or_append(oracleCtx *c, int retainBasis, int obsIdx, int ck, int feasFlag, int phiLength,
		const int32_t *sigmaIdx, const int32_t *omegaIdx, int *newBasisFlag) {
	if (newBasisFlag) *newBasisFlag = 1;
	if (!retainBasis) {                                        /* :101-113 */
		for (int64_t b = 0; b < c->basisCnt; b++) {
			if (phiLength == c->bPhiLen[b] && c->obsFeasible[b] && c->obsFeasible[b][obsIdx]) {
				int same = 1;
				for (int t = 0; t <= phiLength; t++)
					if (c->tSigma[c->bTermStart[b] + t] != sigmaIdx[t]) { same = 0; break; }
				if (same) {
					c->bWeight[b]++;
					if (newBasisFlag) *newBasisFlag = 0;
					return (int) b;
				}
			}
		}
	}
	return sdo_basis_append(c, ck, feasFlag, phiLength, sigmaIdx, omegaIdx);
}

int sdo_basis_set_obs_feasible(oracleCtx *c, int basisIdx, int obsIdx, int flag) {
	if (basisIdx < 0 || basisIdx >= c->basisCnt || !c->obsFeasible[basisIdx]) return fail("set_obs_feasible: bad basis");
	if (obsIdx < 0 || obsIdx >= c->caps.maxOmega) return fail("set_obs_feasible: bad observation");
	c->obsFeasible[basisIdx][obsIdx] = flag != 0;
	return 0;
}
int sdo_basis_set_obs_feasible_row(oracleCtx *c, int basisIdx, const uint8_t *flags) {
	if (basisIdx < 0 || basisIdx >= c->basisCnt || !c->obsFeasible[basisIdx]) return fail("set_obs_feasible_row: bad basis");
	for (int64_t o = 0; o < c->omegaCnt; o++) c->obsFeasible[basisIdx][o] = flags[o] != 0;
	return 0;
}
int sdo_basis_set_obs_feasible_col(oracleCtx *c, int obsIdx, const uint8_t *flags) {
	if (obsIdx < 0 || obsIdx >= c->caps.maxOmega) return fail("set_obs_feasible_col: bad observation");
	for (int64_t b = 0; b < c->basisCnt; b++) if (c->obsFeasible[b]) c->obsFeasible[b][obsIdx] = flags[b] != 0;
	return 0;
}

/* ---- checkBasisFeasibility randCost.c:202-258 ---------------------------------------------------------------- */
int sdo_set_cost_coords(oracleCtx *c, const int32_t *rvdOmCols, const char *senx) {
	free(c->rvdOmCols); free(c->senx);
	c->rvdOmCols = copyI(rvdOmCols, c->num.rvdOmCnt);
	c->senx = (char *) malloc((size_t) c->num.rows);
	memcpy(c->senx, senx, (size_t) c->num.rows);
	return 0;
}

int sdo_basis_set_feas_data(oracleCtx *c, int b, const double *piDet, const double *phi, const double *gBar, const double *psiVal,
		const int32_t *cstat) {
	if (b < 0 || b >= c->basisCnt) return fail("basis_set_feas_data: bad basis");
	int rows = c->num.rows, cols = c->num.cols, pl = c->bPhiLen[b];
	free(c->fPiDet[b]); free(c->fPhi[b]); free(c->fGBar[b]); free(c->fPsi[b]); free(c->fCstat[b]);
	c->fPiDet[b] = copyD(piDet, rows); c->fGBar[b] = copyD(gBar, cols); c->fCstat[b] = copyI(cstat, cols);
	c->fPhi[b] = (double *) malloc(((size_t) pl * (rows + 1) + 1) * sizeof(double));
	c->fPsi[b] = (double *) malloc(((size_t) pl * cols + 1) * sizeof(double));
	if (pl) { memcpy(c->fPhi[b], phi, (size_t) pl * (rows + 1) * sizeof(double)); memcpy(c->fPsi[b], psiVal, (size_t) pl * cols * sizeof(double)); }
	return 0;
}

static int pairFeasible(oracleCtx *c, int b, int64_t obs, double tol) {
	int rows = c->num.rows, cols = c->num.cols, rvd = c->num.rvdOmCnt, pl = c->bPhiLen[b];
	const double *val = omegaRow(c, obs) + c->rvOffset[2];               /* dOmega.val, 1-based (stocUpdate.c:28) */
	if (rvd == 0) return 1;                                              /* randCost.c:208 */
	if (pl > 0) {                                                        /* :213-224 */
		for (int r = 1; r <= rows; r++) {
			double theta = 0.0;
			for (int n = 0; n < pl; n++) theta += c->fPhi[b][(size_t) n * (rows + 1) + r] * val[c->tOmega[c->bTermStart[b] + 1 + n]];
			double v = c->fPiDet[b][r] + theta;
			if ((v < -tol && c->senx[r - 1] == 'G') || (v > tol && c->senx[r - 1] == 'L')) return 0;
		}
	}
	double *rc = copyD(c->fGBar[b], cols);                               /* copyVector :239 */
	for (int j = 1; j <= rvd; j++) rc[c->rvdOmCols[j]] += val[j];        /* addVectors :240 */
	for (int i = 1; i <= cols; i++)                                      /* MSparsexvSub :242, entry order (i, j) */
		for (int n = 0; n < pl; n++) rc[i] -= c->fPsi[b][(size_t) (i - 1) * pl + n] * val[c->tOmega[c->bTermStart[b] + 1 + n]];
	int ok = 1;
	for (int i = 1; i <= cols; i++)                                      /* :246-252, AT_UPPER == 2 */
		if (rc[i] < -tol && c->fCstat[b][i] != 2) { ok = 0; break; }
	free(rc);
	return ok;
}

int sdo_check_feasibility_obs(oracleCtx *c, int obsIdx, double tol, uint8_t *flagsOut) {
	if (obsIdx < 0 || obsIdx >= c->omegaCnt) return fail("check_feasibility_obs: bad observation");
	for (int64_t b = 0; b < c->basisCnt; b++) {
		if (c->obsFeasible[b] && c->fPiDet[b]) c->obsFeasible[b][obsIdx] = (uint8_t) pairFeasible(c, (int) b, obsIdx, tol);   /* stocUpdate.c:29-30 */
		if (flagsOut) flagsOut[b] = c->obsFeasible[b] ? c->obsFeasible[b][obsIdx] : 0;
	}
	return 0;
}

int sdo_check_feasibility_basis(oracleCtx *c, int b, double tol, uint8_t *flagsOut) {
	if (b < 0 || b >= c->basisCnt || !c->obsFeasible[b] || !c->fPiDet[b]) return fail("check_feasibility_basis: bad basis");
	for (int64_t o = 0; o < c->omegaCnt; o++) {                           /* stocUpdate.c:123-126 */
		c->obsFeasible[b][o] = (uint8_t) pairFeasible(c, b, o, tol);
		if (flagsOut) flagsOut[o] = c->obsFeasible[b][o];
	}
	return 0;
}

/* ---- raw feasibility cuts: the body of updtFeasCutPool cuts.c:473-490 ------------------------------------------------ */
int sdo_feas_cuts(oracleCtx *c, int obsFirst, int obsLast, int basisFirst, int basisLast, int maxOut, double *alpha, double *beta) {
	int n = 0, n1 = c->num.prevCols, n1c = c->num.cntCcols, Q = c->num.rvCOmCnt;
	if (obsFirst < 0 || obsLast > c->omegaCnt || basisFirst < 0 || basisLast > c->basisCnt) return fail("feas_cuts: range out of bounds");
	for (int o = obsFirst; o < obsLast; o++)
		for (int b = basisFirst; b < basisLast; b++) {
			if (c->bFeas[b]) continue;
			if (n >= maxOut) return fail("feas_cuts: output buffer too small");
			int s = c->tSigma[c->bTermStart[b]], l = c->sigmaLambda[s];
			double *bt = beta + (size_t) n * (n1 + 1);
			for (int i = 0; i <= n1; i++) bt[i] = 0.0;
			alpha[n] = c->sigmaPib[s] + c->deltaPib[l][o];                                   /* cuts.c:481 */
			for (int k = 1; k <= n1c; k++) bt[c->CCols[k]] += sigmaPiCRow(c, s)[k];           /* :483-484 */
			for (int k = 1; k <= Q; k++) bt[c->rvCols[k]] += c->deltaPiC[l][(size_t) o * (Q + 1) + k];   /* :485-486 */
			n++;
		}
	return n;
}

/* ---- the feasibility-cut pool: updtFeasCutPool cuts.c:465-517, addCut2Pool(FEASIBILITY) cuts.c:643-655, checkFeasCutPool :521-567 */
static int sameCut(double aA, const double *aB, double bA, const double *bB, int n1, double tol) {
	if (!(ABSV(aA - bA) < tol)) return 0;                                                    /* cuts.c:645 */
	for (int c = 1; c <= n1; c++) if (ABSV(aB[c] - bB[c]) > tol) return 0;                   /* equalVector :646 */
	return 1;
}

static int poolAdd(oracleCtx *c, double alpha, const double *beta, double tol) {             /* cuts.c:643-655 */
	int n1 = c->num.prevCols;
	for (int64_t i = 0; i < c->fpCnt; i++)
		if (sameCut(alpha, beta, c->fpAlpha[i], c->fpBeta + (size_t) i * (n1 + 1), n1, tol)) return 0;
	if (c->fpCnt == c->fpCap) {
		c->fpCap = c->fpCap ? 2 * c->fpCap : 256;
		c->fpAlpha = (double *) realloc(c->fpAlpha, (size_t) c->fpCap * sizeof(double));
		c->fpBeta = (double *) realloc(c->fpBeta, (size_t) c->fpCap * (n1 + 1) * sizeof(double));
	}
	c->fpAlpha[c->fpCnt] = alpha;
	memcpy(c->fpBeta + (size_t) c->fpCnt * (n1 + 1), beta, ((size_t) n1 + 1) * sizeof(double));
	c->fpCnt++;
	return 1;
}

int sdo_feas_pool_update(oracleCtx *c, int *fUpdt, double tol) {
	int n1 = c->num.prevCols;
	double a, *b = (double *) calloc((size_t) n1 + 1, sizeof(double));
	for (int o = fUpdt[1]; o < c->omegaCnt; o++)                                             /* cuts.c:472-490 */
		for (int i = 0; i < fUpdt[0]; i++)
			if (!c->bFeas[i] && sdo_feas_cuts(c, o, o + 1, i, i + 1, 1, &a, b) == 1) poolAdd(c, a, b, tol);
	fUpdt[1] = (int) c->omegaCnt;
	for (int o = 0; o < c->omegaCnt; o++)                                                    /* cuts.c:494-512 */
		for (int i = fUpdt[0]; i < c->basisCnt; i++)
			if (!c->bFeas[i] && sdo_feas_cuts(c, o, o + 1, i, i + 1, 1, &a, b) == 1) poolAdd(c, a, b, tol);
	fUpdt[0] = (int) c->basisCnt;
	free(b);
	return (int) c->fpCnt;
}

int sdo_feas_pool_size(oracleCtx *c) { return (int) c->fpCnt; }

int sdo_feas_pool_get(oracleCtx *c, int first, int count, double *alpha, double *beta) {
	int n1 = c->num.prevCols;
	if (first < 0 || count < 0 || first + count > c->fpCnt) return fail("feas_pool_get: range out of bounds");
	memcpy(alpha, c->fpAlpha + first, (size_t) count * sizeof(double));
	memcpy(beta, c->fpBeta + (size_t) first * (n1 + 1), (size_t) count * (n1 + 1) * sizeof(double));
	return count;
}

int sdo_feas_pool_check(oracleCtx *c, int nFcuts, const double *fAlpha, const double *fBeta, const double *incumbX,
		const double *candidX, double tol, int32_t *action, int *infeasIncumb) {
	int n1 = c->num.prevCols;
	if (infeasIncumb) *infeasIncumb = 0;
	for (int64_t idx = 0; idx < c->fpCnt; idx++) {                                           /* cuts.c:526-560 */
		const double alpha = c->fpAlpha[idx], *beta = c->fpBeta + (size_t) idx * (n1 + 1);
		int dup = 0;
		for (int f = 0; f < nFcuts && !dup; f++) dup = sameCut(alpha, beta, fAlpha[f], fBeta + (size_t) f * (n1 + 1), n1, tol);
		action[idx] = 0;
		if (dotIdx(beta, incumbX, NULL, n1) < alpha) {
			if (infeasIncumb) *infeasIncumb = 1;
			action[idx] = dup ? 2 : 1;
		}
		else if (!dup && dotIdx(beta, candidX, NULL, n1) < alpha) action[idx] = 3;
	}
	return (int) c->fpCnt;
}

/* ---- argmax: computeIstar stocUpdate.c:142-190 ---------------------------------------------------------- */
static void piCbarXAll(oracleCtx *c, const double *X, double *out) {                        /* cuts.c:105-106 */
	for (int64_t s = 0; s < c->sigmaCnt; s++) out[s] = dotIdx(sigmaPiCRow(c, s), X, c->CCols, c->num.cntCcols);
}

static int istarOne(oracleCtx *c, const double *piCbarX, const double *X, int64_t obs, int numSamples, int pi_eval,
		int isNew, double *argmax) {
	int Q = c->num.rvCOmCnt;
	const double *w = omegaRow(c, obs);
	int up, low, best = 0;
	if (pi_eval) numSamples -= (int) (0.1 * numSamples + 1);                                 /* :147-148 */
	if (!isNew) { up = numSamples; low = -INT_MAX; } else { up = INT_MAX; low = numSamples; } /* :151-156 */
	*argmax = -DBL_MAX;                                                                      /* :158 */
	for (int64_t b = 0; b < c->basisCnt; b++) {                                              /* :161-184 */
		if (!(c->bFeas[b] && c->bCk[b] > low && c->bCk[b] <= up)) continue;
		if (!c->obsFeasible[b][obs]) continue;
		double arg = 0.0;
		for (int t = 0; t <= c->bPhiLen[b]; t++) {
			int s = c->tSigma[c->bTermStart[b] + t];
			int l = c->sigmaLambda[s];
			double m = (t == 0) ? 1.0 : w[c->rvOffset[2] + c->tOmega[c->bTermStart[b] + t]];
			arg += m * (c->sigmaPib[s] + c->deltaPib[l][obs] - piCbarX[s]);                  /* :174 */
			double dx = Q ? dotIdx(c->deltaPiC[l] + (size_t) obs * (Q + 1), X, c->rvCOmCols, Q) : 0.0;
			arg -= m * dx;                                                                   /* :175 */
		}
		if (arg > *argmax) { *argmax = arg; best = (int) b; }                                /* :178-181 */
	}
	return (*argmax == -DBL_MAX) ? SDGPU_NONE : best;                                        /* :186-189 */
}

int sdo_compute_istar(oracleCtx *c, const double *X, int obs, int numSamples, int pi_eval, int isNew, double *argmax) {
	if (obs < 0 || obs >= c->omegaCnt) return fail("compute_istar: observation out of range");
	double *pcx = (double *) calloc((size_t) c->sigmaCnt + 1, sizeof(double));
	piCbarXAll(c, X, pcx);
	int r = istarOne(c, pcx, X, obs, numSamples, pi_eval, isNew, argmax);
	free(pcx);
	return r;
}

/* ---- the cut: SDCut cuts.c:91-194 ------------------------------------------------------------------------ */
/* one observation's istar + window sums (cuts.c:118-134); returns istar */
static int cutObsIstar(oracleCtx *c, const double *pcx, const double *X, int64_t obs, int numSamples, int pi_eval, double lb,
		double *cummOld, double *cummAll) {
	double aOld, aNew, a;
	int istar;
	if (pi_eval) {
		int iOld = istarOne(c, pcx, X, obs, numSamples, 1, 0, &aOld);
		int iNew = istarOne(c, pcx, X, obs, numSamples, 1, 1, &aNew);
		a = fmax(aOld, aNew);                                                               /* :124 */
		istar = (aNew > aOld) ? iNew : iOld;                                                /* :125 */
		*cummOld += fmax(aOld - lb, 0) * c->omegaW[obs];                                    /* :127 */
		*cummAll += fmax(a - lb, 0) * c->omegaW[obs];                                       /* :128 */
	}
	else
		istar = istarOne(c, pcx, X, obs, numSamples, 0, 0, &a);                             /* :132 */
	return istar;
}

/* one observation's contribution to alpha / beta (cuts.c:142-168); beta is the full [prevCols+1] accumulator */
static int cutObsAccumulate(oracleCtx *c, int64_t obs, int istar, double *alpha, double *beta) {
	int n1c = c->num.cntCcols, Q = c->num.rvCOmCnt;
	int wgt = c->omegaW[obs];
	if (c->num.rvdOmCnt > 0) {                                                              /* :142-159 */
		const double *w = omegaRow(c, obs);
		for (int t = 0; t <= c->bPhiLen[istar]; t++) {
			int s = c->tSigma[c->bTermStart[istar] + t];
			int l = c->sigmaLambda[s];
			double m = (t == 0) ? 1.0 : w[c->rvOffset[2] + c->tOmega[c->bTermStart[istar] + t]];
			*alpha += wgt * m * (c->sigmaPib[s] + c->deltaPib[l][obs]);
			const double *pc = sigmaPiCRow(c, s);
			for (int k = 1; k <= n1c; k++) beta[c->CCols[k]] += wgt * m * pc[k];
			for (int k = 1; k <= Q; k++) beta[c->rvCOmCols[k]] += wgt * m * c->deltaPiC[l][(size_t) obs * (Q + 1) + k];
		}
	}
	else {                                                                                  /* :160-168: the BASIS index is used as the sigma index */
		if (istar >= c->sigmaCnt) return fail("sd_cut: iStar used as a sigma index is out of range (cuts.c:161)");
		int l = c->sigmaLambda[istar];
		*alpha += c->sigmaPib[istar] * wgt;
		*alpha += c->deltaPib[l][obs] * wgt;
		const double *pc = sigmaPiCRow(c, istar);
		for (int k = 1; k <= n1c; k++) beta[c->CCols[k]] += pc[k] * wgt;
		for (int k = 1; k <= Q; k++) beta[c->rvCols[k]] += c->deltaPiC[l][(size_t) obs * (Q + 1) + k] * wgt;
	}
	return 0;
}

/* un-normalised sums over observations [first, last): partial = [alpha, beta[1..n1], cummOld, cummAll, missing] */
static int cutPartial(oracleCtx *c, const double *X, int numSamples, int pi_eval, double lb, int64_t first, int64_t last,
		const double *pcx, double *alpha, double *beta, double *cummOld, double *cummAll, int32_t *iStar, int64_t *missing) {
	for (int64_t obs = first; obs < last; obs++) {
		int istar = cutObsIstar(c, pcx, X, obs, numSamples, pi_eval, lb, cummOld, cummAll);
		if (iStar) iStar[obs] = istar;
		if (istar < 0) { (*missing)++; continue; }                                          /* :136-139 */
		if (cutObsAccumulate(c, obs, istar, alpha, beta) < 0) return SDGPU_ERR;
	}
	return 0;
}

static int cutFinish(oracleCtx *c, int numSamples, double alpha, const double *beta, double cummOld, double cummAll,
		int64_t missing, sdgpu_cut *cut) {
	cut->omegaCnt = (int32_t) c->omegaCnt; cut->numSamples = numSamples;
	cut->cummOld = cummOld; cut->cummAll = cummAll;
	if (missing > 0) { fail("sd_cut: failed to identify maximal Pi for an observation"); return SDGPU_NONE; }
	cut->alpha = alpha / numSamples;                                                        /* :184 */
	for (int k = 1; k <= c->num.prevCols; k++) cut->beta[k] = beta[k] / numSamples;         /* :186-187 */
	cut->beta[0] = 1.0;                                                                     /* :188 */
	return 0;
}

int sdo_sd_cut(oracleCtx *c, const double *X, int numSamples, int pi_eval_flag, double lb, sdgpu_cut *cut) {
	double *pcx = (double *) calloc((size_t) c->sigmaCnt + 1, sizeof(double));
	double *beta = (double *) calloc((size_t) c->num.prevCols + 1, sizeof(double));
	double alpha = 0.0, cummOld = 0.0, cummAll = 0.0;
	int64_t missing = 0;
	piCbarXAll(c, X, pcx);
	int r = cutPartial(c, X, numSamples, pi_eval_flag, lb, 0, c->omegaCnt, pcx, &alpha, beta, &cummOld, &cummAll, cut->iStar, &missing);
	if (r == 0) r = cutFinish(c, numSamples, alpha, beta, cummOld, cummAll, missing, cut);
	free(pcx); free(beta);
	return r;
}

/* shard form used by the multi-rank host tests: partial[0] = alpha, [1..n1] = beta, [n1+1] = cummOld,
 * [n1+2] = cummAll, [n1+3] = missing; iStar (may be NULL) has omegaCnt entries. */
int sdo_sd_cut_partial_host(oracleCtx *c, const double *X, int numSamples, int pi_eval_flag, double lb, double *partial, int32_t *iStar) {
	int n1 = c->num.prevCols;
	double *pcx = (double *) calloc((size_t) c->sigmaCnt + 1, sizeof(double));
	double *beta = (double *) calloc((size_t) n1 + 1, sizeof(double));
	double alpha = 0.0, cummOld = 0.0, cummAll = 0.0;
	int64_t missing = 0;
	piCbarXAll(c, X, pcx);
	int r = cutPartial(c, X, numSamples, pi_eval_flag, lb, 0, c->omegaCnt, pcx, &alpha, beta, &cummOld, &cummAll, iStar, &missing);
	partial[0] = alpha;
	for (int k = 1; k <= n1; k++) partial[k] = beta[k];
	partial[n1 + 1] = cummOld; partial[n1 + 2] = cummAll; partial[n1 + 3] = (double) missing;
	free(pcx); free(beta);
	return r;
}

/* "fastcpu" flavour: identical per-element arithmetic and tie-breaks, observations
 * split over OpenMP threads, per-thread sums combined in thread order.  iStar is bit-identical to sdo_sd_cut;
 * alpha/beta agree to rounding.  Returns the thread count used through *threads. */
int sdo_sd_cut_omp(oracleCtx *c, const double *X, int numSamples, int pi_eval_flag, double lb, sdgpu_cut *cut, int *threads) {
	int n1 = c->num.prevCols, nt = 1, bad = 0;
	double *pcx = (double *) calloc((size_t) c->sigmaCnt + 1, sizeof(double));
	piCbarXAll(c, X, pcx);
#ifdef _OPENMP
	nt = omp_get_max_threads();
#endif
	double *acc = (double *) calloc((size_t) nt * (n1 + 4), sizeof(double));
	int64_t *miss = (int64_t *) calloc((size_t) nt, sizeof(int64_t));
#pragma omp parallel num_threads(nt)
	{
		int t = 0;
#ifdef _OPENMP
		t = omp_get_thread_num();
#endif
		int64_t first = c->omegaCnt * t / nt, last = c->omegaCnt * (t + 1) / nt;
		double *a = acc + (size_t) t * (n1 + 4);
		/* a[0] = alpha, a[1..n1] = beta (a itself serves as the 1-based beta accumulator), a[n1+1], a[n1+2] = cumm */
		double alpha = 0.0;
		if (cutPartial(c, X, numSamples, pi_eval_flag, lb, first, last, pcx, &alpha, a, &a[n1 + 1], &a[n1 + 2], cut->iStar, &miss[t]) < 0)
			bad = 1;
		a[0] = alpha;
	}
	double alpha = 0.0, cummOld = 0.0, cummAll = 0.0;
	double *beta = (double *) calloc((size_t) n1 + 1, sizeof(double));
	int64_t missing = 0;
	for (int t = 0; t < nt; t++) {
		const double *a = acc + (size_t) t * (n1 + 4);
		alpha += a[0];
		for (int k = 1; k <= n1; k++) beta[k] += a[k];
		cummOld += a[n1 + 1]; cummAll += a[n1 + 2]; missing += miss[t];
	}
	int r = bad ? SDGPU_ERR : cutFinish(c, numSamples, alpha, beta, cummOld, cummAll, missing, cut);
	if (threads) *threads = nt;
	free(pcx); free(acc); free(miss); free(beta);
	return r;
}

/* cuts.c:171-182 with calcVariance cuts.c:366-396 (mean_value == NULL form: length = SCAN_LEN) */
static double scanVariance(const double *x, int length) {
	double mean = x[0], vari = 0.0, temp;
	for (int count = 1; count < length; count++) {
		temp = mean;
		mean = mean + (x[count] - mean) / (double) (count + 1);
		vari = (1 - 1 / (double) count) * vari + (count + 1) * (mean - temp) * (mean - temp);
	}
	return vari;
}
double sdo_calc_variance(const double *x, int scanLen) { return scanVariance(x, scanLen); }

int sdo_dual_stability(double cummOld, double cummAll, int numSamples, int piEvalStart, int scanLen, double *pi_ratio) {
	double variance;
	pi_ratio[numSamples % scanLen] = cummOld / cummAll;                                      /* :172 */
	if (numSamples - piEvalStart > scanLen) variance = scanVariance(pi_ratio, scanLen);      /* :173-176 */
	else variance = 1.0;
	if (ABSV(variance) >= .000002 || pi_ratio[numSamples % scanLen] < 0.95) return 0;        /* :178-181 */
	return 1;
}

/* ---- cut heights and aging: cuts.c:197-227, master.c:152,174 ---------------------------------------------- */
int sdo_cut_heights(oracleCtx *c, int n, const double *alpha, const double *beta, const int32_t *numSamples,
		const double *alphaIncumb, int currIter, const double *xk, double lb, double *height, double *etaCoef, double *rhs) {
	int n1 = c->num.prevCols, best = SDGPU_NONE;
	double Sm = -1.0e20;                                                                     /* -INF of utils.h (shim value) */
	for (int i = 0; i < n; i++) {
		const double *b = beta + (size_t) i * (n1 + 1);
		double t_over_k = ((double) numSamples[i] / (double) currIter);                     /* cuts.c:215 */
		double h = alpha[i] - dotIdx(b, xk, NULL, n1);                                       /* :218 */
		h *= t_over_k;                                                                       /* :221 */
		h += (1 - t_over_k) * lb;                                                            /* :224 */
		if (height) height[i] = h;
		if (Sm < h) { Sm = h; best = i; }                                                    /* :203-205 */
		if (etaCoef) etaCoef[i] = (double) (currIter) / (double) numSamples[i];             /* master.c:152 */
		if (rhs) rhs[i] = (alphaIncumb ? alphaIncumb[i] : 0.0) + ((double) currIter / (double) numSamples[i] - 1) * lb;   /* master.c:174 */
	}
	return best;
}

/* ---- reformCuts optimal.c:187-236 for one cut ------------------------------------------------------------- */
int sdo_reform_cut(oracleCtx *c, const int32_t *iStar, int omegaCnt, const int32_t *observ, int k, int lbType, int lb,
		double *alphaOut, double *beta) {
	int n1 = c->num.prevCols, n1c = c->num.cntCcols, Q = c->num.rvCOmCnt, count = 0;
	double alpha = 0.0;
	for (int i = 0; i <= n1; i++) beta[i] = 0.0;                                             /* :197-199 */
	for (int n = 0; n < k; n++) {                                                            /* :203-226 */
		int o = observ[n];
		if (o < omegaCnt) {
			int istar = iStar[o];
			const double *w = omegaRow(c, o);
			for (int t = 0; t <= c->bPhiLen[istar]; t++) {
				int s = c->tSigma[c->bTermStart[istar] + t];
				int l = c->sigmaLambda[s];
				double m = (t == 0) ? 1.0 : w[c->rvOffset[2] + c->tOmega[c->bTermStart[istar] + t]];
				alpha += m * (c->sigmaPib[s] + c->deltaPib[l][o]);
				const double *pc = sigmaPiCRow(c, s);
				for (int j = 1; j <= n1c; j++) beta[c->CCols[j]] += m * pc[j];
				for (int j = 1; j <= Q; j++) beta[c->rvCOmCols[j]] += m * c->deltaPiC[l][(size_t) o * (Q + 1) + j];
			}
			count++;
		}
	}
	for (int i = 0; i <= n1; i++) beta[i] /= (double) k;                                     /* :229-230 */
	alpha /= (double) k;                                                                     /* :232 */
	if (lbType == 1) alpha += (1 - (double) count / (double) k) * lb;                       /* :234-235, NONTRIVIAL == 1 */
	*alphaOut = alpha;
	return 0;
}

int sdo_reform_cuts_batch(oracleCtx *c, int nCuts, const int32_t *iStar, int istarStride, const int32_t *omegaCnt,
		int nReps, const int32_t *observ, int k, int lbType, int lb, double *alpha, double *beta) {
	int n1 = c->num.prevCols;
	for (int r = 0; r < nReps; r++)
		for (int i = 0; i < nCuts; i++) {
			int st = sdo_reform_cut(c, iStar + (size_t) i * istarStride, omegaCnt[i], observ + (size_t) r * k, k, lbType, lb,
					alpha + (size_t) r * nCuts + i, beta + ((size_t) r * nCuts + i) * (n1 + 1));
			if (st < 0) return st;
		}
	return 0;
}

/* ---- readers ------------------------------------------------------------------------------------------------ */
int sdo_get_omega(oracleCtx *c, int idx, double *vals, int *weight) {
	if (idx < 0 || idx >= c->omegaCnt) return fail("get_omega: index out of range");
	if (vals) for (int j = 1; j <= c->num.numRV; j++) vals[j] = omegaRow(c, idx)[j];
	if (weight) *weight = c->omegaW[idx];
	return 0;
}
int sdo_get_lambda(oracleCtx *c, int idx, double *vals) {
	if (idx < 0 || idx >= c->lambdaCnt) return fail("get_lambda: index out of range");
	for (int i = 1; i <= c->num.rvRowCnt; i++) vals[i] = lambdaRow(c, idx)[i];
	return 0;
}
int sdo_get_sigma(oracleCtx *c, int idx, double *pib, double *piC, int *lambdaIdx, int *ck) {
	if (idx < 0 || idx >= c->sigmaCnt) return fail("get_sigma: index out of range");
	if (pib) *pib = c->sigmaPib[idx];
	if (piC) for (int k = 1; k <= c->num.cntCcols; k++) piC[k] = sigmaPiCRow(c, idx)[k];
	if (lambdaIdx) *lambdaIdx = c->sigmaLambda[idx];
	if (ck) *ck = c->sigmaCk[idx];
	return 0;
}
int sdo_get_delta_block(oracleCtx *c, int64_t l0, int64_t l1, int64_t o0, int64_t o1, int plane, double *out) {
	int Q = c->num.rvCOmCnt;
	if (l0 < 0 || l1 > c->lambdaCnt || o0 < 0 || o1 > c->omegaCnt || l0 > l1 || o0 > o1 || plane < 0 || plane > Q) return fail("get_delta_block: out of range");
	for (int64_t l = l0; l < l1; l++)
		for (int64_t o = o0; o < o1; o++)
			out[(size_t) (l - l0) * (o1 - o0) + (o - o0)] = plane == 0 ? c->deltaPib[l][o] : c->deltaPiC[l][(size_t) o * (Q + 1) + plane];
	return 0;
}

int sdo_get_delta(oracleCtx *c, int lambdaIdx, int obsIdx, double *pib, double *piC) {
	int Q = c->num.rvCOmCnt;
	if (lambdaIdx < 0 || lambdaIdx >= c->lambdaCnt || obsIdx < 0 || obsIdx >= c->omegaCnt) return fail("get_delta: index out of range");
	if (pib) *pib = c->deltaPib[lambdaIdx][obsIdx];
	if (piC && Q) for (int k = 1; k <= Q; k++) piC[k] = c->deltaPiC[lambdaIdx][(size_t) obsIdx * (Q + 1) + k];
	return 0;
}
