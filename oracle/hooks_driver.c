/*
 * oracle/hooks_driver.c  --  TEST INFRASTRUCTURE ONLY.
 *
 * Flat C entry points (`sdhk_*`) on top of the PATCHED reference host: /root/reference/twoSD_src with
 * integration/twoSD_src.patch applied (stocUpdate.c, cuts.c, optimal.c, randCost.c compiled from a scratch copy by
 * oracle/Makefile, target `hooks`) plus integration/sdgpu_hooks.c, linked against libsdgpu.so -- or, for the CPU suite,
 * against the restated oracle through oracle/shim/sdgpu_as_sdo.h.  This is the host a maintainer gets after applying the
 * patch, minus CPLEX: the solver is the replay LP of oracle/shim/solver_cplex.h, fed with recorded solves.
 *
 * Nothing here re-implements host logic: it builds the reference's structs (probType, cellType pieces) and forwards to the
 * patched calcOmega call site (algo.c:152), stochasticUpdates (stocUpdate.c:14), SDCut (cuts.c:91), updtFeasCutPool
 * (cuts.c:465) and reformCuts (optimal.c:187), with the reference's clock() bracketing of the "argmax" time
 * (subprob.c:68-73, cuts.c:55-62) and its summary line (inout.c:57).
 */
#include "twoSD.h"
#include "sdgpu.h"

configType config;   /* twoSD.c:17 defines it in the real program */

typedef struct {
	numType      num;
	coordType    coord;
	sparseVector bBar, dBar;
	sparseMatrix Cbar;
	oneProblem   sp;
	probType     prob;
	cellType    *cell;          /* omega, basis, gpu, fcutsPool, fUpdt, time, k */
} hostCtx;

static iVector dupInts(const int32_t *src, int n) {
	iVector d = arr_alloc(n + 1, int);
	int i;
	if (src) for (i = 0; i <= n; i++) d[i] = src[i];
	return d;
}

static dVector dupDbls(const double *src, int n) {
	dVector d = arr_alloc(n + 1, double);
	int i;
	if (src) for (i = 0; i <= n; i++) d[i] = src[i];
	return d;
}

/* newCell (setup.c:67-170) as far as the hot path goes: host observation list, host basis list, device tables */
int sdhk_create(const sdgpu_problem *p, const int32_t *rvdOmCols, const char *senx, int dBarCnt, const int32_t *dBarCol,
		const double *dBarVal, int maxIter, int tau, int device, void **out) {
	hostCtx *c = (hostCtx *) calloc(1, sizeof(hostCtx));
	*out = NULL;
	config.MAX_ITER = maxIter; config.TAU = tau;
	c->num.rows = p->num.rows;         c->num.cols = p->num.cols;
	c->num.prevCols = p->num.prevCols; c->num.cntCcols = p->num.cntCcols;
	c->num.rvRowCnt = p->num.rvRowCnt; c->num.rvbOmCnt = p->num.rvbOmCnt;
	c->num.rvCOmCnt = p->num.rvCOmCnt; c->num.rvdOmCnt = p->num.rvdOmCnt;
	c->num.numRV = p->num.numRV;
	c->coord.CCols     = dupInts(p->coord.CCols, p->num.cntCcols);
	c->coord.rvRows    = dupInts(p->coord.rvRows, p->num.rvRowCnt);
	c->coord.rvbOmRows = dupInts(p->coord.rvbOmRows, p->num.rvbOmCnt);
	c->coord.rvCOmCols = dupInts(p->coord.rvCOmCols, p->num.rvCOmCnt);
	c->coord.rvCOmRows = dupInts(p->coord.rvCOmRows, p->num.rvCOmCnt);
	c->coord.rvCols    = dupInts(p->coord.rvCols, p->num.rvCOmCnt);
	c->coord.rvdOmCols = dupInts(rvdOmCols, p->num.rvdOmCnt);
	c->coord.rvOffset  = arr_alloc(3, int);
	c->coord.rvOffset[0] = p->coord.rvOffset[0]; c->coord.rvOffset[1] = p->coord.rvOffset[1];
	c->coord.rvOffset[2] = p->coord.rvOffset[2];
	c->bBar.cnt = p->bBar.cnt; c->bBar.col = dupInts(p->bBar.col, p->bBar.cnt); c->bBar.val = dupDbls(p->bBar.val, p->bBar.cnt);
	c->Cbar.cnt = p->Cbar.cnt; c->Cbar.col = dupInts(p->Cbar.col, p->Cbar.cnt); c->Cbar.row = dupInts(p->Cbar.row, p->Cbar.cnt);
	c->Cbar.val = dupDbls(p->Cbar.val, p->Cbar.cnt);
	c->dBar.cnt = dBarCnt; c->dBar.col = dupInts(dBarCol, dBarCnt); c->dBar.val = dupDbls(dBarVal, dBarCnt);
	c->sp.senx = (char *) calloc((size_t) p->num.rows + 1, 1);
	if (senx) memcpy(c->sp.senx, senx, (size_t) p->num.rows);
	c->prob.num = &c->num; c->prob.coord = &c->coord; c->prob.sp = &c->sp;
	c->prob.bBar = &c->bBar; c->prob.Cbar = &c->Cbar; c->prob.dBar = &c->dBar;

	c->cell = (cellType *) calloc(1, sizeof(cellType));
	c->cell->basis = newBasisType(2 * maxIter + 1, c->num.cols, c->num.rows, WORDLENGTH);      /* setup.c:140 as patched */
	c->cell->omega = newOmega(c->num.numRV, maxIter);                                            /* setup.c:144 */
	c->cell->fcutsPool = newCuts(4096);
	c->cell->gpu = newGpuTables(&c->prob, device);                                               /* sdgpu_hooks.c */
	if (c->cell->gpu == NULL) return SDGPU_ERR;
	*out = c;
	return 0;
}

const char *sdhk_last_error(void) { return sdgpu_last_error(); }

/* cleanCellType (setup.c:195-268), the table part as patched */
int sdhk_reset(void *vc) {
	hostCtx *c = (hostCtx *) vc;
	int n;
	for (n = 0; n < c->cell->basis->cnt; n++) {
		freeOneBasis(c->cell->basis->vals[n]);
		if (c->cell->basis->obsFeasible[n]) mem_free(c->cell->basis->obsFeasible[n]);
		c->cell->basis->vals[n] = NULL; c->cell->basis->obsFeasible[n] = NULL;
	}
	c->cell->basis->cnt = 0;
	freeCutsType(c->cell->fcutsPool, true);
	c->cell->fUpdt[0] = c->cell->fUpdt[1] = 0;
	cleanGpuTables(c->cell->gpu);
	freeOmegaType(c->cell->omega, true);
	c->cell->k = 0;
	c->cell->time.argmaxIter = c->cell->time.argmaxAccumTime = 0.0;
	return 0;
}

void sdhk_destroy(void *vc) {
	hostCtx *c = (hostCtx *) vc;
	if (!c) return;
	sdhk_reset(c);
	freeGpuTables(c->cell->gpu);
	freeBasisType(c->cell->basis, false);
	freeOmegaType(c->cell->omega, false);
	freeCutsType(c->cell->fcutsPool, false);
	free(c->cell);
	free(c->coord.CCols); free(c->coord.rvRows); free(c->coord.rvbOmRows); free(c->coord.rvCOmCols);
	free(c->coord.rvCOmRows); free(c->coord.rvCols); free(c->coord.rvdOmCols); free(c->coord.rvOffset);
	free(c->bBar.col); free(c->bBar.val); free(c->Cbar.col); free(c->Cbar.row); free(c->Cbar.val);
	free(c->dBar.col); free(c->dBar.val); free(c->sp.senx);
	free(c);
}

int sdhk_get_counts(void *vc, sdgpu_counts *out) {
	hostCtx *c = (hostCtx *) vc;
	if (sdgpu_get_counts(c->cell->gpu, out)) return SDGPU_ERR;
	if (out->omega != c->cell->omega->cnt || out->basis != c->cell->basis->cnt) return SDGPU_ERR;     /* host lists and device records in step */
	return 0;
}

/* algo.c:152 as patched: the observation goes to the host list (computeRHS reads it) and to the device */
int sdhk_calc_omega(void *vc, const double *observ, double tol, int *newOmegaFlag) {
	hostCtx *c = (hostCtx *) vc;
	bool flag = false;
	int idx = calcOmega((dVector) observ, 0, c->num.numRV, c->cell->omega, &flag, tol);
	if (calcOmega_gpu(c->cell->gpu, (dVector) observ, &flag, tol) != idx) return SDGPU_ERR;
	if (newOmegaFlag) *newOmegaFlag = flag;
	return idx;
}

/* the tail of solveSubprob (subprob.c:66-74 as patched) on a replayed solve */
int sdhk_stochastic_updates(void *vc, const sdReplayLP *lp, int omegaIdx, int newOmegaFlag, int currentIter, double tol,
		int subFeasFlag, int *newBasisFlag) {
	hostCtx *c = (hostCtx *) vc;
	bool nb = newBasisFlag ? (*newBasisFlag != 0) : true;
	clock_t tic = clock();
	int status = stochasticUpdates(&c->prob, (LPptr) lp, c->cell->basis, c->cell->gpu, c->cell->omega, omegaIdx, newOmegaFlag != 0,
			currentIter, tol, &nb, subFeasFlag != 0);
	c->cell->time.argmaxIter += ((double) (clock() - tic)) / CLOCKS_PER_SEC;
	if (newBasisFlag) *newBasisFlag = nb;
	return status;
}

/* formSDCut (b) (cuts.c:54-62 as patched) with the reference's own config gates.  pi_ratio has scanLen entries. */
int sdhk_sd_cut_cfg(void *vc, const double *Xvect, int numSamples, int dualStability, int piEvalStart, int piCycle, int scanLen,
		double lb, sdgpu_cut *cut, double *pi_ratio, int *dualStableFlag) {
	hostCtx *c = (hostCtx *) vc;
	oneCut *rc;
	bool stable = dualStableFlag ? (*dualStableFlag != 0) : false;
	clock_t tic;
	int i;
	config.DUAL_STABILITY = dualStability; config.PI_EVAL_START = piEvalStart; config.PI_CYCLE = piCycle; config.SCAN_LEN = scanLen;
	c->cell->k = numSamples;
	tic = clock();
	rc = SDCut(&c->num, c->cell->gpu, c->cell->omega->cnt, (dVector) Xvect, numSamples, &stable, pi_ratio, lb);
	c->cell->time.argmaxIter += ((double) (clock() - tic)) / CLOCKS_PER_SEC;
	if (rc == NULL) return SDGPU_NONE;
	cut->alpha = rc->alpha; cut->omegaCnt = rc->omegaCnt; cut->numSamples = rc->numSamples;
	for (i = 0; i <= c->num.prevCols; i++) cut->beta[i] = rc->beta[i];
	if (cut->iStar) for (i = 0; i < rc->omegaCnt; i++) cut->iStar[i] = rc->iStar[i];
	if (dualStableFlag) *dualStableFlag = stable;
	freeOneCut(rc);
	return 0;
}

/* the per-iteration roll-up of algo.c:179-182 and the summary line of inout.c:57 */
double sdhk_end_iteration(void *vc) {
	hostCtx *c = (hostCtx *) vc;
	c->cell->time.argmaxAccumTime += c->cell->time.argmaxIter;
	c->cell->time.argmaxIter = 0.0;
	return c->cell->time.argmaxAccumTime;
}

void sdhk_print_summary(void *vc) {
	hostCtx *c = (hostCtx *) vc;
	fprintf(stdout, "Number of unique observations      : %d\n", c->cell->omega->cnt);
	fprintf(stdout, "Total time for argmax operation    : %f\n", c->cell->time.argmaxAccumTime);
	fflush(stdout);
}

int sdhk_basis_info(void *vc, int b, int *ck, int *weight, int *phiLength, int *feasFlag, double *mubBar, int32_t *sigmaIdx,
		int32_t *omegaIdx, uint8_t *obsFeasible) {
	hostCtx *c = (hostCtx *) vc;
	oneBasis *B;
	int i;
	if (b < 0 || b >= c->cell->basis->cnt) return SDGPU_ERR;
	B = c->cell->basis->vals[b];
	*ck = B->ck; *weight = B->weight; *phiLength = B->phiLength; *feasFlag = B->feasFlag; *mubBar = B->mubBar;
	for (i = 0; i <= B->phiLength; i++) sigmaIdx[i] = B->sigmaIdx[i];
	for (i = 1; i <= B->phiLength; i++) omegaIdx[i] = B->omegaIdx[i];
	if (obsFeasible)
		for (i = 0; i < c->cell->omega->cnt; i++) obsFeasible[i] = c->cell->basis->obsFeasible[b] ? c->cell->basis->obsFeasible[b][i] : 2;
	return 0;
}

int updtFeasCutPool(numType *num, coordType *coord, cellType *cell);      /* cuts.c:465 (file-local prototype in cuts.c:18) */

/* the patched updtFeasCutPool (cuts.c:465-517; gathers on the device, pool de-duplication cuts.c:643-655 on the host) */
int sdhk_updt_feas_cut_pool(void *vc, int *fUpdt, double tol, int maxOut, double *alpha, double *beta) {
	hostCtx *c = (hostCtx *) vc;
	int i, j, n1 = c->num.prevCols;
	c->cell->fUpdt[0] = fUpdt[0]; c->cell->fUpdt[1] = fUpdt[1];
	config.TOLERANCE = tol;
	updtFeasCutPool(&c->num, &c->coord, c->cell);
	fUpdt[0] = c->cell->fUpdt[0]; fUpdt[1] = c->cell->fUpdt[1];
	if (c->cell->fcutsPool->cnt > maxOut) return SDGPU_ERR;
	for (i = 0; i < c->cell->fcutsPool->cnt; i++) {
		alpha[i] = c->cell->fcutsPool->vals[i]->alpha;
		for (j = 0; j <= n1; j++) beta[(size_t) i * (n1 + 1) + j] = c->cell->fcutsPool->vals[i]->beta[j];
	}
	return c->cell->fcutsPool->cnt;
}

/* the patched reformCuts (optimal.c:187-236) for one cut */
int sdhk_reform_cut(void *vc, const int32_t *iStar, int omegaCnt, const int32_t *observ, int k, int lbType, int lb,
		double *alpha, double *beta) {
	hostCtx *c = (hostCtx *) vc;
	cutsType *g = newCuts(1);
	int i;
	g->vals[0] = newCut(c->num.prevCols, omegaCnt, k);
	for (i = 0; i < omegaCnt; i++) g->vals[0]->iStar[i] = iStar[i];
	g->cnt = 1;
	reformCuts(c->cell->gpu, &c->num, g, (int *) observ, k, lbType, lb, c->num.prevCols);
	*alpha = g->vals[0]->alpha;
	for (i = 0; i <= c->num.prevCols; i++) beta[i] = g->vals[0]->beta[i];
	freeCutsType(g, false);
	return 0;
}

/* ---- link-time stubs for host functions cuts.c / optimal.c name but this harness never reaches -------- */
#define NOT_IN_HARNESS(name) do { fprintf(stderr, "sdhk :: %s() is host/CPLEX code outside the harness\n", name); abort(); } while (0)
int solveSubprob(probType *prob, oneProblem *subproblem, dVector Xvect, basisType *basis, sdgpu_ctx *gpu,
		omegaType *omega, int omegaIdx, bool *newOmegaFlag, int currentIter, double TOLERANCE,
		bool *subFeasFlag, bool *newBasisFlag, double *subprobTime, double *argmaxTime) { NOT_IN_HARNESS("solveSubprob"); return 1; }
int solveQPMaster(numType *num, sparseVector *dBar, cellType *cell, double lb) { NOT_IN_HARNESS("solveQPMaster"); return 1; }
int addCut2Master(oneProblem *master, oneCut *cut, dVector vectX, int lenX) { NOT_IN_HARNESS("addCut2Master"); return 1; }
int replaceIncumbent(probType *prob, cellType *cell, double candidEst) { NOT_IN_HARNESS("replaceIncumbent"); return 1; }
int changeQPproximal(LPptr lp, int numCols, double sigma) { NOT_IN_HARNESS("changeQPproximal"); return 1; }
