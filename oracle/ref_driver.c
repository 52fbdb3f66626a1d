/*
 * oracle/ref_driver.c  --  TEST INFRASTRUCTURE ONLY.
 *
 * Flat C entry points (`sdref_*`, same shapes as include/sdgpu.h) on top of the REFERENCE'S OWN
 * functions, compiled from the sources where they lie under /root/reference/twoSD_src
 * (stocUpdate.c, cuts.c, optimal.c, randCost.c) against the header shim in oracle/shim/.  The recipe is
 * oracle/Makefile; the output is oracle/_ref/libsdref.so (git-ignored, travels to the GPU box).
 *
 * Nothing here re-implements table or cut arithmetic: it only builds the reference's structs
 * (stoc.h:22-97, twoSD.h:69-85) from flat arrays and forwards to
 *   calcOmega  stocUpdate.c:326   calcLambda stocUpdate.c:264   calcSigma stocUpdate.c:286
 *   calcDelta  stocUpdate.c:196   computeIstar stocUpdate.c:142 SDCut cuts.c:91
 *   cutHeight  cuts.c:213         maxCutHeight cuts.c:197       calcVariance cuts.c:366
 *   reformCuts optimal.c:187      checkBasisFeasibility randCost.c:202
 * The basis bookkeeping of stochasticUpdates (stocUpdate.c:101-131) needs CPLEX for everything before
 * it, so the append / dedup step is replayed here from its inputs.
 */
#include "twoSD.h"
#include "../include/sdgpu.h"

configType config;   /* twoSD.c:17 defines it in the real program */

/* checkFeasCutPool (cuts.c:521-567) hands the cuts it wants in the master to addCut2Master: record them instead of calling CPLEX */
static oneCut **g_addedCuts = NULL;
static int g_addedCnt = 0, g_addedCap = 0, g_recordAdds = 0;

typedef struct {
	numType      num;
	coordType    coord;
	sparseVector bBar;
	sparseMatrix Cbar;
	sparseVector dBar;          /* second-stage cost vector, read by calcBasis (randCost.c:70); sdref_set_dbar */
	sdgpu_caps   caps;
	iVector      rvdOmCols; char *senx;
	cellType    *cell;          /* only the fields updtFeasCutPool / addCut2Pool(FEASIBILITY) read */
	lambdaType  *lambda;
	sigmaType   *sigma;
	deltaType   *delta;
	omegaType   *omega;
	basisType   *basis;
} refCtx;

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

int sdref_create(const sdgpu_problem *p, const sdgpu_caps *caps, int device, void **out) {
	refCtx *c = (refCtx *) calloc(1, sizeof(refCtx));
	(void) device;
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
	c->coord.rvOffset  = arr_alloc(3, int);
	c->coord.rvOffset[0] = p->coord.rvOffset[0]; c->coord.rvOffset[1] = p->coord.rvOffset[1];
	c->coord.rvOffset[2] = p->coord.rvOffset[2];
	c->bBar.cnt = p->bBar.cnt; c->bBar.col = dupInts(p->bBar.col, p->bBar.cnt); c->bBar.val = dupDbls(p->bBar.val, p->bBar.cnt);
	c->Cbar.cnt = p->Cbar.cnt; c->Cbar.col = dupInts(p->Cbar.col, p->Cbar.cnt); c->Cbar.row = dupInts(p->Cbar.row, p->Cbar.cnt);
	c->Cbar.val = dupDbls(p->Cbar.val, p->Cbar.cnt);
	c->caps = *caps;
	/* setup.c:140-144 */
	c->basis  = newBasisType((int) caps->maxBasis, c->num.cols, c->num.rows, WORDLENGTH);
	c->lambda = newLambda((int) caps->maxLambda, 0, c->num.rvRowCnt);
	c->sigma  = newSigma((int) caps->maxSigma, c->num.cntCcols, 0);
	c->delta  = newDelta((int) caps->maxLambda);
	c->omega  = newOmega(c->num.numRV, (int) caps->maxOmega);
	*out = c;
	return 0;
}

static void dropBases(refCtx *c) {
	int n;
	for (n = 0; n < c->basis->cnt; n++) {
		freeOneBasis(c->basis->vals[n]);
		if (c->basis->obsFeasible[n]) mem_free(c->basis->obsFeasible[n]);
		c->basis->vals[n] = NULL; c->basis->obsFeasible[n] = NULL;
	}
	c->basis->cnt = 0;
}

int sdref_reset(void *vc) {
	refCtx *c = (refCtx *) vc;
	int n;
	/* setup.c:236,242-246 */
	if (c->cell && c->cell->fcutsPool) { freeCutsType(c->cell->fcutsPool, true); c->cell->fUpdt[0] = c->cell->fUpdt[1] = 0; }
	dropBases(c);
	freeDeltaType(c->delta, c->lambda->cnt, c->omega->cnt, true);
	for (n = 0; n < c->lambda->cnt; n++) c->delta->vals[n] = NULL;
	freeLambdaType(c->lambda, true);
	freeSigmaType(c->sigma, true);
	freeOmegaType(c->omega, true);
	return 0;
}

void sdref_destroy(void *vc) {
	refCtx *c = (refCtx *) vc;
	if (!c) return;
	sdref_reset(c);
	freeBasisType(c->basis, false);
	freeDeltaType(c->delta, 0, 0, false);
	freeLambdaType(c->lambda, false);
	freeSigmaType(c->sigma, false);
	freeOmegaType(c->omega, false);
	free(c->coord.CCols); free(c->coord.rvRows); free(c->coord.rvbOmRows); free(c->coord.rvCOmCols);
	free(c->coord.rvCOmRows); free(c->coord.rvCols); free(c->coord.rvOffset);
	free(c->bBar.col); free(c->bBar.val); free(c->Cbar.col); free(c->Cbar.row); free(c->Cbar.val);
	free(c);
}

int sdref_get_counts(void *vc, sdgpu_counts *out) {
	refCtx *c = (refCtx *) vc;
	out->omega = c->omega->cnt; out->lambda = c->lambda->cnt; out->sigma = c->sigma->cnt; out->basis = c->basis->cnt;
	return 0;
}

int sdref_calc_omega(void *vc, const double *observ, double tol, int *newOmegaFlag) {
	refCtx *c = (refCtx *) vc;
	bool flag = false;
	int idx = calcOmega((dVector) observ, 0, c->num.numRV, c->omega, &flag, tol);   /* algo.c:152 */
	if (newOmegaFlag) *newOmegaFlag = flag;
	return idx;
}

int sdref_calc_lambda(void *vc, const double *Pi, double tol, int *newLambdaFlag) {
	refCtx *c = (refCtx *) vc;
	bool flag = false;
	int idx = calcLambda(&c->num, &c->coord, (dVector) Pi, c->lambda, &flag, tol);
	if (newLambdaFlag) *newLambdaFlag = flag;
	return idx;
}

int sdref_calc_sigma(void *vc, const double *pi, double mubBar, int idxLambda, int newLambdaFlag, int currentIter,
		double tol, int *newSigmaFlag) {
	refCtx *c = (refCtx *) vc;
	bool flag = false;
	int idx = calcSigma(&c->num, &c->coord, &c->bBar, &c->Cbar, (dVector) pi, mubBar, idxLambda, newLambdaFlag != 0,
			currentIter, c->sigma, &flag, tol);
	if (newSigmaFlag) *newSigmaFlag = flag;
	return idx;
}

int sdref_calc_delta(void *vc, int newOmegaFlag, int elemIdx) {
	refCtx *c = (refCtx *) vc;
	return calcDelta(&c->num, &c->coord, c->lambda, c->delta, (int) c->caps.maxOmega, c->omega, newOmegaFlag != 0, elemIdx);
}

/* stocUpdate.c:78-85 for one dual vector */
int sdref_update_dual(void *vc, const double *pi, double mubBar, int currentIter, double tol,
		int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag) {
	int nl = 0, ns = 0, li, si;
	li = sdref_calc_lambda(vc, pi, tol, &nl);
	si = sdref_calc_sigma(vc, pi, mubBar, li, nl, currentIter, tol, &ns);
	if (nl) sdref_calc_delta(vc, 0, li);
	if (lambdaIdx) *lambdaIdx = li;
	if (newLambdaFlag) *newLambdaFlag = nl;
	if (sigmaIdx) *sigmaIdx = si;
	if (newSigmaFlag) *newSigmaFlag = ns;
	return 0;
}

int sdref_update_dual_col(void *vc, int newOmegaIdx, const double *pi, double mubBar, int currentIter, double tol,
		int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag) {
	if (newOmegaIdx >= 0) sdref_calc_delta(vc, 1, newOmegaIdx);                 /* stocUpdate.c:24-25 */
	return sdref_update_dual(vc, pi, mubBar, currentIter, tol, lambdaIdx, newLambdaFlag, sigmaIdx, newSigmaFlag);
}

/* Bulk loader for the timing harness (bench.py): every table entry is still produced by the reference's own
 * calcLambda / calcSigma / calcDelta; only the loop over NEW lambda rows of calcDelta case II (rows are
 * independent, stocUpdate.c:230-254) is spread over OpenMP threads so that building a timing sample does not
 * take longer than timing it.  Observations are appended without the calcOmega scan. */
int sdref_bulk_load(void *vc, int nObs, const double *obsVals, const int32_t *weights, int nDuals, const double *pis,
		const double *mubBar, const int32_t *iters, double tol, int32_t *lambdaIdx, int32_t *sigmaIdx) {
	refCtx *c = (refCtx *) vc;
	int i, nNew = 0, *newRows = arr_alloc(nDuals + 1, int);
	for (i = 0; i < nObs; i++) {
		if (c->omega->cnt >= (int) c->caps.maxOmega) return SDGPU_ERR;
		c->omega->vals[c->omega->cnt] = duplicVector((dVector) (obsVals + (size_t) i * (c->num.numRV + 1)), c->num.numRV + 1);
		c->omega->weights[c->omega->cnt++] = weights ? weights[i] : 1;
	}
	for (i = 0; i < nDuals; i++) {
		const double *pi = pis + (size_t) i * (c->num.rows + 1);
		bool nl = false, ns = false;
		int li = calcLambda(&c->num, &c->coord, (dVector) pi, c->lambda, &nl, tol);
		int si = calcSigma(&c->num, &c->coord, &c->bBar, &c->Cbar, (dVector) pi, mubBar ? mubBar[i] : 0.0, li, nl,
				iters ? iters[i] : i + 1, c->sigma, &ns, tol);
		if (nl) newRows[nNew++] = li;
		if (lambdaIdx) lambdaIdx[i] = li;
		if (sigmaIdx) sigmaIdx[i] = si;
	}
#pragma omp parallel for schedule(dynamic, 4)
	for (i = 0; i < nNew; i++)
		calcDelta(&c->num, &c->coord, c->lambda, c->delta, (int) c->caps.maxOmega, c->omega, false, newRows[i]);
	mem_free(newRows);
	return 0;
}

static oneBasis *makeBasis(int ck, int feasFlag, int phiLength, const int32_t *sigmaIdx, const int32_t *omegaIdx) {
	oneBasis *B = (oneBasis *) calloc(1, sizeof(oneBasis));
	int i;
	B->ck = ck; B->weight = 1; B->phiLength = phiLength; B->feasFlag = feasFlag != 0;
	B->sigmaIdx = arr_alloc(phiLength + 1, int);
	for (i = 0; i <= phiLength; i++) B->sigmaIdx[i] = sigmaIdx[i];
	if (phiLength > 0) {
		B->omegaIdx = arr_alloc(phiLength + 1, int);
		for (i = 1; i <= phiLength; i++) B->omegaIdx[i] = omegaIdx[i];
	}
	return B;
}

static int pushBasis(refCtx *c, oneBasis *B) {
	int cnt;
	/* stocUpdate.c:117-131 with checkBasisFeasibility == true (the caller overrides entries afterwards) */
	c->basis->vals[c->basis->cnt] = B;
	if (B->feasFlag) {
		c->basis->obsFeasible[c->basis->cnt] = (bool *) arr_alloc((int) c->caps.maxOmega, bool);
		for (cnt = 0; cnt < (int) c->caps.maxOmega; cnt++) c->basis->obsFeasible[c->basis->cnt][cnt] = true;
	}
	else
		c->basis->obsFeasible[c->basis->cnt] = NULL;
	return c->basis->cnt++;
}

int sdref_basis_append(void *vc, int ck, int feasFlag, int phiLength, const int32_t *sigmaIdx, const int32_t *omegaIdx) {
	return pushBasis((refCtx *) vc, makeBasis(ck, feasFlag, phiLength, sigmaIdx, omegaIdx));
}

int sdref_basis_find_or_append(void *vc, int retainBasis, int obsIdx, int ck, int feasFlag, int phiLength,
		const int32_t *sigmaIdx, const int32_t *omegaIdx, int *newBasisFlag) {
	refCtx *c = (refCtx *) vc;
	oneBasis *B = makeBasis(ck, feasFlag, phiLength, sigmaIdx, omegaIdx);
	int cnt;
	if (newBasisFlag) *newBasisFlag = 1;
	if (!retainBasis) {
		/* replay of stocUpdate.c:101-113 (obsFeasible of an infeasible basis is NULL there: treated as false) */
		for (cnt = 0; cnt < c->basis->cnt; cnt++) {
			if (B->phiLength == c->basis->vals[cnt]->phiLength && c->basis->obsFeasible[cnt] && c->basis->obsFeasible[cnt][obsIdx]) {
				if (equalIntvec(B->sigmaIdx - 1, c->basis->vals[cnt]->sigmaIdx - 1, B->phiLength + 1)) {
					freeOneBasis(B);
					c->basis->vals[cnt]->weight++;
					if (newBasisFlag) *newBasisFlag = 0;
					return cnt;
				}
			}
		}
	}
	return pushBasis(c, B);
}

int sdref_basis_set_obs_feasible(void *vc, int basisIdx, int obsIdx, int flag) {
	refCtx *c = (refCtx *) vc;
	if (basisIdx < 0 || basisIdx >= c->basis->cnt || !c->basis->obsFeasible[basisIdx]) return SDGPU_ERR;
	c->basis->obsFeasible[basisIdx][obsIdx] = flag != 0;
	return 0;
}

int sdref_basis_set_obs_feasible_row(void *vc, int basisIdx, const uint8_t *flags) {
	refCtx *c = (refCtx *) vc;
	int o;
	if (basisIdx < 0 || basisIdx >= c->basis->cnt || !c->basis->obsFeasible[basisIdx]) return SDGPU_ERR;
	for (o = 0; o < c->omega->cnt; o++) c->basis->obsFeasible[basisIdx][o] = flags[o] != 0;
	return 0;
}

int sdref_basis_set_obs_feasible_col(void *vc, int obsIdx, const uint8_t *flags) {
	refCtx *c = (refCtx *) vc;
	int b;
	for (b = 0; b < c->basis->cnt; b++)
		if (c->basis->obsFeasible[b]) c->basis->obsFeasible[b][obsIdx] = flags[b] != 0;
	return 0;
}

/* checkBasisFeasibility randCost.c:202-258: fill the oneBasis fields calcBasis / decomposeDualSolution would have set */
int sdref_set_cost_coords(void *vc, const int32_t *rvdOmCols, const char *senx) {
	refCtx *c = (refCtx *) vc;
	c->rvdOmCols = dupInts(rvdOmCols, c->num.rvdOmCnt);
	c->senx = (char *) malloc(c->num.rows + 1);
	memcpy(c->senx, senx, c->num.rows);
	return 0;
}

int sdref_basis_set_feas_data(void *vc, int b, const double *piDet, const double *phi, const double *gBar, const double *psiVal,
		const int32_t *cstat) {
	refCtx *c = (refCtx *) vc;
	oneBasis *B = c->basis->vals[b];
	int rows = c->num.rows, cols = c->num.cols, i, j, n;
	iVector cs = dupInts(cstat, cols);
	B->piDet = dupDbls(piDet, rows);
	B->gBar = dupDbls(gBar, cols);
	B->cCode = encodeIntvec(cs, cols, WORDLENGTH, 3);                      /* randCost.c:171 */
	mem_free(cs);
	if (B->phiLength > 0) {
		B->phi = (dVector *) arr_alloc(B->phiLength, dVector);
		for (n = 0; n < B->phiLength; n++) B->phi[n] = dupDbls(phi + (size_t) n * (rows + 1), rows);
		B->psi = (sparseMatrix *) mem_malloc(sizeof(sparseMatrix));         /* randCost.c:56-61,83-88 */
		B->psi->val = (dVector) arr_alloc(cols * B->phiLength + 1, double);
		B->psi->col = (iVector) arr_alloc(cols * B->phiLength + 1, int);
		B->psi->row = (iVector) arr_alloc(cols * B->phiLength + 1, int);
		B->psi->cnt = 0;
		for (i = 1; i <= cols; i++)
			for (j = 1; j <= B->phiLength; j++) {
				B->psi->row[B->psi->cnt + 1] = i;
				B->psi->col[B->psi->cnt + 1] = B->omegaIdx[j];
				B->psi->val[B->psi->cnt + 1] = psiVal[(size_t) (i - 1) * B->phiLength + (j - 1)];
				B->psi->cnt++;
			}
	}
	return 0;
}

static bool refPairFeasible(refCtx *c, int b, int obs, double tol) {
	sparseVector dOmega;
	dOmega.cnt = c->num.rvdOmCnt; dOmega.col = c->rvdOmCols;
	dOmega.val = c->coord.rvOffset[2] + c->omega->vals[obs];                /* stocUpdate.c:28 */
	return checkBasisFeasibility(c->basis->vals[b], dOmega, c->senx, c->num.cols, c->num.rows, tol);
}

int sdref_check_feasibility_obs(void *vc, int obsIdx, double tol, uint8_t *flagsOut) {
	refCtx *c = (refCtx *) vc;
	int b;
	for (b = 0; b < c->basis->cnt; b++) {
		if (c->basis->obsFeasible[b] && c->basis->vals[b]->piDet) c->basis->obsFeasible[b][obsIdx] = refPairFeasible(c, b, obsIdx, tol);
		if (flagsOut) flagsOut[b] = c->basis->obsFeasible[b] ? c->basis->obsFeasible[b][obsIdx] : 0;
	}
	return 0;
}

int sdref_check_feasibility_basis(void *vc, int b, double tol, uint8_t *flagsOut) {
	refCtx *c = (refCtx *) vc;
	int o;
	if (b < 0 || b >= c->basis->cnt || !c->basis->obsFeasible[b] || !c->basis->vals[b]->piDet) return SDGPU_ERR;
	for (o = 0; o < c->omega->cnt; o++) {
		c->basis->obsFeasible[b][o] = refPairFeasible(c, b, o, tol);
		if (flagsOut) flagsOut[o] = c->basis->obsFeasible[b][o];
	}
	return 0;
}

/* ---- the reference's own stochasticUpdates (stocUpdate.c:14-133) on a replayed solve ---------------------------------------
 * Everything before the table arithmetic asks the solver for the basis, the duals and (random cost) rows of the basis inverse:
 * the replay LP of oracle/shim/solver_cplex.h answers from one recorded solve.  Used by tests/test_host_patch.py as the unpatched
 * side of the lock step with the patched host (oracle/hooks_driver.c). */
int sdref_set_dbar(void *vc, int cnt, const int32_t *col, const double *val) {
	refCtx *c = (refCtx *) vc;
	c->dBar.cnt = cnt; c->dBar.col = dupInts(col, cnt); c->dBar.val = dupDbls(val, cnt);
	return 0;
}

int sdref_stochastic_updates(void *vc, const sdReplayLP *lp, int omegaIdx, int newOmegaFlag, int currentIter, double tol,
		int subFeasFlag, int *newBasisFlag) {
	refCtx *c = (refCtx *) vc;
	probType prob;
	oneProblem sp;
	coordType coord = c->coord;
	bool nb = newBasisFlag ? (*newBasisFlag != 0) : true;
	int status;
	memset(&prob, 0, sizeof prob); memset(&sp, 0, sizeof sp);
	coord.rvdOmCols = c->rvdOmCols;
	sp.senx = c->senx;
	prob.num = &c->num; prob.coord = &coord; prob.sp = &sp; prob.bBar = &c->bBar; prob.Cbar = &c->Cbar; prob.dBar = &c->dBar;
	status = stochasticUpdates(&prob, (LPptr) lp, c->basis, c->lambda, c->sigma, c->delta, (int) c->caps.maxOmega, c->omega, omegaIdx,
			newOmegaFlag != 0, currentIter, tol, &nb, subFeasFlag != 0);
	if (newBasisFlag) *newBasisFlag = nb;
	return status;
}

/* one record of the host's basis list (stoc.h:72-97); obsFeasible may be NULL, its first omega->cnt entries are copied */
int sdref_basis_info(void *vc, int b, int *ck, int *weight, int *phiLength, int *feasFlag, double *mubBar, int32_t *sigmaIdx,
		int32_t *omegaIdx, uint8_t *obsFeasible) {
	refCtx *c = (refCtx *) vc;
	oneBasis *B;
	int i;
	if (b < 0 || b >= c->basis->cnt) return SDGPU_ERR;
	B = c->basis->vals[b];
	*ck = B->ck; *weight = B->weight; *phiLength = B->phiLength; *feasFlag = B->feasFlag; *mubBar = B->mubBar;
	for (i = 0; i <= B->phiLength; i++) sigmaIdx[i] = B->sigmaIdx[i];
	for (i = 1; i <= B->phiLength; i++) omegaIdx[i] = B->omegaIdx[i];
	if (obsFeasible)
		for (i = 0; i < c->omega->cnt; i++) obsFeasible[i] = c->basis->obsFeasible[b] ? c->basis->obsFeasible[b][i] : 2;
	return 0;
}

int updtFeasCutPool(numType *num, coordType *coord, cellType *cell);      /* cuts.c:465 (file-local prototype in cuts.c:18) */

/* The reference's own updtFeasCutPool (cuts.c:465-517) including the pool de-duplication of addCut2Pool (cuts.c:643-655).
 * fUpdt is cell->fUpdt (in/out); the whole pool is exported after the call.  Returns the pool size. */
int sdref_updt_feas_cut_pool(void *vc, int *fUpdt, double tol, int maxOut, double *alpha, double *beta) {
	refCtx *c = (refCtx *) vc;
	int i, j, n1 = c->num.prevCols;
	if (!c->cell) {
		c->cell = (cellType *) calloc(1, sizeof(cellType));
		c->cell->fcutsPool = newCuts(4096);
	}
	c->cell->omega = c->omega; c->cell->basis = c->basis; c->cell->sigma = c->sigma; c->cell->delta = c->delta;
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

/* same shapes as sdgpu_feas_pool_* (include/sdgpu.h), on the reference's own updtFeasCutPool / addCut2Pool / checkFeasCutPool */
static void ensureCell(refCtx *c) {
	if (!c->cell) {
		c->cell = (cellType *) calloc(1, sizeof(cellType));
		c->cell->fcutsPool = newCuts(65536);
	}
	c->cell->omega = c->omega; c->cell->basis = c->basis; c->cell->sigma = c->sigma; c->cell->delta = c->delta;
}

int sdref_feas_pool_update(void *vc, int *fUpdt, double tol) {
	refCtx *c = (refCtx *) vc;
	ensureCell(c);
	c->cell->fUpdt[0] = fUpdt[0]; c->cell->fUpdt[1] = fUpdt[1];
	config.TOLERANCE = tol;
	updtFeasCutPool(&c->num, &c->coord, c->cell);
	fUpdt[0] = c->cell->fUpdt[0]; fUpdt[1] = c->cell->fUpdt[1];
	return c->cell->fcutsPool->cnt;
}

int sdref_feas_pool_size(void *vc) { refCtx *c = (refCtx *) vc; return c->cell ? c->cell->fcutsPool->cnt : 0; }

int sdref_feas_pool_get(void *vc, int first, int count, double *alpha, double *beta) {
	refCtx *c = (refCtx *) vc;
	int i, j, n1 = c->num.prevCols;
	if (!c->cell || first < 0 || count < 0 || first + count > c->cell->fcutsPool->cnt) return count == 0 ? 0 : SDGPU_ERR;
	for (i = 0; i < count; i++) {
		alpha[i] = c->cell->fcutsPool->vals[first + i]->alpha;
		for (j = 0; j <= n1; j++) beta[(size_t) i * (n1 + 1) + j] = c->cell->fcutsPool->vals[first + i]->beta[j];
	}
	return count;
}

/* action[i]: 1 = the reference added pool cut i to the master (cuts.c:545 or :556), 0 = it did not; the reference does not tell the two
 * "add" cases apart, nor "violated but already there" from "nothing" other than through infeasIncumb */
int sdref_feas_pool_check(void *vc, int nFcuts, const double *fAlpha, const double *fBeta, const double *incumbX,
		const double *candidX, double tol, int32_t *action, int *infeasIncumb) {
	refCtx *c = (refCtx *) vc;
	int i, j, n1 = c->num.prevCols;
	ensureCell(c);
	c->cell->fcuts = newCuts(nFcuts > 0 ? nFcuts : 1);
	for (i = 0; i < nFcuts; i++) {
		c->cell->fcuts->vals[i] = newCut(n1, 0, 1);
		c->cell->fcuts->vals[i]->alpha = fAlpha[i];
		for (j = 0; j <= n1; j++) c->cell->fcuts->vals[i]->beta[j] = fBeta[(size_t) i * (n1 + 1) + j];
		c->cell->fcuts->cnt++;
	}
	c->cell->incumbX = (dVector) incumbX; c->cell->candidX = (dVector) candidX; c->cell->infeasIncumb = false;
	config.TOLERANCE = tol;
	g_recordAdds = 1; g_addedCnt = 0;
	checkFeasCutPool(c->cell, n1);
	g_recordAdds = 0;
	for (i = 0; i < c->cell->fcutsPool->cnt; i++) {
		action[i] = 0;
		for (j = 0; j < g_addedCnt; j++) if (g_addedCuts[j] == c->cell->fcutsPool->vals[i]) action[i] = 1;
	}
	if (infeasIncumb) *infeasIncumb = c->cell->infeasIncumb;
	freeCutsType(c->cell->fcuts, false); c->cell->fcuts = NULL;
	c->cell->incumbX = c->cell->candidX = NULL;
	return c->cell->fcutsPool->cnt;
}

int sdref_compute_istar(void *vc, const double *Xvect, int obs, int numSamples, int pi_eval, int isNew, double *argmax) {
	refCtx *c = (refCtx *) vc;
	dVector piCbarX = arr_alloc(c->sigma->cnt + 1, double);
	int s, idx;
	for (s = 0; s < c->sigma->cnt; s++)     /* cuts.c:105-106 */
		piCbarX[s] = vXv(c->sigma->vals[s].piC, (dVector) Xvect, c->coord.CCols, c->num.cntCcols);
	idx = computeIstar(&c->num, &c->coord, c->basis, c->sigma, c->delta, piCbarX, (dVector) Xvect, c->omega->vals[obs],
			obs, numSamples, pi_eval != 0, argmax, isNew != 0);
	mem_free(piCbarX);
	return idx;
}

/* SDCut with the reference's own config gates.  piRatioOut receives pi_ratio[numSamples % SCAN_LEN]. */
int sdref_sd_cut_cfg(void *vc, const double *Xvect, int numSamples, int dualStability, int piEvalStart, int piCycle,
		int scanLen, double lb, sdgpu_cut *cut, double *pi_ratio, int *dualStableFlag) {
	refCtx *c = (refCtx *) vc;
	oneCut *rc;
	bool stable = dualStableFlag ? (*dualStableFlag != 0) : false;
	int i;
	config.DUAL_STABILITY = dualStability; config.PI_EVAL_START = piEvalStart; config.PI_CYCLE = piCycle;
	config.SCAN_LEN = scanLen;
	rc = SDCut(&c->num, &c->coord, c->basis, c->sigma, c->delta, c->omega, (dVector) Xvect, numSamples, &stable, pi_ratio, lb);
	if (rc == NULL) return SDGPU_NONE;
	cut->alpha = rc->alpha; cut->omegaCnt = rc->omegaCnt; cut->numSamples = rc->numSamples;
	for (i = 0; i <= c->num.prevCols; i++) cut->beta[i] = rc->beta[i];
	if (cut->iStar) for (i = 0; i < rc->omegaCnt; i++) cut->iStar[i] = rc->iStar[i];
	if (dualStableFlag) *dualStableFlag = stable;
	freeOneCut(rc);
	return 0;
}

/* Same shape as sdgpu_sd_cut.  The reference only exposes cummOld/cummAll as their ratio, so on return
 * cut->cummOld holds that ratio and cut->cummAll is 1.0 (0/0 stays NaN as in cuts.c:172). */
int sdref_sd_cut(void *vc, const double *Xvect, int numSamples, int pi_eval_flag, double lb, sdgpu_cut *cut) {
	double ratio[1] = { 0.0 };
	int st;
	/* SCAN_LEN = 1: slot numSamples % 1 == 0; "numSamples - start > SCAN_LEN" would call calcVariance over one
	 * element (a no-op loop), harmless. */
	st = sdref_sd_cut_cfg(vc, Xvect, numSamples, pi_eval_flag != 0, -1, 1, 1, lb, cut, ratio, NULL);
	cut->cummOld = pi_eval_flag ? ratio[0] : 0.0;
	cut->cummAll = pi_eval_flag ? 1.0 : 0.0;
	return st;
}

double sdref_calc_variance(double *x, int scanLen) {
	config.SCAN_LEN = scanLen;
	return calcVariance(x, NULL, NULL, 0);
}

int sdref_cut_heights(void *vc, int n, const double *alpha, const double *beta, const int32_t *numSamples,
		const double *alphaIncumb, int currIter, const double *xk, double lb, double *height, double *etaCoef, double *rhs) {
	refCtx *c = (refCtx *) vc;
	cutsType *cuts = newCuts(n > 0 ? n : 1);
	int i, j, best = SDGPU_NONE, n1 = c->num.prevCols;
	double Sm = -INF;
	for (i = 0; i < n; i++) {
		cuts->vals[i] = newCut(n1, 0, numSamples[i]);
		cuts->vals[i]->alpha = alpha[i];
		for (j = 0; j <= n1; j++) cuts->vals[i]->beta[j] = beta[(size_t) i * (n1 + 1) + j];
		cuts->cnt++;
	}
	for (i = 0; i < n; i++) {
		double ht = cutHeight(cuts->vals[i], currIter, (dVector) xk, n1, lb);
		if (height) height[i] = ht;
		if (Sm < ht) { Sm = ht; best = i; }     /* order of maxCutHeight cuts.c:201-206 */
		/* the two aging formulas are inline expressions in master.c:152 and master.c:174 (CPLEX calls around them) */
		if (etaCoef) etaCoef[i] = (double) (currIter) / (double) numSamples[i];
		if (rhs) rhs[i] = (alphaIncumb ? alphaIncumb[i] : 0.0) + ((double) currIter / (double) numSamples[i] - 1) * lb;
	}
	if (n > 0) {
		double m = maxCutHeight(cuts, currIter, (dVector) xk, n1, lb);
		if (best >= 0 && height && m != height[best]) best = SDGPU_ERR;
	}
	freeCutsType(cuts, false);
	return best;
}

int sdref_reform_cut(void *vc, const int32_t *iStar, int omegaCnt, const int32_t *observ, int k, int lbType, int lb,
		double *alpha, double *beta) {
	refCtx *c = (refCtx *) vc;
	cutsType *g = newCuts(1);
	int i;
	g->vals[0] = newCut(c->num.prevCols, omegaCnt, k);
	for (i = 0; i < omegaCnt; i++) g->vals[0]->iStar[i] = iStar[i];
	g->cnt = 1;
	reformCuts(c->basis, c->sigma, c->delta, c->omega, &c->num, &c->coord, g, (int *) observ, k, lbType, lb, c->num.prevCols);
	*alpha = g->vals[0]->alpha;
	for (i = 0; i <= c->num.prevCols; i++) beta[i] = g->vals[0]->beta[i];
	freeCutsType(g, false);
	return 0;
}

/* the bootstrap loop of fullTest (optimal.c:96-103): reformCuts over all cuts at once, once per replication */
int sdref_reform_cuts_batch(void *vc, int nCuts, const int32_t *iStar, int istarStride, const int32_t *omegaCnt,
		int nReps, const int32_t *observ, int k, int lbType, int lb, double *alpha, double *beta) {
	refCtx *c = (refCtx *) vc;
	cutsType *g = newCuts(nCuts);
	int i, j, r, n1 = c->num.prevCols;
	for (i = 0; i < nCuts; i++) {
		g->vals[i] = newCut(n1, omegaCnt[i], k);
		for (j = 0; j < omegaCnt[i]; j++) g->vals[i]->iStar[j] = iStar[(size_t) i * istarStride + j];
		g->cnt++;
	}
	for (r = 0; r < nReps; r++) {
		reformCuts(c->basis, c->sigma, c->delta, c->omega, &c->num, &c->coord, g, (int *) (observ + (size_t) r * k), k, lbType, lb, n1);
		for (i = 0; i < nCuts; i++) {
			alpha[(size_t) r * nCuts + i] = g->vals[i]->alpha;
			for (j = 0; j <= n1; j++) beta[((size_t) r * nCuts + i) * (n1 + 1) + j] = g->vals[i]->beta[j];
		}
	}
	freeCutsType(g, false);
	return 0;
}

int sdref_get_omega(void *vc, int idx, double *vals, int *weight) {
	refCtx *c = (refCtx *) vc;
	int i;
	if (idx < 0 || idx >= c->omega->cnt) return SDGPU_ERR;
	if (vals) for (i = 1; i <= c->num.numRV; i++) vals[i] = c->omega->vals[idx][i];
	if (weight) *weight = c->omega->weights[idx];
	return 0;
}

int sdref_get_lambda(void *vc, int idx, double *vals) {
	refCtx *c = (refCtx *) vc;
	int i;
	if (idx < 0 || idx >= c->lambda->cnt) return SDGPU_ERR;
	for (i = 1; i <= c->num.rvRowCnt; i++) vals[i] = c->lambda->vals[idx][i];
	return 0;
}

int sdref_get_sigma(void *vc, int idx, double *pib, double *piC, int *lambdaIdx, int *ck) {
	refCtx *c = (refCtx *) vc;
	int i;
	if (idx < 0 || idx >= c->sigma->cnt) return SDGPU_ERR;
	if (pib) *pib = c->sigma->vals[idx].pib;
	if (piC) for (i = 1; i <= c->num.cntCcols; i++) piC[i] = c->sigma->vals[idx].piC[i];
	if (lambdaIdx) *lambdaIdx = c->sigma->lambdaIdx[idx];
	if (ck) *ck = c->sigma->ck[idx];
	return 0;
}

int sdref_get_delta(void *vc, int lambdaIdx, int obsIdx, double *pib, double *piC) {
	refCtx *c = (refCtx *) vc;
	int i;
	if (lambdaIdx < 0 || lambdaIdx >= c->lambda->cnt || obsIdx < 0 || obsIdx >= c->omega->cnt) return SDGPU_ERR;
	if (pib) *pib = c->delta->vals[lambdaIdx][obsIdx].pib;
	if (piC && c->num.rvCOmCnt > 0) for (i = 1; i <= c->num.rvCOmCnt; i++) piC[i] = c->delta->vals[lambdaIdx][obsIdx].piC[i];
	return 0;
}

/* ---- link-time stubs for host functions cuts.c / optimal.c name but the oracle never reaches -------- */
#define NOT_IN_ORACLE(name) do { fprintf(stderr, "sdref :: %s() is host/CPLEX code outside the oracle\n", name); abort(); } while (0)
int solveSubprob(probType *prob, oneProblem *subproblem, dVector Xvect, basisType *basis, lambdaType *lambda, sigmaType *sigma,
		deltaType *delta, int deltaRowLength, omegaType *omega, int omegaIdx, bool *newOmegaFlag, int currentIter, double TOLERANCE,
		bool *subFeasFlag, bool *newBasisFlag, double *subprobTime, double *argmaxTime) { NOT_IN_ORACLE("solveSubprob"); return 1; }
int solveQPMaster(numType *num, sparseVector *dBar, cellType *cell, double lb) { NOT_IN_ORACLE("solveQPMaster"); return 1; }
int addCut2Master(oneProblem *master, oneCut *cut, dVector vectX, int lenX) {
	(void) master; (void) vectX; (void) lenX;
	if (!g_recordAdds) NOT_IN_ORACLE("addCut2Master");
	if (g_addedCnt == g_addedCap) { g_addedCap = g_addedCap ? 2 * g_addedCap : 64; g_addedCuts = (oneCut **) realloc(g_addedCuts, (size_t) g_addedCap * sizeof(oneCut *)); }
	g_addedCuts[g_addedCnt++] = cut;
	return 0;
}
int replaceIncumbent(probType *prob, cellType *cell, double candidEst) { NOT_IN_ORACLE("replaceIncumbent"); return 1; }
int changeQPproximal(LPptr lp, int numCols, double sigma) { NOT_IN_ORACLE("changeQPproximal"); return 1; }
