/*
 * integration/sdgpu_hooks.c -- the host-side patch for SMU-SODA/stochasticDecomposition (twoSD_src/).
 *
 * Drop this file into twoSD_src/, add `sdgpu_ctx *gpu;` to cellType (twoSD.h:132-136, beside lambda/sigma/delta),
 * link with -lsdgpu, and make the four one-line substitutions listed in INTEGRATION.md.  Everything that talks to
 * CPLEX stays exactly where it is (newBasis / calcBasis / decomposeDualSolution / checkBasisFeasibility in
 * randCost.c, solveSubprob in subprob.c, the master in master.c); SMPS input, config.sd and the output files are
 * untouched.  What moves to the GPU is the table arithmetic of stocUpdate.c (calcOmega, calcLambda, calcSigma,
 * calcDelta) and the numeric core of cuts.c (computeIstar + the SDCut accumulation).
 *
 * The functions below keep the reference's names with a _gpu suffix, the reference's argument meaning and its
 * error convention (0 / 1 / -1 / NULL after errMsg).  Line references are to the reference's sources.
 *
 * This file is compile-checked against the reference headers (tests/test_integration_compiles.py); it cannot be
 * linked or run here because CPLEX and the spAlgorithms utilities are absent.
 */
#include "twoSD.h"
#include "sdgpu.h"

extern configType config;

/* ---- setup.c:136-144: allocate the device tables with the reference's capacities ------------------------- */
sdgpu_ctx *newGpuTables(probType *prob, int device) {
	sdgpu_problem p;
	sdgpu_caps caps;
	sdgpu_ctx *ctx = NULL;
	numType *num = prob->num;
	coordType *coord = prob->coord;
	int length;

	p.num.rows = num->rows;         p.num.cols = num->cols;         p.num.prevCols = num->prevCols;
	p.num.cntCcols = num->cntCcols; p.num.rvRowCnt = num->rvRowCnt; p.num.rvbOmCnt = num->rvbOmCnt;
	p.num.rvCOmCnt = num->rvCOmCnt; p.num.rvdOmCnt = num->rvdOmCnt; p.num.numRV = num->numRV;
	p.coord.CCols = coord->CCols;         p.coord.rvRows = coord->rvRows;       p.coord.rvbOmRows = coord->rvbOmRows;
	p.coord.rvCOmCols = coord->rvCOmCols; p.coord.rvCOmRows = coord->rvCOmRows; p.coord.rvCols = coord->rvCols;
	p.coord.rvOffset[0] = coord->rvOffset[0]; p.coord.rvOffset[1] = coord->rvOffset[1]; p.coord.rvOffset[2] = coord->rvOffset[2];
	p.bBar.cnt = prob->bBar->cnt; p.bBar.col = prob->bBar->col; p.bBar.val = prob->bBar->val;
	p.Cbar.cnt = prob->Cbar->cnt; p.Cbar.col = prob->Cbar->col; p.Cbar.row = prob->Cbar->row; p.Cbar.val = prob->Cbar->val;

	if ( num->rvdOmCnt > 0 )                                                    /* setup.c:136-139 */
		length = num->rvdOmCnt*config.MAX_ITER + config.MAX_ITER / config.TAU + 1;
	else
		length = config.MAX_ITER + config.MAX_ITER / config.TAU + 1;
	caps.maxLambda = length; caps.maxSigma = length;
	caps.maxBasis = 2*config.MAX_ITER + 1;       /* the reference sizes this MAX_ITER although two solves per iteration can each add a basis */
	caps.maxOmega = config.MAX_ITER;             /* setup.c:144, cuts.c:28 */
	caps.maxTerms = 1 + num->rvdOmCnt;

	if ( sdgpu_create(&p, &caps, device, &ctx) ) {
		errMsg("allocation", "newGpuTables", sdgpu_last_error(), 0);
		return NULL;
	}
	return ctx;
}//END newGpuTables()

/* ---- algo.c:152 ------------------------------------------------------------------------------------------ */
int calcOmega_gpu(sdgpu_ctx *gpu, dVector observ, bool *newOmegaFlag, double TOLERANCE) {
	int flag = 0, idx;

	idx = sdgpu_calc_omega(gpu, observ, TOLERANCE, &flag);
	if ( idx < 0 )
		errMsg("algorithm", "calcOmega_gpu", sdgpu_last_error(), 0);
	(*newOmegaFlag) = (flag != 0);
	return idx;
}//END calcOmega_gpu()

/* ---- stocUpdate.c:14-133 ------------------------------------------------------------------------------------
 * basis (the host list of oneBasis records) is still maintained: CPLEX-side code reads cCode/rCode, phi, piDet,
 * gBar, psi from it.  lambda, sigma, delta are gone from the host: their arithmetic happens in the library. */
int stochasticUpdates_gpu(probType *prob, LPptr lp, basisType *basis, sdgpu_ctx *gpu, omegaType *omega, int omegaIdx,
		bool newOmegaFlag, int currentIter, double TOLERANCE, bool *newBasisFlag, bool subFeasFlag) {
	oneBasis *B;
	sparseVector dOmega;
	unsigned char *flags;
	int 	cnt, idx, lambdaIdx, newLambda, newSigma, isNew = 1;
	bool	retainBasis;

	dOmega.cnt = prob->num->rvdOmCnt; dOmega.col = prob->coord->rvdOmCols;

	if ( newOmegaFlag ) {
		/* stocUpdate.c:25 -- new column of delta */
		if ( sdgpu_calc_delta(gpu, 1, omegaIdx) ) {
			errMsg("algorithm", "stochasticUpdates", sdgpu_last_error(), 0);
			return -1;
		}
		/* stocUpdate.c:28-30 -- feasibility of the stored bases at the new observation (host: needs phi/psi) */
		if ( prob->num->rvdOmCnt > 0 && basis->cnt > 0 ) {
			flags = (unsigned char *) arr_alloc(basis->cnt, unsigned char);
			dOmega.val = prob->coord->rvOffset[2]+omega->vals[omegaIdx];
			for ( cnt = 0; cnt < basis->cnt; cnt++ ) {
				basis->obsFeasible[cnt][omegaIdx] = checkBasisFeasibility(basis->vals[cnt], dOmega, prob->sp->senx, prob->num->cols, prob->num->rows, TOLERANCE);
				flags[cnt] = basis->obsFeasible[cnt][omegaIdx];
			}
			sdgpu_basis_set_obs_feasible_col(gpu, omegaIdx, flags);
			mem_free(flags);
		}
	}

	if ( (B = newBasis(lp, prob->num->cols, prob->num->rows, currentIter, subFeasFlag)) == NULL ) {
		errMsg("algorithm", "stochasticUpdates", "failed to create a new basis type structure", 0);
		return -1;
	}

	/* stocUpdate.c:39-53 -- basis seen before? (host: compares CPLEX basis codes) */
	for ( cnt = 0; cnt < basis->cnt; cnt++ ) {
		if ( B->feasFlag ) {
			if ( equalLongIntvec(B->cCode, basis->vals[cnt]->cCode, basis->cCodeLen) && equalLongIntvec(B->rCode,
					basis->vals[cnt]->rCode, basis->rCodeLen) ) {
				freeOneBasis(B);
				basis->vals[cnt]->weight++;
				(*newBasisFlag) = false;
				return cnt;
			}
		}
	}

	/* stocUpdate.c:55-75 -- CPLEX basis inverse / duals (host) */
	if ( B->feasFlag ) {
		if ( prob->num->rvdOmCnt > 0 )
			calcBasis(lp, prob->num, prob->coord, prob->dBar, B, basis->basisDim);
		if ( decomposeDualSolution(lp, B, omega->vals[omegaIdx]+prob->coord->rvOffset[2], prob->num->rows) ) {
			errMsg("algorithm", "stochasticUpdates", "failed to decompose the dual solution", 0);
			return -1;
		}
	}
	else {
		if ( !(B->piDet = (dVector) arr_alloc(prob->num->rows+1, double)) )
			errMsg("allocation", "decomposeDualSolution", "piS", 0);
		if ( getDual(lp, B->piDet, prob->num->rows) ) {
			errMsg("algorithm", "stochasticUpdates", "failed to get the dual", 0);
			return 1;
		}
	}

	/* stocUpdate.c:78-85 -- calcLambda + calcSigma + calcDelta(row) for the deterministic dual, one device round trip */
	B->sigmaIdx = (iVector) mem_realloc(B->sigmaIdx, (B->phiLength+1)*sizeof(int));
	if ( sdgpu_update_dual(gpu, B->piDet, B->mubBar, currentIter, TOLERANCE, &lambdaIdx, &newLambda, &B->sigmaIdx[0], &newSigma) ) {
		errMsg("algorithm", "stochasticUpdates", sdgpu_last_error(), 0);
		return -1;
	}
	retainBasis = (newSigma != 0);

	/* stocUpdate.c:88-99 -- the same triple for every phi column (random cost only) */
	for (cnt = 0; cnt < B->phiLength; cnt++ ) {
		if ( sdgpu_update_dual(gpu, B->phi[cnt], 0, currentIter, TOLERANCE, &lambdaIdx, &newLambda, &B->sigmaIdx[cnt+1], &newSigma) ) {
			errMsg("algorithm", "stochasticUpdates", sdgpu_last_error(), 0);
			return -1;
		}
		retainBasis = (retainBasis || (newSigma != 0));
	}

	/* stocUpdate.c:101-131 -- dedup by sigma list, else append; the library mirrors the host list index for index */
	idx = sdgpu_basis_find_or_append(gpu, retainBasis, omegaIdx, B->ck, B->feasFlag, B->phiLength, B->sigmaIdx, B->omegaIdx, &isNew);
	if ( idx < 0 ) {
		errMsg("algorithm", "stochasticUpdates", sdgpu_last_error(), 0);
		return -1;
	}
	if ( !isNew ) {
		freeOneBasis(B);
		basis->vals[idx]->weight++;
		(*newBasisFlag) = false;
		return idx;
	}

	basis->vals[basis->cnt] = B;
	if ( B->feasFlag ) {
		if ( !(basis->obsFeasible[basis->cnt] = (bool*) arr_alloc(config.MAX_ITER, bool)) )
			errMsg("allocation", "stochasticUpdates", "basis->obsFeasibility[n]", 0);
		flags = (unsigned char *) arr_alloc(omega->cnt + 1, unsigned char);
		for ( cnt = 0; cnt < omega->cnt; cnt++ ) {
			dOmega.val = prob->coord->rvOffset[2]+omega->vals[cnt];
			basis->obsFeasible[basis->cnt][cnt] = checkBasisFeasibility(B, dOmega, prob->sp->senx, prob->num->cols, prob->num->rows, TOLERANCE);
			flags[cnt] = basis->obsFeasible[basis->cnt][cnt];
		}
		if ( prob->num->rvdOmCnt > 0 )
			sdgpu_basis_set_obs_feasible_row(gpu, idx, flags);
		mem_free(flags);
	}
	else
		basis->obsFeasible[basis->cnt] = NULL;

	return basis->cnt++;
}//END stochasticUpdates_gpu()

/* ---- cuts.c:91-194 ---------------------------------------------------------------------------------------- */
oneCut *SDCut_gpu(numType *num, sdgpu_ctx *gpu, int omegaCnt, dVector Xvect, int numSamples, bool *dualStableFlag,
		dVector pi_ratio, double lb) {
	oneCut *cut;
	sdgpu_cut res;
	bool pi_eval_flag = false;
	int status;

	cut = newCut(num->prevCols, omegaCnt, numSamples);                          /* cuts.c:100 */

	if (config.DUAL_STABILITY && numSamples > config.PI_EVAL_START && !(numSamples % config.PI_CYCLE))   /* cuts.c:112 */
		pi_eval_flag = true;

	res.beta = cut->beta; res.iStar = cut->iStar;
	status = sdgpu_sd_cut(gpu, Xvect, numSamples, pi_eval_flag, lb, &res);      /* cuts.c:105-169 and :184-188 */
	if ( status != 0 ) {
		errMsg("algorithm", "SDCut", status == SDGPU_NONE ? "failed to identify maximal Pi for an observation" : sdgpu_last_error(), 0);
		freeOneCut(cut);
		return NULL;
	}
	cut->alpha = res.alpha;

	if (pi_eval_flag == true)                                                   /* cuts.c:171-182 */
		*dualStableFlag = sdgpu_dual_stability(res.cummOld, res.cummAll, numSamples, config.PI_EVAL_START, config.SCAN_LEN, pi_ratio) != 0;

	return cut;
}//END SDCut_gpu()

/* ---- soln.c:24 / cuts.c:197-209: highest cut at xk, on the device ------------------------------------------ */
double maxCutHeight_gpu(sdgpu_ctx *gpu, cutsType *cuts, int currIter, dVector xk, int betaLen, double lb) {
	dVector alpha, beta, height;
	iVector numSamples;
	double Sm = -INF;
	int cnt, c, best;

	if ( cuts->cnt == 0 )
		return Sm;
	alpha = (dVector) arr_alloc(cuts->cnt, double); height = (dVector) arr_alloc(cuts->cnt, double);
	beta = (dVector) arr_alloc(cuts->cnt*(betaLen+1), double); numSamples = (iVector) arr_alloc(cuts->cnt, int);
	for (cnt = 0; cnt < cuts->cnt; cnt++) {
		alpha[cnt] = cuts->vals[cnt]->alpha; numSamples[cnt] = cuts->vals[cnt]->numSamples;
		for (c = 0; c <= betaLen; c++)
			beta[cnt*(betaLen+1)+c] = cuts->vals[cnt]->beta[c];
	}
	best = sdgpu_cut_heights(gpu, cuts->cnt, alpha, beta, numSamples, NULL, currIter, xk, lb, height, NULL, NULL);
	if ( best >= 0 )
		Sm = height[best];
	mem_free(alpha); mem_free(beta); mem_free(height); mem_free(numSamples);
	return Sm;
}//END maxCutHeight_gpu()

/* ---- cuts.c:465-517: feasibility-cut pool update; the gathers run on the device, the pool de-duplication
 * (addCut2Pool, cuts.c:643-655) stays here ---------------------------------------------------------------------- */
int addCut2Pool(cellType *cell, oneCut *cut, int lenX, double lb, typeOfCut type);     /* cuts.c:616 */

static int feasCutsFromDevice(numType *num, cellType *cell, sdgpu_ctx *gpu, int obsFirst, int obsLast, int basisFirst, int basisLast) {
	int n, cnt, c, maxOut = (obsLast - obsFirst) * (basisLast - basisFirst);
	dVector alpha, beta;
	oneCut *cut;

	if ( maxOut <= 0 )
		return 0;
	alpha = (dVector) arr_alloc(maxOut, double); beta = (dVector) arr_alloc(maxOut*(num->prevCols+1), double);
	n = sdgpu_feas_cuts(gpu, obsFirst, obsLast, basisFirst, basisLast, maxOut, alpha, beta);
	for ( cnt = 0; cnt < n; cnt++ ) {
		cut = newCut(num->prevCols, 0, 1);
		cut->alpha = alpha[cnt];
		for ( c = 0; c <= num->prevCols; c++ )
			cut->beta[c] = beta[cnt*(num->prevCols+1)+c];
		addCut2Pool(cell, cut, num->prevCols, 0.0, FEASIBILITY);
	}
	mem_free(alpha); mem_free(beta);
	return n < 0 ? -1 : 0;
}

int updtFeasCutPool_gpu(numType *num, cellType *cell, sdgpu_ctx *gpu) {
	int initCutsCnt = cell->fcutsPool->cnt;

	feasCutsFromDevice(num, cell, gpu, cell->fUpdt[1], cell->omega->cnt, 0, cell->fUpdt[0]);        /* cuts.c:472-490 */
	cell->fUpdt[1] = cell->omega->cnt;
	feasCutsFromDevice(num, cell, gpu, 0, cell->omega->cnt, cell->fUpdt[0], cell->basis->cnt);      /* cuts.c:494-512 */
	cell->fUpdt[0] = cell->basis->cnt;
	return (cell->fcutsPool->cnt - initCutsCnt);
}//END updtFeasCutPool_gpu()

/* ---- randCost.c:202-258 on the device: hand a new basis' CPLEX-derived vectors over once (after calcBasis /
 * decomposeDualSolution), then let the library evaluate obsFeasible for it --------------------------------------- */
int basisFeasibility_gpu(probType *prob, sdgpu_ctx *gpu, oneBasis *B, int basisIdx, iVector cstat, double TOLERANCE) {
	dVector phi = NULL, psiVal = NULL;
	int n, i;

	if ( prob->num->rvdOmCnt == 0 )
		return 0;
	if ( B->phiLength > 0 ) {
		phi = (dVector) arr_alloc(B->phiLength*(prob->num->rows+1), double);
		for ( n = 0; n < B->phiLength; n++ )
			for ( i = 0; i <= prob->num->rows; i++ )
				phi[n*(prob->num->rows+1)+i] = B->phi[n][i];
		psiVal = (dVector) arr_alloc(B->psi->cnt+1, double);
		for ( i = 1; i <= B->psi->cnt; i++ )
			psiVal[i-1] = B->psi->val[i];                                 /* entry order of randCost.c:83-88 */
	}
	if ( sdgpu_basis_set_feas_data(gpu, basisIdx, B->piDet, phi, B->gBar, psiVal, cstat) ||
			sdgpu_check_feasibility_basis(gpu, basisIdx, TOLERANCE, NULL) ) {
		errMsg("algorithm", "basisFeasibility_gpu", sdgpu_last_error(), 0);
		return 1;
	}
	if ( phi ) mem_free(phi);
	if ( psiVal ) mem_free(psiVal);
	return 0;
}//END basisFeasibility_gpu()

/* ---- setup.c:242-246 and :282-286 ---------------------------------------------------------------------------- */
void cleanGpuTables(sdgpu_ctx *gpu) { sdgpu_reset(gpu); }
void freeGpuTables(sdgpu_ctx *gpu)  { sdgpu_destroy(gpu); }
