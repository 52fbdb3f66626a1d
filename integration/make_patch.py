#!/usr/bin/env python
"""Regenerates integration/twoSD_src.patch: the substitutions that wire libsdgpu.so into the reference host, as a unified diff
against /root/reference/twoSD_src (run where the reference sources exist).  The edits are applied to a scratch copy under /tmp;
only the diff is written into the repository.

    python integration/make_patch.py [--keep /tmp/dir]     # --keep leaves the patched tree behind (the oracle build uses it)
"""
import os
import shutil
import subprocess
import sys

REF = "/root/reference/twoSD_src"
HERE = os.path.dirname(os.path.abspath(__file__))


def sub(path, old, new, count=1):
    s = open(path).read()
    assert s.count(old) >= 1, (path, old[:60])
    assert count == 0 or s.count(old) == count, (path, old[:60], s.count(old))
    open(path, "w").write(s.replace(old, new))


def between(path, start, end, new):
    """replace everything from the first `start` up to and including the next `end`"""
    s = open(path).read()
    i = s.index(start)
    j = s.index(end, i) + len(end)
    open(path, "w").write(s[:i] + new + s[j:])


def apply(dst):
    f = lambda n: os.path.join(dst, n)
    # ---- stoc.h / twoSD.h: the tables become one device context ------------------------------------------------------------
    sub(f("stoc.h"), '#include "prob.h"\n', '#include "prob.h"\n#include "sdgpu.h"\n')
    sub(f("stoc.h"), "int solveSubprob(probType *prob, oneProblem *subproblem, dVector Xvect, basisType *basis, lambdaType *lambda, sigmaType *sigma, deltaType *delta, int deltaRowLength,\n",
        "int solveSubprob(probType *prob, oneProblem *subproblem, dVector Xvect, basisType *basis, sdgpu_ctx *gpu,\n")
    sub(f("stoc.h"), "int stochasticUpdates(probType *prob, LPptr spLP, basisType *basis, lambdaType *lambda, sigmaType *sigma, deltaType *delta, int deltaRowLength,\n",
        "int stochasticUpdates(probType *prob, LPptr spLP, basisType *basis, sdgpu_ctx *gpu,\n")
    sub(f("twoSD.h"), "	lambdaType 	*lambda;			/* holds dual solutions corresponding to rows effected by randomness */\n"
                      "	sigmaType 	*sigma;				/* holds $\\pi \\times \\bar{b}$ and $\\pi \\times \\bar{C} $ values */\n"
                      "	deltaType   *delta;				/* calculations based on realization and dual solutions observed */\n",
        "	sdgpu_ctx	*gpu;				/* lambda, sigma and delta live on the GPU (libsdgpu.so) */\n")
    sub(f("twoSD.h"), "oneCut *SDCut(numType *num, coordType *coord, basisType *basis, sigmaType *sigma, deltaType *delta, omegaType *omega, dVector Xvect, int numSamples,\n",
        "oneCut *SDCut(numType *num, sdgpu_ctx *gpu, int omegaCnt, dVector Xvect, int numSamples,\n")
    sub(f("twoSD.h"), "void reformCuts(basisType *basis, sigmaType *sigma, deltaType *delta, omegaType *omega, numType *num, coordType *coord,\n",
        "void reformCuts(sdgpu_ctx *gpu, numType *num,\n")
    sub(f("twoSD.h"), "/* cuts.c */\n",
        "/* sdgpu_hooks.c */\n"
        "sdgpu_ctx *newGpuTables(probType *prob, int device);\n"
        "int calcOmega_gpu(sdgpu_ctx *gpu, dVector observ, bool *newOmegaFlag, double TOLERANCE);\n"
        "int feasCutsFromDevice(numType *num, cellType *cell, int obsFirst, int obsLast, int basisFirst, int basisLast);\n"
        "void cleanGpuTables(sdgpu_ctx *gpu);\n"
        "void freeGpuTables(sdgpu_ctx *gpu);\n\n"
        "/* cuts.c */\n")
    # ---- setup.c: allocation, reset between replications, free ---------------------------------------------------------------
    sub(f("setup.c"), "	cell->lambda = NULL; cell->sigma = NULL; cell->delta = NULL; cell->omega = NULL;\n", "	cell->gpu = NULL; cell->omega = NULL;\n")
    sub(f("setup.c"), "	cell->basis  = newBasisType(config.MAX_ITER, prob[1]->num->cols, prob[1]->num->rows, WORDLENGTH);\n"
                      "	cell->lambda = newLambda(length, 0, prob[1]->num->rvRowCnt);\n"
                      "	cell->sigma  = newSigma(length, prob[1]->num->cntCcols, 0);\n"
                      "	cell->delta  = newDelta(length);\n",
        "	cell->basis  = newBasisType(2*config.MAX_ITER+1, prob[1]->num->cols, prob[1]->num->rows, WORDLENGTH);\n"
        "	if ( (cell->gpu = newGpuTables(prob[1], 0)) == NULL ) {\n"
        "		errMsg(\"setup\", \"newCell\", \"failed to create the device tables\", 0);\n"
        "		return NULL;\n"
        "	}\n")
    sub(f("setup.c"), "	if (cell->delta) freeDeltaType(cell->delta, cell->lambda->cnt, cell->omega->cnt, true);\n"
                      "	if (cell->lambda) freeLambdaType(cell->lambda, true);\n"
                      "	if (cell->sigma) freeSigmaType(cell->sigma, true);\n",
        "	if (cell->gpu) cleanGpuTables(cell->gpu);\n")
    sub(f("setup.c"), "		if (cell->delta) freeDeltaType(cell->delta, cell->lambda->cnt, cell->omega->cnt, false);\n", "		if (cell->gpu) freeGpuTables(cell->gpu);\n")
    sub(f("setup.c"), "		if (cell->lambda) freeLambdaType(cell->lambda, false);\n		if (cell->sigma) freeSigmaType(cell->sigma, false);\n", "")
    # ---- algo.c: the observation goes to the host list (computeRHS reads it) and to the device --------------------------------
    sub(f("algo.c"), "		omegaIdx = calcOmega(observ, 0, prob[1]->num->numRV, cell->omega, &newOmegaFlag, config.TOLERANCE);\n",
        "		omegaIdx = calcOmega(observ, 0, prob[1]->num->numRV, cell->omega, &newOmegaFlag, config.TOLERANCE);\n"
        "		if ( calcOmega_gpu(cell->gpu, observ, &newOmegaFlag, config.TOLERANCE) != omegaIdx ) {\n"
        "			errMsg(\"algorithm\", \"solveCell\", \"host and device observation lists disagree\", 0);\n"
        "			goto TERMINATE;\n"
        "		}\n")
    # ---- subprob.c -------------------------------------------------------------------------------------------------------------
    sub(f("subprob.c"), "int solveSubprob(probType *prob, oneProblem *subproblem, dVector Xvect, basisType *basis, lambdaType *lambda, sigmaType *sigma, deltaType *delta, int deltaRowLength,\n",
        "int solveSubprob(probType *prob, oneProblem *subproblem, dVector Xvect, basisType *basis, sdgpu_ctx *gpu,\n")
    sub(f("subprob.c"), "		status = stochasticUpdates(prob, subproblem->lp, basis, lambda, sigma, delta, deltaRowLength,\n",
        "		status = stochasticUpdates(prob, subproblem->lp, basis, gpu,\n")
    between(f("subprob.c"), "#ifdef STOCH_CHECK\n		obj = sigma->vals[status].pib", "#endif\n", "")
    # ---- stocUpdate.c: stochasticUpdates keeps the CPLEX-side bookkeeping, the table arithmetic becomes library calls ----------
    p = f("stocUpdate.c")
    sub(p, "int stochasticUpdates(probType *prob, LPptr lp, basisType *basis, lambdaType *lambda, sigmaType *sigma, deltaType *delta, int deltaRowLength,\n",
        "int stochasticUpdates(probType *prob, LPptr lp, basisType *basis, sdgpu_ctx *gpu,\n")
    sub(p, "	int 	cnt, lambdaIdx;\n	bool	newSigmaFlag, newLambdaFlag, retainBasis;\n",
        "	int 	cnt, idx, lambdaIdx, newSigmaFlag, newLambdaFlag, isNew = 1;\n	bool	retainBasis;\n	unsigned char *flags;\n")
    sub(p, "		calcDelta(prob->num, prob->coord, lambda, delta, deltaRowLength, omega, newOmegaFlag, omegaIdx);\n",
        "		/* (the new column of delta is computed on the device by the launch that scans lambda for the first dual below, or on its\n"
        "		 * own if the basis turns out to be an old one) */\n")
    sub(p, "				(*newBasisFlag) = false;\n#if defined (STOCH_CHECK)\n",
        "				(*newBasisFlag) = false;\n"
        "				if ( newOmegaFlag && sdgpu_calc_delta(gpu, 1, omegaIdx) ) {\n"
        "					errMsg(\"algorithm\", \"stochasticUpdates\", sdgpu_last_error(), 0);\n"
        "					return -1;\n"
        "				}\n"
        "#if defined (STOCH_CHECK)\n")
    sub(p, "		for ( cnt = 0; cnt < basis->cnt; cnt++ )\n"
           "			basis->obsFeasible[cnt][omegaIdx] = checkBasisFeasibility(basis->vals[cnt], dOmega, prob->sp->senx, prob->num->cols, prob->num->rows, TOLERANCE);\n",
        "		for ( cnt = 0; cnt < basis->cnt; cnt++ )\n"
        "			basis->obsFeasible[cnt][omegaIdx] = checkBasisFeasibility(basis->vals[cnt], dOmega, prob->sp->senx, prob->num->cols, prob->num->rows, TOLERANCE);\n"
        "		if ( prob->num->rvdOmCnt > 0 && basis->cnt > 0 ) {\n"
        "			flags = (unsigned char *) arr_alloc(basis->cnt, unsigned char);\n"
        "			for ( cnt = 0; cnt < basis->cnt; cnt++ )\n"
        "				flags[cnt] = basis->obsFeasible[cnt][omegaIdx];\n"
        "			sdgpu_basis_set_obs_feasible_col(gpu, omegaIdx, flags);\n"
        "			mem_free(flags);\n"
        "		}\n")
    between(p, "	/* Elements of deterministic component of dual solution corresponding to rows with random elements in them */\n",
            "	retainBasis = newSigmaFlag;\n",
            "	/* calcDelta(column, if the observation is new) + calcLambda + calcSigma + calcDelta(row) for the deterministic component of the\n"
            "	 * dual solution: one device round trip */\n"
            "	B->sigmaIdx = (iVector) mem_realloc(B->sigmaIdx, (B->phiLength+1)*sizeof(int));\n"
            "	if ( sdgpu_update_dual_col(gpu, newOmegaFlag ? omegaIdx : -1, B->piDet, B->mubBar, currentIter, TOLERANCE, &lambdaIdx, &newLambdaFlag,\n"
            "			&B->sigmaIdx[0], &newSigmaFlag) ) {\n"
            "		errMsg(\"algorithm\", \"stochasticUpdates\", sdgpu_last_error(), 0);\n"
            "		return -1;\n"
            "	}\n\n"
            "	retainBasis = newSigmaFlag;\n")
    between(p, "		/* Elements of basis column corresponding to rows with random elements in them */\n",
            "		retainBasis = (retainBasis || newSigmaFlag);\n",
            "		/* the same triple for the basis column */\n"
            "		if ( sdgpu_update_dual(gpu, B->phi[cnt], 0, currentIter, TOLERANCE, &lambdaIdx, &newLambdaFlag, &B->sigmaIdx[cnt+1], &newSigmaFlag) ) {\n"
            "			errMsg(\"algorithm\", \"stochasticUpdates\", sdgpu_last_error(), 0);\n"
            "			return -1;\n"
            "		}\n"
            "		retainBasis = (retainBasis || newSigmaFlag);\n")
    between(p, "	if ( !retainBasis ) {\n		/* All the sigmas computed were encountered before */\n", "	/* Add the basis to the structure */\n",
            "	/* All the sigmas computed were encountered before: the library looks the basis up in its records (kept index for index\n"
            "	 * with this list), else appends it. */\n"
            "	idx = sdgpu_basis_find_or_append(gpu, retainBasis, omegaIdx, B->ck, B->feasFlag, B->phiLength, B->sigmaIdx, B->omegaIdx, &isNew);\n"
            "	if ( idx < 0 ) {\n"
            "		errMsg(\"algorithm\", \"stochasticUpdates\", sdgpu_last_error(), 0);\n"
            "		return -1;\n"
            "	}\n"
            "	if ( !isNew ) {\n"
            "		/* The basis was encountered before */\n"
            "		freeOneBasis(B);\n"
            "		basis->vals[idx]->weight++;\n"
            "		(*newBasisFlag) = false;\n"
            "		return idx;\n"
            "	}\n\n"
            "	/* Add the basis to the structure */\n")
    sub(p, "		if ( !(basis->obsFeasible[basis->cnt] = (bool*) arr_alloc(deltaRowLength, bool)) )\n",
        "		if ( !(basis->obsFeasible[basis->cnt] = (bool*) arr_alloc(config.MAX_ITER, bool)) )\n")
    sub(p, "			basis->obsFeasible[basis->cnt][cnt] = checkBasisFeasibility(B, dOmega, prob->sp->senx, prob->num->cols, prob->num->rows, TOLERANCE);\n		}\n",
        "			basis->obsFeasible[basis->cnt][cnt] = checkBasisFeasibility(B, dOmega, prob->sp->senx, prob->num->cols, prob->num->rows, TOLERANCE);\n		}\n"
        "		if ( prob->num->rvdOmCnt > 0 && omega->cnt > 0 ) {\n"
        "			flags = (unsigned char *) arr_alloc(omega->cnt, unsigned char);\n"
        "			for ( cnt = 0; cnt < omega->cnt; cnt++ )\n"
        "				flags[cnt] = basis->obsFeasible[basis->cnt][cnt];\n"
        "			sdgpu_basis_set_obs_feasible_row(gpu, idx, flags);\n"
        "			mem_free(flags);\n"
        "		}\n")
    sub(p, '#include "stoc.h"\n', '#include "stoc.h"\n#include "twoSD.h"\n\nextern configType config;\n')
    # ---- cuts.c: the call sites, SDCut's numeric core and the gathers of updtFeasCutPool -----------------------------------------
    p = f("cuts.c")
    sub(p, "	if ( solveSubprob(prob[1], cell->subprob, Xvect, cell->basis, cell->lambda, cell->sigma, cell->delta, config.MAX_ITER,\n			cell->omega, omegaIdx, newOmegaFlag, cell->k, config.TOLERANCE, &cell->spFeasFlag, &newBasisFlag,\n",
        "	if ( solveSubprob(prob[1], cell->subprob, Xvect, cell->basis, cell->gpu,\n			cell->omega, omegaIdx, newOmegaFlag, cell->k, config.TOLERANCE, &cell->spFeasFlag, &newBasisFlag,\n")
    sub(p, "	cut = SDCut(prob[1]->num, prob[1]->coord, cell->basis, cell->sigma, cell->delta, cell->omega, Xvect, cell->k, &cell->dualStableFlag, cell->pi_ratio, cell->lb);\n",
        "	cut = SDCut(prob[1]->num, cell->gpu, cell->omega->cnt, Xvect, cell->k, &cell->dualStableFlag, cell->pi_ratio, cell->lb);\n")
    sub(p, "		if ( solveSubprob(prob[1], cell->subprob, Xvect, cell->basis, cell->lambda, cell->sigma, cell->delta, config.MAX_ITER,\n				cell->omega, cnt, newOmegaFlag,",
        "		if ( solveSubprob(prob[1], cell->subprob, Xvect, cell->basis, cell->gpu,\n				cell->omega, cnt, newOmegaFlag,")
    sub(p, "		if ( solveSubprob(prob[1], cell->subprob->lp, cell->candidX, cell->basis, cell->lambda, cell->sigma, cell->delta, config.MAX_ITER,\n",
        "		if ( solveSubprob(prob[1], cell->subprob->lp, cell->candidX, cell->basis, cell->gpu,\n")
    between(p, "oneCut *SDCut(numType *num, coordType *coord, basisType *basis,", "}//END SDCut\n",
            "oneCut *SDCut(numType *num, sdgpu_ctx *gpu, int omegaCnt, dVector Xvect, int numSamples,\n"
            "		bool *dualStableFlag, dVector pi_ratio, double lb) {\n"
            "	oneCut *cut;\n"
            "	sdgpu_cut res;\n"
            "	bool    pi_eval_flag = false;\n"
            "	int     status;\n\n"
            "	/* allocate memory to hold a new cut */\n"
            "	cut = newCut(num->prevCols, omegaCnt, numSamples);\n\n"
            "	/* Calculate pi_eval_flag to determine the way of computing argmax */\n"
            "	if (config.DUAL_STABILITY && numSamples > config.PI_EVAL_START && !(numSamples % config.PI_CYCLE))\n"
            "		pi_eval_flag = true;\n\n"
            "	/* For each observation, find the Pi which maximizes height at X and average the maximizers: on the device */\n"
            "	res.beta = cut->beta; res.iStar = cut->iStar;\n"
            "	status = sdgpu_sd_cut(gpu, Xvect, numSamples, pi_eval_flag, lb, &res);\n"
            "	if ( status != 0 ) {\n"
            "		errMsg(\"algorithm\", \"SDCut\", status == SDGPU_NONE ? \"failed to identify maximal Pi for an observation\" : sdgpu_last_error(), 0);\n"
            "		return NULL;\n"
            "	}\n"
            "	cut->alpha = res.alpha;\n\n"
            "	if (pi_eval_flag == true)\n"
            "		*dualStableFlag = sdgpu_dual_stability(res.cummOld, res.cummAll, numSamples, config.PI_EVAL_START, config.SCAN_LEN, pi_ratio) != 0;\n\n"
            "	return cut;\n"
            "}//END SDCut\n")
    between(p, "	oneCut	*cut;\n	int		idx, obs, c, initCutsCnt, sigmaIdx, lambdaIdx;\n", "	cell->fUpdt[0] = cell->basis->cnt;\n",
            "	int		initCutsCnt;\n\n"
            "	initCutsCnt = cell->fcutsPool->cnt;\n\n"
            "	/* Update computations with respect to the newly discovered observations and all the elements of the stochastic structures. */\n"
            "	feasCutsFromDevice(num, cell, cell->fUpdt[1], cell->omega->cnt, 0, cell->fUpdt[0]);\n"
            "	cell->fUpdt[1] = cell->omega->cnt;\n\n"
            "	/* Update computations with respect to the newly discovered stochastic structures and all the observations discovered until now. */\n"
            "	feasCutsFromDevice(num, cell, 0, cell->omega->cnt, cell->fUpdt[0], cell->basis->cnt);\n"
            "	cell->fUpdt[0] = cell->basis->cnt;\n")
    # ---- optimal.c: reformCuts gathers on the device -----------------------------------------------------------------------------
    p = f("optimal.c")
    sub(p, "		reformCuts(cell->basis, cell->sigma, cell->delta, cell->omega, prob[1]->num, prob[1]->coord,\n", "		reformCuts(cell->gpu, prob[1]->num,\n")
    between(p, "void reformCuts(basisType *basis, sigmaType *sigma, deltaType *delta, omegaType *omega, numType *num, coordType *coord,\n", "}//END reform_cuts\n",
            "void reformCuts(sdgpu_ctx *gpu, numType *num,\n"
            "		cutsType *gCuts, int *observ, int k, int lbType, int lb, int lenX) {\n"
            "	int cnt;\n\n"
            "	/* Loop through all the cuts and reform them: the gathers over the stored istar's run on the device */\n"
            "	for (cnt = 0; cnt < gCuts->cnt; cnt++) {\n"
            "		if ( sdgpu_reform_cut(gpu, gCuts->vals[cnt]->iStar, gCuts->vals[cnt]->omegaCnt, observ, k, lbType == NONTRIVIAL, lb,\n"
            "				&gCuts->vals[cnt]->alpha, gCuts->vals[cnt]->beta) )\n"
            "			errMsg(\"algorithm\", \"reformCuts\", sdgpu_last_error(), 0);\n"
            "	}\n\n"
            "}//END reform_cuts\n")


def main():
    keep = sys.argv[sys.argv.index("--keep") + 1] if "--keep" in sys.argv else None
    work = keep or "/tmp/sdgpu_patch_work"
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(os.path.join(work, "a"))
    os.makedirs(os.path.join(work, "b"))
    shutil.copytree(REF, os.path.join(work, "a", "twoSD_src"))
    shutil.copytree(REF, os.path.join(work, "b", "twoSD_src"))
    apply(os.path.join(work, "b", "twoSD_src"))
    out = subprocess.run(["diff", "-u", "-r", "a/twoSD_src", "b/twoSD_src"], cwd=work, capture_output=True, text=True)
    assert out.returncode in (0, 1), out.stderr
    lines = [ln for ln in out.stdout.splitlines(keepends=True) if not ln.startswith("diff -u -r")]
    # drop the timestamps of the ---/+++ lines: the patch must not change from run to run
    lines = [ln.split("\t")[0] + "\n" if ln.startswith(("--- ", "+++ ")) else ln for ln in lines]
    if "--no-write" not in sys.argv:
        with open(os.path.join(HERE, "twoSD_src.patch"), "w") as fh:
            fh.writelines(lines)
    if not keep:
        shutil.rmtree(work, ignore_errors=True)
    print(f"{sum(1 for ln in lines if ln.startswith('@@'))} hunks")


if __name__ == "__main__":
    main()
