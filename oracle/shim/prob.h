/*
 * oracle/shim/prob.h  --  TEST INFRASTRUCTURE ONLY.
 * Stand-in for spAlgorithms' stage-decomposition header.  Field names follow the reference's uses
 * (SURVEY.md §8c field census: num->{rows,cols,prevCols,prevRows,cntCcols,rvRowCnt,rvbOmCnt,rvCOmCnt,
 * rvdOmCnt,numRV}, coord->{CCols,rvRows,rvbOmRows,rvCOmCols,rvCOmRows,rvCols,rvdOmCols,rvOffset}).
 */
#ifndef SDREF_SHIM_PROB_H
#define SDREF_SHIM_PROB_H

#include "utils.h"
#include "smps.h"

typedef struct {
	int rows, cols, intCols, binCols;
	int prevRows, prevCols;
	int cntCcols, cntCrows;
	int numRV, rvRowCnt, rvColCnt;
	int rvaOmCnt, rvbOmCnt, rvcOmCnt, rvdOmCnt, rvAOmCnt, rvBOmCnt, rvCOmCnt, rvDOmCnt;
} numType;

typedef struct {
	iVector allRVRows, allRVCols;
	iVector CCols, CRows;
	iVector rvCols, rvRows;
	iVector rvbOmRows, rvdOmCols;
	iVector rvCOmCols, rvCOmRows;
	iVector rvDOmCols, rvDOmRows;
	iVector rvOffset;
} coordType;

typedef struct {
	cString       name;
	oneProblem   *sp;
	numType      *num;
	coordType    *coord;
	sparseVector *aBar, *bBar, *cBar, *dBar;
	sparseMatrix *Abar, *Bbar, *Cbar, *Dbar;
	dVector       mean;
	double        lb;
} probType;

void freeProbType(probType **prob, int T);

#endif
