/*
 * oracle/shim/solver_cplex.h  --  TEST INFRASTRUCTURE ONLY.
 * Stand-in for spAlgorithms' CPLEX wrapper header.  Only names the reference's hot-path translation
 * units mention.  The oracle never solves an LP: the get* entry points replay one recorded solve (sdReplayLP below) when the
 * LPptr is not NULL and abort otherwise; everything else is a stub that aborts if reached (oracle/shim/shim.c).
 */
#ifndef SDREF_SHIM_SOLVER_H
#define SDREF_SHIM_SOLVER_H

#include "utils.h"

typedef void *LPptr;

/* Replay "solver" (tests of the host patch): an LPptr of the harness points at ONE recorded solve -- what CPLEX would hand back
 * after solveProblem() -- and the get* entry points below copy from it.  Arrays are 0-based, as a solver returns them; the
 * wrappers shift where the reference's call sites expect the un-vendored wrapper to (1-based vectors). */
typedef struct {
	int rows, cols, basisDim;
	const int    *cstat, *rstat;     /* [cols], [rows]: AT_LOWER / BASIC / AT_UPPER / FREE_SUPER                  */
	const double *x, *dj, *pi;       /* [cols] primal, [cols] reduced costs, [rows] duals                         */
	const int    *head;              /* [basisDim] basic variable of each row: column j >= 0, or -(row+1) for a slack */
	const double *binvRows;          /* [basisDim][rows] rows of the basis inverse (only the rows asked for are read) */
	const double *binvACols;         /* [cols][rows]     B^-1 A_j (the tableau columns)                              */
} sdReplayLP;

#define ON  1
#define OFF 0
#define PROB_LP 0
#define PROB_QP 5
#define ALG_PRIMAL 1
#define PARAM_PREIND 1030
#define STAT_INFEASIBLE 3
#define AT_LOWER 0
#define BASIC    1
#define AT_UPPER 2
#define FREE_SUPER 3
#define GE 'G'
#define LE 'L'
#define EQ 'E'

int    getDual(LPptr lp, dVector pi, int length);
int    getPrimal(LPptr lp, dVector x, int length);
int    getDualSlacks(LPptr lp, dVector dj, int length);
int    getBasis(LPptr lp, iVector cstat, iVector rstat);
int    getBasisHead(LPptr lp, iVector head, dVector x);
int    getBasisInvRow(LPptr lp, int i, dVector y);
int    getBasisInvACol(LPptr lp, int i, dVector y);
double getObjective(LPptr lp, int type);
int    removeRows(LPptr lp, int begin, int end);
int    addRow(LPptr lp, int nzcnt, double rhs, char sense, int matbeg, iVector rmatind, dVector rmatval, cString rowname);
int    changeRHS(LPptr lp, int cnt, iVector indices, dVector rhs);
int    changeCol(LPptr lp, int column, dVector coef, int start, int stop);
int    writeProblem(LPptr lp, cString fname);
int    setIntParam(int paramname, int paramvalue);
void   changeLPSolverType(int method);
int    solveProblem(LPptr lp, cString pname, int type, int *status);

#endif
