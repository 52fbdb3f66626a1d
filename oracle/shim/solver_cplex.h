/*
 * oracle/shim/solver_cplex.h  --  TEST INFRASTRUCTURE ONLY.
 * Stand-in for spAlgorithms' CPLEX wrapper header.  Only names the reference's hot-path translation
 * units mention; every solver entry point is a stub that aborts if reached (oracle/shim/shim.c): the
 * oracle never solves an LP, it is fed recorded dual vertices.
 */
#ifndef SDREF_SHIM_SOLVER_H
#define SDREF_SHIM_SOLVER_H

#include "utils.h"

typedef void *LPptr;

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
