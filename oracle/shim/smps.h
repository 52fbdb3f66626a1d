/*
 * oracle/shim/smps.h  --  TEST INFRASTRUCTURE ONLY.
 * Stand-in for spAlgorithms' SMPS reader header: just the aggregate types the reference's prototypes
 * mention (oneProblem, timeType, stocType).  No SMPS parsing exists here.
 */
#ifndef SDREF_SHIM_SMPS_H
#define SDREF_SHIM_SMPS_H

#include "utils.h"
#include "solver_cplex.h"

typedef struct {
	int     type;
	LPptr   lp;
	cString name;
	int     mac, mar, macsz, marsz, numnz, numInt, matsz, cstorsz, rstorsz, objsen;
	cString objname, senx, ctype, cstore, rstore;
	cString *cname, *rname;
	dVector objx, rhsx, bdl, bdu, matval;
	iVector matbeg, matcnt, matind;
} oneProblem;

typedef struct { int numStages; } timeType;

typedef struct {
	int     numOmega;
	dVector mean;
} stocType;

int generateOmega(stocType *stoc, dVector observ, long long *seed, void *unused);

#endif
