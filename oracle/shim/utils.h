/*
 * oracle/shim/utils.h  --  TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for the un-vendored SMU-SODA/spAlgorithms `spUtils/utils.h` (no version pin exists;
 * reference README.md:32).  It lets the reference's own translation units
 * (/root/reference/twoSD_src/{stocUpdate,cuts,optimal,randCost}.c) compile *where they lie* into
 * oracle/_ref/libsdref.so so that the restated oracle (oracle/sd_oracle.c) can be checked against the
 * reference's real control flow.  Everything declared here is written from scratch; the semantics of
 * each helper are inferred from the reference's call sites and are the one part of the
 * parity chain that cannot be verified against an executable upstream.
 */
#ifndef SDREF_SHIM_UTILS_H
#define SDREF_SHIM_UTILS_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#include <time.h>
#include <stdbool.h>

typedef double *dVector;
typedef int    *iVector;
typedef char   *cString;

typedef struct { int cnt; iVector col; dVector val; } sparseVector;
typedef struct { int cnt; iVector col; iVector row; dVector val; } sparseMatrix;

#define NAMESIZE   32
#define BLOCKSIZE  256
#define WORDLENGTH 64
#define INF        1.0e20
#define DBL_ABS(x) ((x) > 0.0 ? (x) : -(x))

/* zero-initialising array allocator (call sites rely on zeroed beta / theta / expandVector output) */
#define arr_alloc(n, type) ((type *) calloc((size_t) ((n) > 0 ? (n) : 1), sizeof(type)))
#define mem_malloc(sz)     malloc((size_t) (sz))
#define mem_realloc(p, sz) realloc((p), (size_t) (sz))
#define mem_free(p)        free((void *) (p))

void   errMsg(const char *type, const char *place, const char *item, int quit);
FILE  *openFile(cString dir, cString name, cString mode);

/* dense / sparse algebra; every dVector is 1-based, slot 0 holds the one-norm */
double  vXv(dVector a, dVector b, iVector idxCol, int len);
double  vXvSparse(dVector v, sparseVector *vSparse);
dVector vxMSparse(dVector v, sparseMatrix *M, int len);
dVector MSparsexvSub(sparseMatrix *M, dVector v, dVector ans);
dVector reduceVector(dVector f_vect, iVector row, int num_elem);
dVector expandVector(dVector red, iVector col, int redElems, int expElems);
dVector duplicVector(dVector a, int len);
void    copyVector(dVector a, dVector b, int len);
void    copyIntvec(iVector a, iVector b, int len);
void    addVectors(dVector a, dVector b, iVector indices, int len);
bool    equalVector(dVector a, dVector b, int len, double tolerance);
bool    equalIntvec(iVector a, iVector b, int len);
bool    equalLongIntvec(unsigned long *a, unsigned long *b, int len);
int     isElementIntvec(iVector intVec, int len, int val);
unsigned long *encodeIntvec(iVector vec, int len, int wordLength, int numBits);
iVector decodeIntvec(unsigned long *code, int len, int wordLength, int numBits);
void    printVector(dVector v, int len, FILE *fp);
void    printIntvec(iVector v, int len, FILE *fp);
void    printSparseVector(dVector v, iVector idx, int len);
void    freeSparseMatrix(sparseMatrix *M);
int     randInteger(long long *seed, int range);

#endif
