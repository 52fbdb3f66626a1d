/*
 * oracle/shim/shim.c  --  TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * From-scratch definitions of the handful of helpers the reference's hot-path translation units import
 * from the un-vendored SMU-SODA/spAlgorithms `spUtils` (unpinned; reference README.md:32), plus abort()
 * stubs for every CPLEX wrapper entry point.  Semantics are inferred from the reference's call sites:
 *
 *   vXv          cuts.c:106 (index vector given), cuts.c:218 (NULL index vector)
 *   vXvSparse    stocUpdate.c:218,244,293
 *   vxMSparse    stocUpdate.c:220,246,295
 *   reduceVector stocUpdate.c:221,247,269,296
 *   expandVector stocUpdate.c:214,236
 *   equalVector  stocUpdate.c:273,302,331
 *   duplicVector stocUpdate.c:338
 *   equalIntvec  stocUpdate.c:105        equalLongIntvec stocUpdate.c:41
 *   copyVector / addVectors / MSparsexvSub / decodeIntvec   randCost.c:236-243
 *
 * All sums run left to right from 0.0 with one rounding per multiply and per add (the build recipe
 * passes -ffp-contract=off so gcc cannot fuse them).
 */
#include "utils.h"
#include "solver_cplex.h"
#include "smps.h"
#include "prob.h"

void errMsg(const char *type, const char *place, const char *item, int quit) {
	fprintf(stderr, "sdref shim :: %s error in %s(): %s\n", type, place, item);
	if (quit) exit(1);
}

FILE *openFile(cString dir, cString name, cString mode) {
	char path[4096];
	snprintf(path, sizeof path, "%s%s", dir ? dir : "", name);
	return fopen(path, mode);
}

double vXv(dVector a, dVector b, iVector idxCol, int len) {
	double sum = 0.0;
	int c;
	if (idxCol == NULL)
		for (c = 1; c <= len; c++) sum += a[c] * b[c];
	else
		for (c = 1; c <= len; c++) sum += a[c] * b[idxCol[c]];
	return sum;
}

double vXvSparse(dVector v, sparseVector *vSparse) {
	double sum = 0.0;
	int c;
	for (c = 1; c <= vSparse->cnt; c++)
		sum += vSparse->val[c] * v[vSparse->col[c]];
	return sum;
}

dVector vxMSparse(dVector v, sparseMatrix *M, int len) {
	dVector ans = arr_alloc(len + 1, double);
	int c;
	for (c = 1; c <= M->cnt; c++)
		ans[M->col[c]] += v[M->row[c]] * M->val[c];
	return ans;
}

dVector MSparsexvSub(sparseMatrix *M, dVector v, dVector ans) {
	int c;
	for (c = 1; c <= M->cnt; c++)
		ans[M->row[c]] -= M->val[c] * v[M->col[c]];
	return ans;
}

static double absSum(dVector a, int len) {
	double s = 0.0;
	int c;
	for (c = 1; c <= len; c++) s += DBL_ABS(a[c]);
	return s;
}

dVector reduceVector(dVector f_vect, iVector row, int num_elem) {
	dVector s = arr_alloc(num_elem + 1, double);
	int c;
	for (c = 1; c <= num_elem; c++) s[c] = f_vect[row[c]];
	s[0] = absSum(s, num_elem);
	return s;
}

dVector expandVector(dVector red, iVector col, int redElems, int expElems) {
	dVector e = arr_alloc(expElems + 1, double);
	int c;
	for (c = 1; c <= redElems; c++) e[col[c]] = red[c];
	e[0] = absSum(e, expElems);
	return e;
}

dVector duplicVector(dVector a, int len) {
	dVector b = arr_alloc(len, double);
	int c;
	for (c = 0; c < len; c++) b[c] = a[c];
	return b;
}

void copyVector(dVector a, dVector b, int len) {
	int c;
	for (c = 0; c < len; c++) b[c] = a[c];
}

void copyIntvec(iVector a, iVector b, int len) {
	int c;
	for (c = 0; c < len; c++) b[c] = a[c];
}

void addVectors(dVector a, dVector b, iVector indices, int len) {
	int c;
	if (indices == NULL)
		for (c = 1; c <= len; c++) a[c] += b[c];
	else
		for (c = 1; c <= len; c++) a[indices[c]] += b[c];
}

bool equalVector(dVector a, dVector b, int len, double tolerance) {
	int c;
	for (c = 1; c <= len; c++)
		if (DBL_ABS(a[c] - b[c]) > tolerance) return false;
	return true;
}

bool equalIntvec(iVector a, iVector b, int len) {
	int c;
	for (c = 1; c <= len; c++)
		if (a[c] != b[c]) return false;
	return true;
}

bool equalLongIntvec(unsigned long *a, unsigned long *b, int len) {
	int c;
	for (c = 0; c < len; c++)
		if (a[c] != b[c]) return false;
	return true;
}

int isElementIntvec(iVector intVec, int len, int val) {
	int c;
	for (c = 1; c <= len; c++)
		if (intVec[c] == val) return c;
	return 0;
}

/* two bits per status value, values 1..len of the input; only has to round-trip with decodeIntvec */
unsigned long *encodeIntvec(iVector vec, int len, int wordLength, int numBits) {
	int perWord = wordLength / 2, words = (len + perWord - 1) / perWord + 1, c;
	unsigned long *code = arr_alloc(words, unsigned long);
	(void) numBits;
	for (c = 1; c <= len; c++)
		code[(c - 1) / perWord] |= ((unsigned long) (vec[c] & 3)) << (2 * ((c - 1) % perWord));
	return code;
}

iVector decodeIntvec(unsigned long *code, int len, int wordLength, int numBits) {
	int perWord = wordLength / 2, c;
	iVector vec = arr_alloc(len + 1, int);
	(void) numBits;
	for (c = 1; c <= len; c++)
		vec[c] = (int) ((code[(c - 1) / perWord] >> (2 * ((c - 1) % perWord))) & 3UL);
	return vec;
}

void printVector(dVector v, int len, FILE *fp) {
	int c;
	if (!fp) fp = stdout;
	for (c = 1; c <= len; c++) fprintf(fp, "%g ", v[c]);
	fprintf(fp, "\n");
}

void printIntvec(iVector v, int len, FILE *fp) {
	int c;
	if (!fp) fp = stdout;
	for (c = 1; c <= len; c++) fprintf(fp, "%d ", v[c]);
	fprintf(fp, "\n");
}

void printSparseVector(dVector v, iVector idx, int len) {
	int c;
	for (c = 1; c <= len; c++) printf("(%d) %g ", idx[c], v[c]);
	printf("\n");
}

void freeSparseMatrix(sparseMatrix *M) {
	if (M) { free(M->col); free(M->row); free(M->val); free(M); }
}

/* SplitMix64-driven uniform integer in [1, range] (only resampleOmega uses it; tests pass observ[] directly) */
int randInteger(long long *seed, int range) {
	unsigned long long z = (unsigned long long) (*seed += 0x9E3779B97F4A7C15LL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	z ^= z >> 31;
	return (int) (z % (unsigned long long) range) + 1;
}

/* ---- solver / driver entry points the compiled reference files name but the oracle never reaches ---- */
#define UNREACHABLE(name) do { fprintf(stderr, "sdref shim :: %s() needs CPLEX; not available in the oracle build\n", name); abort(); } while (0)

/* replay of one recorded solve (sdReplayLP, solver_cplex.h); a NULL LPptr still means "no solver here" */
#define REPLAY(name) const sdReplayLP *r = (const sdReplayLP *) lp; if (!r) UNREACHABLE(name)
int getDual(LPptr lp, dVector pi, int length) {
	int i; REPLAY("getDual");
	for (i = 0; i < length && i < r->rows; i++) pi[i + 1] = r->pi[i];
	return 0;
}
int getPrimal(LPptr lp, dVector x, int length) {
	int i; REPLAY("getPrimal");
	for (i = 0; i < length && i < r->cols; i++) x[i + 1] = r->x[i];
	return 0;
}
int getDualSlacks(LPptr lp, dVector dj, int length) {
	int i; REPLAY("getDualSlacks");
	for (i = 0; i < length && i < r->cols; i++) dj[i + 1] = r->dj[i];
	return 0;
}
int getBasis(LPptr lp, iVector cstat, iVector rstat) {
	int i; REPLAY("getBasis");
	for (i = 0; i < r->cols; i++) cstat[i + 1] = r->cstat[i];      /* computeMU reads cstat[1..numCols] (stocUpdate.c:372) */
	for (i = 0; i < r->rows; i++) rstat[i + 1] = r->rstat[i];
	return 0;
}
int getBasisHead(LPptr lp, iVector head, dVector x) {
	int i; REPLAY("getBasisHead");
	(void) x;
	for (i = 0; i < r->basisDim; i++) head[i] = r->head[i];        /* the caller passes basisHead+1 (randCost.c:31) */
	return 0;
}
int getBasisInvRow(LPptr lp, int i, dVector y) {
	int c; REPLAY("getBasisInvRow");
	for (c = 0; c < r->rows; c++) y[c] = r->binvRows[(size_t) i * r->rows + c];      /* caller passes phi+1 (randCost.c:47) */
	return 0;
}
int getBasisInvACol(LPptr lp, int i, dVector y) {
	int c; REPLAY("getBasisInvACol");
	for (c = 0; c < r->rows; c++) y[c] = r->binvACols[(size_t) i * r->rows + c];     /* caller passes tempPsiRow+1 (randCost.c:79) */
	return 0;
}
double getObjective(LPptr lp, int type)                     { (void) lp; (void) type; UNREACHABLE("getObjective"); return 0.0; }
int    removeRows(LPptr lp, int begin, int end)             { (void) lp; (void) begin; (void) end; UNREACHABLE("removeRows"); return 1; }
int    writeProblem(LPptr lp, cString fname)                { (void) lp; (void) fname; UNREACHABLE("writeProblem"); return 1; }
