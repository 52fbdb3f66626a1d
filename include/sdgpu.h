/*
 * sdgpu.h -- C ABI of the B200-native stochastic-decomposition cut-formation library (libsdgpu.so).
 *
 * The library owns, in HBM, the dual-vertex x observation tables of two-stage SD (omega, lambda, sigma,
 * delta, the basis descriptors and the basis x observation feasibility mask) and forms the SD cut
 * (per-observation argmax over all stored bases fused with the weighted reduction into alpha/beta and the
 * dual-stability sums).  It replaces, call for call, the table half of the reference's stocUpdate.c and
 * the numeric core of cuts.c; the C host keeps CPLEX, the SMPS reader, config.sd and every output file.
 * Each entry point below names the reference function (file:line under /root/reference/twoSD_src) it
 * replaces.  INTEGRATION.md shows the host-side patch.
 *
 * Conventions (the reference's, unchanged):
 *   - every double vector handed in or out is 1-BASED with slot 0 reserved (the one-norm slot), every
 *     coordinate array is 1-based and holds 1-based positions;
 *   - observation / lambda / sigma / basis indices are 0-based table slots, as in the reference;
 *   - functions returning an index return >= 0, or SDGPU_NONE (-1) where the reference returns -1
 *     (computeIstar stocUpdate.c:186-189), or a value <= SDGPU_ERR (-2) on failure; int status
 *     functions return 0 on success and < 0 on failure.  Nothing exits or throws across the ABI; the
 *     message for the last failure on the calling thread is available from sdgpu_last_error();
 *   - every call is synchronous on return (the host feeds the results straight into CPLEX,
 *     master.c:110); one host thread drives one context; buffers passed in are borrowed for the call.
 *   - there is NO CPU fallback: without a CUDA device sdgpu_create() fails.
 */
#ifndef SDGPU_H_
#define SDGPU_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDGPU_ABI_VERSION 2
#define SDGPU_NONE (-1)
#define SDGPU_ERR  (-2)

typedef struct sdgpu_ctx sdgpu_ctx;

/* numType fields the path reads (prob.h of spAlgorithms; census in SURVEY.md section 8c). */
typedef struct {
	int32_t rows;      /* num->rows      second-stage rows                                     */
	int32_t cols;      /* num->cols      second-stage columns                                  */
	int32_t prevCols;  /* num->prevCols  n1, first-stage columns = length of beta              */
	int32_t cntCcols;  /* num->cntCcols  n1c, columns of Cbar with non-zeros                   */
	int32_t rvRowCnt;  /* num->rvRowCnt  R, rows with random elements = length of a lambda     */
	int32_t rvbOmCnt;  /* num->rvbOmCnt  Rb, random right-hand-side elements                   */
	int32_t rvCOmCnt;  /* num->rvCOmCnt  Q, random technology-matrix elements                  */
	int32_t rvdOmCnt;  /* num->rvdOmCnt  random cost coefficients (selects the cuts.c:142 branch) */
	int32_t numRV;     /* num->numRV     length of an observation                              */
} sdgpu_num;

/* coordType arrays the path reads; all 1-based (slot 0 ignored), copied to the device at create. */
typedef struct {
	const int32_t *CCols;      /* [cntCcols+1]  cuts.c:106,155,165; stocUpdate.c:296            */
	const int32_t *rvRows;     /* [rvRowCnt+1]  stocUpdate.c:214,236,269                        */
	const int32_t *rvbOmRows;  /* [rvbOmCnt+1]  stocUpdate.c:203                                */
	const int32_t *rvCOmCols;  /* [rvCOmCnt+1]  stocUpdate.c:175,204; cuts.c:157                */
	const int32_t *rvCOmRows;  /* [rvCOmCnt+1]  stocUpdate.c:204                                */
	const int32_t *rvCols;     /* [rvCOmCnt+1]  cuts.c:167 (plain-branch beta scatter)          */
	int32_t        rvOffset[3];/* rvOffset[2] = start of the cost deltas in an observation      */
} sdgpu_coord;

/* sparseVector / sparseMatrix of utils.h: cnt entries at positions 1..cnt. */
typedef struct { int32_t cnt; const int32_t *col; const double *val; } sdgpu_sparse_vec;
typedef struct { int32_t cnt; const int32_t *col; const int32_t *row; const double *val; } sdgpu_sparse_mat;

/* Capacity contract of setup.c:136-144 and cuts.c:28 (the reference never bounds-checks; we do). */
typedef struct {
	int64_t maxLambda;  /* rows of lambda and of delta  (setup.c:139,141,143)                    */
	int64_t maxSigma;   /* rows of sigma               (setup.c:142)                            */
	int64_t maxBasis;   /* basis records               (setup.c:140)                            */
	int64_t maxOmega;   /* distinct observations = delta row length (setup.c:144, cuts.c:28)    */
	int32_t maxTerms;   /* max (1 + phiLength) of one basis; 1 unless rvdOmCnt > 0              */
} sdgpu_caps;

typedef struct {
	sdgpu_num        num;
	sdgpu_coord      coord;
	sdgpu_sparse_vec bBar;   /* prob->bBar  stocUpdate.c:293 */
	sdgpu_sparse_mat Cbar;   /* prob->Cbar  stocUpdate.c:295 */
} sdgpu_problem;

typedef struct { int64_t omega, lambda, sigma, basis; } sdgpu_counts;

/* Result of one cut (oneCut of twoSD.h:69-80 without the solver bookkeeping). */
typedef struct {
	double   alpha;      /* cut->alpha  cuts.c:184                                               */
	double  *beta;       /* [prevCols+1] caller buffer; beta[0] = 1.0  cuts.c:186-188            */
	int32_t *iStar;      /* [omegaCnt] caller buffer or NULL (stays device-resident)  cuts.c:140 */
	int32_t  omegaCnt;   /* observations the cut was formed on  cuts.c:100                       */
	int32_t  numSamples; /* cut->numSamples                                                      */
	double   cummOld;    /* cuts.c:127 (0 unless pi_eval)                                        */
	double   cummAll;    /* cuts.c:128                                                           */
} sdgpu_cut;

/* ---- lifecycle: newLambda/newSigma/newDelta/newOmega/newBasisType (setup.c:140-144) --------------- */
int  sdgpu_abi_version(void);
int  sdgpu_create(const sdgpu_problem *prob, const sdgpu_caps *caps, int device, sdgpu_ctx **out);
/* free*Type(..., partial = true) between replications (setup.c:242-246): counts to 0, memory kept */
int  sdgpu_reset(sdgpu_ctx *ctx);
/* free*Type(..., false) (setup.c:282-286) */
void sdgpu_destroy(sdgpu_ctx *ctx);
const char *sdgpu_last_error(void);
int  sdgpu_get_counts(sdgpu_ctx *ctx, sdgpu_counts *out);

/* ---- stochastic tables (stocUpdate.c) ------------------------------------------------------------ */
/* calcOmega stocUpdate.c:326-348: first stored observation within tol of observ[1..numRV] gets its
 * weight bumped, else observ is appended with weight 1.  Returns the observation index. */
int  sdgpu_calc_omega(sdgpu_ctx *ctx, const double *observ, double tol, int *newOmegaFlag);
/* the three steps of calcOmega on their own (used when observations are sharded across GPUs) */
int  sdgpu_omega_find(sdgpu_ctx *ctx, const double *observ, double tol);        /* idx or SDGPU_NONE */
int  sdgpu_omega_append(sdgpu_ctx *ctx, const double *observ, int weight);      /* idx               */
int  sdgpu_omega_bump(sdgpu_ctx *ctx, int idx, int by);
/* bulk load of n observations without the dedup scan (synthetic sweeps; vals is n rows of numRV+1) */
int  sdgpu_omega_append_bulk(sdgpu_ctx *ctx, int64_t n, const double *vals, const int32_t *weights);

/* calcLambda stocUpdate.c:264-284: Pi is the full dual [rows+1]; returns the lambda index. */
int  sdgpu_calc_lambda(sdgpu_ctx *ctx, const double *Pi, double tol, int *newLambdaFlag);
/* calcSigma stocUpdate.c:286-320: returns the sigma index. */
int  sdgpu_calc_sigma(sdgpu_ctx *ctx, const double *pi, double mubBar, int idxLambda, int newLambdaFlag,
                      int currentIter, double tol, int *newSigmaFlag);
/* calcDelta stocUpdate.c:196-257: newOmegaFlag != 0 fills column elemIdx for every lambda (case I),
 * else fills row elemIdx for every observation (case II). */
int  sdgpu_calc_delta(sdgpu_ctx *ctx, int newOmegaFlag, int elemIdx);
/* calcDelta for every (lambda, observation) pair of the block [l0,l1) x [o0,o1): the bulk form used after the
 * bulk loaders below (same per-cell arithmetic and order as the one-at-a-time appends). */
int  sdgpu_calc_delta_block(sdgpu_ctx *ctx, int64_t l0, int64_t l1, int64_t o0, int64_t o1);
/* calcLambda + calcSigma + calcDelta(row) for one dual vector in one device round trip
 * (stocUpdate.c:78-85 / :90-97).  Outputs may be NULL. */
int  sdgpu_update_dual(sdgpu_ctx *ctx, const double *pi, double mubBar, int currentIter, double tol,
                       int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag);
/* the same with the delta column of a NEW observation (calcDelta case I, stocUpdate.c:24-25) folded into the launch that scans lambda:
 * newOmegaIdx = the index sdgpu_calc_omega just returned with *newOmegaFlag set, or -1 when the observation was not new (then this
 * is sdgpu_update_dual).  One launch for {column, lambda, sigma}, one programmatically dependent launch for the delta row. */
int  sdgpu_update_dual_col(sdgpu_ctx *ctx, int newOmegaIdx, const double *pi, double mubBar, int currentIter, double tol,
                           int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag);
/* bulk load of n dual vectors (rows of rows+1 doubles).  tol >= 0: the same find-or-append chain as
 * sdgpu_update_dual, vector after vector, without host round trips in between.  tol < 0: synthetic loader,
 * no dedup scan -- vector i becomes lambda row and sigma row (count + i) and NO delta row is computed
 * (call sdgpu_calc_delta_block afterwards).  idx outputs and mubBar / iters may be NULL. */
int  sdgpu_update_dual_bulk(sdgpu_ctx *ctx, int64_t n, const double *pis, const double *mubBar,
                            const int32_t *iters, double tol, int32_t *lambdaIdx, int32_t *sigmaIdx);

/* basis records read by the argmax (stoc.h:72-97): append as stocUpdate.c:117-131 does.
 * sigmaIdx[0..phiLength] are sigma slots; omegaIdx[1..phiLength] are 1-based positions in the cost
 * block of an observation (omegaIdx[0] unused; may be NULL when phiLength == 0).  A feasible basis gets
 * an obsFeasible row initialised to all-true (checkBasisFeasibility with rvdOmCnt == 0, randCost.c:208). */
int  sdgpu_basis_append(sdgpu_ctx *ctx, int ck, int feasFlag, int phiLength, const int32_t *sigmaIdx,
                        const int32_t *omegaIdx);
/* n single-term bases (phiLength == 0) at once: basis i gets ck[i], feasFlag feas[i] (NULL = all feasible) and
 * sigmaIdx[i].  Bulk form of sdgpu_basis_append for table loaders.  Returns the index of the first one. */
int  sdgpu_basis_append_bulk(sdgpu_ctx *ctx, int64_t n, const int32_t *ck, const int32_t *feas, const int32_t *sigmaIdx);
/* stocUpdate.c:101-131: when retainBasis == 0 return the first stored basis with the same phiLength,
 * obsFeasible[b][obsIdx] set and the same sigmaIdx list (newBasisFlag = 0), else append. */
int  sdgpu_basis_find_or_append(sdgpu_ctx *ctx, int retainBasis, int obsIdx, int ck, int feasFlag,
                                int phiLength, const int32_t *sigmaIdx, const int32_t *omegaIdx,
                                int *newBasisFlag);
/* basis->obsFeasible[basisIdx][obsIdx] = flag (stocUpdate.c:30,125) and the row / column forms */
int  sdgpu_basis_set_obs_feasible(sdgpu_ctx *ctx, int basisIdx, int obsIdx, int flag);
int  sdgpu_basis_set_obs_feasible_row(sdgpu_ctx *ctx, int basisIdx, const uint8_t *flags /* [omegaCnt] */);
int  sdgpu_basis_set_obs_feasible_col(sdgpu_ctx *ctx, int obsIdx, const uint8_t *flags /* [basisCnt] */);

/* ---- basis feasibility under random costs: checkBasisFeasibility randCost.c:202-258 on the device -------------
 * The host still extracts piDet / phi / gBar / psi / cstat from the CPLEX basis (randCost.c:19-180); it hands them over
 * once per new basis and the library evaluates the basis x observation mask itself, for a new observation against
 * all stored bases (stocUpdate.c:28-30) and for a new basis against all stored observations (stocUpdate.c:123-126). */
int  sdgpu_set_cost_coords(sdgpu_ctx *ctx, const int32_t *rvdOmCols /* [rvdOmCnt+1], 1-based */, const char *senx /* [rows] */);
/* phi: phiLength vectors of rows+1 doubles back to back; psiVal: psi->val in the order calcBasis writes it
 * (randCost.c:83-88), i.e. [cols][phiLength] row-major, 0-based; gBar [cols+1]; cstat [cols+1] (AT_UPPER == 2). */
int  sdgpu_basis_set_feas_data(sdgpu_ctx *ctx, int basisIdx, const double *piDet, const double *phi, const double *gBar,
                               const double *psiVal, const int32_t *cstat);
/* flagsOut (may be NULL) receives the computed flags; the device mask and the library's host mirror are updated. */
int  sdgpu_check_feasibility_obs(sdgpu_ctx *ctx, int obsIdx, double tol, uint8_t *flagsOut /* [basisCnt] */);
int  sdgpu_check_feasibility_basis(sdgpu_ctx *ctx, int basisIdx, double tol, uint8_t *flagsOut /* [omegaCnt] */);

/* ---- feasibility cuts: the numeric core of updtFeasCutPool cuts.c:465-517 ------------------------------------------
 * For every observation in [obsFirst, obsLast) (outer loop) and every INFEASIBLE basis in [basisFirst, basisLast)
 * (inner loop, feasFlag == 0) one raw cut: alpha = sigma.pib + delta.pib, beta[CCols] += sigma.piC, beta[rvCols] +=
 * delta.piC of the basis' first sigma (cuts.c:478-486).  Returns the number of cuts written (<= maxOut) in that loop
 * order; the pool de-duplication of addCut2Pool (cuts.c:643-655) stays with the host. */
int  sdgpu_feas_cuts(sdgpu_ctx *ctx, int obsFirst, int obsLast, int basisFirst, int basisLast, int maxOut,
                     double *alpha /* [maxOut] */, double *beta /* [maxOut][prevCols+1] */);

/* ---- the feasibility-cut pool (cell->fcutsPool) on the device ---------------------------------------------------------------
 * updtFeasCutPool cuts.c:465-517 INCLUDING the duplicate test of addCut2Pool cuts.c:643-655: the raw cuts of the two loops are
 * generated in the reference's order and appended to a device-resident pool unless an equal cut (|alpha - alpha'| < tol and every
 * |beta - beta'| <= tol) is already in it -- first come, first kept, exactly the sequential rule.  fUpdt is cell->fUpdt (in/out:
 * [0] bases, [1] observations the pool has been updated for).  Returns the pool size.  sdgpu_reset() empties the pool. */
int  sdgpu_feas_pool_update(sdgpu_ctx *ctx, int *fUpdt /* [2] */, double tol);
int  sdgpu_feas_pool_size(sdgpu_ctx *ctx);
int  sdgpu_feas_pool_get(sdgpu_ctx *ctx, int first, int count, double *alpha /* [count] */, double *beta /* [count][prevCols+1] */);
/* checkFeasCutPool cuts.c:521-567 for the whole pool in one launch.  fAlpha / fBeta: the nFcuts feasibility cuts already in the master
 * (cell->fcuts).  action[i] for pool cut i: 0 nothing to do; 1 the incumbent violates it (beta.x < alpha) and no equal cut is in the
 * master: add it (cuts.c:545); 2 the incumbent violates it but an equal cut is in the master (cuts.c:541); 3 the candidate violates it:
 * add it (cuts.c:556).  *infeasIncumb = 1 iff some action is 1 or 2 (cuts.c:540).  The host calls addCut2Master for actions 1 and 3 in
 * pool order.  Returns the pool size. */
int  sdgpu_feas_pool_check(sdgpu_ctx *ctx, int nFcuts, const double *fAlpha, const double *fBeta /* [nFcuts][prevCols+1] */,
                           const double *incumbX, const double *candidX /* [prevCols+1] */, double tol, int32_t *action /* [pool] */,
                           int *infeasIncumb);

/* ---- cut formation (cuts.c, stocUpdate.c:142-190) ------------------------------------------------- */
/* computeIstar stocUpdate.c:142-190 for ONE observation (debug / STOCH_CHECK use; the cut path below
 * never calls it).  Returns the basis index or SDGPU_NONE; *argmax as the reference sets it. */
int  sdgpu_compute_istar(sdgpu_ctx *ctx, const double *Xvect, int obs, int numSamples, int pi_eval,
                         int isNew, double *argmax);
/* SDCut cuts.c:91-194 minus the config-dependent tail (pi_ratio / variance, cuts.c:171-182, stays with
 * the host: see sdgpu_dual_stability()).  pi_eval_flag is the value cuts.c:112 computes.  Returns 0, or
 * SDGPU_NONE when some observation has no eligible basis (the reference returns NULL, cuts.c:136-139). */
int  sdgpu_sd_cut(sdgpu_ctx *ctx, const double *Xvect, int numSamples, int pi_eval_flag, double lb,
                  sdgpu_cut *cut);
/* cuts.c:171-182 + calcVariance cuts.c:366-396: updates pi_ratio[numSamples % scanLen] and returns the
 * new dualStableFlag (host arithmetic on <= SCAN_LEN doubles; provided so the host patch is one line). */
int  sdgpu_dual_stability(double cummOld, double cummAll, int numSamples, int piEvalStart, int scanLen,
                          double *pi_ratio);

/* Sharded form of sdgpu_sd_cut for observation-parallel multi-GPU runs: _partial leaves the
 * un-normalised sums [alpha, beta[1..prevCols], cummOld, cummAll, missing] (prevCols + 4 doubles) in a
 * device buffer, the caller all-reduces that buffer (NCCL sum), _finish divides by numSamples and copies
 * the cut (and this shard's iStar) back. */
int  sdgpu_sd_cut_partial(sdgpu_ctx *ctx, const double *Xvect, int numSamples, int pi_eval_flag, double lb);
int  sdgpu_sd_cut_partial_buffer(sdgpu_ctx *ctx, void **devPtr, int *len);
int  sdgpu_sd_cut_finish(sdgpu_ctx *ctx, int numSamples, sdgpu_cut *cut);
/* Attach an NCCL communicator (ncclComm_t, created by the caller for this context's device); after this
 * sdgpu_sd_cut() performs the all-reduce itself, on the context's stream, between _partial and _finish. */
int  sdgpu_attach_nccl(sdgpu_ctx *ctx, void *ncclComm);
int  sdgpu_nccl_unique_id(void *id128);                                  /* ncclGetUniqueId            */
int  sdgpu_nccl_init(sdgpu_ctx *ctx, int nranks, int rank, const void *id128); /* ncclCommInitRank + attach */

/* NVLink peer-memory all-reduce fused into the cut kernel (the latency-optimal form of the one exchange step): every rank
 * exports a small exchange buffer (CUDA IPC), attaches everyone else's, and from then on the last block of the cut's merge
 * kernel writes its n1+4 partial sums straight into every peer's buffer over NVLink, waits for the peers' flags, adds the
 * rank slots in rank order (so all ranks get bit-identical cuts, run to run) and normalises -- no NCCL call, no extra launch.
 * sdgpu_peer_export fills handle64 (64 bytes, cudaIpcMemHandle_t); handles = nranks x 64 bytes in rank order. */
int  sdgpu_peer_export(sdgpu_ctx *ctx, int nranks, void *handle64);
int  sdgpu_peer_attach(sdgpu_ctx *ctx, int nranks, int rank, const void *handles);
/* Which exchange sdgpu_sd_cut() uses when both an NCCL communicator and a peer exchange are attached (the two may be attached in either
 * order and are torn down independently): 0 = automatic (the peer exchange if attached, else NCCL), 1 = NCCL, 2 = peer exchange.
 * All ranks must choose the same one for a given cut. */
int  sdgpu_set_collective(sdgpu_ctx *ctx, int mode);

/* ---- one host thread, several GPUs (the reference's host is a single process) -------------------------------------------
 * A group owns one context per device of this process.  Observations are dealt round-robin (global observation o lives on
 * member o % G at local slot o / G); lambda / sigma / basis records are replicated.  The group calls below have the
 * single-context meaning over the union of the shards; the cut's all-reduce is the NVLink peer-memory exchange inside the cut
 * kernel (direct peer pointers, no IPC needed in one process).  rvdOmCnt must be 0. */
typedef struct sdgpu_group sdgpu_group;
int  sdgpu_group_create(const sdgpu_problem *prob, const sdgpu_caps *capsPerDevice, int nDevices, const int *devices, sdgpu_group **out);
void sdgpu_group_destroy(sdgpu_group *grp);
int  sdgpu_group_reset(sdgpu_group *grp);
int  sdgpu_group_size(sdgpu_group *grp);
sdgpu_ctx *sdgpu_group_member(sdgpu_group *grp, int i);
int  sdgpu_group_get_counts(sdgpu_group *grp, sdgpu_counts *out);                      /* omega = total over the shards */
int  sdgpu_group_calc_omega(sdgpu_group *grp, const double *observ, double tol, int *newOmegaFlag);   /* global index; fills the delta column when new */
int  sdgpu_group_update_dual(sdgpu_group *grp, const double *pi, double mubBar, int currentIter, double tol,
                             int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag);
int  sdgpu_group_basis_find_or_append(sdgpu_group *grp, int retainBasis, int ck, int feasFlag, int sigmaIdx, int *newBasisFlag);
/* SDCut over all shards; cut->iStar (if not NULL) receives the maximisers in GLOBAL observation order */
int  sdgpu_group_sd_cut(sdgpu_group *grp, const double *Xvect, int numSamples, int pi_eval_flag, double lb, sdgpu_cut *cut);

/* cutHeight cuts.c:213-227 / maxCutHeight cuts.c:197-209 and the aging coefficients of
 * changeEtaCol master.c:152 and updateRHS master.c:174 for a batch of cuts held by the host:
 * alpha[n], beta = n rows of prevCols+1, numSamples[n]; outputs (each may be NULL) height[n],
 * etaCoef[n] = k/numSamples, rhs[n] = alphaIncumb + (k/numSamples - 1)*lb.  Returns the index of the
 * highest cut (first on ties, as the strict '<' of cuts.c:203 keeps it) or SDGPU_NONE when n == 0. */
int  sdgpu_cut_heights(sdgpu_ctx *ctx, int n, const double *alpha, const double *beta,
                       const int32_t *numSamples, const double *alphaIncumb, int currIter,
                       const double *xk, double lb, double *height, double *etaCoef, double *rhs);

/* reformCuts optimal.c:187-236 for one cut: re-averages the stored maximisers iStar[] over the
 * resampled observation list observ[0..k).  iStar == NULL uses the device-resident iStar of the most
 * recent sdgpu_sd_cut().  lb is truncated to int as optimal.c:188 does. */
int  sdgpu_reform_cut(sdgpu_ctx *ctx, const int32_t *iStar, int omegaCnt, const int32_t *observ, int k,
                      int lbType, int lb, double *alpha, double *beta /* [prevCols+1] */);

/* The bootstrap of fullTest (optimal.c:96-103) in one call: nCuts cuts (iStar rows of istarStride ints, omegaCnt[i] valid
 * entries each) x nReps resampled observation lists (observ = nReps rows of k).  Outputs alpha[nReps][nCuts] and
 * beta[nReps][nCuts][prevCols+1]. */
int  sdgpu_reform_cuts_batch(sdgpu_ctx *ctx, int nCuts, const int32_t *iStar, int istarStride, const int32_t *omegaCnt,
                             int nReps, const int32_t *observ, int k, int lbType, int lb, double *alpha, double *beta);

/* ---- readers for the host code that still walks the tables (optimal.c:203-221, cuts.c:472-513) ---- */
int  sdgpu_get_omega(sdgpu_ctx *ctx, int idx, double *vals /* [numRV+1] */, int *weight);
int  sdgpu_get_lambda(sdgpu_ctx *ctx, int idx, double *vals /* [rvRowCnt+1] */);
int  sdgpu_get_sigma(sdgpu_ctx *ctx, int idx, double *pib, double *piC /* [cntCcols+1] */, int *lambdaIdx, int *ck);
int  sdgpu_get_delta(sdgpu_ctx *ctx, int lambdaIdx, int obsIdx, double *pib, double *piC /* [rvCOmCnt+1] */);
/* one plane (0 = pib, 1..rvCOmCnt = piC) of the block [l0,l1) x [o0,o1) of delta, row-major [l1-l0][o1-o0] */
int  sdgpu_get_delta_block(sdgpu_ctx *ctx, int64_t l0, int64_t l1, int64_t o0, int64_t o1, int plane, double *out);
/* device-resident iStar of the most recent cut (int32[omegaCnt]) for callers that keep it on the GPU */
int  sdgpu_last_istar_device(sdgpu_ctx *ctx, void **devPtr, int *len);

/* ---- instrumentation ------------------------------------------------------------------------------ */
typedef struct {
	double  last_cut_ms;        /* device time of the most recent sd_cut (CUDA events on the ctx stream) */
	double  last_sweep_ms;      /* ... of its argmax sweep kernel alone                                 */
	int64_t last_cut_launches;  /* kernels launched by the most recent sd_cut                           */
	int64_t total_launches;     /* kernels launched since create                                        */
	int64_t last_sweep_bytes;   /* algorithmic bytes of the most recent sweep    */
	int64_t last_sweep_variant; /* 1 = LDG streaming, 2 = bulk-copy (UBLKCP) ring, 3 = per-term gathers (random cost), 4 = term-linear bulk-copy ring
	                             * (random cost), 5 = recompute from (lambda, omega), no delta stream (Rb <= 8), 6 = bulk-copy ring over bases grouped by lambda row */
	/* ABI 2: the split of last_cut_ms (same events; all zero unless sdgpu_set_timing is on) */
	double  last_prep_ms;       /* prologue: piCbarX + basis descriptors (0 when fused into the sweep)            */
	double  last_merge_ms;      /* merge + accumulate kernel (includes the NVLink peer exchange when that is the collective) */
	double  last_collective_ms; /* everything after the merge kernel: NCCL all-reduce + normalise, iStar copy    */
} sdgpu_stats;
int  sdgpu_get_stats(sdgpu_ctx *ctx, sdgpu_stats *out);
/* CUDA-event timing of the cut (last_cut_ms / last_sweep_ms) costs four event records per cut; off by default. */
int  sdgpu_set_timing(sdgpu_ctx *ctx, int on);
/* 0 = automatic by size (TMA bulk rings for large sweeps, LDG streaming / per-term gathers for small ones);
 * 1 forces the load-based kernels, 2 the TMA rings wherever one exists for the problem's shape, 3 the recompute sweep
 * where the problem qualifies (RHS-only, at most 8 random right-hand sides; automatic up to 4), 4 the grouped ring
 * (RHS-only; automatic when several bases share lambda rows). */
int  sdgpu_set_sweep_variant(sdgpu_ctx *ctx, int variant);
/* host-only: the 2-D sweep grid (observation tiles x basis chunks) the library would launch for a table of this shape on a
 * GPU with smCount SMs -- the wave-aware chunk count of DESIGN.md section 4; needs no device */
int  sdgpu_plan_sweep_grid(int smCount, int64_t observations, int64_t bases, int maxChunks, int *tiles, int *chunkSize, int *nChunks);
/* host-only: the sweep family (the value sdgpu_stats.last_sweep_variant would report) the library picks for a problem of this
 * shape and these table sizes under sdgpu_set_sweep_variant(variant), and whether the cut skips the separate prologue launch */
int  sdgpu_plan_sweep_kind(int rvCOmCnt, int rvdOmCnt, int rvbOmCnt, int maxPhiLength, int costColumns, int n1, int n1c,
                           int64_t bases, int64_t terms, int64_t distinctLambdaRows, int64_t observations, int variant, int *fusedPrologue);
/* micro-benchmark of the FP64 pipe for the instruction mix this library issues: separately rounded DMUL + DADD pairs (counted as two
 * operations per pair; the library never fuses them) and, for comparison, DFMA flops.  Best of `reps` launches.
 * The roofline denominator of the FP64-bound kernels (recompute sweep, bulk delta build). */
int  sdgpu_fp64_peak(int device, int reps, double *mulAddOpsPerSec, double *fmaFlopsPerSec);
/* The delta table (8 (1+Q) bytes x dual rows x observations, the one table whose CAPACITY can exceed the GPU) is allocated whole by default.
 * With the environment variable SDGPU_VMM=1 -- or automatically when the capacity passed to sdgpu_create does not fit the free device
 * memory -- the library reserves the table's address range and maps physical memory only under the part in use (512 dual rows of an
 * observation tile at a time; nothing is ever copied or moved).  An append that cannot get memory fails with SDGPU_ERR and a message.
 * Returns 1 when the table is mapped on demand, 0 when it was allocated whole; the byte counts are optional outputs. */
int  sdgpu_delta_memory(sdgpu_ctx *ctx, int64_t *reservedBytes, int64_t *mappedBytes);
/* instrumentation: median wall time (us) of `launches` empty kernels in a row followed by mode 0: a stream synchronise, mode 1: the host
 * spinning on a word the last kernel writes into mapped pinned memory -- the floor under every synchronous call of this library */
int  sdgpu_launch_roundtrip(int device, int mode, int launches, int reps, double *medianUs);
/* run subsequent work on an existing CUDA stream (cudaStream_t) instead of the context's own */
int  sdgpu_set_stream(sdgpu_ctx *ctx, void *cudaStream);

#ifdef __cplusplus
}
#endif
#endif /* SDGPU_H_ */
