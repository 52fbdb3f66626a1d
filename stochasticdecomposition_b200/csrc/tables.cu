// tables.cu -- device-resident stochastic tables of two-stage SD: omega, lambda, sigma, delta, basis records.
// Replaces the table half of the reference's stocUpdate.c.  Citations are file:line under
// /root/reference/twoSD_src.  Built with -fmad=false: every product and every sum is rounded separately and
// accumulated left to right exactly as the reference's scalar loops do, so table entries are bit-identical
// to the CPU path and the argmax downstream sees the same scores.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <climits>
#include <algorithm>

#include "sdgpu_internal.cuh"

thread_local std::string g_sdgpu_err;

int sdgpu_fail(const char *fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	g_sdgpu_err = buf;
	return SDGPU_ERR;
}

extern "C" const char *sdgpu_last_error(void) { return g_sdgpu_err.c_str(); }
extern "C" int sdgpu_abi_version(void) { return SDGPU_ABI_VERSION; }

// 8-byte asynchronous copies global -> shared (LDGSTS): a thread queues a whole batch of operands without holding a register
// per load, waits once, and reads them back from shared memory.  (With plain loads ptxas interleaves the loads with the
// dependent add chain, three or four deep, whatever the source says: one memory round trip per few terms.)
__device__ __forceinline__ void sd_cp_async8(double *smemDst, const double *gmemSrc) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t) __cvta_generic_to_shared(smemDst)), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void sd_cp_async_wait_all() {
	asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void sd_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` (0 or 1) committed groups of this thread are still in flight
__device__ __forceinline__ void sd_cp_async_wait(bool onePending) {
	if (onePending) asm volatile("cp.async.wait_group 1;" ::: "memory");
	else asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// A small host vector (observation, dual vector) travelling as a KERNEL PARAMETER: no DMA, and no zero-copy read over PCIe by every
// block either (~2 us per kernel at real-problem sizes).  Longer vectors go through the mapped staging buffer / one DMA copy.
#define SD_VEC_PARAM_MAX 640
struct SdVecParam { double v[SD_VEC_PARAM_MAX]; };

// DBL_ABS of the reference: (x) > 0 ? (x) : -(x); a NaN difference therefore never counts as a mismatch
__device__ __forceinline__ double sd_abs(double x) { return x > 0.0 ? x : -x; }

// ======================================================================================================
// kernels
// ======================================================================================================

// ---- fused find-or-append kernels (one launch per table, no host round trip) ---------------------------------
// Every block scans its rows; the block that draws the last ticket commits the result and publishes the device
// state into mapped pinned host memory, so the host only has to wait for the stream.
__device__ __forceinline__ void sd_publish(const SdDevState *st, SdDevState *host) {
	*host = *st;           // mapped pinned memory; the host reads it after it has waited for the stream, which makes the write visible
}

// calcLambda (stocUpdate.c:264-284) in one launch, followed in the same launch by the staging part of calcSigma
// (stocUpdate.c:293-296) so that the sigma scan can start right after.  `pi` may be a mapped host pointer.
__global__ void k_lambda_fused(const double *__restrict__ pi, int rows, const int32_t *__restrict__ rvRows, int R,
		double *__restrict__ lambda, int64_t LP, int64_t cap, double tol,
		const int32_t *__restrict__ bCol, const double *__restrict__ bVal, int bCnt, double mubBar,
		const int32_t *__restrict__ cbStart, const int32_t *__restrict__ cbRow, const double *__restrict__ cbVal, int n1c, int cbStage,
		double *__restrict__ vecDev, double *__restrict__ candC, SdDevState *st, SdDevState *hst, int publish) {
	extern __shared__ double s_cand[];              // [R] reduced candidate, [rows + 1] the whole vector, [bCnt] and [cbStage] products
	double *s_pi = s_cand + R;
	double *s_prod = s_pi + rows + 1;
	double *s_cprod = s_prod + bCnt;
	for (int i = threadIdx.x; i <= rows; i += blockDim.x) s_pi[i] = pi[i];                  // one trip to the (possibly host-mapped) vector
	__syncthreads();
	for (int i = threadIdx.x; i < R; i += blockDim.x) s_cand[i] = s_pi[rvRows[i]];          // reduceVector :269
	__syncthreads();
	const int cnt = st->lambdaCnt;
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < cnt) {                                                                           // :272-277
		bool same = true;
		for (int i = 0; i < R; i++)
			if (sd_abs(s_cand[i] - lambda[(size_t) i * LP + r]) > tol) { same = false; break; }
		if (same) atomicMin(&st->foundLambda, r);
	}
	if (!sd_is_last_block(&st->ticket)) return;
	const int found = *(volatile int *) &st->foundLambda;
	int idx = found, isNew = 0;
	if (found == INT_MAX) {
		if (cnt >= cap) { idx = -1; if (threadIdx.x == 0) st->overflow = 1; }
		else {
			for (int i = threadIdx.x; i < R; i += blockDim.x) lambda[(size_t) i * LP + cnt] = s_cand[i];   // :280
			idx = cnt; isNew = 1;
		}
	}
	// calcSigma's candidate from the same vector.  The products bBar.val[e] * pi[bBar.col[e]] and pi[Cbar.row[e]] * Cbar.val[e] are
	// formed in parallel (each is rounded on its own in the reference too) and then added up left to right per sum, so the sums
	// have the reference's bits without one dependent memory round trip per term.
	for (int i = threadIdx.x; i <= rows; i += blockDim.x) vecDev[i] = s_pi[i];              // keep pi on the device for calcSigma
	for (int e = threadIdx.x; e < bCnt; e += blockDim.x) s_prod[e] = __dmul_rn(bVal[e], s_pi[bCol[e]]);
	const int nnz = cbStart[n1c];
	const bool staged = nnz <= cbStage;
	if (staged) for (int e = threadIdx.x; e < nnz; e += blockDim.x) s_cprod[e] = __dmul_rn(s_pi[cbRow[e]], cbVal[e]);
	__syncthreads();
	for (int k = threadIdx.x; k < n1c; k += blockDim.x) {                                    // vxMSparse + reduceVector :295-296
		double t = 0.0;
		if (staged) for (int e = cbStart[k]; e < cbStart[k + 1]; e++) t = __dadd_rn(t, s_cprod[e]);
		else for (int e = cbStart[k]; e < cbStart[k + 1]; e++) t = __dadd_rn(t, __dmul_rn(s_pi[cbRow[e]], cbVal[e]));
		candC[k] = t;
	}
	if (threadIdx.x == 0) {
		double sum = 0.0;                                                                    // vXvSparse :293
		for (int e = 0; e < bCnt; e++) sum = __dadd_rn(sum, s_prod[e]);
		st->pibBar = __dadd_rn(sum, mubBar);
		st->lambdaIdx = idx; st->newLambda = isNew; st->foundLambda = INT_MAX;
		if (isNew) st->lambdaCnt = cnt + 1;
		if (publish) sd_publish(st, hst);       // chained with calcSigma (sdgpu_update_dual): the sigma kernel publishes the whole state
	}
}

// calcSigma scan + commit (stocUpdate.c:299-318) in one launch
__global__ void k_sigma_fused(double *__restrict__ pib, double *__restrict__ piCk, double *__restrict__ piCr, int32_t *__restrict__ lam,
		int32_t *__restrict__ ck, int64_t SP, int n1c, int n1cP, const double *__restrict__ candC, double tol, int iter, int64_t cap,
		SdDevState *st, SdDevState *hst) {
	const int cnt = st->sigmaCnt;
	if (!st->newLambda) {
		const int s = blockIdx.x * blockDim.x + threadIdx.x;
		if (s < cnt && sd_abs(st->pibBar - pib[s]) <= tol) {
			bool same = true;
			for (int k = 0; k < n1c; k++)
				if (sd_abs(candC[k] - piCk[(size_t) k * SP + s]) > tol) { same = false; break; }
			if (same && lam[s] == st->lambdaIdx) atomicMin(&st->foundSigma, s);
		}
	}
	if (!sd_is_last_block(&st->ticket)) return;
	const int found = *(volatile int *) &st->foundSigma;
	int idx = found, isNew = 0;
	if (found == INT_MAX) {
		if (cnt >= cap || st->lambdaIdx < 0) { idx = -1; if (threadIdx.x == 0) st->overflow = 1; }
		else {
			for (int k = threadIdx.x; k < n1c; k += blockDim.x) {
				piCk[(size_t) k * SP + cnt] = candC[k];
				piCr[(size_t) cnt * n1cP + k] = candC[k];
			}
			idx = cnt; isNew = 1;
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (isNew) { pib[cnt] = st->pibBar; lam[cnt] = st->lambdaIdx; ck[cnt] = iter; st->sigmaCnt = cnt + 1; }
		st->sigmaIdx = idx; st->newSigma = isNew; st->foundSigma = INT_MAX;
		sd_publish(st, hst);
	}
}

// calcOmega (stocUpdate.c:326-348) in one launch.  mode 0: find, then bump or append; 1: find only; 2: append only
__global__ void k_omega_fused(const double *__restrict__ observ, SdVecParam vp, int numRV, double *__restrict__ omega, int32_t *__restrict__ w, int64_t NP,
		int64_t cap, double tol, int mode, int weight, SdDevState *st, SdDevState *hst) {
	extern __shared__ double s_cand[];
	for (int j = threadIdx.x; j < numRV; j += blockDim.x) s_cand[j] = observ ? observ[1 + j] : vp.v[1 + j];
	__syncthreads();
	const int cnt = st->omegaCnt;
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (mode != 2 && r < cnt) {                                                              // :330-335
		bool same = true;
		for (int j = 0; j < numRV; j++)
			if (sd_abs(s_cand[j] - omega[(size_t) j * NP + r]) > tol) { same = false; break; }
		if (same) atomicMin(&st->foundOmega, r);
	}
	if (!sd_is_last_block(&st->ticket)) return;
	const int found = *(volatile int *) &st->foundOmega;
	if (mode == 1) {
		if (threadIdx.x == 0) { st->omegaIdx = found == INT_MAX ? -1 : found; st->newOmega = 0; sd_publish(st, hst); st->foundOmega = INT_MAX; }
		return;
	}
	int idx = found, isNew = 0;
	if (found == INT_MAX) {
		if (cnt >= cap) { idx = -1; if (threadIdx.x == 0) st->overflow = 1; }
		else {
			for (int j = threadIdx.x; j < numRV; j += blockDim.x) omega[(size_t) j * NP + cnt] = s_cand[j];   // :338
			idx = cnt; isNew = 1;
		}
	}
	if (threadIdx.x == 0) {
		if (isNew) { w[cnt] = weight; st->omegaCnt = cnt + 1; }
		else if (idx >= 0) w[idx] += 1;                                                      // :333
		st->omegaIdx = idx; st->newOmega = isNew; st->foundOmega = INT_MAX;
		sd_publish(st, hst);
	}
}

struct SdBasisRec { int b, ck, feas, phiLen, termStart, nT; int sigma[16]; int omega[16]; };

__global__ void k_basis_commit(SdBasisRec r, int32_t *bCk, int32_t *bFeas, int32_t *bPhiLen, int32_t *bTermStart, int32_t *tSigma, int32_t *tOmega,
		SdDevState *st, SdDevState *hst) {
	bCk[r.b] = r.ck; bFeas[r.b] = r.feas; bPhiLen[r.b] = r.phiLen;
	bTermStart[r.b] = r.termStart; bTermStart[r.b + 1] = r.termStart + r.nT;
	for (int t = 0; t < r.nT; t++) { tSigma[r.termStart + t] = r.sigma[t]; tOmega[r.termStart + t] = r.omega[t]; }
	st->basisCnt = r.b + 1;
	sd_publish(st, hst);
}

// calcSigma stocUpdate.c:293-296: pibBar = (sum_e bBar.val[e] * pi[bBar.col[e]]) + mubBar, and for each kept column
// CCols[k] the entries of pi x Cbar that land in it, accumulated in nnz order from 0.0.
__global__ void k_sigma_prepare(const double *__restrict__ pi, const int32_t *__restrict__ bCol, const double *__restrict__ bVal,
		int bCnt, double mubBar, const int32_t *__restrict__ cbStart, const int32_t *__restrict__ cbRow,
		const double *__restrict__ cbVal, int n1c, double *__restrict__ candC, SdDevState *st,
		int overrideNewLambda, int overrideLambdaIdx) {
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n1c) {
		double t = 0.0;
		for (int e = cbStart[k]; e < cbStart[k + 1]; e++) t += pi[cbRow[e]] * cbVal[e];
		candC[k] = t;
	}
	if (k == n1c || (n1c == 0 && k == 0)) {
		double s = 0.0;
		for (int e = 0; e < bCnt; e++) s += bVal[e] * pi[bCol[e]];
		st->pibBar = s + mubBar;
		st->foundSigma = INT_MAX;
		if (overrideNewLambda >= 0) { st->newLambda = overrideNewLambda; st->lambdaIdx = overrideLambdaIdx; }
	}
}

// One delta cell, the arithmetic of stocUpdate.c:218-221 / :244-247:
//   pib    = sum_{j=1..Rb} omega_o[j] * lambdaFull[rvbOmRows[j]]          (vXvSparse, index order)
//   piC[c] = sum over the e with rvCOmCols[e] == rvCOmCols[c], in e order, of lambdaFull[rvCOmRows[e]] * omega_o[Rb+e]
// lambdaFull is the lambda expanded to all rows, zero where the row carries no lambda entry (expandVector).
// LamAt(p) / OmAt(j) fetch lambda position p and observation entry j (0-based) for the cell at hand.
// The random-T planes of one cell (Q is small): piC[c] as above.
template <class LamAt, class OmAt>
__device__ __forceinline__ void sd_delta_cell_T(int Rb, int Q, const int32_t *__restrict__ cLamPos, const int32_t *__restrict__ cListStart,
		const int32_t *__restrict__ cList, LamAt lamAt, OmAt omAt, double *__restrict__ out, size_t planeStride) {
	for (int c = 0; c < Q; c++) {
		double t = 0.0;
		for (int n = cListStart[c]; n < cListStart[c + 1]; n++) {
			int e = cList[n], p = cLamPos[e];
			t += (p >= 0 ? lamAt(p) : 0.0) * omAt(Rb + e);
		}
		out[(size_t) (1 + c) * planeStride] = t;
	}
}

#define DC_THREADS 128     // threads per CTA of the row / column kernels
#define DC_BATCH   32      // operands a thread keeps in flight (DC_BATCH x DC_THREADS x 8 bytes of shared memory)

// calcDelta case II stocUpdate.c:230-254: a new dual -> one delta row, one thread per observation (coalesced
// reads of omega, contiguous W-segment writes).  The lambda row sits in shared memory.
__global__ void __launch_bounds__(DC_THREADS) k_delta_row(const double *__restrict__ lambda, int64_t LP, int R, const double *__restrict__ omega, int64_t NP,
		int Rb, int Q, const int32_t *__restrict__ bLamPos, const int32_t *__restrict__ cLamPos, const int32_t *__restrict__ cListStart,
		const int32_t *__restrict__ cList, double *__restrict__ delta, int64_t Dcap, const SdDevState *st, int forcedRow) {
	__shared__ double s_buf[DC_BATCH][DC_THREADS];
	extern __shared__ double s_lam[];            // [Rb] the lambda entry each random RHS row meets (0.0 where it meets none), then [R] the row itself
	double *s_row = s_lam + Rb;
	sd_pdl_wait();                               // (launched as the programmatic dependent of the find-or-append kernel)
	int l = forcedRow >= 0 ? forcedRow : (st->newLambda ? st->lambdaIdx : -1);
	if (l < 0) return;
	for (int j = threadIdx.x; j < Rb; j += blockDim.x) { const int p = bLamPos[j]; s_lam[j] = p >= 0 ? lambda[(size_t) p * LP + l] : 0.0; }
	for (int i = threadIdx.x; i < R; i += blockDim.x) s_row[i] = lambda[(size_t) i * LP + l];
	__syncthreads();
	int64_t o = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (o >= st->omegaCnt) return;
	double s = 0.0;                                                                        // vXvSparse :244, index order
	// two half-batches alternate: the loads of group g+1 are in flight while the (sequential) adds of group g run
	constexpr int H = DC_BATCH / 2;
	const int G = (Rb + H - 1) / H;
	auto issue = [&](int g) {
		const int j0 = g * H, n = min(H, Rb - j0);
		for (int u = 0; u < n; u++) sd_cp_async8(&s_buf[(g & 1) * H + u][threadIdx.x], omega + (size_t) (j0 + u) * NP + o);
		sd_cp_async_commit();
	};
	if (G > 0) issue(0);
	for (int g = 0; g < G; g++) {
		if (g + 1 < G) issue(g + 1);
		sd_cp_async_wait(g + 1 < G);
		const int j0 = g * H, n = min(H, Rb - j0);
		for (int u = 0; u < n; u++) s = __dadd_rn(s, __dmul_rn(s_buf[(g & 1) * H + u][threadIdx.x], s_lam[j0 + u]));
	}
	double *out = delta + sd_delta_off(Dcap, Q, l, 0, o);
	out[0] = s;
	sd_delta_cell_T(Rb, Q, cLamPos, cListStart, cList, [&](int p) { return s_row[p]; }, [&](int j) { return omega[(size_t) j * NP + o]; }, out, SD_TILE_W);
}

// calcDelta case I stocUpdate.c:206-229: a new observation -> one delta column, one thread per dual (coalesced
// reads of lambda, strided 8-byte writes).  The observation sits in shared memory.
__global__ void __launch_bounds__(DC_THREADS) k_delta_col(const double *__restrict__ lambda, int64_t LP, const double *__restrict__ omega, int64_t NP, int numRV,
		int Rb, int Q, const int32_t *__restrict__ bLamPos, const int32_t *__restrict__ cLamPos, const int32_t *__restrict__ cListStart,
		const int32_t *__restrict__ cList, double *__restrict__ delta, int64_t Dcap, const SdDevState *st, int forcedCol) {
	__shared__ double s_buf[DC_BATCH][DC_THREADS];
	extern __shared__ double s_om[];             // [numRV] the observation, then [Rb] ints: position of each random RHS row inside a lambda
	int32_t *s_pos = reinterpret_cast<int32_t *>(s_om + numRV);
	int o = forcedCol >= 0 ? forcedCol : (st->newOmega ? st->omegaIdx : -1);
	if (o < 0) return;
	for (int j = threadIdx.x; j < numRV; j += blockDim.x) s_om[j] = omega[(size_t) j * NP + o];
	for (int j = threadIdx.x; j < Rb; j += blockDim.x) s_pos[j] = bLamPos[j];
	__syncthreads();
	int64_t l = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (l >= st->lambdaCnt) return;
	double s = 0.0;                                                                        // vXvSparse :218, index order
	// two half-batches alternate: the loads of group g+1 are in flight while the (sequential) adds of group g run
	constexpr int H = DC_BATCH / 2;
	const int G = (Rb + H - 1) / H;
	auto issue = [&](int g) {
		const int j0 = g * H, n = min(H, Rb - j0);
		for (int u = 0; u < n; u++) { const int p = s_pos[j0 + u]; if (p >= 0) sd_cp_async8(&s_buf[(g & 1) * H + u][threadIdx.x], lambda + (size_t) p * LP + l); }
		sd_cp_async_commit();
	};
	if (G > 0) issue(0);
	for (int g = 0; g < G; g++) {
		if (g + 1 < G) issue(g + 1);
		sd_cp_async_wait(g + 1 < G);
		const int j0 = g * H, n = min(H, Rb - j0);
		for (int u = 0; u < n; u++) s = __dadd_rn(s, __dmul_rn(s_om[j0 + u], s_pos[j0 + u] >= 0 ? s_buf[(g & 1) * H + u][threadIdx.x] : 0.0));
	}
	double *out = delta + sd_delta_off(Dcap, Q, l, 0, o);
	out[0] = s;
	sd_delta_cell_T(Rb, Q, cLamPos, cListStart, cList, [&](int p) { return lambda[(size_t) p * LP + l]; }, [&](int j) { return s_om[j]; }, out, SD_TILE_W);
}

// ---- one launch for stocUpdate.c:24-25 + :78-81: the delta column of a new observation (blocks [nbScan, gridDim.x)) next to the
// lambda scan (blocks [0, nbScan)); the block that draws the last ticket commits lambda (:280-283), forms calcSigma's candidate
// (:293-296), scans sigma by itself (:299-310 -- a pib compare per row, the piC compare only where pib matches), commits sigma
// (:312-318) and publishes the state.  The delta ROW of a new lambda follows as a second, programmatically dependent launch.
// Same arithmetic per element as k_lambda_fused / k_sigma_fused / k_delta_col, hence the same bits.
#define UF_THREADS 256
#define UF_COLB    16      // column role: operands a thread keeps in flight
struct UpdArgs {
	const double *pi; int rows; const int32_t *rvRows; int R; double *lambda; int64_t LP, lambdaCap; double tol;
	const int32_t *bCol; const double *bVal; int bCnt; double mubBar;
	const int32_t *cbStart, *cbRow; const double *cbVal; int n1c, n1cP, cbStage;
	double *vecDev;
	double *sigPib, *sigPiCk, *sigPiCr; int32_t *sigLam, *sigCk; int64_t SP, sigmaCap; int iter;
	int nbScan, colObs, lambdaCntHost;
	const double *omega; int64_t NP; int numRV, Rb, Q; const int32_t *bLamPos, *cLamPos, *cListStart, *cList; double *delta; int64_t Dcap;
	SdDevState *st, *hst;
};

__global__ void __launch_bounds__(UF_THREADS) k_update_fused(UpdArgs a, SdVecParam vp) {
	extern __shared__ double s_dyn[];
	__shared__ int s_found;
	sd_pdl_launch_dependents();                               // the delta-row kernel may be scheduled; it waits for this grid
	const int tid = threadIdx.x;
	if ((int) blockIdx.x >= a.nbScan) {
		// ------------- delta column (calcDelta case I, stocUpdate.c:206-229): one thread per stored lambda row -------------------
		double *s_om = s_dyn;                                 // [numRV] the observation
		int32_t *s_pos = reinterpret_cast<int32_t *>(s_om + a.numRV);                 // [Rb]
		double *s_buf = reinterpret_cast<double *>(s_pos + ((a.Rb + 1) & ~1));        // [UF_COLB][UF_THREADS]
		const int o = a.colObs;
		for (int j = tid; j < a.numRV; j += blockDim.x) s_om[j] = a.omega[(size_t) j * a.NP + o];
		for (int j = tid; j < a.Rb; j += blockDim.x) s_pos[j] = a.bLamPos[j];
		__syncthreads();
		const int64_t l = (int64_t) (blockIdx.x - a.nbScan) * blockDim.x + tid;
		if (l < a.lambdaCntHost) {                            // rows stored BEFORE this call (a row appended by it gets its whole delta row next)
			double sum = 0.0;                                 // vXvSparse :218, index order; two half-batches alternate (see k_delta_col)
			constexpr int H = UF_COLB / 2;
			const int G = (a.Rb + H - 1) / H;
			auto issue = [&](int g) {
				const int j0 = g * H, n = min(H, a.Rb - j0);
				for (int u = 0; u < n; u++) { const int p = s_pos[j0 + u]; if (p >= 0) sd_cp_async8(&s_buf[((g & 1) * H + u) * UF_THREADS + tid], a.lambda + (size_t) p * a.LP + l); }
				sd_cp_async_commit();
			};
			if (G > 0) issue(0);
			for (int g = 0; g < G; g++) {
				if (g + 1 < G) issue(g + 1);
				sd_cp_async_wait(g + 1 < G);
				const int j0 = g * H, n = min(H, a.Rb - j0);
				for (int u = 0; u < n; u++) sum = __dadd_rn(sum, __dmul_rn(s_om[j0 + u], s_pos[j0 + u] >= 0 ? s_buf[((g & 1) * H + u) * UF_THREADS + tid] : 0.0));
			}
			double *out = a.delta + sd_delta_off(a.Dcap, a.Q, l, 0, o);
			out[0] = sum;
			sd_delta_cell_T(a.Rb, a.Q, a.cLamPos, a.cListStart, a.cList, [&](int p) { return a.lambda[(size_t) p * a.LP + l]; }, [&](int j) { return s_om[j]; }, out, SD_TILE_W);
		}
	}
	else {
		// ------------- lambda scan (calcLambda, stocUpdate.c:269-277): one thread per stored row -----------------------------------
		double *s_cand = s_dyn, *s_pi = s_dyn + a.R;
		for (int i = tid; i <= a.rows; i += blockDim.x) s_pi[i] = a.pi ? a.pi[i] : vp.v[i];        // kernel parameter, or one trip to the staged vector
		__syncthreads();
		if (blockIdx.x == 0) for (int i = tid; i <= a.rows; i += blockDim.x) a.vecDev[i] = s_pi[i];     // for the committing block, whichever it is
		for (int i = tid; i < a.R; i += blockDim.x) s_cand[i] = s_pi[a.rvRows[i]];      // reduceVector :269
		__syncthreads();
		const int r = blockIdx.x * blockDim.x + tid;
		if (r < a.lambdaCntHost) {
			bool same = true;
			for (int i = 0; i < a.R; i++)
				if (sd_abs(s_cand[i] - a.lambda[(size_t) i * a.LP + r]) > a.tol) { same = false; break; }
			if (same) atomicMin(&a.st->foundLambda, r);
		}
	}
	if (!sd_is_last_block(&a.st->ticket)) return;

	// ------------- commit (one block; every other block's writes are visible after the ticket) ---------------------------------------
	double *s_cand = s_dyn, *s_pi = s_cand + a.R, *s_prod = s_pi + a.rows + 1, *s_cprod = s_prod + a.bCnt, *s_candC = s_cprod + a.cbStage;
	for (int i = tid; i <= a.rows; i += blockDim.x) s_pi[i] = __ldcg(a.vecDev + i);
	if (tid == 0) s_found = INT_MAX;
	__syncthreads();
	for (int i = tid; i < a.R; i += blockDim.x) s_cand[i] = s_pi[a.rvRows[i]];
	__syncthreads();
	const int cnt = a.lambdaCntHost;
	const int found = *(volatile int *) &a.st->foundLambda;
	int idx = found, isNew = 0;
	if (found == INT_MAX) {                                                             // :280-283
		if (cnt >= a.lambdaCap) { idx = -1; if (tid == 0) a.st->overflow = 1; }
		else {
			for (int i = tid; i < a.R; i += blockDim.x) a.lambda[(size_t) i * a.LP + cnt] = s_cand[i];
			idx = cnt; isNew = 1;
		}
	}
	// calcSigma's candidate (:293-296): products in parallel, sums left to right per sum (see k_lambda_fused)
	for (int e = tid; e < a.bCnt; e += blockDim.x) s_prod[e] = __dmul_rn(a.bVal[e], s_pi[a.bCol[e]]);
	const int nnz = a.cbStart[a.n1c];
	const bool staged = nnz <= a.cbStage;
	if (staged) for (int e = tid; e < nnz; e += blockDim.x) s_cprod[e] = __dmul_rn(s_pi[a.cbRow[e]], a.cbVal[e]);
	__syncthreads();
	for (int k = tid; k < a.n1c; k += blockDim.x) {
		double t = 0.0;
		if (staged) for (int e = a.cbStart[k]; e < a.cbStart[k + 1]; e++) t = __dadd_rn(t, s_cprod[e]);
		else for (int e = a.cbStart[k]; e < a.cbStart[k + 1]; e++) t = __dadd_rn(t, __dmul_rn(s_pi[a.cbRow[e]], a.cbVal[e]));
		s_candC[k] = t;
	}
	__shared__ double s_pibBar;
	if (tid == 0) {
		double sum = 0.0;                                                               // vXvSparse :293
		for (int e = 0; e < a.bCnt; e++) sum = __dadd_rn(sum, s_prod[e]);
		s_pibBar = __dadd_rn(sum, a.mubBar);
	}
	__syncthreads();
	const double pibBar = s_pibBar;
	const int scnt = a.st->sigmaCnt;
	if (!isNew && idx >= 0) {                                                           // :299-310
		for (int sg = tid; sg < scnt; sg += blockDim.x) {
			if (!(sd_abs(pibBar - a.sigPib[sg]) <= a.tol)) continue;
			bool same = true;
			for (int k = 0; k < a.n1c; k++)
				if (sd_abs(s_candC[k] - a.sigPiCk[(size_t) k * a.SP + sg]) > a.tol) { same = false; break; }
			if (same && a.sigLam[sg] == idx) atomicMin(&s_found, sg);
		}
	}
	__syncthreads();
	const int sfound = s_found;
	int sidx = sfound, sNew = 0;
	if (sfound == INT_MAX) {                                                            // :312-318
		if (scnt >= a.sigmaCap || idx < 0) { sidx = -1; if (tid == 0) a.st->overflow = 1; }
		else {
			for (int k = tid; k < a.n1c; k += blockDim.x) {
				a.sigPiCk[(size_t) k * a.SP + scnt] = s_candC[k];
				a.sigPiCr[(size_t) scnt * a.n1cP + k] = s_candC[k];
			}
			sidx = scnt; sNew = 1;
		}
	}
	__syncthreads();
	if (tid == 0) {
		SdDevState *st = a.st;
		st->pibBar = pibBar;
		st->lambdaIdx = idx; st->newLambda = isNew; st->foundLambda = INT_MAX;
		if (isNew) st->lambdaCnt = cnt + 1;
		if (sNew) { a.sigPib[scnt] = pibBar; a.sigLam[scnt] = idx; a.sigCk[scnt] = a.iter; st->sigmaCnt = scnt + 1; }
		st->sigmaIdx = sidx; st->newSigma = sNew; st->foundSigma = INT_MAX;
		sd_publish(st, a.hst);
	}
}

// calcDelta for a whole block of (dual, observation) pairs -- the bulk loader of synthetic sweeps.  A CTA owns
// 32 duals x 128 observations; lambda and omega slices for the block are staged in shared memory Rb-chunk by
// Rb-chunk and every thread carries 4 duals x 4 observations of running sums, each still accumulated in index
// order j = 1..Rb with separate multiply and add, so the result is bit-identical to the one-at-a-time appends.
#define DB_L 32
#define DB_O 128
#define DB_K 32
__global__ void __launch_bounds__(256) k_delta_block_rhs(const double *__restrict__ lambda, int64_t LP, const double *__restrict__ omega,
		int64_t NP, int Rb, const int32_t *__restrict__ bLamPos, double *__restrict__ delta, int64_t Dcap, int Q,
		int64_t l0, int64_t l1, int64_t o0, int64_t o1) {
	__shared__ double s_l[DB_K][DB_L + 1];
	__shared__ double s_o[DB_K][DB_O];
	int64_t lb = l0 + (int64_t) blockIdx.y * DB_L, ob = o0 + (int64_t) blockIdx.x * DB_O;
	int tx = threadIdx.x % 32, ty = threadIdx.x / 32;      // tx -> 4 observations (strided by 32), ty -> 4 duals (strided by 8)
	double acc[4][4];
#pragma unroll
	for (int a = 0; a < 4; a++)
#pragma unroll
		for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
	for (int j0 = 0; j0 < Rb; j0 += DB_K) {
		int jn = min(DB_K, Rb - j0);
		for (int n = threadIdx.x; n < DB_K * DB_L; n += 256) {
			int j = n / DB_L, li = n % DB_L;
			double v = 0.0;
			if (j < jn && lb + li < l1) { int p = bLamPos[j0 + j]; v = p >= 0 ? lambda[(size_t) p * LP + lb + li] : 0.0; }
			s_l[j][li] = v;
		}
		for (int n = threadIdx.x; n < DB_K * DB_O; n += 256) {
			int j = n / DB_O, oi = n % DB_O;
			s_o[j][oi] = (j < jn && ob + oi < o1) ? omega[(size_t) (j0 + j) * NP + ob + oi] : 0.0;
		}
		__syncthreads();
		for (int j = 0; j < jn; j++) {
			double lv[4], ov[4];
#pragma unroll
			for (int a = 0; a < 4; a++) lv[a] = s_l[j][ty + 8 * a];
#pragma unroll
			for (int b = 0; b < 4; b++) ov[b] = s_o[j][tx + 32 * b];
#pragma unroll
			for (int a = 0; a < 4; a++)
#pragma unroll
				for (int b = 0; b < 4; b++) acc[a][b] = __dadd_rn(acc[a][b], __dmul_rn(ov[b], lv[a]));
		}
		__syncthreads();
	}
#pragma unroll
	for (int a = 0; a < 4; a++)
#pragma unroll
		for (int b = 0; b < 4; b++) {
			int64_t l = lb + ty + 8 * a, o = ob + tx + 32 * b;
			if (l < l1 && o < o1) delta[sd_delta_off(Dcap, Q, l, 0, o)] = acc[a][b];
		}
}

// the piC planes of a block (Q is small): one thread per (dual, observation) pair
__global__ void k_delta_block_T(const double *__restrict__ lambda, int64_t LP, const double *__restrict__ omega, int64_t NP, int Rb, int Q,
		const int32_t *__restrict__ cLamPos, const int32_t *__restrict__ cListStart, const int32_t *__restrict__ cList,
		double *__restrict__ delta, int64_t Dcap, int64_t l0, int64_t l1, int64_t o0, int64_t o1) {
	int64_t o = o0 + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t l = l0 + blockIdx.y;
	if (o >= o1 || l >= l1) return;
	for (int c = 0; c < Q; c++) {
		double t = 0.0;
		for (int n = cListStart[c]; n < cListStart[c + 1]; n++) {
			int e = cList[n], p = cLamPos[e];
			t += (p >= 0 ? lambda[(size_t) p * LP + l] : 0.0) * omega[(size_t) (Rb + e) * NP + o];
		}
		delta[sd_delta_off(Dcap, Q, l, 1 + c, o)] = t;
	}
}

// bulk loaders (no dedup scan): n observations / n dual vectors straight into the tables
__global__ void k_omega_bulk(const double *__restrict__ vals, int64_t n, int numRV, const int32_t *__restrict__ weights,
		double *__restrict__ omega, int32_t *__restrict__ w, int64_t NP, int64_t base) {
	int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double *v = vals + (size_t) i * (numRV + 1);
	for (int j = 0; j < numRV; j++) omega[(size_t) j * NP + base + i] = v[1 + j];
	w[base + i] = weights ? weights[i] : 1;
}

__global__ void k_dual_bulk(const double *__restrict__ pis, int64_t n, int rows, const double *__restrict__ mub, const int32_t *__restrict__ iters,
		const int32_t *__restrict__ rvRows, int R, const int32_t *__restrict__ bCol, const double *__restrict__ bVal, int bCnt,
		const int32_t *__restrict__ cbStart, const int32_t *__restrict__ cbRow, const double *__restrict__ cbVal, int n1c, int n1cP,
		double *__restrict__ lambda, int64_t LP, double *__restrict__ pib, double *__restrict__ piCk, double *__restrict__ piCr,
		int32_t *__restrict__ lam, int32_t *__restrict__ ck, int64_t SP, int64_t lbase, int64_t sbase) {
	int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double *pi = pis + (size_t) i * (rows + 1);
	for (int r = 0; r < R; r++) lambda[(size_t) r * LP + lbase + i] = pi[rvRows[r]];
	double s = 0.0;
	for (int e = 0; e < bCnt; e++) s += bVal[e] * pi[bCol[e]];
	pib[sbase + i] = s + (mub ? mub[i] : 0.0);
	for (int k = 0; k < n1c; k++) {
		double t = 0.0;
		for (int e = cbStart[k]; e < cbStart[k + 1]; e++) t += pi[cbRow[e]] * cbVal[e];
		piCk[(size_t) k * SP + sbase + i] = t;
		piCr[(size_t) (sbase + i) * n1cP + k] = t;
	}
	lam[sbase + i] = (int32_t) (lbase + i);
	ck[sbase + i] = iters ? iters[i] : (int32_t) (i + 1);
}

__global__ void k_bump_counts(SdDevState *st, SdDevState *hst, int dOmega, int dLambda, int dSigma, int dBasis) {
	st->omegaCnt += dOmega; st->lambdaCnt += dLambda; st->sigmaCnt += dSigma; st->basisCnt += dBasis;
	sd_publish(st, hst);
}

__global__ void k_record_pair(const SdDevState *st, int32_t *lamOut, int32_t *sigOut, int64_t i) {
	if (lamOut) lamOut[i] = st->lambdaIdx;
	if (sigOut) sigOut[i] = st->sigmaIdx;
}

__global__ void k_omega_bump(int32_t *w, int idx, int by) { w[idx] += by; }

// the mask is bit-packed: one thread per 32-bit word of the basis' row (observations 32 j .. 32 j + 31); flags beyond n keep their bits
__global__ void k_mask_fill_row(uint32_t *mask, int64_t Bcap, int64_t b, int64_t NP, const uint8_t *flags, int64_t n, int fillAll) {
	const int64_t j = (int64_t) blockIdx.x * blockDim.x + threadIdx.x, o0 = j * 32;
	if (o0 >= NP) return;
	uint32_t *w = mask + sd_mask_word(Bcap, b, o0);
	if (fillAll) { *w = 0xffffffffu; return; }
	if (o0 >= n) return;
	uint32_t v = *w;
	for (int i = 0; i < 32 && o0 + i < n; i++) v = flags[o0 + i] ? (v | (1u << i)) : (v & ~(1u << i));
	*w = v;
}

// one observation, every feasible basis: thread b owns the word of (b, o) (no other thread of this launch touches it)
__global__ void k_mask_fill_col(uint32_t *mask, int64_t Bcap, int64_t o, const uint8_t *flags, const int32_t *feas, int64_t nb) {
	int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (b < nb && feas[b]) {
		uint32_t *w = mask + sd_mask_word(Bcap, b, o);
		const uint32_t bit = 1u << (o & 31);
		*w = flags[b] ? (*w | bit) : (*w & ~bit);
	}
}

__global__ void k_mask_set_bit(uint32_t *mask, int64_t Bcap, int64_t b, int64_t o, int v) {
	uint32_t *w = mask + sd_mask_word(Bcap, b, o);
	const uint32_t bit = 1u << (o & 31);
	*w = v ? (*w | bit) : (*w & ~bit);
}

// checkBasisFeasibility randCost.c:202-258 for one (basis, observation) pair, evaluated by a whole CTA: every row's
// reconstructed dual (piDet + sum_n phi_n * dOmega) against its sense and every column's reduced cost
// (gBar + dOmega - psi * dOmega) against its bound status; the pair is feasible iff no thread finds a violation.
struct FeasArgs {
	const double *omega; int64_t NP; int rvOffset2, rvd;
	const int32_t *rvdOmCols; const char *senx; int rows, cols;
	const int32_t *bPhiLen, *bTermStart, *bFeas, *tOmega;
	const double *piDet, *phi, *gBar, *psi; const int8_t *cstat; const uint8_t *has;
	double tol; uint32_t *mask; int64_t Bcap; uint8_t *flags;
};

__device__ __forceinline__ bool sd_pair_violates(const FeasArgs &a, int b, const double *s_val) {
	const int phiLen = a.bPhiLen[b], t0 = a.bTermStart[b];
	bool bad = false;
	if (phiLen > 0) {                                                   // randCost.c:213-224
		for (int c = threadIdx.x; c < a.rows; c += blockDim.x) {
			double theta = 0.0;
			for (int n = 0; n < phiLen; n++)
				theta += a.phi[(size_t) (t0 + 1 + n) * a.rows + c] * s_val[a.tOmega[t0 + 1 + n] - 1];
			const double v = a.piDet[(size_t) b * a.rows + c] + theta;
			const char sense = a.senx[c];
			if ((v < -a.tol && sense == 'G') || (v > a.tol && sense == 'L')) bad = true;
		}
	}
	for (int c = threadIdx.x; c < a.cols; c += blockDim.x) {            // randCost.c:236-252
		double rc = a.gBar[(size_t) b * a.cols + c];
		for (int j = 0; j < a.rvd; j++) if (a.rvdOmCols[j] == c + 1) rc += s_val[j];          // addVectors
		for (int n = 0; n < phiLen; n++)                                                       // MSparsexvSub, entry order
			rc -= a.psi[(size_t) (t0 + 1 + n) * a.cols + c] * s_val[a.tOmega[t0 + 1 + n] - 1];
		if (rc < -a.tol && a.cstat[(size_t) b * a.cols + c] != 2) bad = true;
	}
	return bad;
}

// a new observation against every stored basis (stocUpdate.c:28-30): one CTA per basis
__global__ void k_feas_obs(FeasArgs a, int obs) {
	extern __shared__ double s_val[];
	const int b = blockIdx.x;
	if (!a.bFeas[b] || !a.has[b]) { if (threadIdx.x == 0) a.flags[b] = 2; return; }             // 2 = untouched
	for (int j = threadIdx.x; j < a.rvd; j += blockDim.x) s_val[j] = a.omega[(size_t) (a.rvOffset2 + j) * a.NP + obs];
	__syncthreads();
	const int bad = __syncthreads_or(sd_pair_violates(a, b, s_val));
	if (threadIdx.x == 0) {                                  // this CTA is the only writer of the word of (b, obs) in this launch
		a.flags[b] = !bad;
		uint32_t *w = a.mask + sd_mask_word(a.Bcap, b, obs);
		const uint32_t bit = 1u << (obs & 31);
		*w = bad ? (*w & ~bit) : (*w | bit);
	}
}

// a new basis against every stored observation (stocUpdate.c:123-126): one CTA per observation
__global__ void k_feas_basis(FeasArgs a, int b) {
	extern __shared__ double s_val[];
	const int obs = blockIdx.x;
	for (int j = threadIdx.x; j < a.rvd; j += blockDim.x) s_val[j] = a.omega[(size_t) (a.rvOffset2 + j) * a.NP + obs];
	__syncthreads();
	const int bad = __syncthreads_or(sd_pair_violates(a, b, s_val));
	if (threadIdx.x == 0) {                                  // 32 CTAs share a mask word: atomic bit update
		a.flags[obs] = !bad;
		uint32_t *w = a.mask + sd_mask_word(a.Bcap, b, obs);
		const uint32_t bit = 1u << (obs & 31);
		if (bad) atomicAnd(w, ~bit); else atomicOr(w, bit);
	}
}

// ======================================================================================================
// host side
// ======================================================================================================

template <class T>
static int sd_alloc(T **p, size_t n) {
	if (n == 0) n = 1;
	cudaError_t e = cudaMalloc((void **) p, n * sizeof(T));
	if (e != cudaSuccess) return sdgpu_fail("cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
	return 0;
}

template <class T>
static int sd_upload(T **p, const std::vector<T> &h) {
	if (sd_alloc(p, h.size())) return SDGPU_ERR;
	if (!h.empty()) SD_CUDA(cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
	return 0;
}

int sd_aux_reserve(sdgpu_ctx *c, size_t bytes) {
	if (bytes <= c->auxCap) return 0;
	SD_CUDA(cudaStreamSynchronize(c->stream));
	if (c->h_aux) cudaFreeHost(c->h_aux);
	c->h_aux = nullptr; c->d_aux = nullptr; c->auxCap = 0;
	size_t cap = std::max<size_t>(bytes * 2, 1 << 16);
	if (cudaHostAlloc((void **) &c->h_aux, cap, cudaHostAllocMapped) != cudaSuccess ||
	    cudaHostGetDevicePointer((void **) &c->d_aux, c->h_aux, 0) != cudaSuccess)
		return sdgpu_fail("pinned scratch of %zu bytes failed", cap);
	c->auxCap = cap;
	return 0;
}

int sd_scratch_reserve(sdgpu_ctx *c, size_t bytes) {
	if (bytes <= c->scratchCap) return 0;
	SD_CUDA(cudaStreamSynchronize(c->stream));
	if (c->d_scratch) cudaFree(c->d_scratch);
	c->d_scratch = nullptr; c->scratchCap = 0;
	size_t cap = std::max<size_t>(bytes * 2, 1 << 20);
	if (cudaMalloc((void **) &c->d_scratch, cap) != cudaSuccess) return sdgpu_fail("device scratch of %zu bytes failed", cap);
	c->scratchCap = cap;
	return 0;
}

int sd_sync_state(sdgpu_ctx *c) {
	SD_CUDA(cudaStreamSynchronize(c->stream));          // the commit kernels already published the state into h_state
	c->omegaCnt = c->h_state->omegaCnt; c->lambdaCnt = c->h_state->lambdaCnt;
	c->sigmaCnt = c->h_state->sigmaCnt;
	// sigma -> lambda row mirror of the grouped sweep: a sigma appended by the call that just finished is mirrored for free (the
	// published state names it); anything else (bulk loads) is read back on demand by sd_refresh_host_lam
	if (c->h_state->newSigma && c->h_state->lambdaIdx >= 0 && c->h_state->sigmaIdx == (int) c->hostLam.size() && (int64_t) c->hostLam.size() < c->sigmaCnt)
		c->hostLam.push_back(c->h_state->lambdaIdx);
	if (c->h_state->overflow) {
		SD_CUDA(cudaMemsetAsync(&c->d_state->overflow, 0, sizeof(int), c->stream));
		c->h_state->overflow = 0;
		return sdgpu_fail("table capacity exceeded (omega %lld/%lld, lambda %lld/%lld, sigma %lld/%lld)",
				(long long) c->omegaCnt, (long long) c->caps.maxOmega, (long long) c->lambdaCnt, (long long) c->caps.maxLambda,
				(long long) c->sigmaCnt, (long long) c->caps.maxSigma);
	}
	return 0;
}

// host vector -> mapped pinned staging; kernels read it through the device alias c->d_pinD (zero copy).  Safe to reuse:
// every caller waits for the stream before returning.
static int sd_stage_vec(sdgpu_ctx *c, const double *h, int n) {
	memcpy(c->h_pinD, h, (size_t) n * sizeof(double));
	return 0;
}

// Where the scan kernels should read the staged vector from: zero-copy over PCIe is the lowest latency for a handful of
// blocks, but every block re-reads the vector, so large tables (many blocks) get one DMA copy into device memory instead.
static const double *sd_staged_source(sdgpu_ctx *c, int n, int64_t rowsToScan) {
	if (rowsToScan <= 32 * 256) return c->d_pinD;
	if (cudaMemcpyAsync(c->d_vecIn, c->h_pinD, (size_t) n * sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return c->d_pinD;
	return c->d_vecIn;
}

static inline int sd_blocks(int64_t n, int t) { return (int) std::max<int64_t>(1, (n + t - 1) / t); }

extern "C" int sdgpu_create(const sdgpu_problem *p, const sdgpu_caps *caps, int device, sdgpu_ctx **out) {
	if (!p || !caps || !out) return sdgpu_fail("sdgpu_create: null argument");
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return sdgpu_fail("sdgpu_create: no CUDA device -- this library has no CPU fallback");
	if (device < 0 || device >= ndev) return sdgpu_fail("sdgpu_create: device %d out of range (%d devices)", device, ndev);
	const sdgpu_num &nm = p->num;
	if (nm.rvRowCnt < 0 || nm.rvbOmCnt < 0 || nm.rvCOmCnt < 0 || nm.cntCcols < 0 || nm.prevCols < 0 || nm.rows <= 0)
		return sdgpu_fail("sdgpu_create: negative dimension");
	if (caps->maxLambda <= 0 || caps->maxSigma <= 0 || caps->maxBasis <= 0 || caps->maxOmega <= 0)
		return sdgpu_fail("sdgpu_create: capacities must be positive");
	SD_CUDA(cudaSetDevice(device));

	sdgpu_ctx *c = new sdgpu_ctx();
	c->device = device; c->num = nm; c->caps = *caps;
	{ const char *e = getenv("SDGPU_PDL"); if (e) c->pdl = atoi(e) != 0; e = getenv("SDGPU_ALTDIR"); if (e) c->altDir = atoi(e) != 0;
	  e = getenv("SDGPU_FUSED_UPDATE"); if (e) c->fusedUpdate = atoi(e) != 0;
	  e = getenv("SDGPU_CHUNKS"); if (e) c->forceChunks = atoi(e);
	  e = getenv("SDGPU_SWEEP_VARIANT"); if (e && atoi(e) >= 0 && atoi(e) <= 4) c->sweepVariant = atoi(e); }   // experiment knobs, read per context
	if (c->caps.maxTerms < 1) c->caps.maxTerms = 1;
	c->n1 = nm.prevCols; c->n1c = nm.cntCcols; c->n1cP = std::max(1, nm.cntCcols); c->R = nm.rvRowCnt; c->Rb = nm.rvbOmCnt;
	c->Q = nm.rvCOmCnt; c->rvd = nm.rvdOmCnt; c->numRV = nm.numRV; c->rows = nm.rows; c->cols = nm.cols;
	memcpy(c->rvOffset, p->coord.rvOffset, sizeof c->rvOffset);
	c->NP = sd_round_up(caps->maxOmega, SD_TILE_W); c->nTiles = c->NP / SD_TILE_W;
	c->LP = sd_round_up(caps->maxLambda, 32); c->SP = sd_round_up(caps->maxSigma, 32); c->BP = sd_round_up(caps->maxBasis, 32);
	c->termCap = caps->maxBasis * (int64_t) c->caps.maxTerms;

#define SD_TRY(x) do { if ((x) != 0) { sdgpu_destroy(c); return SDGPU_ERR; } } while (0)
	cudaError_t ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
	if (ce != cudaSuccess) { sdgpu_fail("cudaStreamCreate: %s", cudaGetErrorString(ce)); sdgpu_destroy(c); return SDGPU_ERR; }
	cudaEventCreate(&c->evA); cudaEventCreate(&c->evB); cudaEventCreate(&c->evC); cudaEventCreate(&c->evD); cudaEventCreate(&c->evE);

	// ---- static maps (host, once) ---------------------------------------------------------------------
	auto lamPosOfRow = [&](int row) {          // expandVector: the LAST lambda entry written to that row wins
		int pos = -1;
		for (int i = 1; i <= c->R; i++) if (p->coord.rvRows[i] == row) pos = i - 1;
		return pos;
	};
	std::vector<int32_t> CCols(c->n1c), rvRows(c->R), bLamPos(c->Rb), cLamPos(c->Q), cListStart(c->Q + 1, 0), cList, rvCOmCols(c->Q), rvCols(c->Q);
	for (int k = 0; k < c->n1c; k++) { CCols[k] = p->coord.CCols[k + 1]; if (CCols[k] < 1 || CCols[k] > c->n1) { sdgpu_fail("CCols[%d]=%d outside 1..%d", k + 1, CCols[k], c->n1); sdgpu_destroy(c); return SDGPU_ERR; } }
	for (int i = 0; i < c->R; i++) { rvRows[i] = p->coord.rvRows[i + 1]; if (rvRows[i] < 1 || rvRows[i] > c->rows) { sdgpu_fail("rvRows[%d]=%d outside 1..%d", i + 1, rvRows[i], c->rows); sdgpu_destroy(c); return SDGPU_ERR; } }
	for (int j = 0; j < c->Rb; j++) bLamPos[j] = lamPosOfRow(p->coord.rvbOmRows[j + 1]);
	for (int e = 0; e < c->Q; e++) {
		cLamPos[e] = lamPosOfRow(p->coord.rvCOmRows[e + 1]);
		rvCOmCols[e] = p->coord.rvCOmCols[e + 1];
		rvCols[e] = p->coord.rvCols ? p->coord.rvCols[e + 1] : rvCOmCols[e];
		if (rvCOmCols[e] < 1 || rvCOmCols[e] > c->n1 || rvCols[e] < 1 || rvCols[e] > c->n1) { sdgpu_fail("random T column outside 1..%d", c->n1); sdgpu_destroy(c); return SDGPU_ERR; }
	}
	for (int k = 0; k < c->Q; k++) {
		for (int e = 0; e < c->Q; e++) if (rvCOmCols[e] == rvCOmCols[k]) cList.push_back(e);
		cListStart[k + 1] = (int32_t) cList.size();
	}
	std::vector<int32_t> bCol(p->bBar.cnt), cbStart(c->n1c + 1, 0), cbRow;
	std::vector<double> bVal(p->bBar.cnt), cbVal;
	for (int e = 0; e < p->bBar.cnt; e++) { bCol[e] = p->bBar.col[e + 1]; bVal[e] = p->bBar.val[e + 1]; }
	for (int k = 0; k < c->n1c; k++) {
		for (int e = 1; e <= p->Cbar.cnt; e++)
			if (p->Cbar.col[e] == CCols[k]) { cbRow.push_back(p->Cbar.row[e]); cbVal.push_back(p->Cbar.val[e]); }
		cbStart[k + 1] = (int32_t) cbRow.size();
	}
	c->bBarCnt = p->bBar.cnt; c->cbNnz = (int) cbRow.size();
	SD_TRY(sd_upload(&c->d_CCols, CCols)); SD_TRY(sd_upload(&c->d_rvRows, rvRows)); SD_TRY(sd_upload(&c->d_bLamPos, bLamPos));
	SD_TRY(sd_upload(&c->d_cLamPos, cLamPos)); SD_TRY(sd_upload(&c->d_cListStart, cListStart)); SD_TRY(sd_upload(&c->d_cList, cList));
	SD_TRY(sd_upload(&c->d_rvCOmCols, rvCOmCols)); SD_TRY(sd_upload(&c->d_rvCols, rvCols));
	SD_TRY(sd_upload(&c->d_bBarCol, bCol)); SD_TRY(sd_upload(&c->d_bBarVal, bVal));
	SD_TRY(sd_upload(&c->d_cbStart, cbStart)); SD_TRY(sd_upload(&c->d_cbRow, cbRow)); SD_TRY(sd_upload(&c->d_cbVal, cbVal));

	// ---- tables -----------------------------------------------------------------------------------------
	SD_TRY(sd_alloc(&c->d_omega, (size_t) c->numRV * c->NP)); SD_TRY(sd_alloc(&c->d_omegaW, (size_t) c->NP));
	SD_TRY(sd_alloc(&c->d_lambda, (size_t) c->R * c->LP));
	SD_TRY(sd_alloc(&c->d_sigmaPib, (size_t) c->SP)); SD_TRY(sd_alloc(&c->d_sigmaPiCk, (size_t) c->n1cP * c->SP));
	SD_TRY(sd_alloc(&c->d_sigmaPiCr, (size_t) c->SP * c->n1cP));
	SD_TRY(sd_alloc(&c->d_sigmaLam, (size_t) c->SP)); SD_TRY(sd_alloc(&c->d_sigmaCk, (size_t) c->SP));
	{   // the delta table: allocated whole, or -- SDGPU_VMM=1, or when its capacity does not fit the free memory -- reserved and mapped as it grows
		const size_t whole = (size_t) c->nTiles * caps->maxLambda * (1 + c->Q) * SD_TILE_W * 8;
		size_t freeB = 0, totalB = 0;
		cudaMemGetInfo(&freeB, &totalB);
		const char *e = getenv("SDGPU_VMM");
		const bool onDemand = e ? atoi(e) != 0 : whole > freeB - std::min(freeB, (size_t) 2 << 30);
		c->Dcap = caps->maxLambda;
		if (onDemand) {
			c->Dcap = sd_round_up(caps->maxLambda, 512);
			if (sd_vm_create(c, (size_t) (1 + c->Q) * SD_TILE_W * 8, c->Dcap, c->nTiles)) { sdgpu_destroy(c); return SDGPU_ERR; }
		}
		else SD_TRY(sd_alloc(&c->d_delta, whole / 8));
	}
	if (c->rvd > 0) SD_TRY(sd_alloc(&c->d_mask, (size_t) c->nTiles * caps->maxBasis * SD_MASK_WORDS));
	SD_TRY(sd_alloc(&c->d_bCk, (size_t) c->BP)); SD_TRY(sd_alloc(&c->d_bFeas, (size_t) c->BP)); SD_TRY(sd_alloc(&c->d_bPhiLen, (size_t) c->BP));
	SD_TRY(sd_alloc(&c->d_bTermStart, (size_t) c->BP + 1)); SD_TRY(sd_alloc(&c->d_tSigma, (size_t) c->termCap)); SD_TRY(sd_alloc(&c->d_tOmega, (size_t) c->termCap));
	SD_TRY(sd_alloc(&c->d_state, 1));
	{ SdDevState init; memset(&init, 0, sizeof init); init.foundLambda = init.foundSigma = init.foundOmega = INT_MAX;
	  cudaMemcpy(c->d_state, &init, sizeof init, cudaMemcpyHostToDevice); }
	cudaMemset(c->d_omegaW, 0, (size_t) c->NP * sizeof(int32_t));
	cudaMemset(c->d_bTermStart, 0, ((size_t) c->BP + 1) * sizeof(int32_t));

	// ---- staging ----------------------------------------------------------------------------------------
	size_t vecLen = (size_t) std::max(std::max(c->rows, c->numRV), c->n1) + 2;
	SD_TRY(sd_alloc(&c->d_vecIn, vecLen)); SD_TRY(sd_alloc(&c->d_cand, (size_t) std::max(c->R, c->numRV) + 1)); SD_TRY(sd_alloc(&c->d_candC, (size_t) c->n1cP));
	c->pinDcap = vecLen + (size_t) c->n1 + 16; c->pinIcap = std::max<size_t>(64, 8 + 2 * (size_t) c->caps.maxTerms);   // basis_append stages 5 + 2 (1 + phiLength) ints
	if (cudaHostAlloc((void **) &c->h_pinD, c->pinDcap * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
	    cudaHostAlloc((void **) &c->h_pinI, c->pinIcap * sizeof(int32_t), cudaHostAllocMapped) != cudaSuccess ||
	    cudaHostAlloc((void **) &c->h_state, sizeof(SdDevState), cudaHostAllocMapped) != cudaSuccess ||
	    cudaHostAlloc((void **) &c->h_cutRes, ((size_t) c->n1 + 8) * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
	    cudaHostAlloc((void **) &c->h_iStar, (size_t) std::min<int64_t>(c->NP, 65536) * sizeof(int32_t), cudaHostAllocMapped) != cudaSuccess ||
	    cudaHostGetDevicePointer((void **) &c->d_iStarHost, c->h_iStar, 0) != cudaSuccess ||
	    cudaHostGetDevicePointer((void **) &c->d_pinD, c->h_pinD, 0) != cudaSuccess ||
	    cudaHostGetDevicePointer((void **) &c->d_hstate, c->h_state, 0) != cudaSuccess ||
	    cudaHostGetDevicePointer((void **) &c->d_cutRes, c->h_cutRes, 0) != cudaSuccess) {
		sdgpu_fail("pinned host allocation failed"); sdgpu_destroy(c); return SDGPU_ERR;
	}
	memset(c->h_state, 0, sizeof(SdDevState));
	c->iStarHostCap = std::min<int64_t>(c->NP, 65536);

	// ---- cut scratch -------------------------------------------------------------------------------------
	c->maxChunks = SD_MAX_CHUNKS;
	{   // the per-chunk partial maxima cost 2 * chunks * NP * 12 bytes; keep them below 1/16 of the delta table or 64 MiB
		size_t deltaBytes = (size_t) c->nTiles * caps->maxLambda * (1 + c->Q) * SD_TILE_W * 8;
		size_t budget = std::max<size_t>(deltaBytes / 16, (size_t) 64 << 20);
		while (c->maxChunks > 1 && (size_t) 2 * c->maxChunks * c->NP * 12 > budget) c->maxChunks /= 2;
	}
	SD_TRY(sd_alloc(&c->d_x, (size_t) c->n1 + 2)); SD_TRY(sd_alloc(&c->d_piCbarX, (size_t) c->SP));
	SD_TRY(sd_alloc(&c->d_descA, (size_t) c->BP)); SD_TRY(sd_alloc(&c->d_descC, (size_t) c->BP));
	SD_TRY(sd_alloc(&c->d_descRow, (size_t) c->BP)); SD_TRY(sd_alloc(&c->d_descWin, (size_t) c->BP));
	SD_TRY(sd_alloc(&c->d_entBasis, (size_t) c->BP));
	SD_TRY(sd_alloc(&c->d_entGroup, (size_t) c->BP)); SD_TRY(sd_alloc(&c->d_groupRow, (size_t) c->BP));
	if (c->rvd > 0) {
		SD_TRY(sd_alloc(&c->d_termA, (size_t) c->termCap)); SD_TRY(sd_alloc(&c->d_termC, (size_t) c->termCap));
		SD_TRY(sd_alloc(&c->d_termRow, (size_t) c->termCap)); SD_TRY(sd_alloc(&c->d_termMeta, (size_t) c->termCap));
		SD_TRY(sd_alloc(&c->d_termBasis, (size_t) c->termCap));
	}
	SD_TRY(sd_alloc(&c->d_partV, (size_t) 2 * c->maxChunks * c->NP)); SD_TRY(sd_alloc(&c->d_partI, (size_t) 2 * c->maxChunks * c->NP));
	SD_TRY(sd_alloc(&c->d_iStar, (size_t) c->NP));
	SD_TRY(sd_alloc(&c->d_tilePart, (size_t) c->nTiles * (SD_TILE_W / 64) * (4 + c->n1c + c->Q)));   // one partial vector per merge CTA (>= 64 observations each)
	SD_TRY(sd_alloc(&c->d_cutPartial, (size_t) c->n1 + 4)); SD_TRY(sd_alloc(&c->d_cutOut, (size_t) c->n1 + 4));
#undef SD_TRY
	if (cudaDeviceSynchronize() != cudaSuccess) { sdgpu_fail("device error during create"); sdgpu_destroy(c); return SDGPU_ERR; }
	*out = c;
	return 0;
}

extern "C" void sdgpu_destroy(sdgpu_ctx *c) {
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	sd_nccl_release(c);
	sd_peer_teardown(c);
	sd_vm_destroy(c);                                      // (clears d_delta when the table was mapped on demand)
	void *dev[] = { c->d_CCols, c->d_rvRows, c->d_bLamPos, c->d_cLamPos, c->d_cListStart, c->d_cList, c->d_rvCOmCols, c->d_rvCols, c->d_bBarCol,
		c->d_bBarVal, c->d_cbStart, c->d_cbRow, c->d_cbVal, c->d_omega, c->d_omegaW, c->d_lambda, c->d_sigmaPib, c->d_sigmaPiCk, c->d_sigmaPiCr,
		c->d_sigmaLam, c->d_sigmaCk, c->d_delta, c->d_mask, c->d_bCk, c->d_bFeas, c->d_bPhiLen, c->d_bTermStart, c->d_tSigma, c->d_tOmega, c->d_state,
		c->d_rvdOmCols, c->d_senx, c->d_fPiDet, c->d_fPhi, c->d_fGBar, c->d_fPsi, c->d_fCstat, c->d_fHas, c->d_fFlags,
		c->d_vecIn, c->d_cand, c->d_candC, c->d_x, c->d_piCbarX, c->d_descA, c->d_descC, c->d_descRow, c->d_descWin, c->d_partV, c->d_partI,
		c->d_iStar, c->d_tilePart, c->d_cutPartial, c->d_cutOut, c->d_termA, c->d_termC, c->d_termRow, c->d_termMeta, c->d_termBasis, c->d_entBasis, c->d_entGroup, c->d_groupRow };
	for (void *p : dev) if (p) cudaFree(p);
	if (c->h_pinD) cudaFreeHost(c->h_pinD);
	if (c->h_pinI) cudaFreeHost(c->h_pinI);
	if (c->h_state) cudaFreeHost(c->h_state);
	if (c->h_cutRes) cudaFreeHost(c->h_cutRes);
	if (c->h_iStar) cudaFreeHost(c->h_iStar);
	if (c->h_aux) cudaFreeHost(c->h_aux);
	if (c->d_scratch) cudaFree(c->d_scratch);
	if (c->d_fpAlpha) cudaFree(c->d_fpAlpha);
	if (c->d_fpBeta) cudaFree(c->d_fpBeta);
	if (c->evA) cudaEventDestroy(c->evA);
	if (c->evB) cudaEventDestroy(c->evB);
	if (c->evC) cudaEventDestroy(c->evC);
	if (c->evD) cudaEventDestroy(c->evD);
	if (c->evE) cudaEventDestroy(c->evE);
	if (c->stream && c->ownStream) cudaStreamDestroy(c->stream);
	delete c;
}

// setup.c:242-246: counts back to zero, every allocation kept
extern "C" int sdgpu_reset(sdgpu_ctx *c) {
	if (!c) return sdgpu_fail("null context");
	SD_CUDA(cudaSetDevice(c->device));
	{ SdDevState init; memset(&init, 0, sizeof init); init.foundLambda = init.foundSigma = init.foundOmega = INT_MAX;
	  SD_CUDA(cudaStreamSynchronize(c->stream));
	  SD_CUDA(cudaMemcpy(c->d_state, &init, sizeof init, cudaMemcpyHostToDevice));
	  *c->h_state = init; }
	c->omegaCnt = c->lambdaCnt = c->sigmaCnt = c->basisCnt = c->termCnt = 0;
	c->maxPhiLen = 0; c->anyInfeasibleBasis = false; c->lastOmegaCnt = 0; c->fpCnt = 0;      // (freeCutsType(cell->fcutsPool, true), setup.c:236)
	c->basis.clear(); c->hostMask.clear();
	c->hostLam.clear(); c->grpRowCount.clear(); c->grpBasis.clear(); c->grpRow.clear(); c->grpGroup.clear(); c->grpGroupRow.clear(); c->grpCounted = c->grpDistinct = c->grpSorted = 0;
	if (c->d_fHas) SD_CUDA(cudaMemset(c->d_fHas, 0, (size_t) c->caps.maxBasis));
	return 0;
}

extern "C" int sdgpu_get_counts(sdgpu_ctx *c, sdgpu_counts *out) {
	if (!c || !out) return sdgpu_fail("null argument");
	out->omega = c->omegaCnt; out->lambda = c->lambdaCnt; out->sigma = c->sigmaCnt; out->basis = c->basisCnt;
	return 0;
}

extern "C" int sdgpu_set_stream(sdgpu_ctx *c, void *stream) {
	if (!c) return sdgpu_fail("null context");
	SD_CUDA(cudaStreamSynchronize(c->stream));
	if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
	c->stream = (cudaStream_t) stream; c->ownStream = false;
	return 0;
}

extern "C" int sdgpu_set_timing(sdgpu_ctx *c, int on) {
	if (!c) return sdgpu_fail("null context");
	c->timing = on != 0;
	return 0;
}

extern "C" int sdgpu_get_stats(sdgpu_ctx *c, sdgpu_stats *out) {
	if (!c || !out) return sdgpu_fail("null argument");
	*out = c->stats;
	return 0;
}

// ---- omega -------------------------------------------------------------------------------------------------
static int sd_launch_omega(sdgpu_ctx *c, const double *observ, double tol, int mode, int weight) {
	if (c->numRV + 1 > SD_VEC_PARAM_MAX && sd_stage_vec(c, observ, c->numRV + 1)) return SDGPU_ERR;
	const int blocks = mode == 2 ? 1 : sd_blocks(c->omegaCnt, 256);
	if (sd_smem_optin(c, k_omega_fused, SD_SMEM_OMEGA, 64, (size_t) std::max(1, c->numRV) * 8, "k_omega_fused")) return SDGPU_ERR;
	SdVecParam vp;
	const bool asParam = c->numRV + 1 <= SD_VEC_PARAM_MAX;
	if (asParam) memcpy(vp.v, observ, ((size_t) c->numRV + 1) * sizeof(double));
	k_omega_fused<<<blocks, 256, (size_t) std::max(1, c->numRV) * 8, c->stream>>>(asParam ? nullptr : sd_staged_source(c, c->numRV + 1, mode == 2 ? 0 : c->omegaCnt), vp, c->numRV, c->d_omega, c->d_omegaW, c->NP,
			c->caps.maxOmega, tol, mode, weight, c->d_state, c->d_hstate);
	SD_LAUNCH_OK("k_omega_fused");
	sd_count_launch(c);
	return sd_sync_state(c);
}

extern "C" int sdgpu_calc_omega(sdgpu_ctx *c, const double *observ, double tol, int *newOmegaFlag) {
	if (!c || !observ) return sdgpu_fail("null argument");
	SD_CUDA(cudaSetDevice(c->device));
	if (sd_launch_omega(c, observ, tol, 0, 1)) return SDGPU_ERR;
	if (newOmegaFlag) *newOmegaFlag = c->h_state->newOmega;
	return c->h_state->omegaIdx;
}

extern "C" int sdgpu_omega_find(sdgpu_ctx *c, const double *observ, double tol) {
	if (!c || !observ) return sdgpu_fail("null argument");
	SD_CUDA(cudaSetDevice(c->device));
	if (sd_launch_omega(c, observ, tol, 1, 0)) return SDGPU_ERR;
	return c->h_state->omegaIdx < 0 ? SDGPU_NONE : c->h_state->omegaIdx;
}

extern "C" int sdgpu_omega_append(sdgpu_ctx *c, const double *observ, int weight) {
	if (!c || !observ) return sdgpu_fail("null argument");
	SD_CUDA(cudaSetDevice(c->device));
	if (sd_launch_omega(c, observ, 0.0, 2, weight)) return SDGPU_ERR;
	return c->h_state->omegaIdx;
}

extern "C" int sdgpu_omega_bump(sdgpu_ctx *c, int idx, int by) {
	if (!c) return sdgpu_fail("null context");
	if (idx < 0 || idx >= c->omegaCnt) return sdgpu_fail("omega_bump: index %d out of range", idx);
	SD_CUDA(cudaSetDevice(c->device));
	k_omega_bump<<<1, 1, 0, c->stream>>>(c->d_omegaW, idx, by);
	sd_count_launch(c);
	SD_CUDA(cudaStreamSynchronize(c->stream));
	return 0;
}

extern "C" int sdgpu_omega_append_bulk(sdgpu_ctx *c, int64_t n, const double *vals, const int32_t *weights) {
	if (!c || (!vals && n > 0)) return sdgpu_fail("null argument");
	if (n <= 0) return 0;
	if (c->omegaCnt + n > c->caps.maxOmega) return sdgpu_fail("omega_append_bulk: %lld + %lld exceeds capacity %lld", (long long) c->omegaCnt, (long long) n, (long long) c->caps.maxOmega);
	SD_CUDA(cudaSetDevice(c->device));
	const int64_t chunk = std::max<int64_t>(1, ((int64_t) 64 << 20) / ((c->numRV + 1) * 8));
	double *d_vals = nullptr; int32_t *d_w = nullptr;
	if (sd_alloc(&d_vals, (size_t) std::min(chunk, n) * (c->numRV + 1))) return SDGPU_ERR;
	if (weights && sd_alloc(&d_w, (size_t) std::min(chunk, n))) { cudaFree(d_vals); return SDGPU_ERR; }
	int rc = 0;
	for (int64_t i0 = 0; i0 < n && rc == 0; i0 += chunk) {
		int64_t m = std::min(chunk, n - i0);
		cudaError_t e = cudaMemcpyAsync(d_vals, vals + (size_t) i0 * (c->numRV + 1), (size_t) m * (c->numRV + 1) * 8, cudaMemcpyHostToDevice, c->stream);
		if (e == cudaSuccess && weights) e = cudaMemcpyAsync(d_w, weights + i0, (size_t) m * 4, cudaMemcpyHostToDevice, c->stream);
		if (e != cudaSuccess) { rc = sdgpu_fail("omega_append_bulk copy: %s", cudaGetErrorString(e)); break; }
		k_omega_bulk<<<sd_blocks(m, 128), 128, 0, c->stream>>>(d_vals, m, c->numRV, weights ? d_w : nullptr, c->d_omega, c->d_omegaW, c->NP, c->omegaCnt + i0);
		sd_count_launch(c);
		if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = sdgpu_fail("omega_append_bulk kernel failed");
	}
	if (rc == 0) {
		k_bump_counts<<<1, 1, 0, c->stream>>>(c->d_state, c->d_hstate, (int) n, 0, 0, 0);
		sd_count_launch(c);
		rc = sd_sync_state(c);
	}
	cudaFree(d_vals); if (d_w) cudaFree(d_w);
	return rc;
}

// ---- lambda / sigma / delta --------------------------------------------------------------------------------
// calcLambda in one launch; its last block also stages calcSigma's candidate (pibBar, piCBar) from the same vector
static int sd_launch_lambda(sdgpu_ctx *c, const double *d_pi, double mubBar, double tol, int64_t lambdaUpper, bool publish) {
	const int cbStage = std::min(c->cbNnz, 2048);
	const size_t smem = ((size_t) std::max(1, c->R) + c->rows + 1 + std::max(1, c->bBarCnt) + std::max(1, cbStage)) * 8;
	if (sd_smem_optin(c, k_lambda_fused, SD_SMEM_LAMBDA, 64, smem, "k_lambda_fused")) return SDGPU_ERR;   // very long dual vectors: opt in (per device, remembered per context)
	k_lambda_fused<<<sd_blocks(lambdaUpper, 256), 256, smem, c->stream>>>(d_pi, c->rows, c->d_rvRows, c->R, c->d_lambda, c->LP,
			c->caps.maxLambda, tol, c->d_bBarCol, c->d_bBarVal, c->bBarCnt, mubBar, c->d_cbStart, c->d_cbRow, c->d_cbVal, c->n1c, cbStage,
			c->d_vecIn, c->d_candC, c->d_state, c->d_hstate, publish ? 1 : 0);
	SD_LAUNCH_OK("k_lambda_fused");
	sd_count_launch(c);
	return 0;
}

static int sd_launch_sigma(sdgpu_ctx *c, int iter, double tol, int64_t sigmaUpper) {
	k_sigma_fused<<<sd_blocks(sigmaUpper, 256), 256, 0, c->stream>>>(c->d_sigmaPib, c->d_sigmaPiCk, c->d_sigmaPiCr, c->d_sigmaLam, c->d_sigmaCk, c->SP,
			c->n1c, c->n1cP, c->d_candC, tol, iter, c->caps.maxSigma, c->d_state, c->d_hstate);
	SD_LAUNCH_OK("k_sigma_fused");
	sd_count_launch(c);
	return 0;
}

static int sd_launch_delta_row(sdgpu_ctx *c, int forcedRow, int64_t omegaUpper) {
	if (omegaUpper <= 0) return 0;
	if (sd_delta_ensure(c, forcedRow >= 0 ? forcedRow + 1 : std::min<int64_t>(c->caps.maxLambda, c->lambdaCnt + 1), omegaUpper)) return SDGPU_ERR;
	if (sd_smem_optin(c, k_delta_row, SD_SMEM_DELTA_ROW, (size_t) DC_BATCH * DC_THREADS * 8, (size_t) std::max(1, c->R + c->Rb) * 8, "k_delta_row")) return SDGPU_ERR;
	k_delta_row<<<sd_blocks(omegaUpper, DC_THREADS), DC_THREADS, (size_t) std::max(1, c->R + c->Rb) * 8, c->stream>>>(c->d_lambda, c->LP, c->R, c->d_omega, c->NP, c->Rb, c->Q,
			c->d_bLamPos, c->d_cLamPos, c->d_cListStart, c->d_cList, c->d_delta, c->Dcap, c->d_state, forcedRow);
	SD_LAUNCH_OK("k_delta_row");
	sd_count_launch(c);
	return 0;
}

static int sd_launch_delta_col(sdgpu_ctx *c, int forcedCol, int64_t lambdaUpper) {
	if (lambdaUpper <= 0) return 0;
	if (sd_delta_ensure(c, lambdaUpper, std::max<int64_t>(c->omegaCnt, forcedCol + 1))) return SDGPU_ERR;
	if (sd_smem_optin(c, k_delta_col, SD_SMEM_DELTA_COL, (size_t) DC_BATCH * DC_THREADS * 8, (size_t) std::max(1, c->numRV) * 8 + (size_t) std::max(1, c->Rb) * 4, "k_delta_col")) return SDGPU_ERR;
	k_delta_col<<<sd_blocks(lambdaUpper, DC_THREADS), DC_THREADS, (size_t) std::max(1, c->numRV) * 8 + (size_t) std::max(1, c->Rb) * 4, c->stream>>>(c->d_lambda, c->LP, c->d_omega, c->NP, c->numRV, c->Rb, c->Q,
			c->d_bLamPos, c->d_cLamPos, c->d_cListStart, c->d_cList, c->d_delta, c->Dcap, c->d_state, forcedCol);
	SD_LAUNCH_OK("k_delta_col");
	sd_count_launch(c);
	return 0;
}

extern "C" int sdgpu_calc_lambda(sdgpu_ctx *c, const double *Pi, double tol, int *newLambdaFlag) {
	if (!c || !Pi) return sdgpu_fail("null argument");
	SD_CUDA(cudaSetDevice(c->device));
	if (sd_stage_vec(c, Pi, c->rows + 1)) return SDGPU_ERR;
	if (sd_launch_lambda(c, sd_staged_source(c, c->rows + 1, c->lambdaCnt), 0.0, tol, c->lambdaCnt, true)) return SDGPU_ERR;
	if (sd_sync_state(c)) return SDGPU_ERR;
	if (newLambdaFlag) *newLambdaFlag = c->h_state->newLambda;
	return c->h_state->lambdaIdx;
}

extern "C" int sdgpu_calc_sigma(sdgpu_ctx *c, const double *pi, double mubBar, int idxLambda, int newLambdaFlag, int currentIter,
		double tol, int *newSigmaFlag) {
	if (!c || !pi) return sdgpu_fail("null argument");
	if (idxLambda < 0 || idxLambda >= c->lambdaCnt) return sdgpu_fail("calc_sigma: lambda index %d out of range", idxLambda);
	SD_CUDA(cudaSetDevice(c->device));
	if (sd_stage_vec(c, pi, c->rows + 1)) return SDGPU_ERR;
	// stand-alone form: stage (pibBar, piCBar) with the caller's lambda index / flag, then scan + commit
	k_sigma_prepare<<<sd_blocks(c->n1c + 1, 128), 128, 0, c->stream>>>(c->d_pinD, c->d_bBarCol, c->d_bBarVal, c->bBarCnt, mubBar, c->d_cbStart,
			c->d_cbRow, c->d_cbVal, c->n1c, c->d_candC, c->d_state, newLambdaFlag != 0, idxLambda);
	SD_LAUNCH_OK("k_sigma_prepare");
	sd_count_launch(c);
	if (sd_launch_sigma(c, currentIter, tol, c->sigmaCnt)) return SDGPU_ERR;
	if (sd_sync_state(c)) return SDGPU_ERR;
	if (newSigmaFlag) *newSigmaFlag = c->h_state->newSigma;
	return c->h_state->sigmaIdx;
}

extern "C" int sdgpu_calc_delta(sdgpu_ctx *c, int newOmegaFlag, int elemIdx) {
	if (!c) return sdgpu_fail("null context");
	SD_CUDA(cudaSetDevice(c->device));
	if (newOmegaFlag) {
		if (elemIdx < 0 || elemIdx >= c->omegaCnt) return sdgpu_fail("calc_delta: observation %d out of range", elemIdx);
		if (sd_launch_delta_col(c, elemIdx, c->lambdaCnt)) return SDGPU_ERR;
	}
	else {
		if (elemIdx < 0 || elemIdx >= c->lambdaCnt) return sdgpu_fail("calc_delta: lambda %d out of range", elemIdx);
		if (sd_launch_delta_row(c, elemIdx, c->omegaCnt)) return SDGPU_ERR;
	}
	// no host wait: later calls are ordered behind this one on the context's stream (a fault surfaces at the next sync)
	SD_CUDA(cudaGetLastError());
	return 0;
}

// the fused form: {delta column || lambda scan -> commit lambda, sigma scan + commit} in one launch, the delta row as its programmatic
// dependent.  Used while the sigma table is small enough for one block to scan (every real problem: <= 7 501 rows, setup.c:139).
static bool sd_fused_update_ok(sdgpu_ctx *c) {
	return c->fusedUpdate && c->sigmaCnt <= 16384 && c->lambdaCnt <= ((int64_t) 1 << 22);
}

static int sd_launch_update_fused(sdgpu_ctx *c, const double *hostPi, double mubBar, int iter, double tol, int colObs) {
	UpdArgs a;
	SdVecParam vp;
	const double *d_pi = nullptr;
	if (c->rows + 1 <= SD_VEC_PARAM_MAX) memcpy(vp.v, hostPi, ((size_t) c->rows + 1) * sizeof(double));
	else { if (sd_stage_vec(c, hostPi, c->rows + 1)) return SDGPU_ERR; d_pi = sd_staged_source(c, c->rows + 1, c->lambdaCnt); }
	const int cbStage = std::min(c->cbNnz, 2048);
	a.pi = d_pi; a.rows = c->rows; a.rvRows = c->d_rvRows; a.R = c->R; a.lambda = c->d_lambda; a.LP = c->LP; a.lambdaCap = c->caps.maxLambda; a.tol = tol;
	a.bCol = c->d_bBarCol; a.bVal = c->d_bBarVal; a.bCnt = c->bBarCnt; a.mubBar = mubBar;
	a.cbStart = c->d_cbStart; a.cbRow = c->d_cbRow; a.cbVal = c->d_cbVal; a.n1c = c->n1c; a.n1cP = c->n1cP; a.cbStage = std::max(1, cbStage);
	a.vecDev = c->d_vecIn;
	a.sigPib = c->d_sigmaPib; a.sigPiCk = c->d_sigmaPiCk; a.sigPiCr = c->d_sigmaPiCr; a.sigLam = c->d_sigmaLam; a.sigCk = c->d_sigmaCk;
	a.SP = c->SP; a.sigmaCap = c->caps.maxSigma; a.iter = iter;
	a.nbScan = sd_blocks(c->lambdaCnt, UF_THREADS);
	const int nbCol = (colObs >= 0 && c->lambdaCnt > 0) ? sd_blocks(c->lambdaCnt, UF_THREADS) : 0;
	a.colObs = colObs; a.lambdaCntHost = (int) c->lambdaCnt;
	a.omega = c->d_omega; a.NP = c->NP; a.numRV = c->numRV; a.Rb = c->Rb; a.Q = c->Q; a.bLamPos = c->d_bLamPos; a.cLamPos = c->d_cLamPos;
	a.cListStart = c->d_cListStart; a.cList = c->d_cList; a.delta = c->d_delta; a.Dcap = c->Dcap;
	a.st = c->d_state; a.hst = c->d_hstate;
	const size_t commitD = (size_t) std::max(1, c->R) + c->rows + 1 + std::max(1, c->bBarCnt) + std::max(1, cbStage) + std::max(1, c->n1c);
	const size_t colD = (size_t) c->numRV + (size_t) ((c->Rb + 1) / 2 + 1) + (size_t) UF_COLB * UF_THREADS;
	const size_t smem = std::max(commitD, colD) * 8;
	if (sd_smem_optin(c, k_update_fused, SD_SMEM_UPD1, 64, smem, "k_update_fused")) return SDGPU_ERR;
	if (sd_delta_ensure(c, std::min<int64_t>(c->caps.maxLambda, c->lambdaCnt + 1), std::max<int64_t>(c->omegaCnt, colObs + 1))) return SDGPU_ERR;
	k_update_fused<<<a.nbScan + nbCol, UF_THREADS, smem, c->stream>>>(a, vp);
	SD_LAUNCH_OK("k_update_fused");
	sd_count_launch(c);
	if (c->omegaCnt > 0) {                                 // :84-85, a no-op kernel unless the lambda was new
		const size_t rs = (size_t) std::max(1, c->R + c->Rb) * 8;
		if (sd_smem_optin(c, k_delta_row, SD_SMEM_DELTA_ROW, (size_t) DC_BATCH * DC_THREADS * 8, rs, "k_delta_row")) return SDGPU_ERR;
		SD_CUDA(sd_launch(k_delta_row, dim3((unsigned) sd_blocks(c->omegaCnt, DC_THREADS)), dim3(DC_THREADS), rs, c->stream, c->pdl,
				(const double *) c->d_lambda, c->LP, c->R, (const double *) c->d_omega, c->NP, c->Rb, c->Q, (const int32_t *) c->d_bLamPos, (const int32_t *) c->d_cLamPos,
				(const int32_t *) c->d_cListStart, (const int32_t *) c->d_cList, c->d_delta, (int64_t) c->Dcap, (const SdDevState *) c->d_state, -1));
		sd_count_launch(c);
	}
	return 0;
}

// stocUpdate.c:24-25 (the delta column of a new observation, newOmegaIdx >= 0) and :78-85 (calcLambda, calcSigma, delta row) in one
// device round trip
extern "C" int sdgpu_update_dual_col(sdgpu_ctx *c, int newOmegaIdx, const double *pi, double mubBar, int currentIter, double tol,
		int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag) {
	if (!c || !pi) return sdgpu_fail("null argument");
	if (newOmegaIdx >= c->omegaCnt) return sdgpu_fail("update_dual_col: observation %d out of range", newOmegaIdx);
	SD_CUDA(cudaSetDevice(c->device));
	if (sd_fused_update_ok(c)) {
		if (sd_launch_update_fused(c, pi, mubBar, currentIter, tol, newOmegaIdx >= 0 ? newOmegaIdx : -1)) return SDGPU_ERR;
	}
	else {
		if (sd_stage_vec(c, pi, c->rows + 1)) return SDGPU_ERR;
		const double *src = sd_staged_source(c, c->rows + 1, c->lambdaCnt);
		if (newOmegaIdx >= 0 && sd_launch_delta_col(c, newOmegaIdx, c->lambdaCnt)) return SDGPU_ERR;
		if (sd_launch_lambda(c, src, mubBar, tol, c->lambdaCnt, false)) return SDGPU_ERR;   // stocUpdate.c:78 (+ staging of :293-296)
		if (sd_launch_sigma(c, currentIter, tol, c->sigmaCnt)) return SDGPU_ERR;                     // :81
		if (sd_launch_delta_row(c, -1, c->omegaCnt)) return SDGPU_ERR;                               // :84-85 (kernel no-op unless the lambda was new)
	}
	if (sd_sync_state(c)) return SDGPU_ERR;
	if (lambdaIdx) *lambdaIdx = c->h_state->lambdaIdx;
	if (newLambdaFlag) *newLambdaFlag = c->h_state->newLambda;
	if (sigmaIdx) *sigmaIdx = c->h_state->sigmaIdx;
	if (newSigmaFlag) *newSigmaFlag = c->h_state->newSigma;
	return 0;
}

extern "C" int sdgpu_update_dual(sdgpu_ctx *c, const double *pi, double mubBar, int currentIter, double tol,
		int *lambdaIdx, int *newLambdaFlag, int *sigmaIdx, int *newSigmaFlag) {
	return sdgpu_update_dual_col(c, -1, pi, mubBar, currentIter, tol, lambdaIdx, newLambdaFlag, sigmaIdx, newSigmaFlag);
}

extern "C" int sdgpu_update_dual_bulk(sdgpu_ctx *c, int64_t n, const double *pis, const double *mubBar, const int32_t *iters,
		double tol, int32_t *lambdaIdx, int32_t *sigmaIdx) {
	if (!c || (!pis && n > 0)) return sdgpu_fail("null argument");
	if (n <= 0) return 0;
	SD_CUDA(cudaSetDevice(c->device));
	const size_t stride = (size_t) c->rows + 1;
	const int64_t chunk = std::max<int64_t>(1, ((int64_t) 64 << 20) / (int64_t) (stride * 8));
	double *d_pis = nullptr, *d_mub = nullptr; int32_t *d_it = nullptr, *d_li = nullptr, *d_si = nullptr;
	int64_t cm = std::min(chunk, n);
	int rc = 0;
	if (sd_alloc(&d_pis, (size_t) cm * stride) || sd_alloc(&d_mub, (size_t) cm) || sd_alloc(&d_it, (size_t) cm) ||
	    sd_alloc(&d_li, (size_t) cm) || sd_alloc(&d_si, (size_t) cm)) rc = SDGPU_ERR;
	std::vector<int32_t> defIters;
	for (int64_t i0 = 0; i0 < n && rc == 0; i0 += chunk) {
		int64_t m = std::min(chunk, n - i0);
		cudaError_t e = cudaMemcpyAsync(d_pis, pis + (size_t) i0 * stride, (size_t) m * stride * 8, cudaMemcpyHostToDevice, c->stream);
		if (e == cudaSuccess && mubBar) e = cudaMemcpyAsync(d_mub, mubBar + i0, (size_t) m * 8, cudaMemcpyHostToDevice, c->stream);
		if (e == cudaSuccess && iters) e = cudaMemcpyAsync(d_it, iters + i0, (size_t) m * 4, cudaMemcpyHostToDevice, c->stream);
		if (e != cudaSuccess) { rc = sdgpu_fail("update_dual_bulk copy: %s", cudaGetErrorString(e)); break; }
		if (tol < 0.0) {
			// synthetic loader: no dedup scan, vector i becomes lambda row / sigma row (count + i); delta rows are the
			// caller's to build (sdgpu_calc_delta_block)
			if (c->lambdaCnt + m > c->caps.maxLambda || c->sigmaCnt + m > c->caps.maxSigma) { rc = sdgpu_fail("update_dual_bulk: capacity exceeded"); break; }
			k_dual_bulk<<<sd_blocks(m, 64), 64, 0, c->stream>>>(d_pis, m, c->rows, mubBar ? d_mub : nullptr, iters ? d_it : nullptr, c->d_rvRows, c->R,
					c->d_bBarCol, c->d_bBarVal, c->bBarCnt, c->d_cbStart, c->d_cbRow, c->d_cbVal, c->n1c, c->n1cP, c->d_lambda, c->LP,
					c->d_sigmaPib, c->d_sigmaPiCk, c->d_sigmaPiCr, c->d_sigmaLam, c->d_sigmaCk, c->SP, c->lambdaCnt, c->sigmaCnt);
			k_bump_counts<<<1, 1, 0, c->stream>>>(c->d_state, c->d_hstate, 0, (int) m, (int) m, 0);
			sd_count_launch(c, 2);
			for (int64_t i = 0; i < m; i++) {
				if (lambdaIdx) lambdaIdx[i0 + i] = (int32_t) (c->lambdaCnt + i);
				if (sigmaIdx) sigmaIdx[i0 + i] = (int32_t) (c->sigmaCnt + i);
			}
			rc = sd_sync_state(c);
		}
		else {
			// the real find-or-append chain, vector after vector, with no host round trip in between: counts and
			// flags live in SdDevState, grids are sized by the host's upper bound on the counts
			if (sd_delta_ensure(c, std::min<int64_t>(c->caps.maxLambda, c->lambdaCnt + m), c->omegaCnt)) { rc = SDGPU_ERR; break; }
			for (int64_t i = 0; i < m; i++) {
				const double *d_pi = d_pis + (size_t) i * stride;
				double mb = mubBar ? mubBar[i0 + i] : 0.0;
				int it = iters ? iters[i0 + i] : (int) (i0 + i + 1);
				if (sd_launch_lambda(c, d_pi, mb, tol, c->lambdaCnt + i, false) || sd_launch_sigma(c, it, tol, c->sigmaCnt + i) ||
				    sd_launch_delta_row(c, -1, c->omegaCnt)) { rc = SDGPU_ERR; break; }
				k_record_pair<<<1, 1, 0, c->stream>>>(c->d_state, d_li, d_si, i);
				sd_count_launch(c);
			}
			if (rc == 0) rc = sd_sync_state(c);
			else cudaStreamSynchronize(c->stream);
			if (rc == 0 && lambdaIdx) if (cudaMemcpy(lambdaIdx + i0, d_li, (size_t) m * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = sdgpu_fail("copy back failed");
			if (rc == 0 && sigmaIdx) if (cudaMemcpy(sigmaIdx + i0, d_si, (size_t) m * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = sdgpu_fail("copy back failed");
		}
	}
	cudaFree(d_pis); cudaFree(d_mub); cudaFree(d_it); cudaFree(d_li); cudaFree(d_si);
	return rc;
}

extern "C" int sdgpu_calc_delta_block(sdgpu_ctx *c, int64_t l0, int64_t l1, int64_t o0, int64_t o1) {
	if (!c) return sdgpu_fail("null context");
	if (l0 < 0 || l1 > c->lambdaCnt || o0 < 0 || o1 > c->omegaCnt || l0 > l1 || o0 > o1) return sdgpu_fail("calc_delta_block: block out of range");
	if (l0 == l1 || o0 == o1) return 0;
	SD_CUDA(cudaSetDevice(c->device));
	if (sd_delta_ensure(c, l1, o1)) return SDGPU_ERR;
	const int64_t maxY = 32768;
	for (int64_t lb = l0; lb < l1; lb += maxY * DB_L) {
		int64_t le = std::min(l1, lb + maxY * DB_L);
		dim3 grid((unsigned) ((o1 - o0 + DB_O - 1) / DB_O), (unsigned) ((le - lb + DB_L - 1) / DB_L));
		k_delta_block_rhs<<<grid, 256, 0, c->stream>>>(c->d_lambda, c->LP, c->d_omega, c->NP, c->Rb, c->d_bLamPos, c->d_delta, c->Dcap, c->Q, lb, le, o0, o1);
		sd_count_launch(c);
	}
	if (c->Q > 0) {
		for (int64_t lb = l0; lb < l1; lb += maxY) {
			int64_t le = std::min(l1, lb + maxY);
			dim3 grid((unsigned) ((o1 - o0 + 127) / 128), (unsigned) (le - lb));
			k_delta_block_T<<<grid, 128, 0, c->stream>>>(c->d_lambda, c->LP, c->d_omega, c->NP, c->Rb, c->Q, c->d_cLamPos, c->d_cListStart, c->d_cList,
					c->d_delta, c->Dcap, lb, le, o0, o1);
			sd_count_launch(c);
		}
	}
	SD_CUDA(cudaStreamSynchronize(c->stream));
	SD_CUDA(cudaGetLastError());
	return 0;
}

// ---- basis records (host bookkeeping mirrored to the device for the argmax) -----------------------------------
extern "C" int sdgpu_basis_append(sdgpu_ctx *c, int ck, int feasFlag, int phiLength, const int32_t *sigmaIdx, const int32_t *omegaIdx) {
	if (!c || !sigmaIdx) return sdgpu_fail("null argument");
	if (phiLength < 0 || phiLength + 1 > c->caps.maxTerms) return sdgpu_fail("basis_append: phiLength %d exceeds maxTerms %d", phiLength, c->caps.maxTerms);
	if (phiLength > 0 && !omegaIdx) return sdgpu_fail("basis_append: omegaIdx required when phiLength > 0");
	if (c->basisCnt >= c->caps.maxBasis) return sdgpu_fail("basis capacity %lld exceeded", (long long) c->caps.maxBasis);
	for (int t = 0; t <= phiLength; t++) {
		if (sigmaIdx[t] < 0 || sigmaIdx[t] >= c->sigmaCnt) return sdgpu_fail("basis_append: sigma index %d out of range", sigmaIdx[t]);
		if (t > 0 && (omegaIdx[t] < 1 || c->rvOffset[2] + omegaIdx[t] > c->numRV)) return sdgpu_fail("basis_append: omegaIdx %d out of range", omegaIdx[t]);
	}
	SD_CUDA(cudaSetDevice(c->device));
	int b = (int) c->basisCnt;
	SdHostBasis hb;
	hb.ck = ck; hb.feas = feasFlag != 0; hb.phiLen = phiLength; hb.weight = 1;
	hb.sigmaIdx.assign(sigmaIdx, sigmaIdx + phiLength + 1);
	hb.omegaIdx.assign(phiLength + 1, 0);
	for (int t = 1; t <= phiLength; t++) hb.omegaIdx[t] = omegaIdx[t];
	const int nT = phiLength + 1;
	if (nT <= 16) {
		SdBasisRec r;
		r.b = b; r.ck = ck; r.feas = hb.feas; r.phiLen = phiLength; r.termStart = (int) c->termCnt; r.nT = nT;
		for (int t = 0; t < nT; t++) { r.sigma[t] = hb.sigmaIdx[t]; r.omega[t] = hb.omegaIdx[t]; }
		k_basis_commit<<<1, 1, 0, c->stream>>>(r, c->d_bCk, c->d_bFeas, c->d_bPhiLen, c->d_bTermStart, c->d_tSigma, c->d_tOmega, c->d_state, c->d_hstate);
		sd_count_launch(c);
	}
	else {
		int32_t *pi = c->h_pinI;      // [ck, feas, phiLen, termStart, termEnd, sigma..., omega...]
		if ((size_t) (5 + 2 * nT) > c->pinIcap) return sdgpu_fail("basis_append: too many terms for the staging buffer");
		pi[0] = ck; pi[1] = hb.feas; pi[2] = phiLength; pi[3] = (int32_t) c->termCnt; pi[4] = (int32_t) (c->termCnt + nT);
		for (int t = 0; t < nT; t++) { pi[5 + t] = hb.sigmaIdx[t]; pi[5 + nT + t] = hb.omegaIdx[t]; }
		SD_CUDA(cudaMemcpyAsync(c->d_bCk + b, pi + 0, 4, cudaMemcpyHostToDevice, c->stream));
		SD_CUDA(cudaMemcpyAsync(c->d_bFeas + b, pi + 1, 4, cudaMemcpyHostToDevice, c->stream));
		SD_CUDA(cudaMemcpyAsync(c->d_bPhiLen + b, pi + 2, 4, cudaMemcpyHostToDevice, c->stream));
		SD_CUDA(cudaMemcpyAsync(c->d_bTermStart + b, pi + 3, 8, cudaMemcpyHostToDevice, c->stream));
		SD_CUDA(cudaMemcpyAsync(c->d_tSigma + c->termCnt, pi + 5, (size_t) nT * 4, cudaMemcpyHostToDevice, c->stream));
		SD_CUDA(cudaMemcpyAsync(c->d_tOmega + c->termCnt, pi + 5 + nT, (size_t) nT * 4, cudaMemcpyHostToDevice, c->stream));
		k_bump_counts<<<1, 1, 0, c->stream>>>(c->d_state, c->d_hstate, 0, 0, 0, 1);
		sd_count_launch(c);
		SD_CUDA(cudaStreamSynchronize(c->stream));
	}
	if (c->rvd > 0) {
		c->hostMask.emplace_back(hb.feas ? std::vector<uint32_t>((size_t) c->NP / 32, 0xffffffffu) : std::vector<uint32_t>());
		if (hb.feas) {
			k_mask_fill_row<<<sd_blocks(c->NP / 32, 256), 256, 0, c->stream>>>(c->d_mask, c->caps.maxBasis, b, c->NP, nullptr, 0, 1);
			sd_count_launch(c);
		}
	}
	c->basis.push_back(std::move(hb));
	c->termCnt += nT; c->basisCnt++;
	c->maxPhiLen = std::max(c->maxPhiLen, phiLength);
	if (!feasFlag) c->anyInfeasibleBasis = true;
	return b;
}

extern "C" int sdgpu_basis_append_bulk(sdgpu_ctx *c, int64_t n, const int32_t *ck, const int32_t *feas, const int32_t *sigmaIdx) {
	if (!c || (n > 0 && (!ck || !sigmaIdx))) return sdgpu_fail("null argument");
	if (n <= 0) return (int) c->basisCnt;
	if (c->basisCnt + n > c->caps.maxBasis) return sdgpu_fail("basis capacity %lld exceeded", (long long) c->caps.maxBasis);
	for (int64_t i = 0; i < n; i++)
		if (sigmaIdx[i] < 0 || sigmaIdx[i] >= c->sigmaCnt) return sdgpu_fail("basis_append_bulk: sigma index %d out of range", sigmaIdx[i]);
	SD_CUDA(cudaSetDevice(c->device));
	const int b0 = (int) c->basisCnt;
	std::vector<int32_t> zeros((size_t) n, 0), ones((size_t) n, 1), starts((size_t) n + 1);
	for (int64_t i = 0; i <= n; i++) starts[i] = (int32_t) (c->termCnt + i);
	SD_CUDA(cudaMemcpyAsync(c->d_bCk + b0, ck, (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(c->d_bFeas + b0, feas ? feas : ones.data(), (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(c->d_bPhiLen + b0, zeros.data(), (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(c->d_bTermStart + b0, starts.data(), ((size_t) n + 1) * 4, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(c->d_tSigma + c->termCnt, sigmaIdx, (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(c->d_tOmega + c->termCnt, zeros.data(), (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
	k_bump_counts<<<1, 1, 0, c->stream>>>(c->d_state, c->d_hstate, 0, 0, 0, (int) n);
	sd_count_launch(c);
	for (int64_t i = 0; i < n; i++) {
		SdHostBasis hb;
		hb.ck = ck[i]; hb.feas = feas ? (feas[i] != 0) : 1; hb.phiLen = 0; hb.weight = 1;
		hb.sigmaIdx.assign(1, sigmaIdx[i]); hb.omegaIdx.assign(1, 0);
		if (!hb.feas) c->anyInfeasibleBasis = true;
		if (c->rvd > 0) {
			c->hostMask.emplace_back(hb.feas ? std::vector<uint32_t>((size_t) c->NP / 32, 0xffffffffu) : std::vector<uint32_t>());
			if (hb.feas) { k_mask_fill_row<<<sd_blocks(c->NP / 32, 256), 256, 0, c->stream>>>(c->d_mask, c->caps.maxBasis, b0 + i, c->NP, nullptr, 0, 1); sd_count_launch(c); }
		}
		c->basis.push_back(std::move(hb));
	}
	SD_CUDA(cudaStreamSynchronize(c->stream));
	c->termCnt += n; c->basisCnt += n;
	return b0;
}

extern "C" int sdgpu_basis_find_or_append(sdgpu_ctx *c, int retainBasis, int obsIdx, int ck, int feasFlag, int phiLength,
		const int32_t *sigmaIdx, const int32_t *omegaIdx, int *newBasisFlag) {
	if (!c || !sigmaIdx) return sdgpu_fail("null argument");
	if (newBasisFlag) *newBasisFlag = 1;
	if (!retainBasis) {                                   // stocUpdate.c:101-113
		if (obsIdx < 0 || obsIdx >= c->omegaCnt) return sdgpu_fail("basis_find_or_append: observation %d out of range", obsIdx);
		for (int64_t b = 0; b < c->basisCnt; b++) {
			const SdHostBasis &hb = c->basis[b];
			bool feasAtObs = hb.feas && (c->rvd == 0 || sd_hm_get(c, b, obsIdx));
			if (hb.phiLen == phiLength && feasAtObs && std::equal(hb.sigmaIdx.begin(), hb.sigmaIdx.end(), sigmaIdx)) {
				c->basis[b].weight++;
				if (newBasisFlag) *newBasisFlag = 0;
				return (int) b;
			}
		}
	}
	return sdgpu_basis_append(c, ck, feasFlag, phiLength, sigmaIdx, omegaIdx);
}

extern "C" int sdgpu_basis_set_obs_feasible_row(sdgpu_ctx *c, int basisIdx, const uint8_t *flags) {
	if (!c || !flags) return sdgpu_fail("null argument");
	if (basisIdx < 0 || basisIdx >= c->basisCnt || !c->basis[basisIdx].feas) return sdgpu_fail("set_obs_feasible_row: bad basis %d", basisIdx);
	if (c->rvd == 0) return 0;          // checkBasisFeasibility is constant true without random costs (randCost.c:208)
	SD_CUDA(cudaSetDevice(c->device));
	if (sd_aux_reserve(c, (size_t) std::max<int64_t>(1, c->omegaCnt))) return SDGPU_ERR;
	memcpy(c->h_aux, flags, (size_t) c->omegaCnt);                       // the kernel reads the flags through the mapped alias
	k_mask_fill_row<<<sd_blocks(c->NP / 32, 256), 256, 0, c->stream>>>(c->d_mask, c->caps.maxBasis, basisIdx, c->NP, c->d_aux, c->omegaCnt, 0);
	sd_count_launch(c);
	SD_CUDA(cudaStreamSynchronize(c->stream));
	for (int64_t o = 0; o < c->omegaCnt; o++) sd_hm_set(c, basisIdx, o, flags[o] != 0);
	return 0;
}

extern "C" int sdgpu_basis_set_obs_feasible_col(sdgpu_ctx *c, int obsIdx, const uint8_t *flags) {
	if (!c || !flags) return sdgpu_fail("null argument");
	if (obsIdx < 0 || obsIdx >= c->caps.maxOmega) return sdgpu_fail("set_obs_feasible_col: bad observation %d", obsIdx);
	if (c->rvd == 0 || c->basisCnt == 0) return 0;
	SD_CUDA(cudaSetDevice(c->device));
	if (sd_aux_reserve(c, (size_t) c->basisCnt)) return SDGPU_ERR;
	memcpy(c->h_aux, flags, (size_t) c->basisCnt);
	k_mask_fill_col<<<sd_blocks(c->basisCnt, 256), 256, 0, c->stream>>>(c->d_mask, c->caps.maxBasis, obsIdx, c->d_aux, c->d_bFeas, c->basisCnt);
	sd_count_launch(c);
	SD_CUDA(cudaStreamSynchronize(c->stream));
	for (int64_t b = 0; b < c->basisCnt; b++) if (c->basis[b].feas) sd_hm_set(c, b, obsIdx, flags[b] != 0);
	return 0;
}

extern "C" int sdgpu_basis_set_obs_feasible(sdgpu_ctx *c, int basisIdx, int obsIdx, int flag) {
	if (!c) return sdgpu_fail("null context");
	if (basisIdx < 0 || basisIdx >= c->basisCnt || !c->basis[basisIdx].feas) return sdgpu_fail("set_obs_feasible: bad basis %d", basisIdx);
	if (obsIdx < 0 || obsIdx >= c->caps.maxOmega) return sdgpu_fail("set_obs_feasible: bad observation %d", obsIdx);
	if (c->rvd == 0) return 0;
	SD_CUDA(cudaSetDevice(c->device));
	k_mask_set_bit<<<1, 1, 0, c->stream>>>(c->d_mask, c->caps.maxBasis, basisIdx, obsIdx, flag != 0);
	SD_LAUNCH_OK("k_mask_set_bit");
	sd_count_launch(c);
	SD_CUDA(cudaStreamSynchronize(c->stream));
	sd_hm_set(c, basisIdx, obsIdx, flag != 0);
	return 0;
}

// ---- readers --------------------------------------------------------------------------------------------------
extern "C" int sdgpu_get_omega(sdgpu_ctx *c, int idx, double *vals, int *weight) {
	if (!c) return sdgpu_fail("null context");
	if (idx < 0 || idx >= c->omegaCnt) return sdgpu_fail("get_omega: index %d out of range", idx);
	SD_CUDA(cudaSetDevice(c->device));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	if (vals && c->numRV > 0) SD_CUDA(cudaMemcpy2D(vals + 1, 8, c->d_omega + idx, (size_t) c->NP * 8, 8, c->numRV, cudaMemcpyDeviceToHost));
	if (weight) { int32_t w; SD_CUDA(cudaMemcpy(&w, c->d_omegaW + idx, 4, cudaMemcpyDeviceToHost)); *weight = w; }
	return 0;
}

extern "C" int sdgpu_get_lambda(sdgpu_ctx *c, int idx, double *vals) {
	if (!c || !vals) return sdgpu_fail("null argument");
	if (idx < 0 || idx >= c->lambdaCnt) return sdgpu_fail("get_lambda: index %d out of range", idx);
	SD_CUDA(cudaSetDevice(c->device));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	if (c->R > 0) SD_CUDA(cudaMemcpy2D(vals + 1, 8, c->d_lambda + idx, (size_t) c->LP * 8, 8, c->R, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int sdgpu_get_sigma(sdgpu_ctx *c, int idx, double *pib, double *piC, int *lambdaIdx, int *ck) {
	if (!c) return sdgpu_fail("null context");
	if (idx < 0 || idx >= c->sigmaCnt) return sdgpu_fail("get_sigma: index %d out of range", idx);
	SD_CUDA(cudaSetDevice(c->device));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	if (pib) SD_CUDA(cudaMemcpy(pib, c->d_sigmaPib + idx, 8, cudaMemcpyDeviceToHost));
	if (piC && c->n1c > 0) SD_CUDA(cudaMemcpy(piC + 1, c->d_sigmaPiCr + (size_t) idx * c->n1cP, (size_t) c->n1c * 8, cudaMemcpyDeviceToHost));
	if (lambdaIdx) { int32_t v; SD_CUDA(cudaMemcpy(&v, c->d_sigmaLam + idx, 4, cudaMemcpyDeviceToHost)); *lambdaIdx = v; }
	if (ck) { int32_t v; SD_CUDA(cudaMemcpy(&v, c->d_sigmaCk + idx, 4, cudaMemcpyDeviceToHost)); *ck = v; }
	return 0;
}

extern "C" int sdgpu_get_delta(sdgpu_ctx *c, int lambdaIdx, int obsIdx, double *pib, double *piC) {
	if (!c) return sdgpu_fail("null context");
	if (lambdaIdx < 0 || lambdaIdx >= c->lambdaCnt || obsIdx < 0 || obsIdx >= c->omegaCnt) return sdgpu_fail("get_delta: (%d, %d) out of range", lambdaIdx, obsIdx);
	SD_CUDA(cudaSetDevice(c->device));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	const double *base = c->d_delta + sd_delta_off(c->Dcap, c->Q, lambdaIdx, 0, obsIdx);
	if (pib) SD_CUDA(cudaMemcpy(pib, base, 8, cudaMemcpyDeviceToHost));
	if (piC && c->Q > 0) SD_CUDA(cudaMemcpy2D(piC + 1, 8, base + SD_TILE_W, (size_t) SD_TILE_W * 8, 8, c->Q, cudaMemcpyDeviceToHost));
	return 0;
}

// a rectangular block of delta.pib (row-major [l1-l0][o1-o0]) for host code that walks the table (optimal.c:203-221)
extern "C" int sdgpu_get_delta_block(sdgpu_ctx *c, int64_t l0, int64_t l1, int64_t o0, int64_t o1, int plane, double *out) {
	if (!c || !out) return sdgpu_fail("null argument");
	if (l0 < 0 || l1 > c->lambdaCnt || o0 < 0 || o1 > c->omegaCnt || l0 > l1 || o0 > o1 || plane < 0 || plane > c->Q) return sdgpu_fail("get_delta_block: out of range");
	SD_CUDA(cudaSetDevice(c->device));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	for (int64_t l = l0; l < l1; l++) {
		int64_t o = o0;
		while (o < o1) {
			int64_t tEnd = std::min<int64_t>(o1, (o / SD_TILE_W + 1) * SD_TILE_W);
			SD_CUDA(cudaMemcpy(out + (size_t) (l - l0) * (o1 - o0) + (o - o0), c->d_delta + sd_delta_off(c->Dcap, c->Q, l, plane, o),
					(size_t) (tEnd - o) * 8, cudaMemcpyDeviceToHost));
			o = tEnd;
		}
	}
	return 0;
}

// ---- checkBasisFeasibility on the device (randCost.c:202-258) ----------------------------------------------------
static int sd_feas_alloc(sdgpu_ctx *c) {
	if (c->d_fPiDet) return 0;
	if (sd_alloc(&c->d_fPiDet, (size_t) c->caps.maxBasis * c->rows) || sd_alloc(&c->d_fPhi, (size_t) c->termCap * c->rows) ||
	    sd_alloc(&c->d_fGBar, (size_t) c->caps.maxBasis * c->cols) || sd_alloc(&c->d_fPsi, (size_t) c->termCap * c->cols) ||
	    sd_alloc(&c->d_fCstat, (size_t) c->caps.maxBasis * c->cols) || sd_alloc(&c->d_fHas, (size_t) c->caps.maxBasis) ||
	    sd_alloc(&c->d_fFlags, (size_t) std::max<int64_t>(c->caps.maxBasis, c->NP)))
		return SDGPU_ERR;
	SD_CUDA(cudaMemset(c->d_fHas, 0, (size_t) c->caps.maxBasis));
	return 0;
}

extern "C" int sdgpu_set_cost_coords(sdgpu_ctx *c, const int32_t *rvdOmCols, const char *senx) {
	if (!c || !senx || (c->rvd > 0 && !rvdOmCols)) return sdgpu_fail("null argument");
	SD_CUDA(cudaSetDevice(c->device));
	if (!c->d_senx && (sd_alloc(&c->d_senx, (size_t) c->rows) || sd_alloc(&c->d_rvdOmCols, (size_t) std::max(1, c->rvd)))) return SDGPU_ERR;
	SD_CUDA(cudaMemcpy(c->d_senx, senx, (size_t) c->rows, cudaMemcpyHostToDevice));
	if (c->rvd > 0) SD_CUDA(cudaMemcpy(c->d_rvdOmCols, rvdOmCols + 1, (size_t) c->rvd * 4, cudaMemcpyHostToDevice));
	return 0;
}

extern "C" int sdgpu_basis_set_feas_data(sdgpu_ctx *c, int basisIdx, const double *piDet, const double *phi, const double *gBar,
		const double *psiVal, const int32_t *cstat) {
	if (!c || !piDet || !gBar || !cstat) return sdgpu_fail("null argument");
	if (basisIdx < 0 || basisIdx >= c->basisCnt) return sdgpu_fail("basis_set_feas_data: bad basis %d", basisIdx);
	if (c->rvd == 0) return 0;                       // constant true without random costs (randCost.c:208)
	const SdHostBasis &hb = c->basis[basisIdx];
	if (hb.phiLen > 0 && (!phi || !psiVal)) return sdgpu_fail("basis_set_feas_data: phi / psi required when phiLength > 0");
	SD_CUDA(cudaSetDevice(c->device));
	if (!c->d_senx) return sdgpu_fail("basis_set_feas_data: call sdgpu_set_cost_coords first");
	if (sd_feas_alloc(c)) return SDGPU_ERR;
	// term offset of this basis: recompute from the host list (terms are appended in basis order)
	int64_t t0 = 0;
	for (int b = 0; b < basisIdx; b++) t0 += c->basis[b].phiLen + 1;
	SD_CUDA(cudaMemcpyAsync(c->d_fPiDet + (size_t) basisIdx * c->rows, piDet + 1, (size_t) c->rows * 8, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(c->d_fGBar + (size_t) basisIdx * c->cols, gBar + 1, (size_t) c->cols * 8, cudaMemcpyHostToDevice, c->stream));
	std::vector<int8_t> cs((size_t) c->cols);
	for (int i = 0; i < c->cols; i++) cs[i] = (int8_t) cstat[i + 1];
	SD_CUDA(cudaMemcpyAsync(c->d_fCstat + (size_t) basisIdx * c->cols, cs.data(), (size_t) c->cols, cudaMemcpyHostToDevice, c->stream));
	std::vector<double> psiT;
	for (int n = 0; n < hb.phiLen; n++) {
		SD_CUDA(cudaMemcpyAsync(c->d_fPhi + (size_t) (t0 + 1 + n) * c->rows, phi + (size_t) n * (c->rows + 1) + 1, (size_t) c->rows * 8, cudaMemcpyHostToDevice, c->stream));
		psiT.resize((size_t) c->cols);
		for (int i = 0; i < c->cols; i++) psiT[i] = psiVal[(size_t) i * hb.phiLen + n];
		SD_CUDA(cudaMemcpyAsync(c->d_fPsi + (size_t) (t0 + 1 + n) * c->cols, psiT.data(), (size_t) c->cols * 8, cudaMemcpyHostToDevice, c->stream));
		SD_CUDA(cudaStreamSynchronize(c->stream));           // psiT is reused
	}
	const uint8_t one = 1;
	SD_CUDA(cudaMemcpyAsync(c->d_fHas + basisIdx, &one, 1, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	return 0;
}

static FeasArgs sd_feas_args(sdgpu_ctx *c, double tol) {
	FeasArgs a;
	a.omega = c->d_omega; a.NP = c->NP; a.rvOffset2 = c->rvOffset[2]; a.rvd = c->rvd;
	a.rvdOmCols = c->d_rvdOmCols; a.senx = c->d_senx; a.rows = c->rows; a.cols = c->cols;
	a.bPhiLen = c->d_bPhiLen; a.bTermStart = c->d_bTermStart; a.bFeas = c->d_bFeas; a.tOmega = c->d_tOmega;
	a.piDet = c->d_fPiDet; a.phi = c->d_fPhi; a.gBar = c->d_fGBar; a.psi = c->d_fPsi; a.cstat = c->d_fCstat; a.has = c->d_fHas;
	a.tol = tol; a.mask = c->d_mask; a.Bcap = c->caps.maxBasis; a.flags = c->d_fFlags;
	return a;
}

extern "C" int sdgpu_check_feasibility_obs(sdgpu_ctx *c, int obsIdx, double tol, uint8_t *flagsOut) {
	if (!c) return sdgpu_fail("null context");
	if (obsIdx < 0 || obsIdx >= c->omegaCnt) return sdgpu_fail("check_feasibility_obs: bad observation %d", obsIdx);
	if (c->rvd == 0 || c->basisCnt == 0) { if (flagsOut) memset(flagsOut, 1, (size_t) c->basisCnt); return 0; }
	if (!c->d_fPiDet) return sdgpu_fail("check_feasibility_obs: no basis carries feasibility data yet");
	SD_CUDA(cudaSetDevice(c->device));
	k_feas_obs<<<(unsigned) c->basisCnt, 128, (size_t) std::max(1, c->rvd) * 8, c->stream>>>(sd_feas_args(c, tol), obsIdx);
	sd_count_launch(c);
	std::vector<uint8_t> f((size_t) c->basisCnt);
	SD_CUDA(cudaMemcpyAsync(f.data(), c->d_fFlags, f.size(), cudaMemcpyDeviceToHost, c->stream));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	for (int64_t b = 0; b < c->basisCnt; b++) {
		if (f[b] == 2) { f[b] = c->basis[b].feas ? (uint8_t) sd_hm_get(c, b, obsIdx) : 0; continue; }   // untouched: keep what the mask holds
		sd_hm_set(c, b, obsIdx, f[b] != 0);
	}
	if (flagsOut) memcpy(flagsOut, f.data(), f.size());
	return 0;
}

extern "C" int sdgpu_check_feasibility_basis(sdgpu_ctx *c, int basisIdx, double tol, uint8_t *flagsOut) {
	if (!c) return sdgpu_fail("null context");
	if (basisIdx < 0 || basisIdx >= c->basisCnt || !c->basis[basisIdx].feas) return sdgpu_fail("check_feasibility_basis: bad basis %d", basisIdx);
	if (c->rvd == 0 || c->omegaCnt == 0) { if (flagsOut) memset(flagsOut, 1, (size_t) c->omegaCnt); return 0; }
	if (!c->d_fPiDet) return sdgpu_fail("check_feasibility_basis: basis %d carries no feasibility data", basisIdx);
	SD_CUDA(cudaSetDevice(c->device));
	k_feas_basis<<<(unsigned) c->omegaCnt, 128, (size_t) std::max(1, c->rvd) * 8, c->stream>>>(sd_feas_args(c, tol), basisIdx);
	sd_count_launch(c);
	std::vector<uint8_t> f((size_t) c->omegaCnt);
	SD_CUDA(cudaMemcpyAsync(f.data(), c->d_fFlags, f.size(), cudaMemcpyDeviceToHost, c->stream));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	for (int64_t o = 0; o < c->omegaCnt; o++) sd_hm_set(c, basisIdx, o, f[o] != 0);
	if (flagsOut) memcpy(flagsOut, f.data(), f.size());
	return 0;
}
