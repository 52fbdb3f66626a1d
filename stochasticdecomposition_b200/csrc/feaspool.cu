// feaspool.cu -- feasibility cuts on the device (problems without relatively complete recourse):
//   sdgpu_feas_cuts          the gathers of updtFeasCutPool                       cuts.c:465-517
//   sdgpu_feas_pool_update   updtFeasCutPool including the duplicate test of addCut2Pool (cuts.c:643-655) on a device-resident pool
//   sdgpu_feas_pool_check    checkFeasCutPool                                      cuts.c:521-567
// Citations are file:line under /root/reference/twoSD_src.
//
// The duplicate test is first-come-first-kept with a tolerance, i.e. NOT transitive: whether raw cut j survives depends on which
// earlier raw cuts survived.  The O(n^2 n1) comparisons are done in parallel (raw x pool, and the lower triangle raw x raw as a bit
// matrix); only the O(n) resolution of "kept" walks the cuts in order, in one warp.
#include <algorithm>
#include <cstring>
#include <vector>

#include "sdgpu_internal.cuh"

__device__ __forceinline__ double fp_abs(double x) { return x > 0.0 ? x : -x; }     // DBL_ABS: a NaN difference is never "> tol"

// raw feasibility cuts (cuts.c:478-486): one thread per (observation, infeasible basis) pair, the scatter into beta done
// in the reference's order (CCols first, then rvCols) so that coinciding columns add up identically
__global__ void k_feas_cuts(int nPairs, const int32_t *__restrict__ pairObs, const int32_t *__restrict__ pairBasis,
		const int32_t *__restrict__ bTermStart, const int32_t *__restrict__ tSigma, const double *__restrict__ sigmaPib,
		const double *__restrict__ sigmaPiCr, const int32_t *__restrict__ sigmaLam, int n1c, int n1cP, const double *__restrict__ delta,
		int64_t Dcap, int Q, const int32_t *__restrict__ CCols, const int32_t *__restrict__ rvCols, int n1,
		double *__restrict__ alpha, double *__restrict__ beta) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= nPairs) return;
	const int o = pairObs[p], b = pairBasis[p];
	const int s = tSigma[bTermStart[b]], l = sigmaLam[s];
	const size_t rowStride = (size_t) (1 + Q) * SD_TILE_W;
	const double *cell = delta + (size_t) (o / SD_TILE_W) * Dcap * rowStride + (size_t) l * rowStride + (o % SD_TILE_W);
	double *bt = beta + (size_t) p * (n1 + 1);
	for (int i = 0; i <= n1; i++) bt[i] = 0.0;
	alpha[p] = __dadd_rn(sigmaPib[s], cell[0]);                                                  // cuts.c:481
	for (int k = 0; k < n1c; k++) bt[CCols[k]] = __dadd_rn(bt[CCols[k]], sigmaPiCr[(size_t) s * n1cP + k]);   // :483-484
	for (int q = 0; q < Q; q++) bt[rvCols[q]] = __dadd_rn(bt[rvCols[q]], cell[(size_t) (1 + q) * SD_TILE_W]);   // :485-486
}


#define FP_BATCH 4096        // raw cuts resolved at a time (triangle bit matrix: 4096 x 128 words = 2 MiB)

// cuts.c:645-647: |alpha - alpha'| < tol (strict) and equalVector(beta, beta', lenX, tol) (every |diff| <= tol)
__device__ __forceinline__ bool fp_same_cut(double aA, const double *__restrict__ aB, double bA, const double *__restrict__ bB, int n1, double tol) {
	if (!(fp_abs(aA - bA) < tol)) return false;
	for (int c = 1; c <= n1; c++) if (fp_abs(aB[c] - bB[c]) > tol) return false;
	return true;
}

// raw cut j against every pool entry: poolMatch[j] != 0 if any matches.  grid (pool chunks of 256, raw cuts)
__global__ void k_fp_match_pool(int nRaw, const double *__restrict__ rawA, const double *__restrict__ rawB, int nPool,
		const double *__restrict__ poolA, const double *__restrict__ poolB, int n1, double tol, int *__restrict__ poolMatch) {
	extern __shared__ double s_b[];
	const int j = blockIdx.y;
	for (int c = threadIdx.x; c <= n1; c += blockDim.x) s_b[c] = rawB[(size_t) j * (n1 + 1) + c];
	__syncthreads();
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < nPool && fp_same_cut(rawA[j], s_b, poolA[i], poolB + (size_t) i * (n1 + 1), n1, tol)) poolMatch[j] = 1;
}

// raw cut j against the earlier raw cuts of its batch: bit j' of row j.  grid (ceil(nRaw / 256), nRaw); a warp writes one word
__global__ void k_fp_match_self(int nRaw, const double *__restrict__ rawA, const double *__restrict__ rawB, int n1, double tol,
		unsigned *__restrict__ tri, int words, int *__restrict__ rowAny) {
	extern __shared__ double s_b[];
	const int j = blockIdx.y;
	if ((int) (blockIdx.x * blockDim.x) >= j) return;                       // nothing of this block lies below the diagonal
	for (int c = threadIdx.x; c <= n1; c += blockDim.x) s_b[c] = rawB[(size_t) j * (n1 + 1) + c];
	__syncthreads();
	const int jp = blockIdx.x * blockDim.x + threadIdx.x;
	const bool m = jp < j && fp_same_cut(rawA[j], s_b, rawA[jp], rawB + (size_t) jp * (n1 + 1), n1, tol);
	const unsigned w = __ballot_sync(0xffffffffu, m);
	if ((threadIdx.x & 31) == 0 && jp / 32 < words) {
		tri[(size_t) j * words + jp / 32] = w;
		if (w) rowAny[j] = 1;
	}
}

// in raw order: kept[j] = no pool match and no EARLIER KEPT raw cut matches; rank[j] = kept cuts before j.  One warp.
__global__ void k_fp_resolve(int nRaw, const int *__restrict__ poolMatch, const unsigned *__restrict__ tri, int words, const int *__restrict__ rowAny,
		int *__restrict__ kept, int *__restrict__ rank, int *__restrict__ keptCount) {
	__shared__ unsigned s_kept[FP_BATCH / 32];
	const int lane = threadIdx.x;
	for (int w = lane; w < FP_BATCH / 32; w += 32) s_kept[w] = 0;
	__syncwarp();
	int count = 0;
	for (int j = 0; j < nRaw; j++) {
		bool dup = poolMatch[j] != 0;
		if (!dup && rowAny[j]) {
			unsigned hit = 0;
			for (int w = lane; w * 32 < j; w += 32) hit |= tri[(size_t) j * words + w] & s_kept[w];      // words holding bits j' < j (all written)
			dup = __any_sync(0xffffffffu, hit != 0);
		}
		if (lane == 0) {
			kept[j] = dup ? 0 : 1; rank[j] = count;
			if (!dup) s_kept[j / 32] |= 1u << (j % 32);
		}
		if (!dup) count++;
		__syncwarp();
	}
	if (lane == 0) *keptCount = count;
}

__global__ void k_fp_append(int nRaw, const double *__restrict__ rawA, const double *__restrict__ rawB, const int *__restrict__ kept,
		const int *__restrict__ rank, int n1, double *__restrict__ poolA, double *__restrict__ poolB, int poolCnt) {
	const int j = blockIdx.x;
	if (!kept[j]) return;
	const size_t dst = (size_t) poolCnt + rank[j];
	if (threadIdx.x == 0) poolA[dst] = rawA[j];
	for (int c = threadIdx.x; c <= n1; c += blockDim.x) poolB[dst * (n1 + 1) + c] = rawB[(size_t) j * (n1 + 1) + c];
}

// checkFeasCutPool cuts.c:521-567, one warp per pool cut.  action: 0 nothing, 1 incumbent violates it and it is not in the master yet
// (add), 2 incumbent violates it but an equal cut is already in the master, 3 candidate violates it (add).
__global__ void k_fp_check(int nPool, const double *__restrict__ poolA, const double *__restrict__ poolB, int n1, int nF,
		const double *__restrict__ fA, const double *__restrict__ fB, const double *__restrict__ incumbX, const double *__restrict__ candidX,
		double tol, int32_t *__restrict__ action) {
	const int idx = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
	if (idx >= nPool) return;
	const double alpha = poolA[idx];
	const double *beta = poolB + (size_t) idx * (n1 + 1);
	bool dup = false;
	for (int c = lane; c < nF; c += 32) if (fp_same_cut(alpha, beta, fA[c], fB + (size_t) c * (n1 + 1), n1, tol)) dup = true;     // :528-535
	dup = __any_sync(0xffffffffu, dup);
	if (lane != 0) return;
	double bx = 0.0;
	for (int c = 1; c <= n1; c++) bx = __dadd_rn(bx, __dmul_rn(beta[c], incumbX[c]));                  // vXv :538
	int act = 0;
	if (bx < alpha) act = dup ? 2 : 1;                                                                 // :539-547
	else if (!dup) {
		double bc = 0.0;
		for (int c = 1; c <= n1; c++) bc = __dadd_rn(bc, __dmul_rn(beta[c], candidX[c]));              // :552
		if (bc < alpha) act = 3;                                                                       // :554-557
	}
	action[idx] = act;
}

static inline int fp_blocks(int64_t n, int t) { return (int) std::max<int64_t>(1, (n + t - 1) / t); }
#define sd_blocks fp_blocks

extern "C" int sdgpu_feas_cuts(sdgpu_ctx *c, int obsFirst, int obsLast, int basisFirst, int basisLast, int maxOut, double *alpha, double *beta) {
	if (!c || (maxOut > 0 && (!alpha || !beta))) return sdgpu_fail("null argument");
	if (obsFirst < 0 || obsLast > c->omegaCnt || basisFirst < 0 || basisLast > c->basisCnt || obsFirst > obsLast || basisFirst > basisLast)
		return sdgpu_fail("feas_cuts: range out of bounds");
	std::vector<int32_t> po, pb;
	for (int o = obsFirst; o < obsLast; o++)                       // cuts.c:473-475 loop order
		for (int b = basisFirst; b < basisLast; b++)
			if (!c->basis[b].feas) { po.push_back(o); pb.push_back(b); }
	const int n = (int) std::min<size_t>(po.size(), (size_t) std::max(0, maxOut));
	if (po.size() > (size_t) std::max(0, maxOut)) return sdgpu_fail("feas_cuts: %zu cuts do not fit maxOut = %d", po.size(), maxOut);
	if (n == 0) return 0;
	SD_CUDA(cudaSetDevice(c->device));
	const size_t n1p = (size_t) c->n1 + 1;
	const size_t outBytes = (size_t) n * (1 + n1p) * 8;
	if (sd_scratch_reserve(c, outBytes + (size_t) 2 * n * 4)) return SDGPU_ERR;
	double *d_o = reinterpret_cast<double *>(c->d_scratch);
	int32_t *d_p = reinterpret_cast<int32_t *>(c->d_scratch + outBytes);
	SD_CUDA(cudaMemcpyAsync(d_p, po.data(), (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(d_p + n, pb.data(), (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
	k_feas_cuts<<<sd_blocks(n, 128), 128, 0, c->stream>>>(n, d_p, d_p + n, c->d_bTermStart, c->d_tSigma, c->d_sigmaPib, c->d_sigmaPiCr, c->d_sigmaLam,
			c->n1c, c->n1cP, c->d_delta, c->Dcap, c->Q, c->d_CCols, c->d_rvCols, c->n1, d_o, d_o + n);
	SD_LAUNCH_OK("k_feas_cuts");
	sd_count_launch(c);
	SD_CUDA(cudaMemcpyAsync(alpha, d_o, (size_t) n * 8, cudaMemcpyDeviceToHost, c->stream));
	SD_CUDA(cudaMemcpyAsync(beta, d_o + n, (size_t) n * n1p * 8, cudaMemcpyDeviceToHost, c->stream));
	cudaError_t e = cudaStreamSynchronize(c->stream);
	if (e != cudaSuccess) return sdgpu_fail("feas_cuts: %s", cudaGetErrorString(e));
	return n;
}

#undef sd_blocks

static int fp_reserve(sdgpu_ctx *c, int64_t need) {
	if (need <= c->fpCap) return 0;
	const int64_t cap = std::max<int64_t>(need * 2, 1024);
	const size_t n1p = (size_t) c->n1 + 1;
	double *a = nullptr, *b = nullptr;
	if (cudaMalloc((void **) &a, (size_t) cap * 8) != cudaSuccess || cudaMalloc((void **) &b, (size_t) cap * n1p * 8) != cudaSuccess) {
		if (a) cudaFree(a);
		return sdgpu_fail("feasibility-cut pool of %lld cuts does not fit device memory", (long long) cap);
	}
	SD_CUDA(cudaStreamSynchronize(c->stream));
	if (c->fpCnt > 0) {
		SD_CUDA(cudaMemcpy(a, c->d_fpAlpha, (size_t) c->fpCnt * 8, cudaMemcpyDeviceToDevice));
		SD_CUDA(cudaMemcpy(b, c->d_fpBeta, (size_t) c->fpCnt * n1p * 8, cudaMemcpyDeviceToDevice));
	}
	if (c->d_fpAlpha) cudaFree(c->d_fpAlpha);
	if (c->d_fpBeta) cudaFree(c->d_fpBeta);
	c->d_fpAlpha = a; c->d_fpBeta = b; c->fpCap = cap;
	return 0;
}

// one batch of (observation, infeasible basis) pairs, in the reference's loop order, through gather -> match -> resolve -> append
static int fp_add_batch(sdgpu_ctx *c, const int32_t *po, const int32_t *pb, int n, double tol) {
	const size_t n1p = (size_t) c->n1 + 1;
	const int words = (n + 31) / 32;
	if (fp_reserve(c, c->fpCnt + n)) return SDGPU_ERR;
	// scratch: rawA [n] | rawB [n][n1+1] | pairs 2n ints | poolMatch n | rowAny n | kept n | rank n | keptCount 1 (+pad) | tri [n][words]
	const size_t oRawB = (size_t) n * 8, oPairs = oRawB + (size_t) n * n1p * 8, oInts = oPairs + (size_t) 2 * n * 4;
	const size_t oTri = (oInts + ((size_t) 4 * n + 2) * 4 + 7) / 8 * 8, total = oTri + (size_t) n * words * 4;
	if (sd_scratch_reserve(c, total)) return SDGPU_ERR;
	double *rawA = reinterpret_cast<double *>(c->d_scratch), *rawB = reinterpret_cast<double *>(c->d_scratch + oRawB);
	int32_t *d_p = reinterpret_cast<int32_t *>(c->d_scratch + oPairs);
	int *poolMatch = reinterpret_cast<int *>(c->d_scratch + oInts), *rowAny = poolMatch + n, *kept = rowAny + n, *rank = kept + n, *keptCount = rank + n;
	unsigned *tri = reinterpret_cast<unsigned *>(c->d_scratch + oTri);
	SD_CUDA(cudaMemcpyAsync(d_p, po, (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(d_p + n, pb, (size_t) n * 4, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemsetAsync(poolMatch, 0, ((size_t) 4 * n + 2) * 4, c->stream));
	k_feas_cuts<<<fp_blocks(n, 128), 128, 0, c->stream>>>(n, d_p, d_p + n, c->d_bTermStart, c->d_tSigma, c->d_sigmaPib, c->d_sigmaPiCr, c->d_sigmaLam,
			c->n1c, c->n1cP, c->d_delta, c->Dcap, c->Q, c->d_CCols, c->d_rvCols, c->n1, rawA, rawB);
	SD_LAUNCH_OK("k_feas_cuts");
	if (c->fpCnt > 0) {
		k_fp_match_pool<<<dim3((unsigned) fp_blocks(c->fpCnt, 256), (unsigned) n), 256, n1p * 8, c->stream>>>(n, rawA, rawB, (int) c->fpCnt, c->d_fpAlpha, c->d_fpBeta, c->n1, tol, poolMatch);
		SD_LAUNCH_OK("k_fp_match_pool");
		sd_count_launch(c);
	}
	if (n > 1) {
		k_fp_match_self<<<dim3((unsigned) fp_blocks(n, 256), (unsigned) n), 256, n1p * 8, c->stream>>>(n, rawA, rawB, c->n1, tol, tri, words, rowAny);
		SD_LAUNCH_OK("k_fp_match_self");
		sd_count_launch(c);
	}
	k_fp_resolve<<<1, 32, 0, c->stream>>>(n, poolMatch, tri, words, rowAny, kept, rank, keptCount);
	k_fp_append<<<n, 64, 0, c->stream>>>(n, rawA, rawB, kept, rank, c->n1, c->d_fpAlpha, c->d_fpBeta, (int) c->fpCnt);
	SD_LAUNCH_OK("k_fp_append");
	sd_count_launch(c, 3);
	int hk = 0;
	SD_CUDA(cudaMemcpyAsync(&hk, keptCount, 4, cudaMemcpyDeviceToHost, c->stream));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	c->fpCnt += hk;
	return 0;
}

// updtFeasCutPool cuts.c:465-517 with the pool on the device.  fUpdt is cell->fUpdt (in / out).  Returns the pool size.
extern "C" int sdgpu_feas_pool_update(sdgpu_ctx *c, int *fUpdt, double tol) {
	if (!c || !fUpdt) return sdgpu_fail("null argument");
	if (fUpdt[0] < 0 || fUpdt[0] > c->basisCnt || fUpdt[1] < 0 || fUpdt[1] > c->omegaCnt) return sdgpu_fail("feas_pool_update: fUpdt out of range");
	SD_CUDA(cudaSetDevice(c->device));
	std::vector<int32_t> po, pb;
	for (int o = fUpdt[1]; o < c->omegaCnt; o++)                    // cuts.c:472-490: new observations x the bases seen by the last update
		for (int b = 0; b < fUpdt[0]; b++)
			if (!c->basis[b].feas) { po.push_back(o); pb.push_back(b); }
	for (int o = 0; o < c->omegaCnt; o++)                           // cuts.c:494-512: every observation x the new bases
		for (int b = fUpdt[0]; b < c->basisCnt; b++)
			if (!c->basis[b].feas) { po.push_back(o); pb.push_back(b); }
	for (size_t i0 = 0; i0 < po.size(); i0 += FP_BATCH) {
		const int n = (int) std::min<size_t>(FP_BATCH, po.size() - i0);
		if (fp_add_batch(c, po.data() + i0, pb.data() + i0, n, tol)) return SDGPU_ERR;
	}
	fUpdt[1] = (int) c->omegaCnt; fUpdt[0] = (int) c->basisCnt;
	return (int) c->fpCnt;
}

extern "C" int sdgpu_feas_pool_size(sdgpu_ctx *c) {
	if (!c) return sdgpu_fail("null context");
	return (int) c->fpCnt;
}

extern "C" int sdgpu_feas_pool_get(sdgpu_ctx *c, int first, int count, double *alpha, double *beta) {
	if (!c || (count > 0 && (!alpha || !beta))) return sdgpu_fail("null argument");
	if (first < 0 || count < 0 || first + count > c->fpCnt) return sdgpu_fail("feas_pool_get: range out of bounds");
	if (count == 0) return 0;
	SD_CUDA(cudaSetDevice(c->device));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	SD_CUDA(cudaMemcpy(alpha, c->d_fpAlpha + first, (size_t) count * 8, cudaMemcpyDeviceToHost));
	SD_CUDA(cudaMemcpy(beta, c->d_fpBeta + (size_t) first * (c->n1 + 1), (size_t) count * (c->n1 + 1) * 8, cudaMemcpyDeviceToHost));
	return count;
}

// checkFeasCutPool cuts.c:521-567: which pool cuts the host has to add to the master (action 1 or 3, in pool order), and infeasIncumb
extern "C" int sdgpu_feas_pool_check(sdgpu_ctx *c, int nFcuts, const double *fAlpha, const double *fBeta, const double *incumbX,
		const double *candidX, double tol, int32_t *action, int *infeasIncumb) {
	if (!c || !incumbX || !candidX || (nFcuts > 0 && (!fAlpha || !fBeta))) return sdgpu_fail("null argument");
	if (infeasIncumb) *infeasIncumb = 0;
	const int n = (int) c->fpCnt;
	if (n == 0) return 0;
	if (!action) return sdgpu_fail("null argument");
	SD_CUDA(cudaSetDevice(c->device));
	const size_t n1p = (size_t) c->n1 + 1;
	const size_t oFB = (size_t) std::max(1, nFcuts) * 8, oX = oFB + (size_t) std::max(1, nFcuts) * n1p * 8, oAct = oX + 2 * n1p * 8, total = oAct + (size_t) n * 4;
	if (sd_scratch_reserve(c, total)) return SDGPU_ERR;
	double *d_fA = reinterpret_cast<double *>(c->d_scratch), *d_fB = reinterpret_cast<double *>(c->d_scratch + oFB);
	double *d_ix = reinterpret_cast<double *>(c->d_scratch + oX), *d_cx = d_ix + n1p;
	int32_t *d_act = reinterpret_cast<int32_t *>(c->d_scratch + oAct);
	if (nFcuts > 0) {
		SD_CUDA(cudaMemcpyAsync(d_fA, fAlpha, (size_t) nFcuts * 8, cudaMemcpyHostToDevice, c->stream));
		SD_CUDA(cudaMemcpyAsync(d_fB, fBeta, (size_t) nFcuts * n1p * 8, cudaMemcpyHostToDevice, c->stream));
	}
	SD_CUDA(cudaMemcpyAsync(d_ix, incumbX, n1p * 8, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(d_cx, candidX, n1p * 8, cudaMemcpyHostToDevice, c->stream));
	k_fp_check<<<fp_blocks((int64_t) n * 32, 128), 128, 0, c->stream>>>(n, c->d_fpAlpha, c->d_fpBeta, c->n1, nFcuts, d_fA, d_fB, d_ix, d_cx, tol, d_act);
	SD_LAUNCH_OK("k_fp_check");
	sd_count_launch(c);
	SD_CUDA(cudaMemcpyAsync(action, d_act, (size_t) n * 4, cudaMemcpyDeviceToHost, c->stream));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	if (infeasIncumb) for (int i = 0; i < n; i++) if (action[i] == 1 || action[i] == 2) *infeasIncumb = 1;
	return n;
}
