// cut.cu -- SD cut formation on the device: per-observation argmax over all stored bases (computeIstar,
// stocUpdate.c:142-190) fused with the weighted reduction into (alpha, beta) and the dual-stability sums
// (SDCut, cuts.c:91-194), cut heights / aging (cuts.c:197-227, master.c:152,174) and reformCuts
// (optimal.c:187-236).  Citations are file:line under /root/reference/twoSD_src.
//
// Pipeline of one cut (three launches on the context's stream, one host wait at the end):
//   k_cut_prep     x as a kernel parameter; per basis (sigma.pib, piCbarX = sigma.piC . x[CCols], lambda row, window)
//                                                                             cuts.c:105-106, stocUpdate.c:151-167
//   k_sweep_*      2-D grid (observation tile x basis chunk): stream the delta tile, keep running (max, first index) per
//                  observation for the old and the new window   stocUpdate.c:161-184.  Five kernels, picked by shape and size:
//                  k_sweep_tma (RHS-only, TMA ring), k_sweep_tma_q (random T elements), k_sweep_tma_gen (random cost: multi-term
//                  bases + obsFeasible mask, term-linear ring), k_sweep_ldg<Q, MASK> (small cuts), k_sweep_general (small random-cost cuts)
//   k_cut_merge    one CTA per 64..512 observations: merge chunk maxima in index order, pick iStar (cuts.c:124-125,136-140),
//                  accumulate w*(sigma.pib + delta.pib), w*sigma.piC, w*delta.piC, cummOld, cummAll   cuts.c:127-168;
//                  the last CTA to finish sums the per-CTA partials in CTA order, scatters into beta (cuts.c:155-167),
//                  applies alpha/k, beta/k (cuts.c:184-188) and writes the cut into mapped pinned host memory
//   [sharded: un-normalised vector -> NCCL all-reduce of n1+4 doubles -> k_cut_normalise]
//
// Scores are evaluated with the reference's operation order and without FMA contraction, so they are
// bit-identical to the CPU path; (max, index) merges keep the LOWEST index among equal maxima (strict '>'
// in stocUpdate.c:178) and the old window wins ties against the new one (cuts.c:125): iStar is bit-exact.
#include <cfloat>
#include <cmath>
#include <climits>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <type_traits>

#include "sdgpu_internal.cuh"

// ======================================================================================================
// small device helpers
// ======================================================================================================
__device__ __forceinline__ double2 ld_stream_f64x2(const double *p) {
	double2 v;
	asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
	return v;
}

__device__ __forceinline__ double sd_warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}

// deterministic block sum (fixed tree); result valid in thread 0
__device__ __forceinline__ double sd_block_sum(double v, double *s_red) {
	int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	v = sd_warp_sum(v);
	__syncthreads();
	if (lane == 0) s_red[warp] = v;
	__syncthreads();
	double t = 0.0;
	if (warp == 0) {
		t = lane < nw ? s_red[lane] : 0.0;
		t = sd_warp_sum(t);
	}
	return t;
}

// four block sums at once: each value goes through exactly the tree of sd_block_sum (so the results have the same bits), but the
// four share one pair of barriers.  s_red holds 4 x 32 doubles.  Results valid in thread 0.
__device__ __forceinline__ void sd_block_sum4(double (&v)[4], double *s_red) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
	for (int u = 0; u < 4; u++) v[u] = sd_warp_sum(v[u]);
	__syncthreads();
	if (lane == 0) {
#pragma unroll
		for (int u = 0; u < 4; u++) s_red[u * 32 + warp] = v[u];
	}
	__syncthreads();
	if (warp == 0) {
#pragma unroll
		for (int u = 0; u < 4; u++) {
			double t = lane < nw ? s_red[u * 32 + lane] : 0.0;
			v[u] = sd_warp_sum(t);
		}
	}
}

// sum_k col[k*stride] * s_x[k], left to right from 0.0 with separate multiply and add (vXv, cuts.c:106); the loads of a
// batch of 32 are issued together (past the end: the last element again, not added) so the chain of dependent adds does not
// also serialise the memory latency -- three round trips for the 89 columns of ssn instead of twelve
template <int BATCH = 32>
__device__ __forceinline__ double sd_dot_strided(const double *__restrict__ col, size_t stride, const double *s_x, int n) {
	double acc = 0.0;
	for (int k = 0; k < n; k += BATCH) {
		double v[BATCH];
#pragma unroll
		for (int u = 0; u < BATCH; u++) v[u] = col[(size_t) min(k + u, n - 1) * stride];
#pragma unroll
		for (int u = 0; u < BATCH; u++) if (k + u < n) acc = __dadd_rn(acc, __dmul_rn(v[u], s_x[k + u]));
	}
	return acc;
}

// Fused prologue of a cut: x arrives as a kernel parameter (no H2D copy), thread i computes piCbarX of sigma i (only
// when the general sweep needs the whole vector) and the descriptor of basis i, whose piCbarX is recomputed from its
// own sigma row with the same left-to-right sum, hence the same bits.
struct SdXParam { double v[256]; };

template <int BATCH>                     // loads in flight per dot: 64 while the grid is a single wave at 146 registers, 32 beyond
__global__ void k_cut_prep(SdXParam xp, const double *__restrict__ xDevIn, double *__restrict__ xDevOut, int n1,
		const double *__restrict__ piCk, int64_t SP, int n1c, const int32_t *__restrict__ CCols, int sigmaCnt, double *__restrict__ piCbarXAll,
		const int32_t *__restrict__ bCk, const int32_t *__restrict__ bFeas, const int32_t *__restrict__ bTermStart,
		const int32_t *__restrict__ tSigma, const double *__restrict__ sigmaPib, const int32_t *__restrict__ sigmaLam,
		int basisCnt, int split, int cutoff,
		double *__restrict__ descA, double *__restrict__ descC, int32_t *__restrict__ descRow, int32_t *__restrict__ descWin,
		const int32_t *__restrict__ tOmega, double *__restrict__ termA, double *__restrict__ termC, int32_t *__restrict__ termRow,
		int32_t *__restrict__ termMeta, int32_t *__restrict__ termBasis) {
	extern __shared__ double s_x[];
	sd_pdl_launch_dependents();              // the sweep may be scheduled now; its CTAs wait in sd_pdl_wait() until this grid has finished
	for (int k = threadIdx.x; k < n1c; k += blockDim.x) s_x[k] = xDevIn ? xDevIn[CCols[k]] : xp.v[CCols[k]];
	if (blockIdx.x == 0 && !xDevIn)
		for (int i = threadIdx.x; i <= n1; i += blockDim.x) xDevOut[i] = xp.v[i];
	__syncthreads();
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (piCbarXAll && i < sigmaCnt) piCbarXAll[i] = sd_dot_strided(piCk + i, (size_t) SP, s_x, n1c);
	if (i < basisCnt) {
		const int s = tSigma[bTermStart[i]];
		const double acc = sd_dot_strided<BATCH>(piCk + s, (size_t) SP, s_x, n1c);   // BATCH = 64: ssn's 89 columns in two round trips
		const int ck = bCk[i];
		int win = 0;
		if (bFeas[i]) {
			if (ck <= cutoff) win = (ck > -INT_MAX) ? 1 : 0;
			else win = split ? 2 : 0;
		}
		descA[i] = sigmaPib[s]; descC[i] = acc; descRow[i] = sigmaLam[s]; descWin[i] = win;
		if (termA) {                                      // term-linear form for the random-cost sweep: one descriptor per (basis, term)
			const int ts = bTermStart[i], te = bTermStart[i + 1];
			for (int t = ts; t < te; t++) {
				const int st = tSigma[t];
				termA[t] = sigmaPib[st];
				termC[t] = (t == ts) ? acc : sd_dot_strided(piCk + st, (size_t) SP, s_x, n1c);
				termRow[t] = sigmaLam[st];
				termMeta[t] = win | (t == te - 1 ? 4 : 0) | ((t == ts ? 0 : tOmega[t]) << 8);
				termBasis[t] = i;
			}
		}
	}
}

// ======================================================================================================
// K6: the sweep.  Variant 1: plain streaming loads (LDG.128, 8 rows in flight per thread).
// ======================================================================================================
struct SweepArgs {
	const double *delta; int64_t Dcap; int Q;
	const double *descA, *descC; const int32_t *descRow, *descWin;
	int basisCnt, chunkSize, nChunks;
	const uint32_t *mask; int64_t Bcap;     // bit-packed obsFeasible: [tile][Bcap][SD_MASK_WORDS]
	const double *x; const int32_t *rvCOmCols;
	double *partV; int32_t *partI; int64_t NP;
	int rev;                        // load-based sweep: walk the chunk's bases in descending order (see k_sweep_ldg)
};

#define SW_BATCH 256    // basis descriptors staged per shared-memory refill
#define SW_UNROLL 8     // delta rows in flight per thread

// FUSED: the descriptors of k_cut_prep are computed here, by the CTA that needs them (small cuts: one launch and ~8 us less; every
// observation tile repeats the piCbarX dots of its chunk, which is nothing next to a launch as long as the grid is a single wave)
struct SweepPrepArgs {
	SdXParam xp;
	const double *piCk; int64_t SP; int n1c; const int32_t *CCols;
	const int32_t *bCk, *bFeas, *bTermStart, *tSigma; const double *sigmaPib; const int32_t *sigmaLam;
	int split, cutoff;
};

// k_cut_prep's arithmetic for one basis (cuts.c:105-106, stocUpdate.c:151-167): (sigma.pib, piCbarX) and (lambda row, window)
__device__ __forceinline__ void sd_basis_descriptor(const SweepPrepArgs &pa, const double *s_xc, int b, double2 &ac, int2 &rw) {
	const int sg = pa.tSigma[pa.bTermStart[b]];
	const double acc = sd_dot_strided(pa.piCk + sg, (size_t) pa.SP, s_xc, pa.n1c);
	const int ck = pa.bCk[b];
	int win = 0;
	if (pa.bFeas[b]) {
		if (ck <= pa.cutoff) win = (ck > -INT_MAX) ? 1 : 0;
		else win = pa.split ? 2 : 0;
	}
	ac = make_double2(pa.sigmaPib[sg], acc);
	rw = make_int2(pa.sigmaLam[sg], win);
}

// REV: the chunk's bases are walked in DESCENDING order.  The library alternates the direction from cut to cut: a delta table
// somewhat larger than the 126 MB L2 (real problems late in a run: 5 000 x 5 000 doubles = 200 MB) is then read back to front right
// after it was read front to back, so the part read last -- still in L2 -- is read first.  Descending order keeps the reference's
// tie rule with '>=' instead of '>': among equal scores the one met LAST, i.e. the lowest basis index, stays (stocUpdate.c:178
// keeps the first one met in ascending order -- the same basis).  A score of exactly -DBL_MAX can then sit in a partial maximum;
// the merge kernel's strict '>' against its -DBL_MAX start drops it, as the reference's loop would.
template <bool HAS_Q, bool HAS_MASK, bool FUSED, bool REV>
__global__ void __launch_bounds__(SD_SWEEP_THREADS) k_sweep_ldg(SweepArgs a, typename std::conditional<FUSED, SweepPrepArgs, int>::type pa_) {
	const SweepPrepArgs &pa = *reinterpret_cast<const SweepPrepArgs *>(&pa_);      // only dereferenced when FUSED
	__shared__ double2 s_ac[SW_BATCH];       // (sigma.pib, piCbarX)
	__shared__ int2 s_rw[SW_BATCH];          // (lambda row, window)
	__shared__ double s_xq[HAS_Q ? SD_MAX_Q : 1];
	extern __shared__ double s_xc[];         // FUSED: x[CCols[k]]
	const int tile = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
	if (FUSED) for (int k = tid; k < pa.n1c; k += blockDim.x) s_xc[k] = pa.xp.v[pa.CCols[k]];
	const int b0 = chunk * a.chunkSize, b1 = min(a.basisCnt, b0 + a.chunkSize), nTot = b1 - b0;
	const size_t rowStride = (size_t) (1 + a.Q) * SD_TILE_W;
	const double *tileBase = a.delta + (size_t) tile * a.Dcap * rowStride + 2 * tid;
	sd_pdl_wait();                           // descriptors and x come from k_cut_prep
	sd_pdl_launch_dependents();              // the merge kernel may be scheduled; it waits for this grid to finish
	if (HAS_Q) {
		for (int j = tid; j < a.Q; j += blockDim.x) s_xq[j] = a.x[a.rvCOmCols[j]];
	}
	double oV0 = -DBL_MAX, oV1 = -DBL_MAX, nV0 = -DBL_MAX, nV1 = -DBL_MAX;
	int oI0 = -1, oI1 = -1, nI0 = -1, nI1 = -1;
#define SD_BASIS_AT(i) (REV ? (b1 - 1 - (i)) : (b0 + (i)))                  // i-th basis of the walk
#define SD_BETTER(sc, best) (REV ? ((sc) >= (best)) : ((sc) > (best)))

	for (int off = 0; off < nTot; off += SW_BATCH) {
		__syncthreads();
		{
			const bool ok = off + tid < nTot;
			const int b = SD_BASIS_AT(off + tid);
			if (!FUSED) {
				s_ac[tid] = ok ? make_double2(a.descA[b], a.descC[b]) : make_double2(0.0, 0.0);
				s_rw[tid] = ok ? make_int2(a.descRow[b], a.descWin[b]) : make_int2(0, 0);
			}
			else if (ok) sd_basis_descriptor(pa, s_xc, b, s_ac[tid], s_rw[tid]);
			else { s_ac[tid] = make_double2(0.0, 0.0); s_rw[tid] = make_int2(0, 0); }
		}
		__syncthreads();
		const int n = min(SW_BATCH, nTot - off);
		for (int j = 0; j < n; j += SW_UNROLL) {
			double2 d[SW_UNROLL];
#pragma unroll
			for (int u = 0; u < SW_UNROLL; u++)          // rows past the end alias row s_rw[..].x == 0 and are ignored (window 0)
				d[u] = ld_stream_f64x2(tileBase + (size_t) s_rw[j + u].x * rowStride);
#pragma unroll
			for (int u = 0; u < SW_UNROLL; u++) {
				const int win = s_rw[j + u].y;
				if (win == 0) continue;
				const double2 ac = s_ac[j + u];
				const int b = SD_BASIS_AT(off + j + u);
				// ((sigma.pib + delta.pib) - piCbarX)   stocUpdate.c:174
				double s0 = __dsub_rn(__dadd_rn(ac.x, d[u].x), ac.y);
				double s1 = __dsub_rn(__dadd_rn(ac.x, d[u].y), ac.y);
				if (HAS_Q) {                         // - delta.piC . x[rvCOmCols]   stocUpdate.c:175
					const double *pc = tileBase + (size_t) s_rw[j + u].x * rowStride;
					double dx0 = 0.0, dx1 = 0.0;
					for (int q = 0; q < a.Q; q++) {
						double2 p = ld_stream_f64x2(pc + (size_t) (1 + q) * SD_TILE_W);
						dx0 = __dadd_rn(dx0, __dmul_rn(p.x, s_xq[q]));
						dx1 = __dadd_rn(dx1, __dmul_rn(p.y, s_xq[q]));
					}
					s0 = __dsub_rn(__dadd_rn(0.0, s0), dx0);
					s1 = __dsub_rn(__dadd_rn(0.0, s1), dx1);
				}
				bool f0 = true, f1 = true;
				if (HAS_MASK) {
					const unsigned mw = a.mask[((size_t) tile * a.Bcap + b) * SD_MASK_WORDS + (tid >> 4)] >> ((2 * tid) & 31);   // 16 threads share a word
					f0 = (mw & 1u) != 0; f1 = (mw & 2u) != 0;
				}
				if (win == 1) {
					if (f0 && SD_BETTER(s0, oV0)) { oV0 = s0; oI0 = b; }
					if (f1 && SD_BETTER(s1, oV1)) { oV1 = s1; oI1 = b; }
				}
				else {
					if (f0 && SD_BETTER(s0, nV0)) { nV0 = s0; nI0 = b; }
					if (f1 && SD_BETTER(s1, nV1)) { nV1 = s1; nI1 = b; }
				}
			}
		}
	}
#undef SD_BASIS_AT
#undef SD_BETTER
	const size_t o = (size_t) tile * SD_TILE_W + 2 * tid;
	const size_t oldAt = ((size_t) 0 * a.nChunks + chunk) * a.NP + o, newAt = ((size_t) 1 * a.nChunks + chunk) * a.NP + o;
	*reinterpret_cast<double2 *>(a.partV + oldAt) = make_double2(oV0, oV1);
	*reinterpret_cast<int2 *>(a.partI + oldAt) = make_int2(oI0, oI1);
	*reinterpret_cast<double2 *>(a.partV + newAt) = make_double2(nV0, nV1);
	*reinterpret_cast<int2 *>(a.partI + newAt) = make_int2(nI0, nI1);
}

// ======================================================================================================
// K6 for problems with very few random right-hand sides (Rb <= 8, e.g. pgp2 with 3): the contraction
// delta.pib[l][o] = sum_j omega_o[j] * lambda_l[rvbOmRows[j]] costs 2 Rb flops, fewer cycles than the 8 bytes of the stored entry cost
// in HBM time, so the sweep recomputes it from the factors instead of streaming the delta table -- the score[lambda, omega]
// contraction fused with the per-observation argmax.  The observation's Rb values live in registers, the lambda entries of 256 bases
// at a time in shared memory (broadcast reads); the sum runs in index order from 0.0 with separate multiply and add, exactly the
// operations calcDelta performed (stocUpdate.c:218, :244), so the recomputed entry has the stored entry's bits and iStar is unchanged.
// FP64-pipe bound (no HBM stream at all).  RHS-only problems without a feasibility mask (Q = 0, rvdOmCnt = 0).
// ======================================================================================================
#ifndef SD_RC_AUTO_MAX
#define SD_RC_AUTO_MAX 4       // automatic use of the recompute sweep up to this many random right-hand sides
#endif
struct SweepRcArgs {
	const double *omega; int64_t NP; const double *lambda; int64_t LP; const int32_t *bLamPos;
	const double *descA, *descC; const int32_t *descRow, *descWin;
	int basisCnt, chunkSize, nChunks;
	double *partV; int32_t *partI;
};

template <int RB, bool FUSED>
__global__ void __launch_bounds__(SD_SWEEP_THREADS) k_sweep_recompute(SweepRcArgs a, typename std::conditional<FUSED, SweepPrepArgs, int>::type pa_) {
	const SweepPrepArgs &pa = *reinterpret_cast<const SweepPrepArgs *>(&pa_);      // only dereferenced when FUSED
	extern __shared__ double s_xc[];                         // FUSED: x[CCols[k]]
	constexpr int RBP = (RB + 1) & ~1;                       // lambda entries per basis in shared memory, padded to a whole number of double2
	__shared__ double2 s_ac[SW_BATCH];                       // (sigma.pib, piCbarX)
	__shared__ int s_win[SW_BATCH];
	__shared__ __align__(16) double s_lam[SW_BATCH][RBP];
	__shared__ int s_pos[RB];
	const int tile = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
	const int b0 = chunk * a.chunkSize, b1 = min(a.basisCnt, b0 + a.chunkSize);
	const size_t o = (size_t) tile * SD_TILE_W + 2 * tid;
	if (tid < RB) s_pos[tid] = a.bLamPos[tid];
	if (FUSED) for (int k = tid; k < pa.n1c; k += blockDim.x) s_xc[k] = pa.xp.v[pa.CCols[k]];
	double om0[RB], om1[RB];
#pragma unroll
	for (int j = 0; j < RB; j++) {
		const double2 v = *reinterpret_cast<const double2 *>(a.omega + (size_t) j * a.NP + o);
		om0[j] = v.x; om1[j] = v.y;
	}
	double oV0 = -DBL_MAX, oV1 = -DBL_MAX, nV0 = -DBL_MAX, nV1 = -DBL_MAX;
	int oI0 = -1, oI1 = -1, nI0 = -1, nI1 = -1;
	sd_pdl_wait();                           // descriptors come from k_cut_prep (the observation values above are table data)
	sd_pdl_launch_dependents();
	for (int base = b0; base < b1; base += SW_BATCH) {
		__syncthreads();
		{
			const int b = base + tid;
			const bool ok = b < b1;
			double2 ac = make_double2(0.0, 0.0);
			int2 rw = make_int2(0, 0);
			if (ok) {
				if (FUSED) sd_basis_descriptor(pa, s_xc, b, ac, rw);
				else { ac = make_double2(a.descA[b], a.descC[b]); rw = make_int2(a.descRow[b], a.descWin[b]); }
			}
			s_ac[tid] = ac;
			s_win[tid] = rw.y;
			const int row = rw.x;
#pragma unroll
			for (int j = 0; j < RBP; j++) {                  // expandVector: 0.0 where the row carries no lambda entry
				const int p = j < RB ? s_pos[j] : -1;
				s_lam[tid][j] = (ok && p >= 0) ? a.lambda[(size_t) p * a.LP + row] : 0.0;
			}
		}
		__syncthreads();
		const int n = min(SW_BATCH, b1 - base);
#pragma unroll 2
		for (int r = 0; r < n; r++) {
			const int win = s_win[r];
			if (win == 0) continue;
			double lam[RBP];
#pragma unroll
			for (int j = 0; j < RBP; j += 2) { const double2 v = *reinterpret_cast<const double2 *>(&s_lam[r][j]); lam[j] = v.x; lam[j + 1] = v.y; }
			double d0 = 0.0, d1 = 0.0;                       // vXvSparse, index order   stocUpdate.c:218
#pragma unroll
			for (int j = 0; j < RB; j++) {
				d0 = __dadd_rn(d0, __dmul_rn(om0[j], lam[j]));
				d1 = __dadd_rn(d1, __dmul_rn(om1[j], lam[j]));
			}
			const double2 ac = s_ac[r];
			const double s0 = __dsub_rn(__dadd_rn(ac.x, d0), ac.y);      // stocUpdate.c:174
			const double s1 = __dsub_rn(__dadd_rn(ac.x, d1), ac.y);
			const int b = base + r;
			if (win == 1) {
				if (s0 > oV0) { oV0 = s0; oI0 = b; }
				if (s1 > oV1) { oV1 = s1; oI1 = b; }
			}
			else {
				if (s0 > nV0) { nV0 = s0; nI0 = b; }
				if (s1 > nV1) { nV1 = s1; nI1 = b; }
			}
		}
	}
	const size_t oldAt = ((size_t) 0 * a.nChunks + chunk) * a.NP + o, newAt = ((size_t) 1 * a.nChunks + chunk) * a.NP + o;
	*reinterpret_cast<double2 *>(a.partV + oldAt) = make_double2(oV0, oV1);
	*reinterpret_cast<int2 *>(a.partI + oldAt) = make_int2(oI0, oI1);
	*reinterpret_cast<double2 *>(a.partV + newAt) = make_double2(nV0, nV1);
	*reinterpret_cast<int2 *>(a.partI + newAt) = make_int2(nI0, nI1);
}

// ======================================================================================================
// K6, variant 2: the same sweep fed by the TMA engine.  One producer warp issues 1-D bulk copies
// (cp.async.bulk.shared::cluster.global, 4 KiB = one dual row of the observation tile each) into a ring of
// shared-memory stages guarded by mbarriers; eight consumer warps read their 16 bytes per row with LDS.128 and
// run the identical score / running-maximum code.  Bytes in flight are set by the ring (TMA_STAGES x TMA_ROWS x
// 4 KiB per CTA), not by registers.  RHS-only bases without a feasibility mask (Q = 0, rvdOmCnt = 0).
// ======================================================================================================
#define TMA_CONSUMERS SD_SWEEP_THREADS
#define TMA_THREADS (TMA_CONSUMERS + 32)
#define TMA_ROW_BYTES (SD_TILE_W * 8)

__device__ __forceinline__ uint32_t sd_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void sd_mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(sd_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sd_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sd_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sd_mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(sd_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sd_mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				: "=r"(done) : "r"(sd_smem_u32(bar)), "r"(parity) : "memory");
	} while (!done);
}
__device__ __forceinline__ void sd_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(sd_smem_u32(dst)), "l"(src), "r"(bytes), "r"(sd_smem_u32(bar)) : "memory");
}

// <ROWS, STAGES, CTAS>: rows per stage, ring depth, resident CTAs per SM.  Default <8, 2, 3>: 64 KiB ring per CTA, three CTAs
// per SM => 192 KiB of bulk copies in flight per SM and 24 consumer warps to hide the FP64 compare chains.
template <int TMA_ROWS, int TMA_STAGES, int TMA_CTAS>
__global__ void __launch_bounds__(TMA_THREADS, TMA_CTAS) k_sweep_tma(SweepArgs a) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *ring = reinterpret_cast<double *>(smem_raw);                                   // [stage][row][512]
	uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t) TMA_STAGES * TMA_ROWS * TMA_ROW_BYTES);
	uint64_t *empty = full + TMA_STAGES;
	double2 *s_ac = reinterpret_cast<double2 *>(empty + TMA_STAGES);                        // [SW_BATCH] (sigma.pib, piCbarX)
	int *s_win = reinterpret_cast<int *>(s_ac + SW_BATCH);                                 // [SW_BATCH]
	const int tile = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
	const int b0 = chunk * a.chunkSize, b1 = min(a.basisCnt, b0 + a.chunkSize);
	const int nRows = b1 - b0, nIter = (nRows + TMA_ROWS - 1) / TMA_ROWS;
	const double *tileBase = a.delta + (size_t) tile * a.Dcap * SD_TILE_W;                  // Q == 0: one plane per row
	if (tid == 0) {
		for (int s = 0; s < TMA_STAGES; s++) { sd_mbar_init(&full[s], 1); sd_mbar_init(&empty[s], TMA_CONSUMERS / 32); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	sd_pdl_wait();                           // the ring is set up; descriptors come from k_cut_prep
	sd_pdl_launch_dependents();

	if (tid >= TMA_CONSUMERS) {
		// ---------------- producer warp: one lane feeds the ring -------------------------------------------------
		if (tid == TMA_CONSUMERS) {
			for (int it = 0; it < nIter; it++) {
				const int s = it % TMA_STAGES;
				sd_mbar_wait(&empty[s], ((it / TMA_STAGES) & 1) ^ 1);
				const int r0 = it * TMA_ROWS, nr = min(TMA_ROWS, nRows - r0);
				sd_mbar_expect_tx(&full[s], (uint32_t) nr * TMA_ROW_BYTES);
				for (int r = 0; r < nr; r++) {
					const int row = a.descRow[b0 + r0 + r];
					sd_bulk_g2s(ring + ((size_t) s * TMA_ROWS + r) * SD_TILE_W, tileBase + (size_t) row * SD_TILE_W, TMA_ROW_BYTES, &full[s]);
				}
			}
		}
		return;
	}

	// ---------------- consumers --------------------------------------------------------------------------------------
	double oV0 = -DBL_MAX, oV1 = -DBL_MAX, nV0 = -DBL_MAX, nV1 = -DBL_MAX;
	int oI0 = -1, oI1 = -1, nI0 = -1, nI1 = -1;
	for (int it = 0; it < nIter; it++) {
		const int r0 = it * TMA_ROWS;
		if (r0 % SW_BATCH == 0) {                                                           // refill the descriptor batch (consumers only)
			asm volatile("bar.sync 1, %0;" :: "n"(TMA_CONSUMERS) : "memory");
			const int b = b0 + r0 + tid;
			const bool ok = b < b1;
			s_ac[tid] = ok ? make_double2(a.descA[b], a.descC[b]) : make_double2(0.0, 0.0);
			s_win[tid] = ok ? a.descWin[b] : 0;
			asm volatile("bar.sync 1, %0;" :: "n"(TMA_CONSUMERS) : "memory");
		}
		const int s = it % TMA_STAGES;
		sd_mbar_wait(&full[s], (it / TMA_STAGES) & 1);
		const double2 *stage = reinterpret_cast<const double2 *>(ring + (size_t) s * TMA_ROWS * SD_TILE_W) + tid;
		double2 d[TMA_ROWS];
#pragma unroll
		for (int r = 0; r < TMA_ROWS; r++) d[r] = stage[(size_t) r * (SD_TILE_W / 2)];
		__syncwarp();
		if ((tid & 31) == 0) sd_mbar_arrive(&empty[s]);                                    // this warp is done with the stage
#pragma unroll
		for (int r = 0; r < TMA_ROWS; r++) {
			const int j = (r0 + r) % SW_BATCH;
			const int win = s_win[j];
			if (win == 0) continue;
			const double2 ac = s_ac[j];
			const int b = b0 + r0 + r;
			const double s0 = __dsub_rn(__dadd_rn(ac.x, d[r].x), ac.y);                     // stocUpdate.c:174
			const double s1 = __dsub_rn(__dadd_rn(ac.x, d[r].y), ac.y);
			if (win == 1) {
				if (s0 > oV0) { oV0 = s0; oI0 = b; }
				if (s1 > oV1) { oV1 = s1; oI1 = b; }
			}
			else {
				if (s0 > nV0) { nV0 = s0; nI0 = b; }
				if (s1 > nV1) { nV1 = s1; nI1 = b; }
			}
		}
	}
	const size_t o = (size_t) tile * SD_TILE_W + 2 * tid;
	const size_t oldAt = ((size_t) 0 * a.nChunks + chunk) * a.NP + o, newAt = ((size_t) 1 * a.nChunks + chunk) * a.NP + o;
	*reinterpret_cast<double2 *>(a.partV + oldAt) = make_double2(oV0, oV1);
	*reinterpret_cast<int2 *>(a.partI + oldAt) = make_int2(oI0, oI1);
	*reinterpret_cast<double2 *>(a.partV + newAt) = make_double2(nV0, nV1);
	*reinterpret_cast<int2 *>(a.partI + newAt) = make_int2(nI0, nI1);
}

// The RHS-only ring for tables in which several bases share a lambda row (calcSigma appends a sigma -- and stochasticUpdates a basis
// -- whenever pi x bBar or pi x Cbar differ, while the entries on the random rows, the lambda, are already stored: stocUpdate.c:299-318).
// The host keeps the bases sorted by (lambda row, basis index) and numbers the groups (= distinct rows) densely; the sweep walks that
// list and copies every distinct row of a chunk ONCE: the delta stream shrinks from one row per basis to one row per distinct lambda
// (SURVEY.md section 8d counts the algorithmic bytes per distinct row).  The walk is no longer in basis order, so the running maximum
// is kept lexicographically -- greater score, or equal score and lower basis index -- which is what the strict '>' of
// stocUpdate.c:178 yields in basis order.
struct SweepGrpArgs {
	const double *delta; int64_t Dcap;
	const double *descA, *descC; const int32_t *descWin;
	const int32_t *entBasis;                             // the bases sorted by (lambda row, basis index)
	const int32_t *entGroup, *groupRow;                  // dense group number of each entry (one group per distinct row), row of each group
	int basisCnt, chunkSize, nChunks;
	double *partV; int32_t *partI; int64_t NP;
};

#ifndef GRP_ROWS
#define GRP_ROWS 8      // distinct rows per ring stage
#define GRP_STAGES 2
#define GRP_BATCH 256   // entry descriptors per shared-memory refill
#define GRP_CTAS 3      // CTAs per SM the kernel is compiled for
#endif

// greater score, or equal score and lower basis index; the equal case is rare, so the common path is one compare
#define SD_LEX_UPDATE(sc, bb, bestV, bestI) do { if ((sc) >= (bestV)) { if ((sc) > (bestV) || (bb) < (bestI)) { (bestV) = (sc); (bestI) = (bb); } } } while (0)

// Shape of the kernel (round 2; the round-1 form had eight ENTRIES per stage and two observations per thread and stopped at ~1.4e12
// pairs/s for every group size: with g bases per row a stage held only 8/g rows, and it spent ~18 instructions per pair, issue slots
// 75-83 % -- profiles/r02_sweep_grp_ncu.md):
//   * a ring stage is eight distinct ROWS with however many entries share them: inside a chunk, group k lives in slot k mod 16 and
//     belongs to stage k / 8; the producer warp copies rows groupRow[G0 + 8 s .. + 7], a consumer warp moves on when an entry's stage
//     number changes;
//   * 128 consumers per CTA, thread t owning observations 2t, 2t+1, 256+2t, 257+2t of the tile (two conflict-free LDS.128 per entry),
//     so the per-entry bookkeeping pays for four scores; window, slot offset, basis index and stage of an entry come pre-decoded in one
//     16-byte broadcast load, and the descriptors of the next 256 entries are fetched into registers while the current batch is consumed;
//   * four entries at a time, sixteen independent score chains per thread, go through one filter: a running maximum only grows, so a
//     score below the maximum as it stood before the batch cannot win after any of its entries; only when some lane of the warp has
//     a candidate are the four entries applied one after the other, exactly.
// 4 096 rows x 131 072 observations: 1.72 / 1.98 / 1.82-2.2 / 1.9-2.3e12 pairs/s at 2 / 3 / 4 / 8 bases per row (6.9 TB/s per distinct
// row at 2 = the HBM roof; before: 1.54 / 1.60 / 1.42 / 1.35e12) -- profiles/r02_group_probe.jsonl, r02_sweep_grp_v2_ncu.md.
#define GRP_CONSUMERS 128
#define GRP_THREADS (GRP_CONSUMERS + 32)

#define SD_LEX_UPDATE4_RARE(sc, bb, V, I) do { \
		const bool h0_ = (sc)[0] >= (V)[0], h1_ = (sc)[1] >= (V)[1], h2_ = (sc)[2] >= (V)[2], h3_ = (sc)[3] >= (V)[3]; \
		if (__any_sync(0xffffffffu, h0_ || h1_ || h2_ || h3_)) { \
			if (h0_ && ((sc)[0] > (V)[0] || (bb) < (I)[0])) { (V)[0] = (sc)[0]; (I)[0] = (bb); } \
			if (h1_ && ((sc)[1] > (V)[1] || (bb) < (I)[1])) { (V)[1] = (sc)[1]; (I)[1] = (bb); } \
			if (h2_ && ((sc)[2] > (V)[2] || (bb) < (I)[2])) { (V)[2] = (sc)[2]; (I)[2] = (bb); } \
			if (h3_ && ((sc)[3] > (V)[3] || (bb) < (I)[3])) { (V)[3] = (sc)[3]; (I)[3] = (bb); } \
		} } while (0)

__global__ void __launch_bounds__(GRP_THREADS, GRP_CTAS) k_sweep_tma_grp(SweepGrpArgs a) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	unsigned char *ring = smem_raw;                                                        // [stage][slot][4 KiB]
	uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t) GRP_STAGES * GRP_ROWS * TMA_ROW_BYTES);
	uint64_t *empty = full + GRP_STAGES;
	double2 *s_ac = reinterpret_cast<double2 *>(empty + GRP_STAGES);                        // [GRP_BATCH] (sigma.pib, piCbarX)
	int4 *s_m = reinterpret_cast<int4 *>(s_ac + GRP_BATCH);                                 // [GRP_BATCH] (window, slot byte offset, basis, stage)
	const int tile = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
	const int e0 = chunk * a.chunkSize, e1 = min(a.basisCnt, e0 + a.chunkSize);
	const int nEnt = e1 - e0;
	const double *tileBase = a.delta + (size_t) tile * a.Dcap * SD_TILE_W;
	if (tid == 0) {
		for (int s = 0; s < GRP_STAGES; s++) { sd_mbar_init(&full[s], 1); sd_mbar_init(&empty[s], GRP_CONSUMERS / 32); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	// sorted entries: first and last group of the chunk (an empty chunk still writes its "nothing found" maxima)
	const int G0 = nEnt > 0 ? a.entGroup[e0] : 0, G1 = nEnt > 0 ? a.entGroup[e1 - 1] : -1;
	const int nGroups = G1 - G0 + 1, nStages = (nGroups + GRP_ROWS - 1) / GRP_ROWS;
	sd_pdl_wait();
	sd_pdl_launch_dependents();

	if (tid >= GRP_CONSUMERS) {
		// ---------------- producer warp: lane r copies the row of group G0 + 8 k + r into slot r of stage k ----------------
		const int lane = tid - GRP_CONSUMERS;
		for (int k = 0; k < nStages; k++) {
			const int s = k % GRP_STAGES;
			const int cnt = min(GRP_ROWS, nGroups - k * GRP_ROWS);
			const int row = lane < cnt ? a.groupRow[G0 + k * GRP_ROWS + lane] : 0;
			sd_mbar_wait(&empty[s], ((k / GRP_STAGES) & 1) ^ 1);
			if (lane == 0) sd_mbar_expect_tx(&full[s], (uint32_t) cnt * TMA_ROW_BYTES);
			__syncwarp();
			if (lane < cnt) sd_bulk_g2s(ring + ((size_t) s * GRP_ROWS + lane) * TMA_ROW_BYTES, tileBase + (size_t) row * SD_TILE_W, TMA_ROW_BYTES, &full[s]);
		}
		return;
	}

	// ---------------- consumers --------------------------------------------------------------------------------------
	double oV[4] = {-DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX}, nV[4] = {-DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
	int oI[4] = {-1, -1, -1, -1}, nI[4] = {-1, -1, -1, -1};
	constexpr int PER = GRP_BATCH / GRP_CONSUMERS;                                         // descriptor entries per thread and batch
	int pB[PER], pW[PER], pG[PER]; double pA[PER], pC[PER];
	// entries past the end of the chunk: window 0 (ignored), last group (no stage change)
#define SD_GRP_FETCH(base) do { _Pragma("unroll") for (int u = 0; u < PER; u++) { \
			const int e_ = e0 + (base) + tid + u * GRP_CONSUMERS; const bool ok_ = e_ < e1; \
			pB[u] = ok_ ? a.entBasis[e_] : 0; pG[u] = ok_ ? a.entGroup[e_] - G0 : nGroups - 1; \
			pW[u] = ok_ ? a.descWin[pB[u]] : 0; pA[u] = ok_ ? a.descA[pB[u]] : 0.0; pC[u] = ok_ ? a.descC[pB[u]] : 0.0; } } while (0)
	if (nEnt > 0) SD_GRP_FETCH(0);
	int cur = -1;                                                                          // stage this warp holds
	const unsigned char *mine = ring + tid * 16;
	for (int j0 = 0; j0 < nEnt; j0 += GRP_BATCH) {
		asm volatile("bar.sync 1, %0;" :: "n"(GRP_CONSUMERS) : "memory");                 // everyone is done with the previous batch
#pragma unroll
		for (int u = 0; u < PER; u++) {
			const int q = tid + u * GRP_CONSUMERS;
			s_ac[q] = make_double2(pA[u], pC[u]);
			s_m[q] = make_int4(pW[u], (pG[u] % (GRP_ROWS * GRP_STAGES)) * TMA_ROW_BYTES, pB[u], pG[u] / GRP_ROWS);
		}
		asm volatile("bar.sync 1, %0;" :: "n"(GRP_CONSUMERS) : "memory");
		SD_GRP_FETCH(j0 + GRP_BATCH);                                                       // in flight while this batch is consumed
		const int n = min(GRP_BATCH, nEnt - j0);
		for (int j = 0; j < n; j += 4) {
			int4 m[4];
#pragma unroll
			for (int r = 0; r < 4; r++) m[r] = s_m[j + r];
			double2 dA[4], dB[4];
			if (m[0].w != cur) {                                                            // the batch starts in the next stage: hand the old one back, wait for the new one
				if (cur >= 0) { __syncwarp(); if ((tid & 31) == 0) sd_mbar_arrive(&empty[cur % GRP_STAGES]); }
				cur = m[0].w;
				sd_mbar_wait(&full[cur % GRP_STAGES], (cur / GRP_STAGES) & 1);
			}
			const int wOr = m[0].x | m[1].x | m[2].x | m[3].x, wAnd = m[0].x & m[1].x & m[2].x & m[3].x;
			if (m[3].w == cur && wOr == wAnd && wOr != 0) {
				// ---- the four entries lie in the stage this warp holds (stage numbers never decrease) and belong to the same window
				// (every entry outside the two-window mode; inside it the windows split by basis age, so mixed batches are few): one
				// filter for all sixteen scores.  A running maximum only grows, so a score that does not reach the maximum as it stood
				// BEFORE these entries cannot replace it after any of them either: if no lane has a candidate the batch is done;
				// otherwise the four entries are applied one after the other, exactly.
				double sc[4][4];
#pragma unroll
				for (int r = 0; r < 4; r++) {
					dA[r] = *reinterpret_cast<const double2 *>(mine + m[r].y);
					dB[r] = *reinterpret_cast<const double2 *>(mine + m[r].y + TMA_ROW_BYTES / 2);
				}
#pragma unroll
				for (int r = 0; r < 4; r++) {
					const double2 ac = s_ac[j + r];
					sc[r][0] = __dsub_rn(__dadd_rn(ac.x, dA[r].x), ac.y);                  // stocUpdate.c:174
					sc[r][1] = __dsub_rn(__dadd_rn(ac.x, dA[r].y), ac.y);
					sc[r][2] = __dsub_rn(__dadd_rn(ac.x, dB[r].x), ac.y);
					sc[r][3] = __dsub_rn(__dadd_rn(ac.x, dB[r].y), ac.y);
				}
#define SD_GRP_BATCH(V, I) do { \
					bool hit = false; \
					_Pragma("unroll") for (int r = 0; r < 4; r++) { _Pragma("unroll") for (int o = 0; o < 4; o++) hit |= sc[r][o] >= (V)[o]; } \
					if (__any_sync(0xffffffffu, hit)) { \
						_Pragma("unroll") for (int r = 0; r < 4; r++) { _Pragma("unroll") for (int o = 0; o < 4; o++) SD_LEX_UPDATE(sc[r][o], m[r].z, (V)[o], (I)[o]); } \
					} } while (0)
				if (wOr == 1) SD_GRP_BATCH(oV, oI); else SD_GRP_BATCH(nV, nI);
#undef SD_GRP_BATCH
				continue;
			}
			// ---- a stage boundary inside the batch, or entries of the new window / of no window: entry by entry
#pragma unroll
			for (int r = 0; r < 4; r++) {
				if (m[r].w != cur) {                                                        // next stage: hand the old one back, wait for the new one
					if (cur >= 0) { __syncwarp(); if ((tid & 31) == 0) sd_mbar_arrive(&empty[cur % GRP_STAGES]); }
					cur = m[r].w;
					sd_mbar_wait(&full[cur % GRP_STAGES], (cur / GRP_STAGES) & 1);
				}
				dA[r] = *reinterpret_cast<const double2 *>(mine + m[r].y);
				dB[r] = *reinterpret_cast<const double2 *>(mine + m[r].y + TMA_ROW_BYTES / 2);
				if (m[r].x == 0) continue;
				const double2 ac = s_ac[j + r];
				const int b = m[r].z;
				double sc[4];
				sc[0] = __dsub_rn(__dadd_rn(ac.x, dA[r].x), ac.y);
				sc[1] = __dsub_rn(__dadd_rn(ac.x, dA[r].y), ac.y);
				sc[2] = __dsub_rn(__dadd_rn(ac.x, dB[r].x), ac.y);
				sc[3] = __dsub_rn(__dadd_rn(ac.x, dB[r].y), ac.y);
				if (m[r].x == 1) { SD_LEX_UPDATE4_RARE(sc, b, oV, oI); }
				else             { SD_LEX_UPDATE4_RARE(sc, b, nV, nI); }
			}
		}
	}
#undef SD_GRP_FETCH
	const size_t o = (size_t) tile * SD_TILE_W + 2 * tid;
	const size_t oldAt = ((size_t) 0 * a.nChunks + chunk) * a.NP + o, newAt = ((size_t) 1 * a.nChunks + chunk) * a.NP + o;
	*reinterpret_cast<double2 *>(a.partV + oldAt) = make_double2(oV[0], oV[1]);
	*reinterpret_cast<double2 *>(a.partV + oldAt + SD_TILE_W / 2) = make_double2(oV[2], oV[3]);
	*reinterpret_cast<int2 *>(a.partI + oldAt) = make_int2(oI[0], oI[1]);
	*reinterpret_cast<int2 *>(a.partI + oldAt + SD_TILE_W / 2) = make_int2(oI[2], oI[3]);
	*reinterpret_cast<double2 *>(a.partV + newAt) = make_double2(nV[0], nV[1]);
	*reinterpret_cast<double2 *>(a.partV + newAt + SD_TILE_W / 2) = make_double2(nV[2], nV[3]);
	*reinterpret_cast<int2 *>(a.partI + newAt) = make_int2(nI[0], nI[1]);
	*reinterpret_cast<int2 *>(a.partI + newAt + SD_TILE_W / 2) = make_int2(nI[2], nI[3]);
}

// The same ring for problems with random technology-matrix elements (Q > 0): in the tiled layout the 1+Q planes of one
// dual row are contiguous, so one bulk copy of (1+Q) x 4 KiB brings delta.pib and all of delta.piC for 512 observations.
// Rows per stage (rps: 1, 2, 4 or 8) and ring depth (stages: 2..4) are picked on the host for the most bytes in flight per SM;
// lane r of the producer warp looks after row r of a stage.
__global__ void __launch_bounds__(TMA_THREADS, 3) k_sweep_tma_q(SweepArgs a, int rps, int stages) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int planes = 1 + a.Q;
	const size_t rowDoubles = (size_t) planes * SD_TILE_W;
	const uint32_t rowBytes = (uint32_t) (rowDoubles * 8);
	const size_t stageDoubles = (size_t) rps * rowDoubles;
	double *ring = reinterpret_cast<double *>(smem_raw);
	uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t) stages * stageDoubles * 8);
	uint64_t *empty = full + 4;
	double2 *s_ac = reinterpret_cast<double2 *>(empty + 4);
	int *s_win = reinterpret_cast<int *>(s_ac + SW_BATCH);
	double *s_xq = reinterpret_cast<double *>(s_win + SW_BATCH);                          // [64]
	const int tile = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
	const int b0 = chunk * a.chunkSize, b1 = min(a.basisCnt, b0 + a.chunkSize);
	const int nRows = b1 - b0, nIter = (nRows + rps - 1) / rps;
	const double *tileBase = a.delta + (size_t) tile * a.Dcap * rowDoubles;
	if (tid == 0) {
		for (int s = 0; s < stages; s++) { sd_mbar_init(&full[s], 1); sd_mbar_init(&empty[s], TMA_CONSUMERS / 32); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	sd_pdl_wait();                           // x and the descriptors come from k_cut_prep
	sd_pdl_launch_dependents();
	for (int q = tid; q < a.Q; q += blockDim.x) s_xq[q] = a.x[a.rvCOmCols[q]];
	__syncthreads();

	if (tid >= TMA_CONSUMERS) {
		const int lane = tid - TMA_CONSUMERS;
		int s = 0, ph = 0;
		for (int it = 0; it < nIter; it++) {
			const int r0 = it * rps, nr = min(rps, nRows - r0);
			const int row = lane < nr ? a.descRow[b0 + r0 + lane] : 0;
			sd_mbar_wait(&empty[s], ph ^ 1);
			if (lane == 0) sd_mbar_expect_tx(&full[s], (uint32_t) nr * rowBytes);
			__syncwarp();
			if (lane < nr)
				sd_bulk_g2s(ring + (size_t) s * stageDoubles + (size_t) lane * rowDoubles, tileBase + (size_t) row * rowDoubles, rowBytes, &full[s]);
			if (++s == stages) { s = 0; ph ^= 1; }
		}
		return;
	}

	double oV0 = -DBL_MAX, oV1 = -DBL_MAX, nV0 = -DBL_MAX, nV1 = -DBL_MAX;
	int oI0 = -1, oI1 = -1, nI0 = -1, nI1 = -1;
	int s = 0, ph = 0;
	for (int it = 0; it < nIter; it++) {
		const int r0 = it * rps;
		if (r0 % SW_BATCH == 0) {
			asm volatile("bar.sync 1, %0;" :: "n"(TMA_CONSUMERS) : "memory");
			const int b = b0 + r0 + tid;
			const bool ok = b < b1;
			s_ac[tid] = ok ? make_double2(a.descA[b], a.descC[b]) : make_double2(0.0, 0.0);
			s_win[tid] = ok ? a.descWin[b] : 0;
			asm volatile("bar.sync 1, %0;" :: "n"(TMA_CONSUMERS) : "memory");
		}
		sd_mbar_wait(&full[s], ph);
		for (int r = 0; r < rps; r++) {
			const int j = (r0 + r) % SW_BATCH;
			const int win = s_win[j];
			if (win == 0) continue;
			const double2 *rowp = reinterpret_cast<const double2 *>(ring + (size_t) s * stageDoubles + (size_t) r * rowDoubles) + tid;
			const double2 d = rowp[0];
			const double2 ac = s_ac[j];
			const int b = b0 + r0 + r;
			double s0 = __dsub_rn(__dadd_rn(ac.x, d.x), ac.y);                              // stocUpdate.c:174
			double s1 = __dsub_rn(__dadd_rn(ac.x, d.y), ac.y);
			double dx0 = 0.0, dx1 = 0.0;                                                    // stocUpdate.c:175
			for (int q = 0; q < a.Q; q++) {
				const double2 p = rowp[(size_t) (1 + q) * (SD_TILE_W / 2)];
				dx0 = __dadd_rn(dx0, __dmul_rn(p.x, s_xq[q]));
				dx1 = __dadd_rn(dx1, __dmul_rn(p.y, s_xq[q]));
			}
			s0 = __dsub_rn(__dadd_rn(0.0, s0), dx0);
			s1 = __dsub_rn(__dadd_rn(0.0, s1), dx1);
			if (win == 1) {
				if (s0 > oV0) { oV0 = s0; oI0 = b; }
				if (s1 > oV1) { oV1 = s1; oI1 = b; }
			}
			else {
				if (s0 > nV0) { nV0 = s0; nI0 = b; }
				if (s1 > nV1) { nV1 = s1; nI1 = b; }
			}
		}
		__syncwarp();
		if ((tid & 31) == 0) sd_mbar_arrive(&empty[s]);
		if (++s == stages) { s = 0; ph ^= 1; }
	}
	const size_t o = (size_t) tile * SD_TILE_W + 2 * tid;
	const size_t oldAt = ((size_t) 0 * a.nChunks + chunk) * a.NP + o, newAt = ((size_t) 1 * a.nChunks + chunk) * a.NP + o;
	*reinterpret_cast<double2 *>(a.partV + oldAt) = make_double2(oV0, oV1);
	*reinterpret_cast<int2 *>(a.partI + oldAt) = make_int2(oI0, oI1);
	*reinterpret_cast<double2 *>(a.partV + newAt) = make_double2(nV0, nV1);
	*reinterpret_cast<int2 *>(a.partI + newAt) = make_int2(nI0, nI1);
}

// General sweep: bases with phi columns (random cost, multi-term score) and/or the feasibility mask.
//   arg = sum_t m_t * ((sigma.pib[s_t] + delta.pib[l_t][o]) - piCbarX[s_t]) - m_t * (delta.piC[l_t][o] . x)   stocUpdate.c:164-176
struct SweepGenArgs {
	const double *delta; int64_t Dcap; int Q;
	const int32_t *bTermStart, *tSigma, *tOmega, *descWin;
	const double *sigmaPib, *piCbarX; const int32_t *sigmaLam;
	const double *omega; int64_t NP; int rvOffset2;
	int basisCnt, chunkSize, nChunks;
	const uint32_t *mask; int64_t Bcap;     // bit-packed obsFeasible: [tile][Bcap][SD_MASK_WORDS]
	const double *x; const int32_t *rvCOmCols;
	double *partV; int32_t *partI;
};

__global__ void __launch_bounds__(SD_SWEEP_THREADS) k_sweep_general(SweepGenArgs a) {
	__shared__ double s_xq[SD_MAX_Q];
	const int tile = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
	const int b0 = chunk * a.chunkSize, b1 = min(a.basisCnt, b0 + a.chunkSize);
	sd_pdl_wait();
	sd_pdl_launch_dependents();
	for (int j = tid; j < a.Q; j += blockDim.x) s_xq[j] = a.x[a.rvCOmCols[j]];
	__syncthreads();
	double bestV[2][2] = {{-DBL_MAX, -DBL_MAX}, {-DBL_MAX, -DBL_MAX}};
	int bestI[2][2] = {{-1, -1}, {-1, -1}};
	const size_t rowStride = (size_t) (1 + a.Q) * SD_TILE_W;
	const double *tileBase = a.delta + (size_t) tile * a.Dcap * rowStride;
	for (int b = b0; b < b1; b++) {
		const int win = a.descWin[b];
		if (win == 0) continue;
		const int ts = a.bTermStart[b], te = a.bTermStart[b + 1];
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const int w = 2 * tid + h;
			const size_t o = (size_t) tile * SD_TILE_W + w;
			if (a.mask && !((a.mask[((size_t) tile * a.Bcap + b) * SD_MASK_WORDS + (w >> 5)] >> (w & 31)) & 1u)) continue;
			double arg = 0.0;
			for (int t = ts; t < te; t++) {
				const int s = a.tSigma[t], l = a.sigmaLam[s];
				const double m = (t == ts) ? 1.0 : a.omega[(size_t) (a.rvOffset2 + a.tOmega[t] - 1) * a.NP + o];
				const double *cell = tileBase + (size_t) l * rowStride + w;
				arg = __dadd_rn(arg, __dmul_rn(m, __dsub_rn(__dadd_rn(a.sigmaPib[s], cell[0]), a.piCbarX[s])));
				double dx = 0.0;
				for (int q = 0; q < a.Q; q++) dx = __dadd_rn(dx, __dmul_rn(cell[(size_t) (1 + q) * SD_TILE_W], s_xq[q]));
				arg = __dsub_rn(arg, __dmul_rn(m, dx));
			}
			if (win == 1) { if (arg > bestV[0][h]) { bestV[0][h] = arg; bestI[0][h] = b; } }
			else          { if (arg > bestV[1][h]) { bestV[1][h] = arg; bestI[1][h] = b; } }
		}
	}
	const size_t o = (size_t) tile * SD_TILE_W + 2 * tid;
#pragma unroll
	for (int wdw = 0; wdw < 2; wdw++) {
		const size_t at = ((size_t) wdw * a.nChunks + chunk) * a.NP + o;
		*reinterpret_cast<double2 *>(a.partV + at) = make_double2(bestV[wdw][0], bestV[wdw][1]);
		*reinterpret_cast<int2 *>(a.partI + at) = make_int2(bestI[wdw][0], bestI[wdw][1]);
	}
}

// Term-linear TMA sweep for random-cost problems (rvdOmCnt > 0): multi-term bases (stocUpdate.c:165-176) and the obsFeasible
// mask (:163).  k_cut_prep has flattened the bases of the chunk into a list of terms (sigma.pib, piCbarX, lambda row, multiplier
// column, window, "last term of its basis").  The producer warp feeds a two-stage ring with one bulk copy per term -- the 1+Q
// planes of that term's delta row for this tile -- plus, at a basis' last term, the 512 mask bits (64 bytes) of that basis; lane r of the
// warp looks after term r of the stage, so the descriptor reads and the copies of a stage go out together.  The cost columns of
// omega for this tile (the multipliers m_c) are copied once per CTA and stay resident in shared memory.  Consumers therefore
// touch only shared memory: LDS.128 of the delta pair, LDS.128 of the multiplier pair, LDS.U16 of the mask pair, and the score
// accumulates in term order with the reference's operation order.  Terms of skipped bases (window 0) are not copied at all.
struct SweepTGArgs {
	const double *delta; int64_t Dcap; int Q;
	const double *termA, *termC; const int32_t *termRow, *termMeta, *termBasis, *bTermStart;
	const double *omegaCost; int64_t NP; int nCost;       // omega rows rvOffset[2].. (cost coefficients), nCost of them
	int basisCnt, chunkSize, nChunks;
	const uint32_t *mask; int64_t Bcap;     // bit-packed obsFeasible: [tile][Bcap][SD_MASK_WORDS]
	const double *x; const int32_t *rvCOmCols;
	double *partV; int32_t *partI;
	int rps, stages;                                      // terms per ring stage (1, 2, 4 or 8) and ring depth (2..4)
};

__global__ void __launch_bounds__(TMA_THREADS, 3) k_sweep_tma_gen(SweepTGArgs a) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int planes = 1 + a.Q, rps = a.rps, stages = a.stages;
	const size_t rowDoubles = (size_t) planes * SD_TILE_W;
	const uint32_t rowBytes = (uint32_t) (rowDoubles * 8);
	const size_t slotBytes = (size_t) rowBytes + SD_MASK_WORDS * 4;                        // delta planes, then the 512 mask bits of the basis
	const size_t stageBytes = (size_t) rps * slotBytes;
	unsigned char *cost = smem_raw;                                                      // [nCost][512] doubles
	unsigned char *ring = smem_raw + (size_t) a.nCost * TMA_ROW_BYTES;
	uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t) stages * stageBytes);
	uint64_t *empty = full + 4;
	uint64_t *costBar = empty + 4;
	double2 *s_ac = reinterpret_cast<double2 *>(costBar + 2);
	int *s_meta = reinterpret_cast<int *>(s_ac + SW_BATCH);
	int *s_basis = s_meta + SW_BATCH;
	double *s_xq = reinterpret_cast<double *>(s_basis + SW_BATCH);                        // [64]
	const int tile = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
	const int b0 = chunk * a.chunkSize, b1 = min(a.basisCnt, b0 + a.chunkSize);
	const int T0 = a.bTermStart[b0], T1 = a.bTermStart[b1];
	const int nTerms = T1 - T0, nIter = (nTerms + rps - 1) / rps;
	const double *tileBase = a.delta + (size_t) tile * a.Dcap * rowDoubles;
	if (tid == 0) {
		for (int s = 0; s < stages; s++) { sd_mbar_init(&full[s], 1); sd_mbar_init(&empty[s], TMA_CONSUMERS / 32); }
		sd_mbar_init(costBar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	sd_pdl_wait();
	sd_pdl_launch_dependents();
	for (int q = tid; q < a.Q; q += blockDim.x) s_xq[q] = a.x[a.rvCOmCols[q]];
	__syncthreads();

	if (tid >= TMA_CONSUMERS) {
		// ---------------- producer warp ---------------------------------------------------------------------------------
		const int lane = tid - TMA_CONSUMERS;
		if (lane == 0) {
			sd_mbar_expect_tx(costBar, (uint32_t) a.nCost * TMA_ROW_BYTES);
			for (int r = 0; r < a.nCost; r++)
				sd_bulk_g2s(cost + (size_t) r * TMA_ROW_BYTES, a.omegaCost + (size_t) r * a.NP + (size_t) tile * SD_TILE_W, TMA_ROW_BYTES, costBar);
		}
		int s = 0, ph = 0;
		for (int it = 0; it < nIter; it++) {
			const int t = T0 + it * rps + lane;
			int meta = 0, row = 0, basis = 0;
			if (lane < rps && t < T1) { meta = a.termMeta[t]; row = a.termRow[t]; basis = a.termBasis[t]; }
			const bool live = (meta & 3) != 0;
			const bool wantMask = live && (meta & 4) && a.mask != nullptr;
			uint32_t bytes = live ? rowBytes + (wantMask ? SD_MASK_WORDS * 4 : 0) : 0;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
			sd_mbar_wait(&empty[s], ph ^ 1);
			if (lane == 0) sd_mbar_expect_tx(&full[s], bytes);
			__syncwarp();
			if (live) {
				unsigned char *slot = ring + (size_t) s * stageBytes + (size_t) lane * slotBytes;
				sd_bulk_g2s(slot, tileBase + (size_t) row * rowDoubles, rowBytes, &full[s]);
				if (wantMask) sd_bulk_g2s(slot + rowBytes, a.mask + ((size_t) tile * a.Bcap + basis) * SD_MASK_WORDS, SD_MASK_WORDS * 4, &full[s]);
			}
			if (++s == stages) { s = 0; ph ^= 1; }
		}
		return;
	}

	// ---------------- consumers ------------------------------------------------------------------------------------------
	double oV0 = -DBL_MAX, oV1 = -DBL_MAX, nV0 = -DBL_MAX, nV1 = -DBL_MAX;
	int oI0 = -1, oI1 = -1, nI0 = -1, nI1 = -1;
	double arg0 = 0.0, arg1 = 0.0;                                                         // stocUpdate.c:164
	sd_mbar_wait(costBar, 0);
	int s = 0, ph = 0;
	for (int it = 0; it < nIter; it++) {
		const int r0 = it * rps;
		if (r0 % SW_BATCH == 0) {
			asm volatile("bar.sync 1, %0;" :: "n"(TMA_CONSUMERS) : "memory");
			const int t = T0 + r0 + tid;
			const bool ok = t < T1;
			s_ac[tid] = ok ? make_double2(a.termA[t], a.termC[t]) : make_double2(0.0, 0.0);
			s_meta[tid] = ok ? a.termMeta[t] : 0;
			s_basis[tid] = ok ? a.termBasis[t] : 0;
			asm volatile("bar.sync 1, %0;" :: "n"(TMA_CONSUMERS) : "memory");
		}
		sd_mbar_wait(&full[s], ph);
		for (int r = 0; r < rps; r++) {
			const int j = (r0 + r) % SW_BATCH;
			const int meta = s_meta[j];
			if ((meta & 3) == 0) continue;                                                 // infeasible basis / outside both windows (or past the end)
			const unsigned char *slot = ring + (size_t) s * stageBytes + (size_t) r * slotBytes;
			const double2 *rowp = reinterpret_cast<const double2 *>(slot) + tid;
			const double2 d = rowp[0];
			const double2 ac = s_ac[j];
			double m0 = 1.0, m1 = 1.0;                                                     // stocUpdate.c:167-170
			const int om = meta >> 8;
			if (om > 0) {
				const double2 mm = reinterpret_cast<const double2 *>(cost + (size_t) (om - 1) * TMA_ROW_BYTES)[tid];
				m0 = mm.x; m1 = mm.y;
			}
			arg0 = __dadd_rn(arg0, __dmul_rn(m0, __dsub_rn(__dadd_rn(ac.x, d.x), ac.y)));    // stocUpdate.c:174
			arg1 = __dadd_rn(arg1, __dmul_rn(m1, __dsub_rn(__dadd_rn(ac.x, d.y), ac.y)));
			double dx0 = 0.0, dx1 = 0.0;                                                   // stocUpdate.c:175
			for (int q = 0; q < a.Q; q++) {
				const double2 p = rowp[(size_t) (1 + q) * (SD_TILE_W / 2)];
				dx0 = __dadd_rn(dx0, __dmul_rn(p.x, s_xq[q]));
				dx1 = __dadd_rn(dx1, __dmul_rn(p.y, s_xq[q]));
			}
			arg0 = __dsub_rn(arg0, __dmul_rn(m0, dx0));
			arg1 = __dsub_rn(arg1, __dmul_rn(m1, dx1));
			if (meta & 4) {                                                                // the basis is complete: stocUpdate.c:163,178-181
				bool f0 = true, f1 = true;
				if (a.mask) {
					const unsigned mw = reinterpret_cast<const unsigned *>(slot + rowBytes)[tid >> 4] >> ((2 * tid) & 31);
					f0 = (mw & 1u) != 0; f1 = (mw & 2u) != 0;
				}
				const int b = s_basis[j];
				if ((meta & 3) == 1) {
					if (f0 && arg0 > oV0) { oV0 = arg0; oI0 = b; }
					if (f1 && arg1 > oV1) { oV1 = arg1; oI1 = b; }
				}
				else {
					if (f0 && arg0 > nV0) { nV0 = arg0; nI0 = b; }
					if (f1 && arg1 > nV1) { nV1 = arg1; nI1 = b; }
				}
				arg0 = 0.0; arg1 = 0.0;
			}
		}
		__syncwarp();
		if ((tid & 31) == 0) sd_mbar_arrive(&empty[s]);
		if (++s == stages) { s = 0; ph ^= 1; }
	}
	const size_t o = (size_t) tile * SD_TILE_W + 2 * tid;
	const size_t oldAt = ((size_t) 0 * a.nChunks + chunk) * a.NP + o, newAt = ((size_t) 1 * a.nChunks + chunk) * a.NP + o;
	*reinterpret_cast<double2 *>(a.partV + oldAt) = make_double2(oV0, oV1);
	*reinterpret_cast<int2 *>(a.partI + oldAt) = make_int2(oI0, oI1);
	*reinterpret_cast<double2 *>(a.partV + newAt) = make_double2(nV0, nV1);
	*reinterpret_cast<int2 *>(a.partI + newAt) = make_int2(nI0, nI1);
}

// ======================================================================================================
// merge + accumulate (one CTA per observation tile, one thread per observation)
// ======================================================================================================
struct MergeArgs {
	const double *partV; const int32_t *partI; int nChunks; int64_t NP;
	int omegaCnt, pi_eval; double lb;
	const int32_t *omegaW; const double *omega; int rvOffset2;
	const double *delta; int64_t Dcap; int Q;
	const double *sigmaPib, *sigmaPiCr; const int32_t *sigmaLam; int sigmaCnt, n1c, n1cP;
	const int32_t *bTermStart, *tSigma, *tOmega;
	int randCost;                 // num->rvdOmCnt > 0: the cuts.c:142-159 branch
	int32_t *iStar; int32_t *iStarHost; int iStarHostCap; double *tilePart; int P;
	int mW;                       // observations per merge CTA: 64, 128, 256 or 512
	int lex;                      // chunk maxima are merged lexicographically (grouped sweep: chunks are not ascending basis ranges)
	// epilogue run by the last block: tile partials -> un-normalised cut [-> normalised cut in mapped host memory]
	int n1; const int32_t *CCols, *qCols; double *partial; int fuseNormalise, numSamples; double *hostRes; SdDevState *st;
	// NVLink peer exchange (peerRanks > 1): every rank's buffer, this rank's index, the sequence number of this cut
	int peerRanks, peerRank; unsigned peerSeq; unsigned char *peerBufs[16];
};

// exchange buffer layout: slots[parity][rank][n1+4] doubles, then flags[parity][rank] uint32
__device__ __forceinline__ double *sd_peer_slot(unsigned char *buf, int G, int n1, int parity, int rank) {
	return reinterpret_cast<double *>(buf) + ((size_t) parity * G + rank) * (n1 + 4);
}
__device__ __forceinline__ unsigned *sd_peer_flag(unsigned char *buf, int G, int n1, int parity, int rank) {
	return reinterpret_cast<unsigned *>(reinterpret_cast<double *>(buf) + (size_t) 2 * G * (n1 + 4)) + parity * G + rank;
}

#define MG_THREADS SD_TILE_W

#ifdef SD_PHASE_CLOCKS
// debug build only (tools/merge_phases.py): global-timer stamps of the merge kernel's phases; phases 0-4 by block 0, 5.. by the last block
__device__ long long g_sd_phase[16];
__device__ __forceinline__ long long sd_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define SD_PHASE(i) do { if ((((i) >= 5 && (i) < 10) || blockIdx.x == 0) && threadIdx.x == 0) g_sd_phase[i] = sd_gtime(); } while (0)
extern "C" int sdgpu_debug_phase_clocks(long long *out) { return cudaMemcpyFromSymbol(out, g_sd_phase, sizeof(long long) * 16) == cudaSuccess ? 0 : -2; }
#else
#define SD_PHASE(i)
#endif

// running (max, first index) over the per-chunk partial maxima [c0, c1) of one observation, chunks in ascending basis order; the
// loads of 8 chunks are issued together (the compare chain is sequential, the memory latency must not be)
__device__ __forceinline__ void sd_merge_chunks(const double *__restrict__ pv, const int32_t *__restrict__ pi, int c0, int c1, int64_t NP,
		bool lex, double &bestV, int &bestI) {
	for (int c = c0; c < c1; c += 8) {
		double v[8]; int ix[8];
#pragma unroll
		for (int u = 0; u < 8; u++) {                        // past the end: re-read the last chunk, it cannot beat itself under strict '>'
			const int cc = min(c + u, c1 - 1);
			v[u] = __ldcg(pv + (size_t) cc * NP); ix[u] = __ldcg(pi + (size_t) cc * NP);
		}
		// lex: the chunks are not ascending basis ranges (grouped sweep), so equal maxima are settled by the lower basis index
#pragma unroll
		for (int u = 0; u < 8; u++) if (v[u] > bestV || (lex && v[u] == bestV && ix[u] >= 0 && ix[u] < bestI)) { bestV = v[u]; bestI = ix[u]; }
	}
}

// both windows at once (pi_eval): the new window's maxima sit `winOff` elements after the old window's; twelve loads in flight
// (six chunks x two windows: one trip for the six chunks a lane owns at real-problem sizes, and few enough registers for two CTAs per SM)
__device__ __forceinline__ void sd_merge_chunks2(const double *__restrict__ pv, const int32_t *__restrict__ pi, size_t winOff, int c0, int c1, int64_t NP,
		bool lex, double &oV, int &oI, double &nV, int &nI) {
	constexpr int CB = 6;
	for (int c = c0; c < c1; c += CB) {
		double v[CB], w[CB]; int ix[CB], iw[CB];
#pragma unroll
		for (int u = 0; u < CB; u++) {
			const size_t at = (size_t) min(c + u, c1 - 1) * NP;
			v[u] = __ldcg(pv + at); ix[u] = __ldcg(pi + at);
			w[u] = __ldcg(pv + winOff + at); iw[u] = __ldcg(pi + winOff + at);
		}
#pragma unroll
		for (int u = 0; u < CB; u++) {
			if (v[u] > oV || (lex && v[u] == oV && ix[u] >= 0 && ix[u] < oI)) { oV = v[u]; oI = ix[u]; }
			if (w[u] > nV || (lex && w[u] == nV && iw[u] >= 0 && iw[u] < nI)) { nV = w[u]; nI = iw[u]; }
		}
	}
}

// Tail of a cut, run by one block with the un-normalised vector [alpha, beta[1..n1], cummOld, cummAll, missing] in shared memory:
// optional all-reduce over NVLink peer memory, then cuts.c:184-188 and the hand-over to the host.
//   peer exchange: push this rank's sums into every rank's slot, raise the flags, wait for everyone's flag, add the slots in
//   rank order (bit-identical on every rank).  Two slot sets alternate by cut parity: a rank can be at most one cut ahead.
__device__ __forceinline__ void sd_cut_exchange_and_store(double *s_cut, int peerRanks, int peerRank, unsigned peerSeq, unsigned char *const *peerBufs,
		int n1, double *partial, int fuseNormalise, int numSamples, double *hostRes) {
	const int tid = threadIdx.x;
	if (peerRanks > 1) {
		const int G = peerRanks, par = peerSeq & 1;
		__shared__ int s_timeout;
		if (tid == 0) s_timeout = 0;
		for (int p = 0; p < G; p++) {
			double *slot = sd_peer_slot(peerBufs[p], G, n1, par, peerRank);
			for (int c = tid; c <= n1 + 3; c += blockDim.x) slot[c] = s_cut[c];
		}
		__threadfence_system();
		__syncthreads();
		if (tid < G) {
			*((volatile unsigned *) sd_peer_flag(peerBufs[tid], G, n1, par, peerRank)) = peerSeq;
			volatile unsigned *mine = sd_peer_flag(peerBufs[peerRank], G, n1, par, tid);
			const long long t0 = clock64();
			while (*mine != peerSeq) {
				if (clock64() - t0 > 20000000000LL) { s_timeout = 1; break; }          // ~10 s: a peer never arrived
				__nanosleep(100);
			}
		}
		__threadfence_system();
		__syncthreads();
		for (int c = tid; c <= n1 + 3; c += blockDim.x) {
			double acc = 0.0;
			for (int r = 0; r < G; r++) acc = __dadd_rn(acc, __ldcv(sd_peer_slot(peerBufs[peerRank], G, n1, par, r) + c));
			s_cut[c] = acc;
		}
		__syncthreads();
		if (tid == 0 && s_timeout) s_cut[n1 + 3] = 2.0e9;                              // reported by sdgpu_sd_cut_finish
		__syncthreads();
	}
	for (int c = tid; c <= n1 + 3; c += blockDim.x) {
		partial[c] = s_cut[c];
		if (fuseNormalise) hostRes[c] = (c <= n1) ? s_cut[c] / numSamples : s_cut[c];
	}
	// no system fence here: the host reads hostRes only after it has waited for this kernel, and kernel completion makes its writes
	// to mapped host memory visible (a fence would only make the last block wait for the PCIe round trip: ~3.5 us per cut)
}

// a rank with nothing to sweep (no local observation yet) still has to take part in the exchange
__global__ void k_cut_exchange(MergeArgs a) {
	extern __shared__ double s_cut[];
	for (int c = threadIdx.x; c <= a.n1 + 3; c += blockDim.x) s_cut[c] = a.partial[c];
	__syncthreads();
	sd_cut_exchange_and_store(s_cut, a.peerRanks, a.peerRank, a.peerSeq, a.peerBufs, a.n1, a.partial, a.fuseNormalise, a.numSamples, a.hostRes);
}

// One CTA of 512 threads looks after W = a.mW observations (64, 128, 256 or 512 -- a whole delta tile only when there are many
// tiles) with L = 512 / W threads per observation: thread (lane, observation) merges its contiguous share of the basis chunks,
// all loads in flight at once, and lane 0 combines the L shares in lane order (strict '>': the lowest chunk wins ties).  With
// few observations (real problems: N <= 5 000) this spreads the merge over ~80 SMs instead of 10 and turns eight dependent
// memory round trips per window into one.
__global__ void __launch_bounds__(MG_THREADS, 2) k_cut_merge(MergeArgs a) {      // two CTAs per SM: 256 CTAs (131 072 observations) stay one wave
	__shared__ int s_istar[SD_TILE_W];
	__shared__ int s_w[SD_TILE_W];
	__shared__ double s_red[4 * 32];
	__shared__ double s_mv[2][MG_THREADS];
	__shared__ int s_mi[2][MG_THREADS];
	extern __shared__ double s_dyn[];        // [groups][n1c] partial sums of the sigma.piC part
	const int W = a.mW, L = MG_THREADS / W;
	const int sub = blockIdx.x, tid = threadIdx.x;
	sd_pdl_wait();                           // the per-chunk maxima come from the sweep
	const int ol = tid & (W - 1), lane = tid / W;
	const int64_t o = (int64_t) sub * W + ol;
	const bool valid = o < a.omegaCnt;
	const size_t rowStride = (size_t) (1 + a.Q) * SD_TILE_W;
	const double *tileBase = a.delta + (size_t) (o / SD_TILE_W) * a.Dcap * rowStride + (o % SD_TILE_W);

	double oldV = -DBL_MAX, newV = -DBL_MAX;
	int oldI = -1, newI = -1, istar = -1, wgt = 0;
	double tAlpha = 0.0, tOld = 0.0, tAll = 0.0, tMiss = 0.0;
	SD_PHASE(0);
	if (valid) {
		const int cpl = (a.nChunks + L - 1) / L, c0 = lane * cpl, c1 = min(a.nChunks, c0 + cpl);
		if (!a.pi_eval) sd_merge_chunks(a.partV + o, a.partI + o, c0, c1, a.NP, a.lex != 0, oldV, oldI);
		else sd_merge_chunks2(a.partV + o, a.partI + o, (size_t) a.nChunks * a.NP, c0, c1, a.NP, a.lex != 0, oldV, oldI, newV, newI);
	}
	SD_PHASE(10);
	if (L > 1) {
		s_mv[0][tid] = oldV; s_mi[0][tid] = oldI; s_mv[1][tid] = newV; s_mi[1][tid] = newI;
		__syncthreads();
		if (lane == 0)
			for (int l = 1; l < L; l++) {                         // ascending lane = ascending basis index
				const double ov = s_mv[0][l * W + ol], nv = s_mv[1][l * W + ol];
				const int oi = s_mi[0][l * W + ol], ni = s_mi[1][l * W + ol];
				if (ov > oldV || (a.lex && ov == oldV && oi >= 0 && oi < oldI)) { oldV = ov; oldI = oi; }
				if (nv > newV || (a.lex && nv == newV && ni >= 0 && ni < newI)) { newV = nv; newI = ni; }
			}
	}
	const bool owner = valid && lane == 0;
	if (owner) {
		wgt = a.omegaW[o];
		if (a.pi_eval) {
			double argmax = fmax(oldV, newV);                     // cuts.c:124
			istar = (newV > oldV) ? newI : oldI;                  // cuts.c:125 (an empty window carries index -1)
			tOld = fmax(oldV - a.lb, 0.0) * wgt;                  // cuts.c:127
			tAll = fmax(argmax - a.lb, 0.0) * wgt;                // cuts.c:128
		}
		else
			istar = oldI;                                         // cuts.c:132
		SD_PHASE(11);
		a.iStar[o] = istar;
		if (o < a.iStarHostCap) a.iStarHost[o] = istar;           // small cuts: iStar lands in mapped host memory, no D2H copy
		if (istar < 0) tMiss = 1.0;                               // cuts.c:136-139
		else if (!a.randCost) {                                   // cuts.c:160-162: the BASIS index doubles as the sigma index
			if (istar >= a.sigmaCnt) { tMiss = 1.0e9; istar = -1; }
			else {
				int l = a.sigmaLam[istar];
				tAlpha = __dadd_rn(__dmul_rn(a.sigmaPib[istar], (double) wgt), __dmul_rn(tileBase[(size_t) l * rowStride], (double) wgt));
			}
		}
		else {                                                    // cuts.c:143-152
			for (int t = a.bTermStart[istar]; t < a.bTermStart[istar + 1]; t++) {
				int s = a.tSigma[t], l = a.sigmaLam[s];
				double m = (t == a.bTermStart[istar]) ? 1.0 : a.omega[(size_t) (a.rvOffset2 + a.tOmega[t] - 1) * a.NP + o];
				tAlpha = __dadd_rn(tAlpha, __dmul_rn(__dmul_rn((double) wgt, m), __dadd_rn(a.sigmaPib[s], tileBase[(size_t) l * rowStride])));
			}
		}
	}
	SD_PHASE(1);
	if (lane == 0) { s_istar[ol] = valid ? istar : -1; s_w[ol] = wgt; }
	double *out = a.tilePart + (size_t) sub * a.P;
	{
		double v4[4] = {tAlpha, tOld, tAll, tMiss};
		sd_block_sum4(v4, s_red);            // the four sums through one pair of barriers; same tree per value as sd_block_sum
		if (tid == 0) { out[0] = v4[0]; out[1] = v4[1]; out[2] = v4[2]; out[3] = v4[3]; }
	}
	SD_PHASE(2);

	// delta.piC part of beta: one block sum per random T element, four elements per pass   cuts.c:156-157 / :166-167
	for (int q0 = 0; q0 < a.Q; q0 += 4) {
		double v4[4] = {0.0, 0.0, 0.0, 0.0};
		if (owner && istar >= 0) {
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const int q = q0 + u;
				if (q >= a.Q) break;
				double v = 0.0;
				if (!a.randCost)
					v = __dmul_rn(tileBase[(size_t) a.sigmaLam[istar] * rowStride + (size_t) (1 + q) * SD_TILE_W], (double) wgt);
				else
					for (int t = a.bTermStart[istar]; t < a.bTermStart[istar + 1]; t++) {
						int s = a.tSigma[t], l = a.sigmaLam[s];
						double m = (t == a.bTermStart[istar]) ? 1.0 : a.omega[(size_t) (a.rvOffset2 + a.tOmega[t] - 1) * a.NP + o];
						v = __dadd_rn(v, __dmul_rn(__dmul_rn((double) wgt, m), tileBase[(size_t) l * rowStride + (size_t) (1 + q) * SD_TILE_W]));
					}
				v4[u] = v;
			}
		}
		sd_block_sum4(v4, s_red);
		if (tid == 0) for (int u = 0; u < 4 && q0 + u < a.Q; u++) out[4 + a.n1c + q0 + u] = v4[u];
	}

	// sigma.piC part of beta: thread (group g, column k) walks the observations of its group in order   cuts.c:154-155 / :164-165
	__syncthreads();
	SD_PHASE(3);
	if (a.n1c > 0) {
		const int kp = min(((a.n1c + 31) / 32) * 32, MG_THREADS);
		const int groups = MG_THREADS / kp;
		const int g = tid / kp;
		const int per = (W + groups - 1) / groups;
		const int w0 = g * per, w1 = min(W, w0 + per);
		for (int k = tid % kp; g < groups && k < a.n1c; k += kp) {
			double acc = 0.0;
			int w = w0;
			if (!a.randCost) {
				for (; w + 8 <= w1; w += 8) {                    // gathers of a batch issued together, adds still in observation order
					double v[8];                                 // (a predicated or clamped last batch instead of the scalar tail below was
#pragma unroll                                               // measured slower at every size: 2.6 against 1.8 us at 5 000, 20 against 15 us at 1M)
					for (int u = 0; u < 8; u++) { const int is = s_istar[w + u]; v[u] = is >= 0 ? a.sigmaPiCr[(size_t) is * a.n1cP + k] : 0.0; }
#pragma unroll
					for (int u = 0; u < 8; u++) if (s_istar[w + u] >= 0) acc = __dadd_rn(acc, __dmul_rn(v[u], (double) s_w[w + u]));
				}
			}
			for (; w < w1; w++) {
				const int is = s_istar[w];
				if (is < 0) continue;
				if (!a.randCost)
					acc = __dadd_rn(acc, __dmul_rn(a.sigmaPiCr[(size_t) is * a.n1cP + k], (double) s_w[w]));
				else {
					const int64_t ow = (int64_t) sub * W + w;
					for (int t = a.bTermStart[is]; t < a.bTermStart[is + 1]; t++) {
						int s = a.tSigma[t];
						double m = (t == a.bTermStart[is]) ? 1.0 : a.omega[(size_t) (a.rvOffset2 + a.tOmega[t] - 1) * a.NP + ow];
						acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double) s_w[w], m), a.sigmaPiCr[(size_t) s * a.n1cP + k]));
					}
				}
			}
			s_dyn[g * a.n1c + k] = acc;
		}
		__syncthreads();
		for (int k = tid; k < a.n1c; k += blockDim.x) {
			double acc = 0.0;
			for (int gg = 0; gg < groups; gg++) acc = __dadd_rn(acc, s_dyn[gg * a.n1c + k]);
			out[4 + k] = acc;
		}
	}

	// ---- epilogue: the last CTA to finish sums the per-CTA partials in CTA order (cuts.c:155-167) and, on a single GPU,
	// applies cuts.c:184-188 and hands the cut to the host through mapped pinned memory
	SD_PHASE(4);
	if (!sd_is_last_block(&a.st->cutTicket)) return;
	SD_PHASE(5);
	// thread (group, column) adds the partials of its contiguous share of the CTAs, loads eight at a time; the group sums are
	// combined in group order: fixed order, a fifth of the dependent round trips (P = 93 columns -> 5 groups)
	const int nT = gridDim.x;
	const int kpP = min(((a.P + 31) / 32) * 32, MG_THREADS), groupsP = MG_THREADS / kpP;
	double *s_grp = s_dyn;                                   // [groupsP][P], then the totals [P], then the cut vector [n1+4]
	double *s_tot = s_dyn + (size_t) groupsP * a.P;
	double *s_cut = s_tot + a.P;
	{
		const int g = tid / kpP, perT = (nT + groupsP - 1) / groupsP;
		const int t0 = g * perT, t1 = min(nT, t0 + perT);
		for (int p = tid % kpP; g < groupsP && p < a.P; p += kpP) {
			double acc = 0.0;
			int t = t0;
			for (; t < t1; t += 16) {                            // sixteen loads in flight; past the end the last partial is read again, not added
				double v[16];
#pragma unroll
				for (int u = 0; u < 16; u++) v[u] = __ldcg(a.tilePart + (size_t) min(t + u, t1 - 1) * a.P + p);
#pragma unroll
				for (int u = 0; u < 16; u++) if (t + u < t1) acc = __dadd_rn(acc, v[u]);
			}
			s_grp[g * a.P + p] = acc;
		}
	}
	__syncthreads();
	for (int p = tid; p < a.P; p += blockDim.x) {
		double acc = 0.0;
		for (int g = 0; g < groupsP; g++) acc = __dadd_rn(acc, s_grp[g * a.P + p]);
		s_tot[p] = acc;
	}
	for (int c = tid; c <= a.n1 + 3; c += blockDim.x) s_cut[c] = 0.0;
	__syncthreads();
	SD_PHASE(6);
	// CCols are distinct columns, so the sigma.piC scatter is conflict free and runs one thread per column (a serial loop here
	// costs one uncached index load per column: 45 us at n1c = 89); the few delta.piC columns may coincide with them and
	// with each other, so they follow in q order, exactly the order of cuts.c:155-157 / :165-167 for every position
	for (int k = tid; k < a.n1c; k += blockDim.x) s_cut[a.CCols[k]] = __dadd_rn(0.0, s_tot[4 + k]);
	__syncthreads();
	if (tid == 0) {
		s_cut[0] = s_tot[0];
		for (int q = 0; q < a.Q; q++) s_cut[a.qCols[q]] = __dadd_rn(s_cut[a.qCols[q]], s_tot[4 + a.n1c + q]);
		s_cut[a.n1 + 1] = s_tot[1]; s_cut[a.n1 + 2] = s_tot[2]; s_cut[a.n1 + 3] = s_tot[3];
	}
	__syncthreads();
	SD_PHASE(7);
	sd_cut_exchange_and_store(s_cut, a.peerRanks, a.peerRank, a.peerSeq, a.peerBufs, a.n1, a.partial, a.fuseNormalise, a.numSamples, a.hostRes);
	SD_PHASE(8);
}

// cuts.c:184-188
__global__ void k_cut_normalise(const double *__restrict__ partial, int n1, int numSamples, double *__restrict__ out) {
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c <= n1 + 3) out[c] = (c <= n1) ? partial[c] / numSamples : partial[c];
	__threadfence_system();
}

// ======================================================================================================
// single-observation computeIstar (debug / STOCH_CHECK path)   stocUpdate.c:142-190
// ======================================================================================================
__global__ void k_istar_one(SweepGenArgs a, int obs, int isNew, double *outV, int32_t *outI) {
	__shared__ double s_v[256];
	__shared__ int s_i[256];
	const int tile = obs / SD_TILE_W, w = obs % SD_TILE_W;
	const size_t rowStride = (size_t) (1 + a.Q) * SD_TILE_W;
	const double *tileBase = a.delta + (size_t) tile * a.Dcap * rowStride;
	double best = -DBL_MAX; int bi = -1;
	for (int b = threadIdx.x; b < a.basisCnt; b += blockDim.x) {
		const int win = a.descWin[b];
		if (win == 0) continue;
		if ((isNew && win != 2) || (!isNew && win != 1)) continue;
		if (a.mask && !((a.mask[((size_t) tile * a.Bcap + b) * SD_MASK_WORDS + (w >> 5)] >> (w & 31)) & 1u)) continue;
		double arg = 0.0;
		const int ts = a.bTermStart[b], te = a.bTermStart[b + 1];
		for (int t = ts; t < te; t++) {
			const int s = a.tSigma[t], l = a.sigmaLam[s];
			const double m = (t == ts) ? 1.0 : a.omega[(size_t) (a.rvOffset2 + a.tOmega[t] - 1) * a.NP + obs];
			const double *cell = tileBase + (size_t) l * rowStride + w;
			arg = __dadd_rn(arg, __dmul_rn(m, __dsub_rn(__dadd_rn(a.sigmaPib[s], cell[0]), a.piCbarX[s])));
			double dx = 0.0;
			for (int q = 0; q < a.Q; q++) dx = __dadd_rn(dx, __dmul_rn(cell[(size_t) (1 + q) * SD_TILE_W], a.x[a.rvCOmCols[q]]));
			arg = __dsub_rn(arg, __dmul_rn(m, dx));
		}
		if (arg > best) { best = arg; bi = b; }          // ascending b within a thread: first maximiser kept
	}
	s_v[threadIdx.x] = best; s_i[threadIdx.x] = bi;
	__syncthreads();
	for (int st = blockDim.x / 2; st > 0; st >>= 1) {
		if (threadIdx.x < st) {
			double v2 = s_v[threadIdx.x + st]; int i2 = s_i[threadIdx.x + st];
			double v1 = s_v[threadIdx.x]; int i1 = s_i[threadIdx.x];
			bool take = (i2 >= 0) && (i1 < 0 || v2 > v1 || (v2 == v1 && i2 < i1));      // greater value, else lower index
			if (take) { s_v[threadIdx.x] = v2; s_i[threadIdx.x] = i2; }
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) { *outV = s_v[0]; *outI = (s_v[0] == -DBL_MAX) ? -1 : s_i[0]; }
}

// ======================================================================================================
// cut heights / aging and reformCuts
// ======================================================================================================
__global__ void k_cut_heights(int n, const double *__restrict__ alpha, const double *__restrict__ beta, const int32_t *__restrict__ numSamples,
		const double *__restrict__ alphaIncumb, int currIter, const double *__restrict__ xk, int n1, double lb,
		double *__restrict__ height, double *__restrict__ eta, double *__restrict__ rhs, int32_t *__restrict__ best) {
	extern __shared__ double s_h[];
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		const double *b = beta + (size_t) i * (n1 + 1);
		double t_over_k = ((double) numSamples[i] / (double) currIter);                 // cuts.c:215
		double dot = 0.0;
		for (int c = 1; c <= n1; c++) dot = __dadd_rn(dot, __dmul_rn(b[c], xk[c]));      // vXv cuts.c:218
		double h = __dsub_rn(alpha[i], dot);
		h = __dmul_rn(h, t_over_k);                                                      // :221
		h = __dadd_rn(h, __dmul_rn(__dsub_rn(1.0, t_over_k), lb));                       // :224
		height[i] = h; s_h[i] = h;
		double r = (double) currIter / (double) numSamples[i];
		eta[i] = r;                                                                      // master.c:152
		rhs[i] = __dadd_rn(alphaIncumb ? alphaIncumb[i] : 0.0, __dmul_rn(__dsub_rn(r, 1.0), lb));   // master.c:174
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		double Sm = -1.0e20; int bi = -1;                                                // -INF, cuts.c:198
		for (int i = 0; i < n; i++) if (Sm < s_h[i]) { Sm = s_h[i]; bi = i; }            // :203-205
		*best = bi;
	}
	__threadfence_system();
}

struct ReformArgs {
	const int32_t *iStar; int istarStride; const int32_t *omegaCnt; int nCuts;      // per cut: iStar row and its length
	const int32_t *observ; int k;                                                    // per replication: k resampled observation indices
	const double *omega; int64_t NP; int rvOffset2;
	const double *delta; int64_t Dcap; int Q;
	const double *sigmaPib, *sigmaPiCr; const int32_t *sigmaLam; int n1c, n1cP;
	const int32_t *bTermStart, *tSigma, *tOmega;
	int n1, lbType, lb; const int32_t *CCols, *qCols;
	double *out;        // [rep][cut][n1+2]: alpha, beta[0..n1]
	int basisCnt; int *errFlag;   // bounds checks on host-supplied observ[] / iStar[] (1: negative observation index, 2: iStar outside the basis list)
};

#define RF_CHUNK 4096

// reformCuts optimal.c:187-236: one CTA per (cut, bootstrap replication).  The k resampled observations are staged through
// shared memory RF_CHUNK at a time (observation index and the basis its iStar names), then thread (group, column) adds its
// share of the samples, gathers issued eight at a time; group sums are combined in group order (deterministic).
__global__ void __launch_bounds__(512) k_reform(ReformArgs a) {
	__shared__ double s_red[32];
	__shared__ int s_o[RF_CHUNK], s_b[RF_CHUNK];
	extern __shared__ double s_part[];      // [groups][nc] partial sums, then [0] alpha, [1] count, [2..2+nc), beta[0..n1]
	const int cut = blockIdx.x, rep = blockIdx.y, tid = threadIdx.x;
	const int32_t *iStar = a.iStar + (size_t) cut * a.istarStride;
	const int32_t *observ = a.observ + (size_t) rep * a.k;
	const int oc = a.omegaCnt[cut];
	const size_t rowStride = (size_t) (1 + a.Q) * SD_TILE_W;
	const int nc = a.n1c + a.Q;
	const int kp = ((max(nc, 1) + 31) / 32) * 32;
	const int groups = max(1, 512 / kp);
	const int g = tid / kp, col = tid % kp;
	double tAlpha = 0.0, tCount = 0.0, acc = 0.0;
	for (int n0 = 0; n0 < a.k; n0 += RF_CHUNK) {
		const int cn = min(RF_CHUNK, a.k - n0);
		__syncthreads();
		for (int i = tid; i < cn; i += blockDim.x) {
			const int o = observ[n0 + i];
			int is = -1;
			if (o < 0) atomicExch(a.errFlag, 1);                            // (the reference would read out of bounds here)
			else if (o < oc) {                                              // optimal.c:205
				is = iStar[o];
				if (is < 0 || is >= a.basisCnt) { atomicExch(a.errFlag, 2); is = -1; s_o[i] = o; s_b[i] = is; continue; }
				const double *cellBase = a.delta + (size_t) (o / SD_TILE_W) * a.Dcap * rowStride + (o % SD_TILE_W);
				for (int t = a.bTermStart[is]; t < a.bTermStart[is + 1]; t++) {
					const int sg = a.tSigma[t], l = a.sigmaLam[sg];
					const double m = (t == a.bTermStart[is]) ? 1.0 : a.omega[(size_t) (a.rvOffset2 + a.tOmega[t] - 1) * a.NP + o];
					tAlpha = __dadd_rn(tAlpha, __dmul_rn(m, __dadd_rn(a.sigmaPib[sg], cellBase[(size_t) l * rowStride])));   // :216
				}
				tCount += 1.0;
			}
			s_o[i] = o; s_b[i] = is;
		}
		__syncthreads();
		if (g < groups && col < nc) {
			const int per = (cn + groups - 1) / groups;
			const int i0 = g * per, i1 = min(cn, i0 + per);
			int i = i0;
			for (; i + 8 <= i1; i += 8) {
				double v[8]; bool single = true;
#pragma unroll
				for (int u = 0; u < 8; u++) {
					const int is = s_b[i + u];
					v[u] = 0.0;
					if (is >= 0) {
						const int t0 = a.bTermStart[is];
						if (a.bTermStart[is + 1] - t0 != 1) { single = false; continue; }
						const int sg = a.tSigma[t0];
						const int o = s_o[i + u];
						v[u] = col < a.n1c ? a.sigmaPiCr[(size_t) sg * a.n1cP + col]
						                   : a.delta[(size_t) (o / SD_TILE_W) * a.Dcap * rowStride + (o % SD_TILE_W) + (size_t) a.sigmaLam[sg] * rowStride + (size_t) (1 + col - a.n1c) * SD_TILE_W];
					}
				}
				if (single) {
#pragma unroll
					for (int u = 0; u < 8; u++) if (s_b[i + u] >= 0) acc = __dadd_rn(acc, v[u]);                         // multiplier 1.0
				}
				else {
					for (int u = 0; u < 8; u++) {
						const int is = s_b[i + u], o = s_o[i + u];
						if (is < 0) continue;
						for (int t = a.bTermStart[is]; t < a.bTermStart[is + 1]; t++) {
							const int sg = a.tSigma[t], l = a.sigmaLam[sg];
							const double m = (t == a.bTermStart[is]) ? 1.0 : a.omega[(size_t) (a.rvOffset2 + a.tOmega[t] - 1) * a.NP + o];
							const double val = col < a.n1c ? a.sigmaPiCr[(size_t) sg * a.n1cP + col]
							    : a.delta[(size_t) (o / SD_TILE_W) * a.Dcap * rowStride + (o % SD_TILE_W) + (size_t) l * rowStride + (size_t) (1 + col - a.n1c) * SD_TILE_W];
							acc = __dadd_rn(acc, __dmul_rn(m, val));                                                    // :218-221
						}
					}
				}
			}
			for (; i < i1; i++) {
				const int is = s_b[i], o = s_o[i];
				if (is < 0) continue;
				for (int t = a.bTermStart[is]; t < a.bTermStart[is + 1]; t++) {
					const int sg = a.tSigma[t], l = a.sigmaLam[sg];
					const double m = (t == a.bTermStart[is]) ? 1.0 : a.omega[(size_t) (a.rvOffset2 + a.tOmega[t] - 1) * a.NP + o];
					const double val = col < a.n1c ? a.sigmaPiCr[(size_t) sg * a.n1cP + col]
					    : a.delta[(size_t) (o / SD_TILE_W) * a.Dcap * rowStride + (o % SD_TILE_W) + (size_t) l * rowStride + (size_t) (1 + col - a.n1c) * SD_TILE_W];
					acc = __dadd_rn(acc, __dmul_rn(m, val));
				}
			}
		}
	}
	__syncthreads();
	if (g < groups && col < nc) s_part[g * nc + col] = acc;
	double *s_fin = s_part + (size_t) groups * max(nc, 1);
	double r = sd_block_sum(tAlpha, s_red); if (tid == 0) s_fin[0] = r;
	r = sd_block_sum(tCount, s_red);        if (tid == 0) s_fin[1] = r;
	__syncthreads();
	if (tid < nc) {
		double t = 0.0;
		for (int gg = 0; gg < groups; gg++) t = __dadd_rn(t, s_part[gg * nc + tid]);
		s_fin[2 + tid] = t;
	}
	__syncthreads();
	// optimal.c:197-199 (zero), :218-221 (scatter: distinct CCols in parallel, then the delta.piC columns in order), :228-235
	double *beta = s_fin + 2 + nc;
	double *out = a.out + ((size_t) rep * a.nCuts + cut) * (a.n1 + 2);
	for (int i = tid; i <= a.n1; i += blockDim.x) beta[i] = 0.0;
	__syncthreads();
	for (int k = tid; k < a.n1c; k += blockDim.x) beta[a.CCols[k]] = __dadd_rn(0.0, s_fin[2 + k]);
	__syncthreads();
	if (tid == 0) {
		for (int q = 0; q < a.Q; q++) beta[a.qCols[q]] = __dadd_rn(beta[a.qCols[q]], s_fin[2 + a.n1c + q]);
		double al = s_fin[0] / (double) a.k;
		if (a.lbType == 1) al = __dadd_rn(al, __dmul_rn(__dsub_rn(1.0, s_fin[1] / (double) a.k), (double) a.lb));
		out[0] = al;
	}
	__syncthreads();
	for (int i = tid; i <= a.n1; i += blockDim.x) out[1 + i] = beta[i] / (double) a.k;
}

// ======================================================================================================
// host side
// ======================================================================================================
static inline int sd_blocks(int64_t n, int t) { return (int) std::max<int64_t>(1, (n + t - 1) / t); }

// piCbarX + basis descriptors (+ x onto the device) in one launch.  x travels as a kernel parameter when it fits
// (n1 < 256, every reference problem), otherwise through one H2D copy.
static int sd_launch_prep(sdgpu_ctx *c, const double *X, int cutoff, int split, bool wantAllPiCbarX, bool wantTerms = false) {
	SdXParam xp;
	const double *xDevIn = nullptr;
	if (c->n1 + 1 <= 256) memcpy(xp.v, X, ((size_t) c->n1 + 1) * sizeof(double));
	else {
		memcpy(c->h_pinD, X, ((size_t) c->n1 + 1) * sizeof(double));
		SD_CUDA(cudaMemcpyAsync(c->d_x, c->h_pinD, ((size_t) c->n1 + 1) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		xDevIn = c->d_x;
	}
	const int64_t n = std::max<int64_t>(std::max<int64_t>(c->basisCnt, wantAllPiCbarX ? c->sigmaCnt : 0), 1);
	// the deeper load batch costs 146 registers (three CTAs of 128 threads per SM): only while every CTA of the grid is resident at once
	// (real-problem sizes: 5 000 bases = 40 CTAs); the 65 536 bases of the bench stay at 32 loads in flight and five CTAs per SM
	const bool deep = sd_blocks(n, 128) <= (int64_t) sd_sm_count(c) * 3;
#define SD_PREP_GO(B) do { \
	if (sd_smem_optin(c, k_cut_prep<B>, (B) == 64 ? SD_SMEM_PREP : SD_SMEM_PREP32, 0, (size_t) std::max(1, c->n1c) * 8, "k_cut_prep")) return SDGPU_ERR; \
	k_cut_prep<B><<<sd_blocks(n, 128), 128, (size_t) std::max(1, c->n1c) * 8, c->stream>>>(xp, xDevIn, c->d_x, c->n1, c->d_sigmaPiCk, c->SP, c->n1c, \
			c->d_CCols, (int) c->sigmaCnt, wantAllPiCbarX ? c->d_piCbarX : nullptr, c->d_bCk, c->d_bFeas, c->d_bTermStart, c->d_tSigma, \
			c->d_sigmaPib, c->d_sigmaLam, (int) c->basisCnt, split, cutoff, c->d_descA, c->d_descC, c->d_descRow, c->d_descWin, \
			c->d_tOmega, wantTerms ? c->d_termA : nullptr, c->d_termC, c->d_termRow, c->d_termMeta, c->d_termBasis); } while (0)
	if (deep) SD_PREP_GO(64); else SD_PREP_GO(32);
#undef SD_PREP_GO
	SD_LAUNCH_OK("k_cut_prep");
	sd_count_launch(c);
	return 0;
}

static int sd_window_cutoff(int numSamples, int pi_eval) {
	if (pi_eval) numSamples -= (int) (0.1 * numSamples + 1);          // stocUpdate.c:147-148
	return numSamples;
}

static SweepGenArgs sd_gen_args(sdgpu_ctx *c, int chunkSize, int nChunks) {
	SweepGenArgs g;
	g.delta = c->d_delta; g.Dcap = c->Dcap; g.Q = c->Q;
	g.bTermStart = c->d_bTermStart; g.tSigma = c->d_tSigma; g.tOmega = c->d_tOmega; g.descWin = c->d_descWin;
	g.sigmaPib = c->d_sigmaPib; g.piCbarX = c->d_piCbarX; g.sigmaLam = c->d_sigmaLam;
	g.omega = c->d_omega; g.NP = c->NP; g.rvOffset2 = c->rvOffset[2];
	g.basisCnt = (int) c->basisCnt; g.chunkSize = chunkSize; g.nChunks = nChunks;
	g.mask = c->rvd > 0 ? c->d_mask : nullptr; g.Bcap = c->caps.maxBasis;
	g.x = c->d_x; g.rvCOmCols = c->d_rvCOmCols;
	g.partV = c->d_partV; g.partI = c->d_partI;
	return g;
}

// pick the basis-chunk count.  The sweep kernels keep three CTAs per SM resident (444 slots on 148 SMs).  Measured on B200
// (tools/chunk_probe.py, profiles/r01_chunk_probe.jsonl): what costs is a grid that ends just past a whole number of waves --
// 256 tiles x 7 chunks = 4.04 waves ran at 6.87 TB/s, 5 / 12 / 26 / 52 chunks (2.9 / 6.9 / 15.0 / 30.0 waves) at 7.23-7.32 TB/s,
// and with 10 tiles 44 chunks (one full wave) beat 64 (1.44 waves) -- and more, shorter CTAs shrink the tail.  So: as many chunks
// as the scratch and a minimum chunk length allow, up to ~32 waves, then step down to the nearest count whose last wave is at
// least 60 % full.
static void sd_pick_chunks_for(int smCount, int tiles, int64_t basisCnt, int maxChunks, int forced, int *chunkSize, int *nChunks) {
	const double slots = (double) smCount * 3;
	int64_t cmax = std::min<int64_t>(maxChunks, std::max<int64_t>(1, (basisCnt + 31) / 32));   // at least four load batches per chunk
	cmax = std::min<int64_t>(cmax, std::max<int64_t>(1, (int64_t) ceil(32.0 * slots / tiles)));
	// with many observation tiles and few bases, short chunks cost twice: every CTA pays its start-up for few rows, and the per-chunk
	// partial maxima (24 bytes per chunk and observation, written by the sweep, read by the merge) grow with the chunk count.  Measured at
	// 8 192 x 131 072 (profiles/r02_chunk_probe.jsonl): 55 chunks of 149 rows 1 310 us per cut, 10-19 chunks (430-820 rows) 1 260-1 267 us.
	// So: at least ~600 rows per chunk as long as that still leaves two and a half waves of CTAs.
	{
		const int64_t cRows = std::max<int64_t>(1, basisCnt / 600);
		if ((double) tiles * (double) cRows >= 2.5 * slots) cmax = std::min(cmax, cRows);
	}
	int64_t want = cmax;
	for (int64_t cc = cmax; cc >= std::max<int64_t>(1, cmax / 2); cc--) {
		const double w = tiles * (double) cc / slots;
		if (ceil(w) - w <= 0.4) { want = cc; break; }
	}
	if (forced > 0) want = std::min<int64_t>(std::min<int64_t>(forced, maxChunks), std::max<int64_t>(1, basisCnt));
	int cs = (int) ((basisCnt + want - 1) / want);
	cs = std::max(cs, 1);
	*chunkSize = cs;
	*nChunks = (int) ((basisCnt + cs - 1) / cs);
}

static void sd_pick_chunks(sdgpu_ctx *c, int tiles, int *chunkSize, int *nChunks) {
	const int smCount = sd_sm_count(c);
	sd_pick_chunks_for(smCount, tiles, c->basisCnt, c->maxChunks, c->forceChunks, chunkSize, nChunks);      // forceChunks: SDGPU_CHUNKS, experiment knob
}

// host-only view of the grid choice (no device needed): what the sweep would launch for a table of this shape
extern "C" int sdgpu_plan_sweep_grid(int smCount, int64_t observations, int64_t bases, int maxChunks, int *tiles, int *chunkSize, int *nChunks) {
	if (smCount <= 0 || observations <= 0 || bases <= 0 || maxChunks <= 0 || !tiles || !chunkSize || !nChunks) return sdgpu_fail("plan_sweep_grid: bad argument");
	*tiles = (int) ((observations + SD_TILE_W - 1) / SD_TILE_W);
	sd_pick_chunks_for(smCount, *tiles, bases, maxChunks, 0, chunkSize, nChunks);
	return 0;
}

template <int ROWS, int STAGES, int CTAS>
static int sd_launch_tma_cfg(sdgpu_ctx *c, dim3 grid, const SweepArgs &a, int slot, bool pdl) {
	const size_t smem = (size_t) STAGES * ROWS * TMA_ROW_BYTES + 2 * STAGES * sizeof(uint64_t) + SW_BATCH * (sizeof(double2) + sizeof(int));
	if (!c->tmaAttrSet[slot]) {             // the attribute is per device: remember it per context, not per process
		SD_CUDA(cudaFuncSetAttribute(k_sweep_tma<ROWS, STAGES, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		c->tmaAttrSet[slot] = true;
	}
	SD_CUDA(sd_launch(k_sweep_tma<ROWS, STAGES, CTAS>, grid, dim3(TMA_THREADS), smem, c->stream, pdl, a));
	return 0;
}

static int sd_launch_tma(sdgpu_ctx *c, dim3 grid, const SweepArgs &a, bool pdl) {
	static int cfg = -1;                      // SDGPU_TMA_CFG = experiment knob (tools/tma_check.py); unset = the measured best
	if (cfg < 0) { const char *e = getenv("SDGPU_TMA_CFG"); cfg = e ? atoi(e) : 0; }
	switch (cfg) {
	// measured on B200, 65 536 x 131 072 (profiles/r01_tma_cfg_sweep.jsonl): <8,2,3> 6.87 TB/s, <8,3,2> 6.73, <4,4,3> 6.67,
	// <4,3,4> 6.62, <8,4,1> 5.71; the LDG variant 6.4-6.5
	case 1: return sd_launch_tma_cfg<8, 3, 2>(c, grid, a, 1, pdl);
	case 2: return sd_launch_tma_cfg<4, 4, 3>(c, grid, a, 2, pdl);
	case 3: return sd_launch_tma_cfg<8, 4, 1>(c, grid, a, 3, pdl);
	case 4: return sd_launch_tma_cfg<4, 3, 4>(c, grid, a, 4, pdl);
	default: return sd_launch_tma_cfg<8, 2, 3>(c, grid, a, 0, pdl);
	}
}

// ring of k_sweep_tma_q: <rows per stage, stages> with the most bytes in flight per SM (three CTAs at most: 72 registers x 288
// threads), a lone CTA per SM discounted like in sd_tma_gen_shape.  SDGPU_Q_RPS / SDGPU_Q_STAGES override (experiment knobs).
static size_t sd_tma_q_smem(int planes, int rps, int stages) {
	return (size_t) stages * rps * planes * TMA_ROW_BYTES + 8 * sizeof(uint64_t) + SW_BATCH * (sizeof(double2) + sizeof(int)) + 64 * sizeof(double);
}

static int sd_launch_tma_q(sdgpu_ctx *c, dim3 grid, const SweepArgs &a, bool pdl) {
	const int planes = 1 + c->Q;
	static int envRps = -1, envStages = -1;
	if (envRps < 0) { const char *e = getenv("SDGPU_Q_RPS"); envRps = e ? atoi(e) : 0; e = getenv("SDGPU_Q_STAGES"); envStages = e ? atoi(e) : 0; }
	const size_t smemPerSM = (size_t) 227 << 10;
	int rps = 1, stages = 2;
	double best = 0.0;
	for (int st = 2; st <= 4; st++)
		for (int r = 8; r >= 1; r >>= 1) {
			const size_t sm = sd_tma_q_smem(planes, r, st) + 1024;
			if (sm > smemPerSM) continue;
			const int ctas = (int) std::min<size_t>(3, smemPerSM / sm);
			const double score = (double) ctas * st * r * planes * TMA_ROW_BYTES * (ctas == 1 ? 0.55 : 1.0) + ctas - 0.1 * st;
			if (score > best) { best = score; rps = r; stages = st; }
		}
	if (envRps > 0 && (envRps == 1 || envRps == 2 || envRps == 4 || envRps == 8)) rps = envRps;
	if (envStages >= 2 && envStages <= 4) stages = envStages;
	const size_t smem = sd_tma_q_smem(planes, rps, stages);
	if (smem + 1024 > smemPerSM) return sdgpu_fail("sd_cut: ring of %zu bytes does not fit shared memory", smem);
	if (smem > c->tmaQAttr) { SD_CUDA(cudaFuncSetAttribute(k_sweep_tma_q, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); c->tmaQAttr = smem; }
	SD_CUDA(sd_launch(k_sweep_tma_q, grid, dim3(TMA_THREADS), smem, c->stream, pdl, a, rps, stages));
	return 0;
}

// shared memory of k_sweep_tma_gen for a ring of `stages` x `rps` term slots
static size_t sd_tma_gen_smem(const sdgpu_ctx *c, int rps, int stages) {
	const int nCost = c->numRV - c->rvOffset[2];
	return (size_t) nCost * TMA_ROW_BYTES + (size_t) stages * rps * ((size_t) (1 + c->Q) * TMA_ROW_BYTES + SD_MASK_WORDS * 4) + 10 * sizeof(uint64_t) +
	       SW_BATCH * (sizeof(double2) + 2 * sizeof(int)) + 64 * sizeof(double);
}

// ring shape <terms per stage, stages>: the one that keeps the most bytes in flight per SM, a lone CTA per SM discounted (8 consumer
// warps alone do not hide the FP64 compare chains).  Measured on B200 (profiles/r01_variant_gen.jsonl): Q = 0 with 4 cost columns
// <8,2> x 2 CTAs 6.25 TB/s, <4,2> x 3 6.09, <8,3> x 1 4.25; Q = 2 with 8 cost columns <2,2> x 2 6.89, <4,3> x 1 6.45, <1,2> x 3 6.15.
// SDGPU_GEN_RPS / SDGPU_GEN_STAGES override (experiment knobs).  false = even the smallest ring does not fit.
static bool sd_tma_gen_shape(const sdgpu_ctx *c, int *rps, int *stages) {
	static int envRps = -1, envStages = -1;
	if (envRps < 0) { const char *e = getenv("SDGPU_GEN_RPS"); envRps = e ? atoi(e) : 0; e = getenv("SDGPU_GEN_STAGES"); envStages = e ? atoi(e) : 0; }
	const size_t smemPerSM = (size_t) 227 << 10, slot = (size_t) (1 + c->Q) * TMA_ROW_BYTES + SD_MASK_WORDS * 4;
	if (envRps > 0 || envStages > 0) {
		int r = envRps > 0 ? envRps : 4, st = envStages > 0 ? envStages : 2;
		if ((r == 1 || r == 2 || r == 4 || r == 8) && st >= 2 && st <= 4 && sd_tma_gen_smem(c, r, st) + 1024 <= smemPerSM) { *rps = r; *stages = st; return true; }
	}
	double best = 0.0;
	for (int st = 2; st <= 3; st++)
		for (int r = 8; r >= 1; r >>= 1) {
			const size_t smem = sd_tma_gen_smem(c, r, st) + 1024;              // + the per-CTA reservation
			if (smem > smemPerSM) continue;
			const int ctas = (int) std::min<size_t>(3, smemPerSM / smem);       // 70 registers x 288 threads: three CTAs at most
			const double score = (double) ctas * st * r * slot * (ctas == 1 ? 0.55 : 1.0) + ctas;
			if (score > best) { best = score; *rps = r; *stages = st; }
		}
	return best > 0.0;
}

// ---- bases grouped by lambda row (host bookkeeping of the grouped sweep) ------------------------------------------------------
// sigma -> lambda row on the host: the tail that is missing is read back from the device (one small copy, only when sigmas were added)
static int sd_refresh_host_lam(sdgpu_ctx *c) {
	const size_t have = c->hostLam.size();
	if ((int64_t) have >= c->sigmaCnt) return 0;
	c->hostLam.resize((size_t) c->sigmaCnt);
	SD_CUDA(cudaMemcpyAsync(c->hostLam.data() + have, c->d_sigmaLam + have, ((size_t) c->sigmaCnt - have) * 4, cudaMemcpyDeviceToHost, c->stream));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	return 0;
}

// number of distinct lambda rows among the bases (all single-term here), kept incrementally
static int sd_group_count(sdgpu_ctx *c) {
	if (c->grpCounted == c->basisCnt) return 0;
	if (sd_refresh_host_lam(c)) return SDGPU_ERR;
	if ((int64_t) c->grpRowCount.size() < c->lambdaCnt) c->grpRowCount.resize((size_t) c->lambdaCnt, 0);
	for (int64_t b = c->grpCounted; b < c->basisCnt; b++) {
		const int r = c->hostLam[c->basis[b].sigmaIdx[0]];
		if (c->grpRowCount[r]++ == 0) c->grpDistinct++;
	}
	c->grpCounted = c->basisCnt;
	return 0;
}

// the bases sorted by (row, basis index) on the device: a few new bases are inserted in place, many are re-sorted
static int sd_group_sort(sdgpu_ctx *c) {
	if (c->grpSorted == c->basisCnt) return 0;
	const int64_t B = c->basisCnt;
	int64_t dirtyFrom = B;
	auto rowOf = [&](int64_t b) { return c->hostLam[c->basis[b].sigmaIdx[0]]; };
	if (B - c->grpSorted > 64 || c->grpSorted == 0) {
		std::vector<int32_t> idx((size_t) B);
		for (int64_t b = 0; b < B; b++) idx[b] = (int32_t) b;
		std::stable_sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) { return rowOf(x) < rowOf(y); });     // stable: ascending basis index inside a row
		c->grpBasis = idx;
		c->grpRow.resize((size_t) B);
		for (int64_t i = 0; i < B; i++) c->grpRow[i] = rowOf(idx[i]);
		dirtyFrom = 0;
	}
	else {
		for (int64_t b = c->grpSorted; b < B; b++) {
			const int r = rowOf(b);
			const int64_t pos = std::upper_bound(c->grpRow.begin(), c->grpRow.end(), r) - c->grpRow.begin();      // end of the row's group: b is the largest index so far
			c->grpRow.insert(c->grpRow.begin() + pos, r);
			c->grpBasis.insert(c->grpBasis.begin() + pos, (int32_t) b);
			dirtyFrom = std::min(dirtyFrom, pos);
		}
	}
	c->grpSorted = B;
	if (dirtyFrom < B) {
		// dense group numbers (one group = one distinct lambda row) from the first changed entry on, and the row of each group
		c->grpGroup.resize((size_t) B);
		int32_t g = dirtyFrom > 0 ? c->grpGroup[dirtyFrom - 1] : -1;       // groups 0..g are untouched (g continues if the next entry shares its row)
		const int32_t gFrom = g + 1;
		c->grpGroupRow.resize((size_t) gFrom);
		for (int64_t i = dirtyFrom; i < B; i++) {
			if (i == 0 || c->grpRow[i] != c->grpRow[i - 1]) { g++; c->grpGroupRow.push_back(c->grpRow[i]); }
			c->grpGroup[i] = g;
		}
		SD_CUDA(cudaMemcpyAsync(c->d_entGroup + dirtyFrom, c->grpGroup.data() + dirtyFrom, (size_t) (B - dirtyFrom) * 4, cudaMemcpyHostToDevice, c->stream));
		if ((int64_t) c->grpGroupRow.size() > gFrom)
			SD_CUDA(cudaMemcpyAsync(c->d_groupRow + gFrom, c->grpGroupRow.data() + gFrom, (c->grpGroupRow.size() - (size_t) gFrom) * 4, cudaMemcpyHostToDevice, c->stream));
		SD_CUDA(cudaMemcpyAsync(c->d_entBasis + dirtyFrom, c->grpBasis.data() + dirtyFrom, (size_t) (B - dirtyFrom) * 4, cudaMemcpyHostToDevice, c->stream));
		SD_CUDA(cudaStreamSynchronize(c->stream));          // the vectors are pageable and may change before the next cut
	}
	return 0;
}

// ---- which sweep kernel, on which grid ----------------------------------------------------------------------------------------
// One decision, taken once per cut from the problem's shape, the table sizes and the measured crossovers (profiles/):
//   random cost (mask and / or multi-term bases)   term-linear TMA ring from ~4M (term, observation) pairs (multi-term: 60 us against
//                                                  139 us already at 6 000 x 5 000) resp. ~50M pairs (single-term with a mask: the LDG
//                                                  kernel wins at 5 000 x 5 000, 49 against 57 us); else per-term gathers / masked LDG
//   RHS-only, Rb <= 4 (5 from 128M pairs)           recompute sweep (16 384 x 131 072: Rb = 1 / 3 / 4 / 5 / 6 -> 2.23 / 1.34 / 1.15 / 1.09 /
//                                                  0.84e12 pairs/s against 0.90e12 streaming)
//   RHS-only, >= 128M pairs                         TMA ring; grouped by lambda row when at least 15 % of the row copies go away
//   random T elements, >= 4M elements               TMA ring with 1+Q planes per row
//   everything smaller                             LDG streaming (5 000 x 5 000: 34 against 39 us for the ring)
// sdgpu_set_sweep_variant: 1 forces the load-based kernels, 2 the TMA rings, 3 the recompute sweep, 4 the grouped ring (each where
// the problem's shape allows it).  Tiny RHS-only cuts on the LDG / recompute kernels skip k_cut_prep (fused prologue).
enum SdSweepKind { SD_SW_LDG, SD_SW_TMA, SD_SW_TMA_Q, SD_SW_GENERAL, SD_SW_TMA_GEN, SD_SW_RECOMPUTE, SD_SW_TMA_GRP };

struct SdSweepPlan {
	SdSweepKind kind; int variantId;          // variantId: what sdgpu_stats.last_sweep_variant reports
	int chunkSize, nChunks;
	int genRps, genStages;                    // ring shape of k_sweep_tma_gen
	bool fusedPrep;                           // the sweep CTAs compute their own descriptors: no k_cut_prep launch
	bool prepAllPiCbarX, prepTerms;           // what k_cut_prep has to produce beyond the per-basis descriptors
};

static int sd_plan_sweep(sdgpu_ctx *c, int N, int tiles, SdSweepPlan *p) {
	const int v = c->sweepVariant;
	const bool multiTerm = c->maxPhiLen > 0, hasMask = c->rvd > 0;
	const int64_t pairs = (int64_t) c->basisCnt * N, big = (int64_t) 128 << 20;
	p->genRps = p->genStages = 0; p->fusedPrep = p->prepAllPiCbarX = p->prepTerms = false;
	sd_pick_chunks(c, tiles, &p->chunkSize, &p->nChunks);
	if (hasMask || multiTerm) {
		const bool genFits = hasMask && sd_tma_gen_shape(c, &p->genRps, &p->genStages);
		const int64_t genFrom = multiTerm ? ((int64_t) 4 << 20) : ((int64_t) 48 << 20);
		if (genFits && (v == 2 || (v == 0 && (int64_t) c->termCnt * N * (1 + c->Q) >= genFrom))) { p->kind = SD_SW_TMA_GEN; p->variantId = 4; p->prepTerms = true; }
		else if (multiTerm) { p->kind = SD_SW_GENERAL; p->variantId = 3; p->prepAllPiCbarX = true; }
		else { p->kind = SD_SW_LDG; p->variantId = 1; }
		return 0;
	}
	if (c->Q == 0) {
		const bool rcOk = c->Rb >= 1 && c->Rb <= 8;
		if (rcOk && (v == 3 || (v == 0 && (c->Rb <= SD_RC_AUTO_MAX || (c->Rb == SD_RC_AUTO_MAX + 1 && pairs >= big))))) { p->kind = SD_SW_RECOMPUTE; p->variantId = 5; }
		else if (c->basisCnt < (1 << 26) && (v == 4 || (v == 0 && pairs >= big))) {
			if (sd_group_count(c)) return SDGPU_ERR;
			if (v == 4 || c->grpDistinct * 100 <= c->basisCnt * 85) { p->kind = SD_SW_TMA_GRP; p->variantId = 6; }
			else { p->kind = SD_SW_TMA; p->variantId = 2; }
		}
		else if (v == 2 || (v == 0 && pairs >= big)) { p->kind = SD_SW_TMA; p->variantId = 2; }
		else { p->kind = SD_SW_LDG; p->variantId = 1; }
		// one CTA per SM at most: the fused kernels need ~128 registers (1 000 x 1 000: cut 44 -> 38 us; at 5 000 x 5 000, 440 CTAs, fusing loses 11 us)
		p->fusedPrep = (p->kind == SD_SW_LDG || p->kind == SD_SW_RECOMPUTE) && c->n1 + 1 <= 256 && c->n1c <= 1024 && (int64_t) tiles * p->nChunks <= 148;
		return 0;
	}
	const bool tmaOk = (1 + c->Q) * TMA_ROW_BYTES <= 48 * 1024;                      // one dual row (all planes) must fit a ring stage
	if (tmaOk && (v == 2 || (v == 0 && pairs * (1 + c->Q) >= ((int64_t) 4 << 20)))) { p->kind = SD_SW_TMA_Q; p->variantId = 2; }
	else { p->kind = SD_SW_LDG; p->variantId = 1; }
	return 0;
}

// host-only view of the kernel choice (no device needed): which sweep family the library would run for a problem of this shape
extern "C" int sdgpu_plan_sweep_kind(int rvCOmCnt, int rvdOmCnt, int rvbOmCnt, int maxPhiLength, int costColumns, int n1, int n1c,
		int64_t bases, int64_t terms, int64_t distinctLambdaRows, int64_t observations, int variant, int *fusedPrologue) {
	if (bases <= 0 || observations <= 0 || terms < bases || variant < 0 || variant > 4) return sdgpu_fail("plan_sweep_kind: bad argument");
	sdgpu_ctx tmp;
	tmp.Q = rvCOmCnt; tmp.rvd = rvdOmCnt; tmp.Rb = rvbOmCnt; tmp.maxPhiLen = maxPhiLength; tmp.n1 = n1; tmp.n1c = n1c;
	tmp.numRV = costColumns; tmp.rvOffset[2] = 0;                    // only the difference (the number of cost columns) matters here
	tmp.basisCnt = bases; tmp.termCnt = terms; tmp.sweepVariant = variant; tmp.maxChunks = SD_MAX_CHUNKS;
	tmp.grpCounted = bases; tmp.grpDistinct = distinctLambdaRows;    // the count the library keeps incrementally
	SdSweepPlan plan;
	const int tiles = (int) ((observations + SD_TILE_W - 1) / SD_TILE_W);
	if (sd_plan_sweep(&tmp, (int) observations, tiles, &plan)) return SDGPU_ERR;
	if (fusedPrologue) *fusedPrologue = plan.fusedPrep ? 1 : 0;
	return plan.variantId;
}

static SweepPrepArgs sd_fused_prep_args(sdgpu_ctx *c, const double *Xvect, int numSamples, int pi_eval_flag) {
	SweepPrepArgs pa;
	memcpy(pa.xp.v, Xvect, ((size_t) c->n1 + 1) * sizeof(double));
	pa.piCk = c->d_sigmaPiCk; pa.SP = c->SP; pa.n1c = c->n1c; pa.CCols = c->d_CCols;
	pa.bCk = c->d_bCk; pa.bFeas = c->d_bFeas; pa.bTermStart = c->d_bTermStart; pa.tSigma = c->d_tSigma;
	pa.sigmaPib = c->d_sigmaPib; pa.sigmaLam = c->d_sigmaLam;
	pa.split = pi_eval_flag != 0; pa.cutoff = sd_window_cutoff(numSamples, pi_eval_flag != 0);
	return pa;
}

static int sd_launch_sweep(sdgpu_ctx *c, const SdSweepPlan &p, int tiles, const double *Xvect, int numSamples, int pi_eval_flag) {
	const dim3 grid((unsigned) tiles, (unsigned) p.nChunks);
	const bool pdl = c->pdl && !c->timing;         // (event records between the kernels would break the programmatic edge anyway)
#define SD_SWEEP_GO(kernel, block, smem, ...) SD_CUDA(sd_launch(kernel, grid, dim3(block), smem, c->stream, pdl, __VA_ARGS__))
	const size_t xs = (size_t) std::max(1, c->n1c) * 8;          // dynamic shared memory of the fused-prologue instantiations
	switch (p.kind) {
	case SD_SW_TMA_GEN: {
		SweepTGArgs g;
		g.delta = c->d_delta; g.Dcap = c->Dcap; g.Q = c->Q;
		g.termA = c->d_termA; g.termC = c->d_termC; g.termRow = c->d_termRow; g.termMeta = c->d_termMeta; g.termBasis = c->d_termBasis;
		g.bTermStart = c->d_bTermStart;
		g.omegaCost = c->d_omega + (size_t) c->rvOffset[2] * c->NP; g.NP = c->NP; g.nCost = c->numRV - c->rvOffset[2];
		g.basisCnt = (int) c->basisCnt; g.chunkSize = p.chunkSize; g.nChunks = p.nChunks;
		g.mask = c->d_mask; g.Bcap = c->caps.maxBasis; g.x = c->d_x; g.rvCOmCols = c->d_rvCOmCols;
		g.partV = c->d_partV; g.partI = c->d_partI; g.rps = p.genRps; g.stages = p.genStages;
		const size_t smem = sd_tma_gen_smem(c, p.genRps, p.genStages);
		if (smem > c->tmaGenAttr) { SD_CUDA(cudaFuncSetAttribute(k_sweep_tma_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); c->tmaGenAttr = smem; }
		SD_SWEEP_GO(k_sweep_tma_gen, TMA_THREADS, smem, g);
		return 0;
	}
	case SD_SW_GENERAL:
		SD_SWEEP_GO(k_sweep_general, SD_SWEEP_THREADS, 0, sd_gen_args(c, p.chunkSize, p.nChunks));
		return 0;
	case SD_SW_RECOMPUTE: {
		SweepRcArgs r;
		r.omega = c->d_omega; r.NP = c->NP; r.lambda = c->d_lambda; r.LP = c->LP; r.bLamPos = c->d_bLamPos;
		r.descA = c->d_descA; r.descC = c->d_descC; r.descRow = c->d_descRow; r.descWin = c->d_descWin;
		r.basisCnt = (int) c->basisCnt; r.chunkSize = p.chunkSize; r.nChunks = p.nChunks; r.partV = c->d_partV; r.partI = c->d_partI;
		SweepPrepArgs pa;
		if (p.fusedPrep) pa = sd_fused_prep_args(c, Xvect, numSamples, pi_eval_flag);
#define SD_RC_LAUNCH(RBV) do { if (p.fusedPrep) SD_SWEEP_GO((k_sweep_recompute<RBV, true>), SD_SWEEP_THREADS, xs, r, pa); \
		else SD_SWEEP_GO((k_sweep_recompute<RBV, false>), SD_SWEEP_THREADS, 0, r, 0); } while (0)
		switch (c->Rb) {
		case 1: SD_RC_LAUNCH(1); break;
		case 2: SD_RC_LAUNCH(2); break;
		case 3: SD_RC_LAUNCH(3); break;
		case 4: SD_RC_LAUNCH(4); break;
		case 5: SD_RC_LAUNCH(5); break;
		case 6: SD_RC_LAUNCH(6); break;
		case 7: SD_RC_LAUNCH(7); break;
		default: SD_RC_LAUNCH(8); break;
		}
#undef SD_RC_LAUNCH
		return 0;
	}
	case SD_SW_TMA_GRP: {
		SweepGrpArgs g;
		g.delta = c->d_delta; g.Dcap = c->Dcap; g.descA = c->d_descA; g.descC = c->d_descC; g.descWin = c->d_descWin;
		g.entBasis = c->d_entBasis; g.entGroup = c->d_entGroup; g.groupRow = c->d_groupRow;
		g.basisCnt = (int) c->basisCnt; g.chunkSize = p.chunkSize; g.nChunks = p.nChunks; g.partV = c->d_partV; g.partI = c->d_partI; g.NP = c->NP;
		const size_t smem = (size_t) GRP_STAGES * GRP_ROWS * TMA_ROW_BYTES + 2 * GRP_STAGES * sizeof(uint64_t) + GRP_BATCH * (sizeof(double2) + sizeof(int4));
		if (sd_smem_optin(c, k_sweep_tma_grp, SD_SMEM_GRP, 0, smem, "k_sweep_tma_grp")) return SDGPU_ERR;
		SD_SWEEP_GO(k_sweep_tma_grp, GRP_THREADS, smem, g);
		return 0;
	}
	default: break;
	}
	SweepArgs a;
	a.delta = c->d_delta; a.Dcap = c->Dcap; a.Q = c->Q;
	a.descA = c->d_descA; a.descC = c->d_descC; a.descRow = c->d_descRow; a.descWin = c->d_descWin;
	a.basisCnt = (int) c->basisCnt; a.chunkSize = p.chunkSize; a.nChunks = p.nChunks;
	a.mask = c->d_mask; a.Bcap = c->caps.maxBasis; a.x = c->d_x; a.rvCOmCols = c->d_rvCOmCols;
	a.partV = c->d_partV; a.partI = c->d_partI; a.NP = c->NP;
	a.rev = 0;
	if (p.kind == SD_SW_TMA) return sd_launch_tma(c, grid, a, pdl);
	if (p.kind == SD_SW_TMA_Q) return sd_launch_tma_q(c, grid, a, pdl);
	// the load-based sweep alternates its direction from cut to cut (L2 re-use on tables a little larger than L2, see k_sweep_ldg)
	const bool rev = c->altDir && (c->sweepFlip ^= 1) == 0;
	a.rev = rev ? 1 : 0;
	const bool hasMask = c->rvd > 0;
#define SD_LDG_GO(HQ, HM, FU, smem, second) do { if (rev) SD_SWEEP_GO((k_sweep_ldg<HQ, HM, FU, true>), SD_SWEEP_THREADS, smem, a, second); \
		else SD_SWEEP_GO((k_sweep_ldg<HQ, HM, FU, false>), SD_SWEEP_THREADS, smem, a, second); } while (0)
	if (c->Q > 0 && hasMask) SD_LDG_GO(true, true, false, 0, 0);
	else if (c->Q > 0)       SD_LDG_GO(true, false, false, 0, 0);
	else if (hasMask)        SD_LDG_GO(false, true, false, 0, 0);
	else if (!p.fusedPrep)   SD_LDG_GO(false, false, false, 0, 0);
	else { const SweepPrepArgs pa = sd_fused_prep_args(c, Xvect, numSamples, pi_eval_flag); SD_LDG_GO(false, false, true, xs, pa); }
#undef SD_LDG_GO
#undef SD_SWEEP_GO
	return 0;
}

static int sd_cut_partial_impl(sdgpu_ctx *c, const double *Xvect, int numSamples, int pi_eval_flag, double lb, bool fuseNormalise) {
	if (!c || !Xvect) return sdgpu_fail("null argument");
	if (numSamples == 0) return sdgpu_fail("sd_cut: numSamples is zero");
	if (c->Q > SD_MAX_Q) return sdgpu_fail("sd_cut: rvCOmCnt %d exceeds the %d random T elements this build stages in shared memory", c->Q, SD_MAX_Q);
	SD_CUDA(cudaSetDevice(c->device));
	int64_t launches0 = c->stats.total_launches;
	if (c->timing) SD_CUDA(cudaEventRecord(c->evA, c->stream));
	const int N = (int) c->omegaCnt;
	const int tiles = (int) ((N + SD_TILE_W - 1) / SD_TILE_W);
	const int P = 4 + c->n1c + c->Q;
	c->lastOmegaCnt = N;
	c->cutFused = false;
	if (N > 0 && c->basisCnt > 0) {
		SdSweepPlan plan;
		if (sd_plan_sweep(c, N, tiles, &plan)) return SDGPU_ERR;
		const int nChunks = plan.nChunks;
		if (!plan.fusedPrep && sd_launch_prep(c, Xvect, sd_window_cutoff(numSamples, pi_eval_flag != 0), pi_eval_flag != 0, plan.prepAllPiCbarX, plan.prepTerms)) return SDGPU_ERR;
		if (plan.kind == SD_SW_TMA_GRP && sd_group_sort(c)) return SDGPU_ERR;      // (host-side; before the sweep's start event)
		if (c->timing) SD_CUDA(cudaEventRecord(c->evC, c->stream));
		if (sd_launch_sweep(c, plan, tiles, Xvect, numSamples, pi_eval_flag)) return SDGPU_ERR;
		SD_LAUNCH_OK("sweep kernel");
		c->stats.last_sweep_variant = plan.variantId;
		const bool lexMerge = plan.kind == SD_SW_TMA_GRP;
		const int64_t sweepRows = plan.kind == SD_SW_TMA_GRP ? c->grpDistinct : c->termCnt;   // delta rows the sweep reads: one per term, one per distinct lambda when grouped
		sd_count_launch(c);
		if (c->timing) SD_CUDA(cudaEventRecord(c->evD, c->stream));
		// algorithmic bytes of the sweep: delta stream + per-observation weight and iStar + per-basis descriptors
		c->stats.last_sweep_bytes = (int64_t) 8 * (1 + c->Q) * sweepRows * (int64_t) N + (c->rvd > 0 ? c->basisCnt * (int64_t) N / 8 : 0) + (int64_t) N * 8 + c->basisCnt * 16;

		MergeArgs m;
		m.partV = c->d_partV; m.partI = c->d_partI; m.nChunks = nChunks; m.NP = c->NP; m.lex = lexMerge ? 1 : 0;
		m.omegaCnt = N; m.pi_eval = pi_eval_flag != 0; m.lb = lb;
		m.omegaW = c->d_omegaW; m.omega = c->d_omega; m.rvOffset2 = c->rvOffset[2];
		m.delta = c->d_delta; m.Dcap = c->Dcap; m.Q = c->Q;
		m.sigmaPib = c->d_sigmaPib; m.sigmaPiCr = c->d_sigmaPiCr; m.sigmaLam = c->d_sigmaLam; m.sigmaCnt = (int) c->sigmaCnt;
		m.n1c = c->n1c; m.n1cP = c->n1cP;
		m.bTermStart = c->d_bTermStart; m.tSigma = c->d_tSigma; m.tOmega = c->d_tOmega;
		m.randCost = c->rvd > 0;
		m.iStar = c->d_iStar; m.tilePart = c->d_tilePart; m.P = P;
		m.iStarHost = c->d_iStarHost; m.iStarHostCap = N <= c->iStarHostCap ? N : 0;
		m.n1 = c->n1; m.CCols = c->d_CCols; m.qCols = c->rvd > 0 ? c->d_rvCOmCols : c->d_rvCols;       // cuts.c:157 vs :167
		const bool peer = sd_use_peer(c);
		m.partial = c->d_cutPartial; m.fuseNormalise = (peer || !sd_use_nccl(c)) && fuseNormalise; m.numSamples = numSamples;
		m.peerRanks = peer ? c->peerRanks : 0; m.peerRank = c->peerRank; m.peerSeq = peer ? c->peerSeq + 1 : 0;   // committed once the launch has succeeded
		for (int r = 0; r < 16; r++) m.peerBufs[r] = peer && r < c->peerRanks ? c->d_peerBufs[r] : nullptr;
		m.hostRes = c->d_cutRes; m.st = c->d_state;
		c->cutFused = m.fuseNormalise != 0;
		// observations per merge CTA: as few as keeps the grid ONE wave -- two CTAs of 512 threads x 64 registers are resident per SM (the
		// rule used to allow four per SM: 512 CTAs = 1.73 waves at 131 072 observations, 86 against 67 us per cut)
		const int smCount = sd_sm_count(c);
		int mW = 64;
		while (mW < SD_TILE_W && (N + mW - 1) / mW > 2 * smCount) mW <<= 1;
		m.mW = mW;
		const int kp = std::min(((c->n1c + 31) / 32) * 32, MG_THREADS);
		const int groups = c->n1c > 0 ? MG_THREADS / kp : 1;
		const int groupsP = MG_THREADS / std::min(((P + 31) / 32) * 32, MG_THREADS);
		const size_t dyn = (size_t) std::max(std::max(1, groups * c->n1c), (groupsP + 1) * P + c->n1 + 4) * 8;
		// static: s_istar, s_w (2 x 2 KiB), s_red, s_mv (8 KiB), s_mi (4 KiB)
		if (sd_smem_optin(c, k_cut_merge, SD_SMEM_MERGE, 17 * 1024, dyn, "k_cut_merge")) return SDGPU_ERR;
		SD_CUDA(sd_launch(k_cut_merge, dim3((unsigned) ((N + mW - 1) / mW)), dim3(MG_THREADS), dyn, c->stream, c->pdl && !c->timing, m));
		if (peer) c->peerSeq++;                   // this rank now takes part in exchange number peerSeq
		sd_count_launch(c);
		if (c->timing) SD_CUDA(cudaEventRecord(c->evE, c->stream));
	}
	else {
		// no observation or no basis at all: nothing to sweep; with observations every one is missing its maximiser (cuts.c:136-139)
		if (c->timing) { SD_CUDA(cudaEventRecord(c->evC, c->stream)); SD_CUDA(cudaEventRecord(c->evD, c->stream)); SD_CUDA(cudaEventRecord(c->evE, c->stream)); }
		c->stats.last_sweep_bytes = 0;
		std::vector<double> zero((size_t) c->n1 + 4, 0.0);
		zero[c->n1 + 3] = (double) N;
		SD_CUDA(cudaMemcpyAsync(c->d_cutPartial, zero.data(), zero.size() * 8, cudaMemcpyHostToDevice, c->stream));
		if (N > 0) SD_CUDA(cudaMemsetAsync(c->d_iStar, 0xff, (size_t) N * 4, c->stream));
		SD_CUDA(cudaStreamSynchronize(c->stream));
		if (sd_use_peer(c)) {                        // the other ranks exchange inside their merge kernel: take part, with zero sums
			MergeArgs m;
			memset(&m, 0, sizeof m);
			m.n1 = c->n1; m.partial = c->d_cutPartial; m.fuseNormalise = fuseNormalise; m.numSamples = numSamples; m.hostRes = c->d_cutRes;
			m.peerRanks = c->peerRanks; m.peerRank = c->peerRank; m.peerSeq = ++c->peerSeq;
			for (int r = 0; r < c->peerRanks; r++) m.peerBufs[r] = c->d_peerBufs[r];
			k_cut_exchange<<<1, 128, ((size_t) c->n1 + 4) * 8, c->stream>>>(m);
			sd_count_launch(c);
			c->cutFused = fuseNormalise;
		}
	}
	c->stats.last_cut_launches = c->stats.total_launches - launches0;
	return 0;
}

// A rank that cannot form its part of a cut still has to take part in the peer exchange the other ranks are waiting in: it
// contributes zero sums and an error marker in the `missing` slot, which makes sdgpu_sd_cut_finish fail on every rank.
int sd_peer_poison_cut(sdgpu_ctx *c) {
	if (!sd_use_peer(c)) return 0;
	SD_CUDA(cudaSetDevice(c->device));
	std::vector<double> v((size_t) c->n1 + 4, 0.0);
	v[c->n1 + 3] = 3.0e9;
	SD_CUDA(cudaMemcpyAsync(c->d_cutPartial, v.data(), v.size() * 8, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	MergeArgs m;
	memset(&m, 0, sizeof m);
	m.n1 = c->n1; m.partial = c->d_cutPartial; m.fuseNormalise = 1; m.numSamples = 1; m.hostRes = c->d_cutRes;
	m.peerRanks = c->peerRanks; m.peerRank = c->peerRank; m.peerSeq = ++c->peerSeq;
	for (int r = 0; r < c->peerRanks; r++) m.peerBufs[r] = c->d_peerBufs[r];
	k_cut_exchange<<<1, 128, ((size_t) c->n1 + 4) * 8, c->stream>>>(m);
	sd_count_launch(c);
	c->cutFused = true;
	return 0;
}

extern "C" int sdgpu_sd_cut_partial(sdgpu_ctx *c, const double *Xvect, int numSamples, int pi_eval_flag, double lb) {
	return sd_cut_partial_impl(c, Xvect, numSamples, pi_eval_flag, lb, false);
}

extern "C" int sdgpu_sd_cut_partial_buffer(sdgpu_ctx *c, void **devPtr, int *len) {
	if (!c || !devPtr || !len) return sdgpu_fail("null argument");
	*devPtr = c->d_cutPartial; *len = c->n1 + 4;
	return 0;
}

extern "C" int sdgpu_sd_cut_finish(sdgpu_ctx *c, int numSamples, sdgpu_cut *cut) {
	if (!c || !cut || !cut->beta) return sdgpu_fail("null argument");
	if (numSamples == 0) return sdgpu_fail("sd_cut: numSamples is zero");
	SD_CUDA(cudaSetDevice(c->device));
	int64_t launches0 = c->stats.total_launches;
	if (!c->cutFused) {          // sharded / split form: normalise after the caller's (or NCCL's) all-reduce, straight into mapped host memory
		k_cut_normalise<<<sd_blocks(c->n1 + 4, 128), 128, 0, c->stream>>>(c->d_cutPartial, c->n1, numSamples, c->d_cutRes);
		sd_count_launch(c);
	}
	double *h = c->h_cutRes;
	const bool istarMapped = c->lastOmegaCnt > 0 && c->lastOmegaCnt <= c->iStarHostCap && c->basisCnt > 0;
	if (cut->iStar && c->lastOmegaCnt > 0 && !istarMapped)
		SD_CUDA(cudaMemcpyAsync(cut->iStar, c->d_iStar, (size_t) c->lastOmegaCnt * 4, cudaMemcpyDeviceToHost, c->stream));
	if (c->timing) SD_CUDA(cudaEventRecord(c->evB, c->stream));
	SD_CUDA(cudaStreamSynchronize(c->stream));
	SD_CUDA(cudaGetLastError());
	if (cut->iStar && istarMapped) memcpy(cut->iStar, c->h_iStar, (size_t) c->lastOmegaCnt * 4);
	if (c->timing) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, c->evA, c->evB) == cudaSuccess) c->stats.last_cut_ms = ms;
		if (cudaEventElapsedTime(&ms, c->evC, c->evD) == cudaSuccess) c->stats.last_sweep_ms = ms;
		if (cudaEventElapsedTime(&ms, c->evA, c->evC) == cudaSuccess) c->stats.last_prep_ms = ms;
		if (cudaEventElapsedTime(&ms, c->evD, c->evE) == cudaSuccess) c->stats.last_merge_ms = ms;
		if (cudaEventElapsedTime(&ms, c->evE, c->evB) == cudaSuccess) c->stats.last_collective_ms = ms;
	}
	c->stats.last_cut_launches += c->stats.total_launches - launches0;
	cut->omegaCnt = c->lastOmegaCnt; cut->numSamples = numSamples;
	cut->cummOld = h[c->n1 + 1]; cut->cummAll = h[c->n1 + 2];
	const double missing = h[c->n1 + 3];
	if (missing >= 3.0e9 && c->peerRanks > 1) return sdgpu_fail("sd_cut: a peer rank failed to form its part of this cut");
	if (missing >= 2.0e9 && c->peerRanks > 1) return sdgpu_fail("sd_cut: peer exchange timed out (a rank did not reach this cut)");
	if (missing >= 1.0e9) return sdgpu_fail("sd_cut: iStar used as a sigma index is out of range (cuts.c:161)");
	if (missing > 0.0) { sdgpu_fail("sd_cut: failed to identify maximal Pi for %g observation(s)", missing); return SDGPU_NONE; }
	cut->alpha = h[0];
	for (int k = 1; k <= c->n1; k++) cut->beta[k] = h[k];
	cut->beta[0] = 1.0;                                              // cuts.c:188
	return 0;
}

extern "C" int sdgpu_sd_cut(sdgpu_ctx *c, const double *Xvect, int numSamples, int pi_eval_flag, double lb, sdgpu_cut *cut) {
	if (!cut || !cut->beta) return sdgpu_fail("null argument");
	int rc = sd_cut_partial_impl(c, Xvect, numSamples, pi_eval_flag, lb, true);
	if (rc != 0) return rc;
	if (sd_use_nccl(c) && !c->cutFused) {           // (with the peer exchange selected the merge kernel has already reduced)
		rc = sd_nccl_allreduce(c, c->d_cutPartial, c->n1 + 4);
		if (rc != 0) return rc;
	}
	return sdgpu_sd_cut_finish(c, numSamples, cut);
}

extern "C" int sdgpu_last_istar_device(sdgpu_ctx *c, void **devPtr, int *len) {
	if (!c || !devPtr || !len) return sdgpu_fail("null argument");
	*devPtr = c->d_iStar; *len = c->lastOmegaCnt;
	return 0;
}

extern "C" int sdgpu_set_sweep_variant(sdgpu_ctx *c, int variant) {
	if (!c) return sdgpu_fail("null context");
	if (variant < 0 || variant > 4) return sdgpu_fail("unknown sweep variant %d", variant);
	c->sweepVariant = variant;
	return 0;
}

extern "C" int sdgpu_compute_istar(sdgpu_ctx *c, const double *Xvect, int obs, int numSamples, int pi_eval, int isNew, double *argmax) {
	if (!c || !Xvect) return sdgpu_fail("null argument");
	if (obs < 0 || obs >= c->omegaCnt) return sdgpu_fail("compute_istar: observation %d out of range", obs);
	if (c->Q > SD_MAX_Q) return sdgpu_fail("compute_istar: rvCOmCnt too large");
	SD_CUDA(cudaSetDevice(c->device));
	// always split at the (possibly shrunk) sample count: isNew asks for the bases above it, !isNew for those at or below
	if (sd_launch_prep(c, Xvect, sd_window_cutoff(numSamples, pi_eval != 0), 1, true)) return SDGPU_ERR;
	double *d_v = c->d_cutOut; int32_t *d_i = (int32_t *) (c->d_cutOut + 1);
	if (c->basisCnt > 0) {
		k_istar_one<<<1, 256, 0, c->stream>>>(sd_gen_args(c, 1, 1), obs, isNew != 0, d_v, d_i);
		sd_count_launch(c);
		SD_CUDA(cudaMemcpyAsync(c->h_pinD, c->d_cutOut, 16, cudaMemcpyDeviceToHost, c->stream));
		SD_CUDA(cudaStreamSynchronize(c->stream));
		SD_CUDA(cudaGetLastError());
		int32_t idx;
		memcpy(&idx, c->h_pinD + 1, 4);
		if (argmax) *argmax = c->h_pinD[0];
		return idx;
	}
	SD_CUDA(cudaStreamSynchronize(c->stream));
	if (argmax) *argmax = -DBL_MAX;
	return SDGPU_NONE;
}

extern "C" int sdgpu_dual_stability(double cummOld, double cummAll, int numSamples, int piEvalStart, int scanLen, double *pi_ratio) {
	// cuts.c:171-182 with calcVariance cuts.c:366-396 (host scalar work on <= SCAN_LEN doubles)
	if (!pi_ratio || scanLen <= 0) return SDGPU_ERR;
	double variance;
	pi_ratio[numSamples % scanLen] = cummOld / cummAll;
	if (numSamples - piEvalStart > scanLen) {
		double mean = pi_ratio[0], vari = 0.0, temp;
		for (int count = 1; count < scanLen; count++) {
			temp = mean;
			mean = mean + (pi_ratio[count] - mean) / (double) (count + 1);
			vari = (1 - 1 / (double) count) * vari + (count + 1) * (mean - temp) * (mean - temp);
		}
		variance = vari;
	}
	else
		variance = 1.0;
	double av = variance > 0.0 ? variance : -variance;
	if (av >= .000002 || pi_ratio[numSamples % scanLen] < 0.95) return 0;
	return 1;
}

extern "C" int sdgpu_cut_heights(sdgpu_ctx *c, int n, const double *alpha, const double *beta, const int32_t *numSamples,
		const double *alphaIncumb, int currIter, const double *xk, double lb, double *height, double *etaCoef, double *rhs) {
	if (!c) return sdgpu_fail("null context");
	if (n <= 0) return SDGPU_NONE;
	if (!alpha || !beta || !numSamples || !xk) return sdgpu_fail("null argument");
	if (currIter == 0) return sdgpu_fail("cut_heights: currIter is zero");
	SD_CUDA(cudaSetDevice(c->device));
	// inputs and outputs live in one block of mapped pinned memory: the kernel reads and writes it directly (<= ~120 KiB for the
	// CUT_MULT * n1 + 3 cuts of setup.c:126), so the call is one launch and one wait, no allocation and no copy
	const size_t n1p = (size_t) c->n1 + 1;
	const size_t nd = (size_t) n * (n1p + 5) + n1p;
	if (sd_aux_reserve(c, nd * 8 + ((size_t) n + 2) * 4)) return SDGPU_ERR;
	double *h = reinterpret_cast<double *>(c->h_aux), *d = reinterpret_cast<double *>(c->d_aux);
	const size_t oAlpha = 0, oBeta = n, oAi = oBeta + (size_t) n * n1p, oXk = oAi + n, oH = oXk + n1p, oE = oH + n, oR = oE + n;
	int32_t *hi = reinterpret_cast<int32_t *>(h + nd), *di = reinterpret_cast<int32_t *>(d + nd);
	memcpy(h + oAlpha, alpha, (size_t) n * 8);
	memcpy(h + oBeta, beta, (size_t) n * n1p * 8);
	if (alphaIncumb) memcpy(h + oAi, alphaIncumb, (size_t) n * 8);
	memcpy(h + oXk, xk, n1p * 8);
	memcpy(hi, numSamples, (size_t) n * 4);
	k_cut_heights<<<1, 128, (size_t) n * 8, c->stream>>>(n, d + oAlpha, d + oBeta, di, alphaIncumb ? d + oAi : nullptr, currIter, d + oXk, c->n1, lb,
			d + oH, d + oE, d + oR, di + n);
	sd_count_launch(c);
	SD_CUDA(cudaStreamSynchronize(c->stream));
	SD_CUDA(cudaGetLastError());
	if (height) memcpy(height, h + oH, (size_t) n * 8);
	if (etaCoef) memcpy(etaCoef, h + oE, (size_t) n * 8);
	if (rhs) memcpy(rhs, h + oR, (size_t) n * 8);
	return hi[n];
}

extern "C" int sdgpu_reform_cuts_batch(sdgpu_ctx *c, int nCuts, const int32_t *iStar, int istarStride, const int32_t *omegaCnt,
		int nReps, const int32_t *observ, int k, int lbType, int lb, double *alpha, double *beta) {
	if (!c || !observ || !alpha || !beta || !omegaCnt) return sdgpu_fail("null argument");
	if (k <= 0 || nCuts <= 0 || nReps <= 0) return sdgpu_fail("reform_cuts_batch: k, nCuts and nReps must be positive");
	SD_CUDA(cudaSetDevice(c->device));
	std::vector<int32_t> oc(omegaCnt, omegaCnt + nCuts);
	if (!iStar) { if (nCuts != 1) return sdgpu_fail("reform_cuts_batch: the device-resident iStar serves one cut"); oc[0] = c->lastOmegaCnt; istarStride = 0; }
	for (int i = 0; i < nCuts; i++)
		if (oc[i] < 0 || oc[i] > c->omegaCnt || (iStar && oc[i] > istarStride)) return sdgpu_fail("reform_cuts_batch: omegaCnt[%d] = %d out of range", i, oc[i]);
	const int nOut = c->n1 + 2;
	// one block of growable device scratch: [out doubles][observ ints][omegaCnt ints][iStar ints] -- no allocation per call
	const size_t outBytes = (size_t) nReps * nCuts * nOut * 8, obBytes = (size_t) nReps * k * 4, ocBytes = (((size_t) nCuts * 4) + 7) / 8 * 8;
	const size_t isBytes = iStar ? (size_t) nCuts * std::max(1, istarStride) * 4 : 0;
	const size_t scratchTotal = outBytes + (obBytes + 7) / 8 * 8 + ocBytes + (isBytes + 7) / 8 * 8 + 8;
	if (sd_scratch_reserve(c, scratchTotal)) return SDGPU_ERR;
	double *d_out = reinterpret_cast<double *>(c->d_scratch);
	int32_t *d_ob = reinterpret_cast<int32_t *>(c->d_scratch + outBytes);
	int32_t *d_oc = reinterpret_cast<int32_t *>(c->d_scratch + outBytes + (obBytes + 7) / 8 * 8);
	int32_t *d_is = iStar ? reinterpret_cast<int32_t *>(c->d_scratch + outBytes + (obBytes + 7) / 8 * 8 + ocBytes) : nullptr;
	SD_CUDA(cudaMemcpyAsync(d_ob, observ, (size_t) nReps * k * 4, cudaMemcpyHostToDevice, c->stream));
	SD_CUDA(cudaMemcpyAsync(d_oc, oc.data(), (size_t) nCuts * 4, cudaMemcpyHostToDevice, c->stream));
	if (iStar) SD_CUDA(cudaMemcpyAsync(d_is, iStar, (size_t) nCuts * istarStride * 4, cudaMemcpyHostToDevice, c->stream));
	int *d_err = reinterpret_cast<int *>(c->d_scratch + scratchTotal - 8);      // bounds-check flag of k_reform, at the end of the scratch block
	SD_CUDA(cudaMemsetAsync(d_err, 0, 4, c->stream));
	ReformArgs a;
	a.iStar = iStar ? d_is : c->d_iStar; a.istarStride = istarStride; a.omegaCnt = d_oc; a.nCuts = nCuts; a.observ = d_ob; a.k = k;
	a.omega = c->d_omega; a.NP = c->NP; a.rvOffset2 = c->rvOffset[2];
	a.delta = c->d_delta; a.Dcap = c->Dcap; a.Q = c->Q;
	a.sigmaPib = c->d_sigmaPib; a.sigmaPiCr = c->d_sigmaPiCr; a.sigmaLam = c->d_sigmaLam; a.n1c = c->n1c; a.n1cP = c->n1cP;
	a.bTermStart = c->d_bTermStart; a.tSigma = c->d_tSigma; a.tOmega = c->d_tOmega;
	a.n1 = c->n1; a.lbType = lbType; a.lb = lb; a.CCols = c->d_CCols; a.qCols = c->d_rvCOmCols;      // optimal.c:220 scatters with rvCOmCols
	a.out = d_out; a.basisCnt = (int) c->basisCnt; a.errFlag = d_err;
	const int nc = c->n1c + c->Q, kp = ((std::max(nc, 1) + 31) / 32) * 32, groups = std::max(1, 512 / kp);
	const size_t dyn = ((size_t) groups * std::max(nc, 1) + 2 + nc + c->n1 + 1) * 8;
	if (nc > 512) { return sdgpu_fail("reform_cuts_batch: more than 512 cut columns is not supported"); }
	if (sd_smem_optin(c, k_reform, SD_SMEM_REFORM, (size_t) 2 * RF_CHUNK * 4 + 512, dyn, "k_reform")) return SDGPU_ERR;
	k_reform<<<dim3((unsigned) nCuts, (unsigned) nReps), 512, dyn, c->stream>>>(a);
	SD_LAUNCH_OK("k_reform");
	sd_count_launch(c);
	std::vector<double> h((size_t) nReps * nCuts * nOut);
	int hErr = 0;
	SD_CUDA(cudaMemcpyAsync(h.data(), d_out, h.size() * 8, cudaMemcpyDeviceToHost, c->stream));
	SD_CUDA(cudaMemcpyAsync(&hErr, d_err, 4, cudaMemcpyDeviceToHost, c->stream));
	cudaError_t e = cudaStreamSynchronize(c->stream);
	if (e != cudaSuccess) return sdgpu_fail("reform_cuts_batch: %s", cudaGetErrorString(e));
	if (hErr == 1) return sdgpu_fail("reform_cuts_batch: observ[] holds a negative observation index");
	if (hErr == 2) return sdgpu_fail("reform_cuts_batch: iStar[] names a basis outside 0..%lld", (long long) c->basisCnt - 1);
	for (size_t i = 0; i < (size_t) nReps * nCuts; i++) {
		alpha[i] = h[i * nOut];
		memcpy(beta + i * (c->n1 + 1), h.data() + i * nOut + 1, ((size_t) c->n1 + 1) * 8);
	}
	return 0;
}

extern "C" int sdgpu_reform_cut(sdgpu_ctx *c, const int32_t *iStar, int omegaCnt, const int32_t *observ, int k, int lbType, int lb,
		double *alpha, double *beta) {
	if (!c) return sdgpu_fail("null context");
	int32_t oc = iStar ? omegaCnt : c->lastOmegaCnt;
	return sdgpu_reform_cuts_batch(c, 1, iStar, omegaCnt, &oc, 1, observ, k, lbType, lb, alpha, beta);
}
