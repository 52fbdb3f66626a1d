// sdgpu_internal.cuh -- context layout and helpers shared by the translation units of libsdgpu.so.
//
// HBM layout.  N = observations, D = lambda rows, S = sigma rows, B = bases.
//   omega    [numRV][NP]            rv-major, NP = observation pitch: thread-per-observation kernels coalesce
//   lambda   [R][LP]                position-major: thread-per-row scans (calcLambda) and K3 coalesce
//   sigmaPib [SP], sigmaLam [SP], sigmaCk [SP]
//   sigmaPiC kept twice: k-major [n1c][SP] (scan + piCbarX, thread per row) and row-major [SP][n1cP]
//            (beta accumulation, thread per column)
//   delta    tiled [nTiles][Dcap][1+Q][W]  W = SD_TILE_W observations: for one observation tile the rows of
//            all duals are back to back, so the sweep of one CTA is a single contiguous HBM stream, a new
//            dual (row append) writes nTiles contiguous W*8-byte segments, and a new observation (column
//            append) writes D strided doubles.
//   mask     tiled [nTiles][Bcap][W/32] 32-bit words, one BIT per (basis, observation), only when rvdOmCnt > 0 (obsFeasible, stoc.h:95)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

#include "sdgpu.h"

#define SD_TILE_W      512          // observations per delta tile (4 KiB of FP64 per dual per tile)
#define SD_SWEEP_THREADS 256        // one double2 (two observations) per thread per dual row
#define SD_MAX_CHUNKS  64           // basis chunks per observation tile in the 2-D sweep grid

extern thread_local std::string g_sdgpu_err;
int sdgpu_fail(const char *fmt, ...);

#define SD_CUDA(call)                                                                              \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess)                                                                     \
			return sdgpu_fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

// Device-resident counters and the result slots of the find-or-append kernels.  Kernels read the counts
// from here, so a chain of table updates needs no host round trip between its kernels.
struct SdDevState {
	int omegaCnt, lambdaCnt, sigmaCnt, basisCnt;
	int foundLambda, lambdaIdx, newLambda;     // calcLambda result
	int foundSigma, sigmaIdx, newSigma;        // calcSigma result
	int foundOmega, omegaIdx, newOmega;        // calcOmega result
	int overflow;                              // a capacity was exceeded
	unsigned int ticket;                       // last-block election of the fused find+commit kernels
	unsigned int cutTicket;                    // ... of the cut merge kernel
	double pibBar;                             // pi x bBar + mubBar of the vector being processed
};

struct SdHostBasis {
	int ck, feas, phiLen, weight;
	std::vector<int32_t> sigmaIdx;   // [phiLen+1]
	std::vector<int32_t> omegaIdx;   // [phiLen+1], slot 0 unused
};

struct SdVm;
struct sdgpu_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool ownStream = true;
	cudaEvent_t evA = nullptr, evB = nullptr, evC = nullptr, evD = nullptr, evE = nullptr;   // cut start / end, sweep start / end, merge end

	sdgpu_num num{};
	sdgpu_caps caps{};
	int n1 = 0, n1c = 0, n1cP = 0, R = 0, Rb = 0, Q = 0, rvd = 0, numRV = 0, rows = 0;
	int32_t rvOffset[3] = {0, 0, 0};
	int64_t NP = 0, LP = 0, SP = 0, BP = 0, nTiles = 0, termCap = 0;

	// static problem description on the device (0-based internally)
	int32_t *d_CCols = nullptr;      // [n1c]  1-based position in X
	int32_t *d_rvRows = nullptr;     // [R]    1-based row in Pi
	int32_t *d_bLamPos = nullptr;    // [Rb]   position (0..R-1) of rvbOmRows[j] inside a lambda, or -1
	int32_t *d_cLamPos = nullptr;    // [Q]    position of rvCOmRows[e] inside a lambda, or -1
	int32_t *d_cListStart = nullptr; // [Q+1]  for output c: the e's with rvCOmCols[e] == rvCOmCols[c], in e order
	int32_t *d_cList = nullptr;
	int32_t *d_rvCOmCols = nullptr;  // [Q]    1-based position in X / beta
	int32_t *d_rvCols = nullptr;     // [Q]    1-based position in beta (plain branch, cuts.c:167)
	int32_t *d_bBarCol = nullptr; double *d_bBarVal = nullptr; int bBarCnt = 0, cbNnz = 0;
	int32_t *d_cbStart = nullptr;    // [n1c+1] per CCols[k]: Cbar entries with that column, in nnz order
	int32_t *d_cbRow = nullptr; double *d_cbVal = nullptr;

	// tables
	double  *d_omega = nullptr;  int32_t *d_omegaW = nullptr;
	double  *d_lambda = nullptr;
	double  *d_sigmaPib = nullptr, *d_sigmaPiCk = nullptr, *d_sigmaPiCr = nullptr;
	int32_t *d_sigmaLam = nullptr, *d_sigmaCk = nullptr;
	double  *d_delta = nullptr;
	uint32_t *d_mask = nullptr;      // bit-packed obsFeasible: [tile][Bcap][SD_MASK_WORDS]
	int32_t *d_bCk = nullptr, *d_bFeas = nullptr, *d_bPhiLen = nullptr, *d_bTermStart = nullptr;
	int32_t *d_tSigma = nullptr, *d_tOmega = nullptr;
	SdDevState *d_state = nullptr;

	// checkBasisFeasibility inputs (allocated on first use, rvdOmCnt > 0 only)
	int      cols = 0;
	int32_t *d_rvdOmCols = nullptr;  // [rvd] 1-based column of each random cost
	char    *d_senx = nullptr;       // [rows]
	double  *d_fPiDet = nullptr;     // [Bcap][rows]
	double  *d_fPhi = nullptr;       // [termCap][rows]   (term t >= 1 of a basis = its phi column t-1)
	double  *d_fGBar = nullptr;      // [Bcap][cols]
	double  *d_fPsi = nullptr;       // [termCap][cols]
	int8_t  *d_fCstat = nullptr;     // [Bcap][cols]
	uint8_t *d_fHas = nullptr;       // [Bcap]
	uint8_t *d_fFlags = nullptr;     // [max(Bcap, NP)] result staging

	// per-call staging
	double  *d_vecIn = nullptr;      // a host vector (Pi / observ / X) of up to max(rows, numRV, n1)+1 doubles
	double  *d_cand = nullptr;       // reduced candidate: lambda [R] / piCBar [n1c] / observation [numRV]
	double  *d_candC = nullptr;
	double  *h_pinD = nullptr;       // pinned doubles (inputs and results)
	int32_t *h_pinI = nullptr;       // pinned ints
	SdDevState *h_state = nullptr;   // pinned + mapped mirror: commit kernels publish the state straight into it
	SdDevState *d_hstate = nullptr;  // device alias of h_state
	double  *d_pinD = nullptr;       // device alias of h_pinD (zero-copy reads of small host vectors)
	double  *h_cutRes = nullptr;     // pinned + mapped [n1+4]: the finished cut, written by the last merge block
	double  *d_cutRes = nullptr;     // device alias of h_cutRes
	size_t   pinDcap = 0, pinIcap = 0;
	unsigned char *h_aux = nullptr, *d_aux = nullptr;   // growable pinned + mapped scratch for the small batched calls
	size_t   auxCap = 0;
	unsigned char *d_scratch = nullptr;                 // growable device scratch of the batched calls (reformCuts, feasibility cuts)
	size_t   scratchCap = 0;
	double  *d_fpAlpha = nullptr, *d_fpBeta = nullptr;  // device-resident feasibility-cut pool (cell->fcutsPool): alpha [fpCap], beta [fpCap][n1+1]
	int64_t  fpCap = 0, fpCnt = 0;

	// cut formation scratch
	double  *d_x = nullptr;          // [n1+1]
	double  *d_piCbarX = nullptr;    // [SP]
	double  *d_descA = nullptr, *d_descC = nullptr;   // per basis: sigma.pib, piCbarX of its first sigma
	int32_t *d_descRow = nullptr, *d_descWin = nullptr; // per basis: lambda row, window (0 skip, 1 old, 2 new)
	// per term of every basis (random-cost problems only, rvdOmCnt > 0): the descriptors the term-linear TMA sweep walks
	double  *d_termA = nullptr, *d_termC = nullptr;     // sigma.pib, piCbarX of the term's sigma
	int32_t *d_termRow = nullptr, *d_termMeta = nullptr, *d_termBasis = nullptr;   // lambda row; window | last-term << 2 | omegaIdx << 8; basis
	size_t   tmaGenAttr = 0;
	SdVm    *vm = nullptr;           // delta table on reserved address space, mapped as it grows (vmem.cu); null: allocated whole
	int64_t  Dcap = 0;               // row stride of the delta table: caps.maxLambda, rounded up to 512 when the table is mapped on demand
	int      smCount = 0;            // multiprocessors of the device (0: not asked yet; sd_sm_count)
	size_t   smemAttr[13] = {};      // dynamic shared memory opted in per kernel on this context's device (index: SdSmemSlot)
	double  *d_partV = nullptr;      // [2][chunks][NP] per-chunk running maxima (old, new)
	int32_t *d_partI = nullptr;      // [2][chunks][NP]
	int32_t *d_iStar = nullptr;      // [NP]
	double  *d_tilePart = nullptr;   // [nTiles][P]  P = 4 + n1c + Q
	double  *d_cutPartial = nullptr; // [n1+4]  alpha, beta[1..n1], cummOld, cummAll, missing (un-normalised)
	double  *d_cutOut = nullptr;     // [n1+4]  alpha, beta[1..n1], cummOld, cummAll, missing (normalised)
	int      maxChunks = 1;
	int      sweepVariant = 0;
	int      lastOmegaCnt = 0;
	bool     timing = false;         // record CUDA events around the cut and its sweep (sdgpu_set_timing)
	int32_t *h_iStar = nullptr;      // pinned + mapped [iStarHostCap]: the merge kernel mirrors iStar here for small N
	int32_t *d_iStarHost = nullptr;  // device alias
	int64_t  iStarHostCap = 0;
	bool     tmaAttrSet[8] = {false, false, false, false, false, false, false, false};   // per context (= per device): opt-in shared memory sizes
	size_t   tmaQAttr = 0;
	bool     cutFused = false;       // the last merge block already normalised the cut into h_cutRes
	bool     pdl = true;             // chain the kernels of a cut with programmatic dependent launch (SDGPU_PDL=0 turns it off)
	bool     fusedUpdate = true;     // {delta column || lambda scan -> sigma} in one launch (SDGPU_FUSED_UPDATE=0: the three-launch chain)
	bool     altDir = true;          // the load-based sweep alternates its row direction from cut to cut (SDGPU_ALTDIR=0 turns it off):
	int      forceChunks = 0;        // SDGPU_CHUNKS at create: forces the basis-chunk count of the sweep grid (experiment knob)
	int      sweepFlip = 0;          //   a table a little larger than the 126 MB L2 then finds its most recently read part still cached

	// host mirrors (bookkeeping only; no table arithmetic happens on the host)
	int64_t omegaCnt = 0, lambdaCnt = 0, sigmaCnt = 0, basisCnt = 0, termCnt = 0;
	int     maxPhiLen = 0;
	bool    anyInfeasibleBasis = false;
	std::vector<SdHostBasis> basis;
	std::vector<std::vector<uint32_t>> hostMask;  // [b][NP / 32] bit-packed mirror of the basis' mask row (empty for infeasible bases), only when rvd > 0
	// bases grouped by lambda row for the grouped sweep (several sigmas / bases on one lambda): host bookkeeping, device copies of the order
	std::vector<int32_t> hostLam;                 // sigma -> lambda row (mirror of d_sigmaLam, refreshed on demand)
	std::vector<int32_t> grpRowCount;             // bases per lambda row (single-term bases only)
	int64_t grpCounted = 0, grpDistinct = 0;      // bases counted so far, distinct rows among them
	std::vector<int32_t> grpBasis, grpRow;        // bases sorted by (row, basis index), and the row of each entry
	int64_t grpSorted = 0;                        // bases present in the sorted arrays
	int32_t *d_entBasis = nullptr;
	std::vector<int32_t> grpGroup, grpGroupRow;   // dense group (= distinct row) number of each sorted entry, and the row of each group
	int32_t *d_entGroup = nullptr, *d_groupRow = nullptr;

	// NVLink peer-memory exchange (sdgpu_peer_export / _attach): slots[2][G][n1+4] doubles then flags[2][G] uint32
	static const int kMaxPeers = 16;
	unsigned char *d_peerLocal = nullptr;          // this rank's exchange buffer (cudaMalloc, exported through CUDA IPC)
	unsigned char *d_peerBufs[kMaxPeers] = {};     // every rank's buffer as seen from this device (own entry = d_peerLocal)
	int      peerRanks = 0, peerRank = -1;
	bool     peerLocalGroup = false;               // peers are contexts of this process (plain peer pointers, nothing to close)
	unsigned peerSeq = 0;                          // monotonic for the life of the exchange buffer (never reset: a stale flag must not match)
	size_t   peerBytes = 0;
	int      collective = 0;                       // sdgpu_set_collective: 0 automatic (peer if attached, else NCCL), 1 NCCL, 2 peer

	// NCCL (resolved with dlopen so that the library loads on a box without NCCL)
	void *ncclComm = nullptr;
	bool  ownComm = false;

	sdgpu_stats stats{};
};

static inline int64_t sd_round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// Kernels whose dynamic shared memory grows with the problem: opt in above 48 KiB (static + dynamic), fail loudly above the 227 KiB
// a CTA can have on sm_100 -- a rejected launch must not pass for a finished call.
enum SdSmemSlot { SD_SMEM_OMEGA, SD_SMEM_LAMBDA, SD_SMEM_DELTA_ROW, SD_SMEM_DELTA_COL, SD_SMEM_MERGE, SD_SMEM_REFORM, SD_SMEM_PREP, SD_SMEM_LDG_FUSED,
                  SD_SMEM_RC_FUSED, SD_SMEM_UPD1, SD_SMEM_UPD2, SD_SMEM_GRP, SD_SMEM_PREP32 };
#define SD_SMEM_LIMIT ((size_t) 227 * 1024)
#define SD_MAX_Q 512             // random T elements (rvCOmCnt): their x entries sit in 4 KiB of shared memory in the load-based sweeps; the rings
                                 // take a dual row with all its planes into one stage and bow out far earlier (1 + Q planes of 4 KiB each)
static inline int sd_sm_count(sdgpu_ctx *c) {
	if (c->smCount <= 0) { c->smCount = 148; cudaDeviceGetAttribute(&c->smCount, cudaDevAttrMultiProcessorCount, c->device); }
	return c->smCount;
}
template <class K>
static inline int sd_smem_optin(sdgpu_ctx *c, K kernel, int slot, size_t staticBytes, size_t dynBytes, const char *what) {
	if (staticBytes + dynBytes > SD_SMEM_LIMIT)
		return sdgpu_fail("%s needs %zu bytes of shared memory per CTA (limit %zu): problem dimension too large for this build", what, staticBytes + dynBytes, SD_SMEM_LIMIT);
	if (staticBytes + dynBytes > 48 * 1024 && dynBytes > c->smemAttr[slot]) {
		cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dynBytes);
		if (e != cudaSuccess) return sdgpu_fail("%s: cudaFuncSetAttribute(%zu bytes) -> %s", what, dynBytes, cudaGetErrorString(e));
		c->smemAttr[slot] = dynBytes;
	}
	return 0;
}
// after a kernel launch: a rejected configuration surfaces here, not at some later synchronisation
#define SD_LAUNCH_OK(what) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return sdgpu_fail("%s launch -> %s", what, cudaGetErrorString(e_)); } while (0)

// delta element (l, plane q, observation o) in the tiled layout
__host__ __device__ static inline size_t sd_delta_off(int64_t Dcap, int Q, int64_t l, int q, int64_t o) {
	int64_t t = o / SD_TILE_W, w = o % SD_TILE_W;
	return (((size_t) t * Dcap + l) * (size_t) (1 + Q) + q) * SD_TILE_W + w;
}
// the mask word holding (basis b, observation o); the bit inside it is o % 32
#define SD_MASK_WORDS (SD_TILE_W / 32)
__host__ __device__ static inline size_t sd_mask_word(int64_t Bcap, int64_t b, int64_t o) {
	int64_t t = o / SD_TILE_W, w = o % SD_TILE_W;
	return ((size_t) t * Bcap + b) * SD_MASK_WORDS + w / 32;
}
// host mirror of the mask (the basis de-duplication of stocUpdate.c:104 reads obsFeasible[b][omegaIdx] on the host)
static inline bool sd_hm_get(const sdgpu_ctx *c, int64_t b, int64_t o) { return (c->hostMask[b][o >> 5] >> (o & 31)) & 1u; }
static inline void sd_hm_set(sdgpu_ctx *c, int64_t b, int64_t o, bool v) {
	uint32_t &w = c->hostMask[b][o >> 5];
	w = v ? (w | (1u << (o & 31))) : (w & ~(1u << (o & 31)));
}

#ifdef __CUDACC__
// Programmatic dependent launch (PDL): the kernels of one cut (prologue -> sweep -> merge) are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so kernel N+1 is scheduled while kernel N still runs and its CTAs block in
// sd_pdl_wait() until kernel N has completed and flushed -- the launch latency and the prologue of N+1 overlap the tail of N.
// Both instructions are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void sd_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void sd_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KA, class... A>
static inline cudaError_t sd_launch(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, A &&...args) {
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
}

// The block that draws the last ticket of a launch (1-D grid) gets `true`; the ticket is reset for the next launch.
__device__ __forceinline__ bool sd_is_last_block(unsigned int *ticket) {
	__shared__ bool s_last;
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned int t = atomicAdd(ticket, 1u);
		s_last = (t == gridDim.x - 1);
		if (s_last) *ticket = 0;
	}
	__syncthreads();
	return s_last;
}
#endif

int sd_aux_reserve(sdgpu_ctx *c, size_t bytes);  // grows h_aux / d_aux (pinned, mapped); contents are not preserved
int sd_scratch_reserve(sdgpu_ctx *c, size_t bytes);  // grows d_scratch (device); contents are not preserved
int sd_vm_create(sdgpu_ctx *c, size_t rowBytes, int64_t Dcap, int64_t nTiles);   // vmem.cu
int sd_delta_ensure(sdgpu_ctx *c, int64_t rows, int64_t obs);                    // physical memory under rows [0, rows) x observations [0, obs)
void sd_vm_destroy(sdgpu_ctx *c);
int sd_sync_state(sdgpu_ctx *c);                 // D2H of SdDevState + stream sync + mirror update
int sd_nccl_allreduce(sdgpu_ctx *c, double *buf, int n);
void sd_nccl_release(sdgpu_ctx *c);              // drops the NCCL communicator only
void sd_peer_teardown(sdgpu_ctx *c);             // closes the peer mappings and frees this rank's exchange buffer
int  sd_peer_poison_cut(sdgpu_ctx *c);           // takes part in the current peer exchange with an error marker (keeps the sequence in step)
// which exchange the next cut uses: the peer exchange (true) or NCCL / none (false)
static inline bool sd_use_peer(const sdgpu_ctx *c) { return c->peerRanks > 1 && (c->collective == 2 || (c->collective == 0)); }
static inline bool sd_use_nccl(const sdgpu_ctx *c) { return c->ncclComm != nullptr && !sd_use_peer(c) && c->collective != 2; }
static inline void sd_count_launch(sdgpu_ctx *c, int n = 1) { c->stats.total_launches += n; }
