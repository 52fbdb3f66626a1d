// nccl_glue.cu -- the one exchange step of the multi-GPU path: an all-reduce (sum) of the n1+4 doubles of the
// un-normalised cut across the GPUs that each hold a shard of the observations.
// NCCL is resolved with dlopen at first use so that libsdgpu.so itself has no link-time NCCL dependency
// (single-GPU users and the symbol-export test on a CPU box never need it).  In a process that already
// imported torch, dlopen("libnccl.so.2") returns the copy torch loaded, so both share one NCCL.
#include <dlfcn.h>
#include <cstring>

#include "sdgpu_internal.cuh"

namespace {
struct NcclUniqueId { char internal[128]; };
typedef int (*fn_getUniqueId)(NcclUniqueId *);
typedef int (*fn_commInitRank)(void **, int, NcclUniqueId, int);
typedef int (*fn_commDestroy)(void *);
typedef int (*fn_allReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_getErrorString)(int);

struct NcclApi {
	void *handle = nullptr;
	fn_getUniqueId getUniqueId = nullptr;
	fn_commInitRank commInitRank = nullptr;
	fn_commDestroy commDestroy = nullptr;
	fn_allReduce allReduce = nullptr;
	fn_getErrorString getErrorString = nullptr;
	bool tried = false;
} g_nccl;

int loadNccl() {
	if (g_nccl.allReduce) return 0;
	if (g_nccl.tried) return sdgpu_fail("NCCL is not available (libnccl.so.2 could not be loaded)");
	g_nccl.tried = true;
	const char *names[] = { "libnccl.so.2", "libnccl.so" };
	for (const char *n : names) {
		g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (g_nccl.handle) break;
	}
	if (!g_nccl.handle) return sdgpu_fail("NCCL is not available: %s", dlerror());
	g_nccl.getUniqueId = (fn_getUniqueId) dlsym(g_nccl.handle, "ncclGetUniqueId");
	g_nccl.commInitRank = (fn_commInitRank) dlsym(g_nccl.handle, "ncclCommInitRank");
	g_nccl.commDestroy = (fn_commDestroy) dlsym(g_nccl.handle, "ncclCommDestroy");
	g_nccl.allReduce = (fn_allReduce) dlsym(g_nccl.handle, "ncclAllReduce");
	g_nccl.getErrorString = (fn_getErrorString) dlsym(g_nccl.handle, "ncclGetErrorString");
	if (!g_nccl.getUniqueId || !g_nccl.commInitRank || !g_nccl.commDestroy || !g_nccl.allReduce) {
		g_nccl.allReduce = nullptr;
		return sdgpu_fail("NCCL library lacks a required symbol");
	}
	return 0;
}

const char *ncclErr(int rc) { return g_nccl.getErrorString ? g_nccl.getErrorString(rc) : "nccl error"; }
}  // namespace

int sd_nccl_allreduce(sdgpu_ctx *c, double *buf, int n) {
	if (!c->ncclComm) return sdgpu_fail("no NCCL communicator attached");
	if (loadNccl()) return SDGPU_ERR;
	int rc = g_nccl.allReduce(buf, buf, (size_t) n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->ncclComm, c->stream);
	if (rc != 0) return sdgpu_fail("ncclAllReduce failed: %s", ncclErr(rc));
	sd_count_launch(c);
	return 0;
}

static void sd_peer_release(sdgpu_ctx *c);

// The two exchanges are independent: attaching / replacing the NCCL communicator leaves an attached peer exchange alone (and
// the other way round); sdgpu_destroy tears both down.
void sd_nccl_release(sdgpu_ctx *c) {
	if (c->ncclComm && c->ownComm && g_nccl.commDestroy) g_nccl.commDestroy(c->ncclComm);
	c->ncclComm = nullptr; c->ownComm = false;
}

void sd_peer_teardown(sdgpu_ctx *c) {
	sd_peer_release(c);
	if (c->d_peerLocal) { cudaFree(c->d_peerLocal); c->d_peerLocal = nullptr; }
}

extern "C" int sdgpu_set_collective(sdgpu_ctx *c, int mode) {
	if (!c) return sdgpu_fail("null context");
	if (mode < 0 || mode > 2) return sdgpu_fail("set_collective: unknown mode %d", mode);
	if (mode == 1 && !c->ncclComm) return sdgpu_fail("set_collective: no NCCL communicator attached");
	if (mode == 2 && c->peerRanks <= 1) return sdgpu_fail("set_collective: no peer exchange attached");
	c->collective = mode;
	return 0;
}

extern "C" int sdgpu_attach_nccl(sdgpu_ctx *c, void *ncclComm) {
	if (!c) return sdgpu_fail("null context");
	if (ncclComm && loadNccl()) return SDGPU_ERR;
	sd_nccl_release(c);
	c->ncclComm = ncclComm; c->ownComm = false;
	return 0;
}

extern "C" int sdgpu_nccl_unique_id(void *id128) {
	if (!id128) return sdgpu_fail("null argument");
	if (loadNccl()) return SDGPU_ERR;
	NcclUniqueId id;
	int rc = g_nccl.getUniqueId(&id);
	if (rc != 0) return sdgpu_fail("ncclGetUniqueId failed: %s", ncclErr(rc));
	memcpy(id128, &id, sizeof id);
	return 0;
}

extern "C" int sdgpu_nccl_init(sdgpu_ctx *c, int nranks, int rank, const void *id128) {
	if (!c || !id128) return sdgpu_fail("null argument");
	if (loadNccl()) return SDGPU_ERR;
	SD_CUDA(cudaSetDevice(c->device));
	NcclUniqueId id;
	memcpy(&id, id128, sizeof id);
	void *comm = nullptr;
	int rc = g_nccl.commInitRank(&comm, nranks, id, rank);
	if (rc != 0) return sdgpu_fail("ncclCommInitRank failed: %s", ncclErr(rc));
	sd_nccl_release(c);
	c->ncclComm = comm; c->ownComm = true;
	return 0;
}

// ---- NVLink peer-memory exchange buffers (CUDA IPC between the one-process-per-GPU ranks) -----------------------------
static size_t sd_peer_bytes(const sdgpu_ctx *c, int nranks) {
	return (size_t) 2 * nranks * (c->n1 + 4) * sizeof(double) + (size_t) 2 * nranks * sizeof(unsigned) + 64;
}

static void sd_peer_release(sdgpu_ctx *c) {
	for (int r = 0; r < c->peerRanks; r++)
		if (c->d_peerBufs[r] && r != c->peerRank && !c->peerLocalGroup) cudaIpcCloseMemHandle(c->d_peerBufs[r]);
	c->peerLocalGroup = false;
	for (int r = 0; r < sdgpu_ctx::kMaxPeers; r++) c->d_peerBufs[r] = nullptr;
	c->peerRanks = 0; c->peerRank = -1;
}

extern "C" int sdgpu_peer_export(sdgpu_ctx *c, int nranks, void *handle64) {
	if (!c || !handle64) return sdgpu_fail("null argument");
	if (nranks < 2 || nranks > sdgpu_ctx::kMaxPeers) return sdgpu_fail("peer_export: nranks %d outside 2..%d", nranks, sdgpu_ctx::kMaxPeers);
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
	SD_CUDA(cudaSetDevice(c->device));
	sd_peer_release(c);
	if (c->d_peerLocal) { cudaFree(c->d_peerLocal); c->d_peerLocal = nullptr; }
	c->peerBytes = sd_peer_bytes(c, nranks);
	SD_CUDA(cudaMalloc((void **) &c->d_peerLocal, c->peerBytes));
	SD_CUDA(cudaMemset(c->d_peerLocal, 0, c->peerBytes));
	cudaIpcMemHandle_t h;
	SD_CUDA(cudaIpcGetMemHandle(&h, c->d_peerLocal));
	memcpy(handle64, &h, sizeof h);
	return 0;
}

extern "C" int sdgpu_peer_attach(sdgpu_ctx *c, int nranks, int rank, const void *handles) {
	if (!c || !handles) return sdgpu_fail("null argument");
	if (!c->d_peerLocal || c->peerBytes != sd_peer_bytes(c, nranks)) return sdgpu_fail("peer_attach: call sdgpu_peer_export(ctx, %d, ...) first", nranks);
	if (rank < 0 || rank >= nranks) return sdgpu_fail("peer_attach: rank %d out of range", rank);
	SD_CUDA(cudaSetDevice(c->device));
	sd_peer_release(c);
	for (int r = 0; r < nranks; r++) {
		if (r == rank) { c->d_peerBufs[r] = c->d_peerLocal; continue; }
		cudaIpcMemHandle_t h;
		memcpy(&h, (const unsigned char *) handles + (size_t) r * 64, sizeof h);
		void *p = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) { sd_peer_release(c); return sdgpu_fail("peer_attach: cannot open rank %d's buffer: %s", r, cudaGetErrorString(e)); }
		c->d_peerBufs[r] = (unsigned char *) p;
	}
	c->peerRanks = nranks; c->peerRank = rank; c->peerSeq = 0;        // fresh, zeroed buffers on every rank: the sequence restarts with them
	return 0;
}

// same-process form: the members of a group see each other's exchange buffers through plain peer pointers
int sd_peer_attach_local(sdgpu_ctx **ctxs, int n) {
	for (int i = 0; i < n; i++) {
		sdgpu_ctx *c = ctxs[i];
		SD_CUDA(cudaSetDevice(c->device));
		sd_peer_release(c);
		if (c->d_peerLocal) { cudaFree(c->d_peerLocal); c->d_peerLocal = nullptr; }
		c->peerBytes = sd_peer_bytes(c, n);
		SD_CUDA(cudaMalloc((void **) &c->d_peerLocal, c->peerBytes));
		SD_CUDA(cudaMemset(c->d_peerLocal, 0, c->peerBytes));
		for (int j = 0; j < n; j++) {
			if (j == i || ctxs[j]->device == c->device) continue;
			int can = 0;
			SD_CUDA(cudaDeviceCanAccessPeer(&can, c->device, ctxs[j]->device));
			if (!can) return sdgpu_fail("devices %d and %d cannot access each other's memory", c->device, ctxs[j]->device);
			cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return sdgpu_fail("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
			cudaGetLastError();
		}
	}
	for (int i = 0; i < n; i++) {
		sdgpu_ctx *c = ctxs[i];
		for (int j = 0; j < n; j++) c->d_peerBufs[j] = ctxs[j]->d_peerLocal;
		c->peerRanks = n; c->peerRank = i; c->peerSeq = 0;
		c->peerLocalGroup = true;
	}
	return 0;
}
