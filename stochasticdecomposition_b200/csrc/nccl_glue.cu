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

void sd_nccl_release(sdgpu_ctx *c) {
	if (c->ncclComm && c->ownComm && g_nccl.commDestroy) g_nccl.commDestroy(c->ncclComm);
	c->ncclComm = nullptr; c->ownComm = false;
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
