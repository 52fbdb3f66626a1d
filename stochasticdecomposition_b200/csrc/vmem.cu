// vmem.cu -- the delta table on reserved virtual address space, physical memory mapped as the table grows.
//
// The reference allocates every table for its capacity contract up front (setup.c:136-144: MAX_ITER + MAX_ITER / TAU + 1 rows, times
// rvdOmCnt with random cost) and SURVEY.md section 7 lists "growth without realloc-copy (reserve VA once, map slabs)" as the hard part of
// a device-resident delta: 8 (1+Q) bytes x duals x observations is the one table whose capacity can exceed the GPU (65 536 x 1 048 576
// doubles = 512 GiB) while the part in use is small.  With SDGPU_VMM=1 -- or automatically when the capacity asked for does not fit
// the free device memory -- sdgpu_create reserves the table's whole address range (cuMemAddressReserve) and maps physical memory
// (cuMemCreate / cuMemMap, 2 MiB granules) only under the part that is in use:
//   layout [tile][Dcap][1+Q][512] (sdgpu_internal.cuh): one dual row of one observation tile is (1+Q) x 4 KiB, so 512 rows of a tile
//   are (1+Q) x 2 MiB -- a whole number of granules when the row stride Dcap is a multiple of 512 (it is rounded up for this);
//   tiles [0, ceil(N / 512)) are mapped over rows [0, roundup(D, 512)); a new dual that crosses a 512-row boundary maps one more slab
//   per tile in use, an observation that opens a new tile maps that tile's rows; bulk builds map their whole range with one
//   allocation per tile.  Addresses never move, so nothing is copied and kernels keep the pointers they were given.
// The driver entry points come from cudaGetDriverEntryPoint: the library still links against nothing but the CUDA runtime.
#include "sdgpu_internal.cuh"
#include <cuda.h>
#include <vector>

struct SdVmMap { CUdeviceptr at; size_t bytes; CUmemGenericAllocationHandle h; };

struct SdVm {
	CUdeviceptr base = 0; size_t reserved = 0, gran = 0, mappedBytes = 0;
	size_t rowBytes = 0, tileBytes = 0;          // one dual row of one tile (all planes); Dcap rows
	int64_t Dcap = 0, nTiles = 0, slabRows = 512;
	std::vector<int64_t> tileRows;               // rows mapped so far, per tile
	std::vector<SdVmMap> maps;
	CUmemAllocationProp prop;
	CUmemAccessDesc access;
	CUresult (*addressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
	CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
	CUresult (*memCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
	CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
	CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
	CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
	CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
	CUresult (*getGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
};

template <class F>
static bool sd_vm_entry(const char *name, F *fn) {
	void *p = nullptr;
	cudaDriverEntryPointQueryResult st = cudaDriverEntryPointSymbolNotFound;
	if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) { cudaGetLastError(); return false; }
	*fn = reinterpret_cast<F>(p);
	return true;
}

void sd_vm_destroy(sdgpu_ctx *c) {
	SdVm *v = c->vm;
	if (!v) return;
	for (const SdVmMap &m : v->maps) { v->memUnmap(m.at, m.bytes); v->memRelease(m.h); }
	if (v->base) v->addressFree(v->base, v->reserved);
	delete v;
	c->vm = nullptr; c->d_delta = nullptr;
}

// reserve the address range of a delta table of nTiles x Dcap rows of rowBytes each; Dcap must be a multiple of 512
int sd_vm_create(sdgpu_ctx *c, size_t rowBytes, int64_t Dcap, int64_t nTiles) {
	SdVm *v = new SdVm;
	c->vm = v;
	if (!sd_vm_entry("cuMemAddressReserve", &v->addressReserve) || !sd_vm_entry("cuMemAddressFree", &v->addressFree) ||
	    !sd_vm_entry("cuMemCreate", &v->memCreate) || !sd_vm_entry("cuMemRelease", &v->memRelease) || !sd_vm_entry("cuMemMap", &v->memMap) ||
	    !sd_vm_entry("cuMemUnmap", &v->memUnmap) || !sd_vm_entry("cuMemSetAccess", &v->memSetAccess) ||
	    !sd_vm_entry("cuMemGetAllocationGranularity", &v->getGranularity)) {
		delete v; c->vm = nullptr;
		return sdgpu_fail("virtual memory management entry points are not available from this driver");
	}
	memset(&v->prop, 0, sizeof v->prop);
	v->prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
	v->prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	v->prop.location.id = c->device;
	v->access.location = v->prop.location;
	v->access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
	CUresult r = v->getGranularity(&v->gran, &v->prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
	if (r != CUDA_SUCCESS || v->gran == 0) { delete v; c->vm = nullptr; return sdgpu_fail("cuMemGetAllocationGranularity failed (%d)", (int) r); }
	v->rowBytes = rowBytes; v->Dcap = Dcap; v->nTiles = nTiles; v->tileBytes = (size_t) Dcap * rowBytes;
	// rows per slab: the smallest multiple of 512 whose bytes are a whole number of granules (512 rows = (1+Q) x 2 MiB: 1 for the usual 2 MiB)
	while (((size_t) v->slabRows * rowBytes) % v->gran != 0 && v->slabRows < Dcap) v->slabRows += 512;
	if (((size_t) v->slabRows * rowBytes) % v->gran != 0 || Dcap % v->slabRows != 0 || v->tileBytes % v->gran != 0) {
		const size_t g = v->gran; delete v; c->vm = nullptr;
		return sdgpu_fail("delta row stride %lld x %zu bytes does not tile the allocation granularity %zu", (long long) Dcap, rowBytes, g);
	}
	v->reserved = v->tileBytes * (size_t) nTiles;
	r = v->addressReserve(&v->base, v->reserved, v->gran, 0, 0);
	if (r != CUDA_SUCCESS) { const size_t n = v->reserved; delete v; c->vm = nullptr; return sdgpu_fail("cuMemAddressReserve of %zu bytes failed (%d)", n, (int) r); }
	v->tileRows.assign((size_t) nTiles, 0);
	c->d_delta = reinterpret_cast<double *>(v->base);
	return 0;
}

static int sd_vm_map(SdVm *v, int64_t tile, int64_t row0, int64_t row1) {
	SdVmMap m;
	m.at = v->base + (size_t) tile * v->tileBytes + (size_t) row0 * v->rowBytes;
	m.bytes = (size_t) (row1 - row0) * v->rowBytes;
	CUresult r = v->memCreate(&m.h, m.bytes, &v->prop, 0);
	if (r != CUDA_SUCCESS)
		return sdgpu_fail("out of device memory while growing the delta table: %zu more bytes for rows [%lld, %lld) of observation tile %lld (%zu bytes mapped so far) (%d)",
				m.bytes, (long long) row0, (long long) row1, (long long) tile, v->mappedBytes, (int) r);
	r = v->memMap(m.at, m.bytes, 0, m.h, 0);
	if (r != CUDA_SUCCESS) { v->memRelease(m.h); return sdgpu_fail("cuMemMap failed (%d)", (int) r); }
	r = v->memSetAccess(m.at, m.bytes, &v->access, 1);
	if (r != CUDA_SUCCESS) { v->memUnmap(m.at, m.bytes); v->memRelease(m.h); return sdgpu_fail("cuMemSetAccess failed (%d)", (int) r); }
	v->maps.push_back(m);
	v->mappedBytes += m.bytes;
	v->tileRows[(size_t) tile] = row1;
	return 0;
}

// physical memory under rows [0, rows) of the tiles that hold observations [0, obs); a no-op for a table that was allocated whole
int sd_delta_ensure(sdgpu_ctx *c, int64_t rows, int64_t obs) {
	SdVm *v = c->vm;
	if (!v) return 0;
	const int64_t tiles = std::min<int64_t>(v->nTiles, (obs + SD_TILE_W - 1) / SD_TILE_W);
	int64_t want = std::min<int64_t>(v->Dcap, ((std::max<int64_t>(rows, 1) + v->slabRows - 1) / v->slabRows) * v->slabRows);
	// a tile that is already in use keeps at least what its neighbours have: rows are uniform over the tiles in use
	for (int64_t t = 0; t < tiles; t++) want = std::max(want, v->tileRows[(size_t) t]);
	for (int64_t t = 0; t < tiles; t++)
		if (v->tileRows[(size_t) t] < want && sd_vm_map(v, t, v->tileRows[(size_t) t], want)) return SDGPU_ERR;
	return 0;
}

extern "C" int sdgpu_delta_memory(sdgpu_ctx *c, int64_t *reservedBytes, int64_t *mappedBytes) {
	if (!c) return sdgpu_fail("null context");
	const int64_t whole = (int64_t) c->nTiles * c->Dcap * (1 + c->Q) * SD_TILE_W * 8;
	if (reservedBytes) *reservedBytes = c->vm ? (int64_t) c->vm->reserved : whole;
	if (mappedBytes) *mappedBytes = c->vm ? (int64_t) c->vm->mappedBytes : whole;
	return c->vm ? 1 : 0;
}
