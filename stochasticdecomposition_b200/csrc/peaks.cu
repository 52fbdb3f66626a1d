// peaks.cu -- instrumentation: the FP64 pipe peak of the device for the instruction mix this library is allowed to use.
//
// Every FP64 operation on the path is a separately rounded multiply or add (-fmad=false, __dadd_rn / __dmul_rn: the reference's
// scalar C rounds each operation), so the roofline denominator of the FP64-bound kernels (k_sweep_recompute,
// k_delta_block_rhs) is the DADD + DMUL issue rate, not the DFMA rate the data sheet quotes.  MEASURED_PEAKS.json has no FP64 figure;
// this micro-benchmark supplies one: register-resident chains, 8 independent chains per thread, every SM full.
#include "sdgpu_internal.cuh"

template <bool FMA>
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double seed) {
	double a[8], m = 1.0 + seed * 1e-9, b = seed * 1e-12;
#pragma unroll
	for (int u = 0; u < 8; u++) a[u] = seed + threadIdx.x + u;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			if (FMA) a[u] = __fma_rn(a[u], m, b);                          // 1 instruction, 2 flops
			else a[u] = __dadd_rn(__dmul_rn(a[u], m), b);                  // 2 instructions, 2 flops: what the library issues
		}
	}
	double s = 0.0;
#pragma unroll
	for (int u = 0; u < 8; u++) s += a[u];
	if (s == 12345.678) out[0] = s;                                        // keeps the chains alive
}

// ops / s of separately rounded DMUL + DADD pairs (counted as 2 operations per pair) and flops / s of DFMA, best of `reps` launches
extern "C" int sdgpu_fp64_peak(int device, int reps, double *mulAddOpsPerSec, double *fmaFlopsPerSec) {
	if (!mulAddOpsPerSec || !fmaFlopsPerSec) return sdgpu_fail("null argument");
	SD_CUDA(cudaSetDevice(device));
	int sms = 0;
	SD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
	double *d = nullptr;
	SD_CUDA(cudaMalloc((void **) &d, 64));
	cudaEvent_t e0, e1;
	SD_CUDA(cudaEventCreate(&e0)); SD_CUDA(cudaEventCreate(&e1));
	const int iters = 1 << 14, blocks = sms * 8;
	const double work = (double) blocks * 256 * 8 * 2 * (double) iters;      // operations (= flops) per launch, either flavour
	double best[2] = {0.0, 0.0};
	for (int f = 0; f < 2; f++)
		for (int r = 0; r < reps + 1; r++) {
			SD_CUDA(cudaEventRecord(e0));
			if (f) k_fp64_peak<true><<<blocks, 256>>>(d, iters, 1.0); else k_fp64_peak<false><<<blocks, 256>>>(d, iters, 1.0);
			SD_CUDA(cudaEventRecord(e1));
			SD_CUDA(cudaEventSynchronize(e1));
			float ms = 0.f;
			SD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
			if (r > 0 && ms > 0.f) best[f] = std::max(best[f], work / (ms * 1e-3));
		}
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
	SD_CUDA(cudaGetLastError());
	*mulAddOpsPerSec = best[0]; *fmaFlopsPerSec = best[1];
	return 0;
}

// ---- the floor under every synchronous call: launch -> kernel start -> completion seen by the host ----------------------------
__global__ void k_roundtrip(volatile unsigned *flag, unsigned seq, int chain) {
	if (chain) sd_pdl_wait();
	if (flag && threadIdx.x == 0) { __threadfence_system(); *flag = seq; }
}

#include <chrono>
#include <vector>
#include <algorithm>
// median wall time in us of `launches` empty kernels in a row (chained with programmatic dependent launch when launches > 1) followed by
//   mode 0: cudaStreamSynchronize        mode 1: the host spinning on a word the last kernel writes into mapped pinned memory
extern "C" int sdgpu_launch_roundtrip(int device, int mode, int launches, int reps, double *medianUs) {
	if (!medianUs || launches < 1 || reps < 1) return sdgpu_fail("launch_roundtrip: bad argument");
	SD_CUDA(cudaSetDevice(device));
	cudaStream_t st;
	SD_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	unsigned *h = nullptr, *d = nullptr;
	SD_CUDA(cudaHostAlloc((void **) &h, 64, cudaHostAllocMapped));
	SD_CUDA(cudaHostGetDevicePointer((void **) &d, h, 0));
	*h = 0;
	std::vector<double> us;
	int rc = 0;
	for (int r = 0; r < reps + 8 && rc == 0; r++) {
		const unsigned seq = (unsigned) r + 1;
		const auto t0 = std::chrono::steady_clock::now();
		for (int l = 0; l < launches && rc == 0; l++) {
			const bool last = l == launches - 1;
			if (sd_launch(k_roundtrip, dim3(1), dim3(32), 0, st, l > 0, (volatile unsigned *) (last && mode == 1 ? d : nullptr), seq, l > 0 ? 1 : 0) != cudaSuccess)
				rc = sdgpu_fail("launch_roundtrip: launch failed");
		}
		if (rc) break;
		if (mode == 1) {
			long spins = 0;
			while (*(volatile unsigned *) h != seq)
				if ((++spins & 0xfffff) == 0 && cudaStreamQuery(st) != cudaErrorNotReady) break;
		}
		else if (cudaStreamSynchronize(st) != cudaSuccess) rc = sdgpu_fail("launch_roundtrip: sync failed");
		const auto t1 = std::chrono::steady_clock::now();
		if (r >= 8) us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
		if (mode == 1) cudaStreamSynchronize(st);
	}
	cudaStreamDestroy(st); cudaFreeHost(h);
	if (rc) return rc;
	std::sort(us.begin(), us.end());
	*medianUs = us[us.size() / 2];
	return 0;
}
