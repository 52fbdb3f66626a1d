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
