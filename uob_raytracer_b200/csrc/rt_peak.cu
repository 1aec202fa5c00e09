// rt_peak.cu — FP32 (non-tensor) pipe microbenchmark: the roofline denominator of the
// brute-force render path.  MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only, and
// this path is neither, so the FFMA peak is measured in the same run, on the
// same clocks, as the kernel it bounds.
#include "rt_internal.h"

namespace rt {

constexpr int kIlp = 16;

__global__ void __launch_bounds__(256) ffma_peak_kernel(float *out, int iters, float a, float b) {
  float acc[kIlp];
#pragma unroll
  for (int k = 0; k < kIlp; k++) acc[k] = (float)(threadIdx.x + k);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < kIlp; k++) acc[k] = fmaf(acc[k], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < kIlp; k++) s += acc[k];
  if (s == 123.456f) out[0] = s;  // never true; keeps the chain alive
}

}  // namespace rt

extern "C" int rt_measure_fp32_peak(rt_ctx *ctx, float *tflops) {
  if (!ctx || !tflops) return RT_ERR_INVALID;
  auto fail = [&](const char *what, cudaError_t e) {
    ctx->err = std::string("CUDA error during '") + what + "': " + cudaGetErrorString(e);
    return RT_ERR_CUDA;
  };
  cudaError_t e;
  if ((e = cudaSetDevice(ctx->cfg.device)) != cudaSuccess) return fail("selecting device", e);
  float *d = nullptr;
  if ((e = cudaMalloc(&d, 4)) != cudaSuccess) return fail("allocating", e);
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 0.0f;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0, ctx->stream);
    rt::ffma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1, ctx->stream);
    if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * (double)blocks * threads * (double)iters * rt::kIlp;
    const float tf = (float)(flops / (ms * 1e-3) / 1e12);
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (e != cudaSuccess) return fail("ffma peak kernel", e);
  *tflops = best;
  return RT_OK;
}
