// rt_peak.cu — small utility kernels: the FP32 (non-tensor) pipe microbenchmark: the roofline denominator of the
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

// ---- frame hand-over flags between GPUs (include/uob_rt.h: rt_peer_signal / rt_peer_wait) ----

__global__ void peer_signal_kernel(uint32_t *flag, uint32_t value) {
  // Everything this stream did before (the draw kernel's stores into the peer's frame) happened-before this
  // kernel; the system-scope fence makes it visible to the other GPU before the flag is.
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void peer_wait_kernel(const uint32_t *flags, int n, uint32_t value, int *status) {
  if ((int)threadIdx.x >= n) return;
  const uint32_t *f = flags + threadIdx.x;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if ((int32_t)(v - value) >= 0) break;
    if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz: give up instead of hanging the GPU
      atomicExch(status, 1);
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

struct PeerCounters {
  uint32_t *c[8];
};
// Launched as a PROGRAMMATIC DEPENDENT of the draw kernel in front of it (cudaLaunchAttributeProgrammaticStreamSerialization):
// its one block is scheduled while the draw kernel is still running and parks in griddepcontrol.wait, which returns once
// that grid has completed and its memory operations are performed — so the delivery is reported without the launch gap
// of an ordinary stream-ordered kernel behind a 40-100 us draw kernel (measured on two B200s: -DRT_PDL_SIGNAL=0 restores
// the plain launch).  Without a kernel in front of it on the stream the wait returns at once.
#ifndef RT_PDL_SIGNAL
#define RT_PDL_SIGNAL 1
#endif
#ifndef RT_SIGNAL_SC_FENCE
#define RT_SIGNAL_SC_FENCE 0
#endif
__global__ void peer_add_kernel(const __grid_constant__ PeerCounters pc) {
#if RT_PDL_SIGNAL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  // Everything this stream did before (the draw kernel's stores into the peers' frames) happened-before this point, and a
  // release at system scope is cumulative: whoever acquires the count sees those stores.  (red.release.sys carries its own
  // MEMBAR.ALL.SYS; a sequentially consistent __threadfence_system() in front of it — -DRT_SIGNAL_SC_FENCE=1, the form
  // of round 1 — is a second, slower system barrier that orders nothing more here.)
#if RT_SIGNAL_SC_FENCE
  __threadfence_system();
#endif
  asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(pc.c[threadIdx.x]) : "memory");
}

cudaError_t launch_peer_add_many(uint32_t *const *counters, int n, cudaStream_t stream) {
  PeerCounters pc{};
  for (int i = 0; i < n && i < 8; i++) pc.c[i] = counters[i];
#if RT_PDL_SIGNAL
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(1);
  lc.blockDim = dim3((unsigned)n);
  lc.dynamicSmemBytes = 0;
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&lc, peer_add_kernel, pc);
  if (e == cudaSuccess) return e;
  cudaGetLastError();  // (attribute not accepted: plain launch)
#endif
  peer_add_kernel<<<1, n, 0, stream>>>(pc);
  return cudaGetLastError();
}

cudaError_t launch_peer_add(uint32_t *counter, cudaStream_t stream) { return launch_peer_add_many(&counter, 1, stream); }

cudaError_t launch_peer_signal(uint32_t *flag, uint32_t value, cudaStream_t stream) {
  peer_signal_kernel<<<1, 1, 0, stream>>>(flag, value);
  return cudaGetLastError();
}

cudaError_t launch_peer_wait(const uint32_t *flags, int n, uint32_t value, int *status, cudaStream_t stream) {
  peer_wait_kernel<<<1, 32, 0, stream>>>(flags, n, value, status);
  return cudaGetLastError();
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
