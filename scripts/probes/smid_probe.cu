// Which SM does block i of a launch land on?  (block scheduler placement, for the launch-order table)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 3) probe(int *smid, long long *t0, int spin) {
  extern __shared__ float4 smem[];
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  if (threadIdx.x == 0) {
    smid[blockIdx.x] = (int)id;
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    t0[blockIdx.x] = t;
  }
  smem[threadIdx.x] = make_float4(1, 2, 3, 4);
  __syncthreads();
  long long start = clock64();
  while (clock64() - start < spin) {}
}
int main() {
  const int n = 1012;
  int *d; long long *t;
  cudaMalloc(&d, n * sizeof(int)); cudaMalloc(&t, n * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  for (int rep = 0; rep < 2; rep++) probe<<<n, 256, 56 * 1024>>>(d, t, 20000);
  cudaDeviceSynchronize();
  int h[n]; long long ht[n];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(ht, t, sizeof ht, cudaMemcpyDeviceToHost);
  printf("first 64 blocks -> SM:");
  for (int i = 0; i < 64; i++) printf(" %d", h[i]);
  printf("\n");
  int distinct[200] = {0}; int nd = 0;
  for (int i = 0; i < 41; i++) if (!distinct[h[i]]++) nd++;
  printf("blocks 0..40 land on %d distinct SMs\n", nd);
  nd = 0; for (int i = 0; i < 200; i++) distinct[i] = 0;
  for (int i = 0; i < 148; i++) if (!distinct[h[i]]++) nd++;
  printf("blocks 0..147 land on %d distinct SMs\n", nd);
  printf("start time of block 0, 147, 148, 443, 444, 600 (ns rel): %lld %lld %lld %lld %lld %lld\n", 0LL, ht[147] - ht[0], ht[148] - ht[0], ht[443] - ht[0], ht[444] - ht[0], ht[600] - ht[0]);
  return 0;
}
