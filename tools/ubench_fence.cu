// Cost of fence flavours for a single thread with nothing outstanding (cycles), on an otherwise busy or idle SM.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k(long long* out, unsigned* g, int busy) {
  __shared__ float sm[1024];
  if (threadIdx.x >= 32) {  // other warps: keep the SM busy with FMA + shared traffic if requested
    float a = threadIdx.x;
    for (int i = 0; i < (busy ? 200000 : 0); ++i) { a = a * 1.0001f + sm[(threadIdx.x + i) & 1023]; }
    sm[threadIdx.x & 1023] = a;
    return;
  }
  if (threadIdx.x != 0) return;
  long long t[8];
  for (int rep = 0; rep < 3; ++rep) {
    long long c0 = clock64();
    asm volatile("fence.sc.gpu;" ::: "memory");
    long long c1 = clock64();
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    long long c2 = clock64();
    asm volatile("fence.acq_rel.cta;" ::: "memory");
    long long c3 = clock64();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(g), "r"(1u) : "memory");
    long long c4 = clock64();
    asm volatile("fence.acq_rel.gpu;" ::: "memory");  // with one RED outstanding
    long long c5 = clock64();
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(g + 32), "r"(1u) : "memory");
    long long c6 = clock64();
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(g + 64) : "memory");
    long long c7 = clock64();
    t[0] = c1 - c0; t[1] = c2 - c1; t[2] = c3 - c2; t[3] = c4 - c3; t[4] = c5 - c4; t[5] = c6 - c5; t[6] = c7 - c6 + (v & 0);
  }
  for (int i = 0; i < 7; ++i) out[blockIdx.x * 8 + i] = t[i];
}
int main() {
  long long* out; unsigned* g;
  cudaMalloc(&out, 8 * 8 * 148 * 4); cudaMalloc(&g, 4096); cudaMemset(g, 0, 4096);
  const char* names[7] = {"fence.sc.gpu", "fence.acq_rel.gpu", "fence.acq_rel.cta", "red.relaxed issue", "fence.acq_rel.gpu after RED", "red.release.gpu", "ld.acquire.gpu"};
  for (int busy = 0; busy < 2; ++busy) {
    k<<<busy ? 296 : 1, 256>>>(out, g, busy);
    cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s:\n", busy ? "busy GPU (296 CTAs)" : "idle GPU (1 CTA)");
    for (int i = 0; i < 7; ++i) printf("  %-30s %6lld cycles\n", names[i], h[i]);
  }
  return 0;
}
