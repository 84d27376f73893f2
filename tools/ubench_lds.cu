// shared-memory load bandwidth per SM for conflict-free LDS.32 / LDS.64 / LDS.128 and warp-uniform (broadcast) LDS.128,
// 16 warps per CTA, one CTA per SM, 8 loads in flight per warp.  Answers: what does one LDS.128 by frame cost the
// log-mel kernel's operand preparation (wfe_logmel_tc.cuh)?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_lds tools/ubench_lds.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
template <int kMode>
__global__ void __launch_bounds__(512, 1) lds_kernel(int iters, long long* cyc, float* sink_g) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 16384; i += 512) reinterpret_cast<float*>(smem)[i] = (float)i;
  __syncthreads();
  float sink = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (kMode == 0) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = reinterpret_cast<const float*>(smem)[lane + 32 * ((it * 8 + u) & 255)];
#pragma unroll
      for (int u = 0; u < 8; ++u) sink += v[u];
    } else if (kMode == 1) {
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = reinterpret_cast<const float2*>(smem)[lane + 32 * ((it * 8 + u) & 127)];
#pragma unroll
      for (int u = 0; u < 8; ++u) sink += v[u].x + v[u].y;
    } else if (kMode == 2) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = reinterpret_cast<const float4*>(smem)[lane + 32 * ((it * 8 + u) & 63)];
#pragma unroll
      for (int u = 0; u < 8; ++u) sink += v[u].x + v[u].w;
    } else if (kMode == 3) {  // frame-strided LDS.128 as in the kernel: lane stride 164 floats
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(smem) + lane * 164 + 4 * ((it * 8 + u) & 31));
#pragma unroll
      for (int u = 0; u < 8; ++u) sink += v[u].x + v[u].w;
    } else {  // warp-uniform LDS.128
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = reinterpret_cast<const float4*>(smem)[(it * 8 + u) & 63];
#pragma unroll
      for (int u = 0; u < 8; ++u) sink += v[u].x + v[u].w;
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (sink == 1.2345f) sink_g[0] = sink;
}
template <int kMode>
void run(const char* name, int bytes_per_lane) {
  long long* d; float* s; cudaMalloc(&d, 8 * 148); cudaMalloc(&s, 4);
  cudaFuncSetAttribute(lds_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int iters = 2000;
  lds_kernel<kMode><<<148, 512, 65536>>>(iters, d, s);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, 8 * 148, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += (double)h[i] / 148;
  const double n = 16.0 * iters * 8;  // warp-instructions per SM
  printf("%-44s %6.2f cycles per warp-instruction per SM, %6.1f B/clk/SM delivered (%s)\n", name, c / n, n * 32 * bytes_per_lane / c, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<0>("LDS.32  conflict-free", 4);
  run<1>("LDS.64  conflict-free", 8);
  run<2>("LDS.128 conflict-free (contiguous)", 16);
  run<3>("LDS.128 by frame (lane stride 164 words)", 16);
  run<4>("LDS.128 warp-uniform (broadcast)", 16);
  return 0;
}
