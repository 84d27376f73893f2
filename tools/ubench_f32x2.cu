// Microbenchmark: issue/pipe throughput of scalar FFMA vs packed FFMA2 (f32x2) on sm_100a, and the mix with LDS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu ; run on a B200.
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096

__global__ void k_ffma(float* out, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
  float2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed math interleaved with shared-memory loads (1 LDS.64 per 4 FFMA2): can issue slots be shared?
__global__ void k_mix(float* out, float a, float b) {
  __shared__ float2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float2(i, -i);
  __syncthreads();
  float2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  const float2 aa = make_float2(a, a);
  float2 acc = make_float2(0, 0);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      x[i] = __ffma2_rn(x[i], aa, acc);
      if ((i & 3) == 0) {
        float2 v = sm[(threadIdx.x + 32 * (i >> 2) + it) & 1023];
        acc = __fadd2_rn(acc, v);
      }
    }
  }
  float s = acc.x + acc.y;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const int blocks = sms * 2, threads = 512;  // 32 warps/SM
  const double warp_instrs = (double)blocks * (threads / 32) * ITERS * 16;
  float ms1 = time_it([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  float ms2 = time_it([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  float ms3 = time_it([&] { k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  printf("SMs=%d clock=%.0f MHz (nominal)\n", sms, khz / 1e3);
  printf("FFMA : %.3f ms  %.2f warp-instr/clk/SM (at nominal clk)  %.1f TFLOP/s\n", ms1,
         warp_instrs / (ms1 * 1e-3) / sms / (khz * 1e3), warp_instrs * 64 / (ms1 * 1e-3) / 1e12);
  printf("FFMA2: %.3f ms  %.2f warp-instr/clk/SM (at nominal clk)  %.1f TFLOP/s\n", ms2,
         warp_instrs / (ms2 * 1e-3) / sms / (khz * 1e3), warp_instrs * 128 / (ms2 * 1e-3) / 1e12);
  printf("MIX  : %.3f ms  (16 FFMA2 + 4 FADD2 + 4 LDS.64 per iter) %.2f warp-instr/clk/SM\n", ms3,
         warp_instrs * 24.0 / 16.0 / (ms3 * 1e-3) / sms / (khz * 1e3));
  return 0;
}
