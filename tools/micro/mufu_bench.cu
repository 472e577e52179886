// Micro-benchmark: MUFU.EX2 throughput per SM for f32 vs packed f16x2 operands (decides whether the flash-attention
// softmax should exponentiate in half2).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__global__ void k_f32(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * 0.001f - 1.f;          // keep the values bounded (FMA pipe, 1 per MUFU)
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_f16x2(float* out, int iters) {
  unsigned a[8];
  for (int i = 0; i < 8; ++i) {
    __half2 h = __floats2half2_rn(threadIdx.x * 1e-3f + i, 0.5f * i);
    a[i] = *reinterpret_cast<unsigned*>(&h);
  }
  const __half2 sc = __floats2half2_rn(0.001f, 0.001f), of = __floats2half2_rn(-1.f, -1.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __half2 h = __hfma2(*reinterpret_cast<__half2*>(&a[i]), sc, of);
      a[i] = *reinterpret_cast<unsigned*>(&h);
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += __low2float(*reinterpret_cast<__half2*>(&a[i]));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const int iters = 20000, blocks = sms * 2, threads = 1024;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_f32<<<blocks, threads>>>(out, iters);
      else k_f16x2<<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double instr = (double)blocks * threads * iters * 8;
      const double elems = instr * (which == 0 ? 1 : 2);
      if (rep == 2)
        printf("%s: %.3f ms  %.2f MUFU instr/ns  = %.1f instr/clk/SM @1.9GHz, %.1f exp/clk/SM\n", which == 0 ? "ex2.f32  " : "ex2.f16x2",
               ms, instr / ms * 1e-6, instr / ms * 1e-6 / 1.9 / sms, elems / ms * 1e-6 / 1.9 / sms);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
