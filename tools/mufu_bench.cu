// Microbenchmark: issue cost of the special-function instructions the GELU epilogues use, per warp instruction and
// SM sub-partition (4 warps per sub-partition, 8 independent chains per thread).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mufu_bench tools/mufu_bench.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int OP>
__global__ void __launch_bounds__(512) k(uint32_t* out, long long* cyc, int iters) {
  uint32_t v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0x38003800u + threadIdx.x * 8 + j;       // ~0.5 as half2 / a small float
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(v[j]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[j]));
      if (OP == 2) { float f = __uint_as_float(v[j]); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f)); v[j] = __float_as_uint(f); }
      if (OP == 3) { float f = __uint_as_float(v[j]); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); v[j] = __float_as_uint(f); }
      if (OP == 4) { __half2 h = *reinterpret_cast<__half2*>(&v[j]); h = __hfma2(h, h, h); v[j] = *reinterpret_cast<uint32_t*>(&h); }
      if (OP == 5) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(v[j]));
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s ^= v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount, iters = 4096;
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(uint32_t) * sms * 512));
  CK(cudaMalloc(&cyc, sizeof(long long) * sms));
  long long h[256];
  const char* names[6] = {"tanh.approx.f16x2", "ex2.approx.f16x2", "tanh.approx.f32", "ex2.approx.ftz.f32", "HFMA2 (reference)", "tanh.approx.bf16x2"};
  for (int op = 0; op < 6; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) k<0><<<sms, 512>>>(out, cyc, iters);
      if (op == 1) k<1><<<sms, 512>>>(out, cyc, iters);
      if (op == 2) k<2><<<sms, 512>>>(out, cyc, iters);
      if (op == 3) k<3><<<sms, 512>>>(out, cyc, iters);
      if (op == 4) k<4><<<sms, 512>>>(out, cyc, iters);
      if (op == 5) k<5><<<sms, 512>>>(out, cyc, iters);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double c = 0;
    for (int b = 0; b < sms; ++b) c += h[b];
    c /= sms;
    // 16 warps per SM = 4 per sub-partition, each issuing iters * 8 instructions
    printf("%-22s %6.2f cycles per warp instruction and sub-partition  (%5.1f lanes / clk / SM)\n", names[op],
           c / (4.0 * iters * 8), 32.0 * 16 * iters * 8 / c);
  }
  return 0;
}
