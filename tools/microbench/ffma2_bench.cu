// Does the packed fp32 FMA of sm_100 (fma.rn.f32x2 -> FFMA2) raise the FMA rate per issue slot?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a ffma2_bench.cu -o ffma2_bench ; run on the B200.
// Prints FMA lanes per clock per SM for scalar FFMA and for FFMA2 (8 independent chains per thread, 4 warps per SMSP),
// and for a 1:1 mix of FFMA(2) with integer adds (the issue-bound case of the edge kernel's epilogue).
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ int iadd(int a, int b) {
  int d;
  asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

template <int kMode>  // 0: FFMA, 1: FFMA2, 2: FFMA + IADD, 3: FFMA2 + IADD
__global__ void __launch_bounds__(512) bench(float* out, int iters, float a, float b) {
  float s[8];
  unsigned long long p[8];
  int n[8];
  for (int j = 0; j < 8; ++j) s[j] = threadIdx.x + j, p[j] = pack(s[j], s[j] + 1.f), n[j] = j;
  const unsigned long long pa = pack(a, a), pb = pack(b, b);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (kMode == 0 || kMode == 2) s[j] = ffma1(s[j], a, b);
      if (kMode == 1 || kMode == 3) p[j] = ffma2(p[j], pa, pb);
      if (kMode >= 2) n[j] = iadd(n[j], i);
    }
  }
  float r = 0.f;
  for (int j = 0; j < 8; ++j) r += s[j] + (float)(p[j] & 0xffff) + n[j];
  if (r == 123.456f) out[0] = r;
}

template <int kMode>
static void run(const char* name, int sms, float clock_ghz) {
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 1 << 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  bench<kMode><<<sms, 512>>>(out, 1024, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  bench<kMode><<<sms, 512>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double insts = (double)iters * 8 * 512;               // FMA instructions per SM (thread level)
  const double lanes = insts * ((kMode & 1) ? 2 : 1);
  printf("%-14s %8.3f ms  %7.1f fma-instr/clk/SM  %7.1f fma-lanes/clk/SM (at %.3f GHz)\n", name, ms,
         insts / (ms * 1e-3 * clock_ghz * 1e9), lanes / (ms * 1e-3 * clock_ghz * 1e9), clock_ghz);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float ghz = khz * 1e-6f;   // nominal max; short kernels run at it
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>("FFMA", p.multiProcessorCount, ghz);
  run<1>("FFMA2", p.multiProcessorCount, ghz);
  run<2>("FFMA+IADD", p.multiProcessorCount, ghz);
  run<3>("FFMA2+IADD", p.multiProcessorCount, ghz);
  return 0;
}
