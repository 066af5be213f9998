// How does mbarrier.try_wait behave on sm_100a when the phase is not complete yet?  The tensor-core kernels wait for
// whole pipeline stages (1-3k cycles) and every poll of a spin loop is an issue slot (and energy) taken from the four
// epilogue warps that share the scheduler, so the question is whether the hardware can hold the waiter:
//   mode 0: try_wait without a time limit, polled back to back
//   mode 1: try_wait + nanosleep(32) between polls (the round-2a loop)
//   mode 2: try_wait with suspendTimeHint = hint_ns, polled back to back
// One waiter warp, one signaller warp that arrives `delay` cycles after the start.  Reported: polls per wait and the
// wake-up latency (cycles from the arrive to the waiter's first instruction after the wait; same SM clock).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o trywait_probe trywait_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(64, 1) probe(int mode, uint32_t hint_ns, int delay, int reps, long long* out) {
  __shared__ uint64_t bar;
  __shared__ long long t_arrive;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long polls = 0, lat = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    const uint32_t parity = r & 1;
    if (warp == 1) {
      if (lane == 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < delay) {}
        t_arrive = clock64();
        __threadfence_block();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
    } else {
      uint32_t done = 0;
      while (!done) {
        if (mode == 2) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
              : "=r"(done)
              : "r"(smem_u32(&bar)), "r"(parity), "r"(hint_ns)
              : "memory");
        } else {
          asm volatile(
              "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
              : "=r"(done)
              : "r"(smem_u32(&bar)), "r"(parity)
              : "memory");
          if (!done && mode == 1) __nanosleep(32);
        }
        ++polls;
      }
      const long long t1 = clock64();
      lat += t1 - *(volatile long long*)&t_arrive;
    }
  }
  if (threadIdx.x == 0) {
    out[0] = polls;
    out[1] = lat;
  }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  const int reps = 200;
  printf("%-34s %10s %12s %14s\n", "mode", "delay", "polls/wait", "wake-up cycles");
  for (int delay : {1000, 3000, 20000, 200000}) {
    struct { int mode; uint32_t hint; const char* name; } cases[] = {
        {0, 0, "try_wait, back to back"},      {1, 0, "try_wait + nanosleep(32)"}, {2, 1000, "try_wait hint 1 us"},
        {2, 20000, "try_wait hint 20 us"}, {2, 1000000, "try_wait hint 1 ms"}};
    for (auto& c : cases) {
      probe<<<1, 64>>>(c.mode, c.hint, delay, reps, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      printf("%-34s %10d %12.1f %14.1f\n", c.name, delay, (double)out[0] / reps, (double)out[1] / reps);
    }
  }
  return 0;
}
