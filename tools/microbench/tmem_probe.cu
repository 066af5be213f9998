// Probe of the data movement the edge kernel's store path relies on (run once on a B200, output kept under profiles/):
//   1. tcgen05.st.32x32b (thread = TMEM lane, registers = columns)  ->  tcgen05.ld.16x256b (matrix-fragment layout):
//      which (lane, column) does register r of thread T receive?
//   2. stmatrix.m8n8.x4.trans.b16 of such fragments: where do the elements land in shared memory?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(uint32_t* frag_out, uint16_t* smem_out) {
  __shared__ uint32_t slot;
  __shared__ __align__(16) uint16_t tile[32 * 8 * 4];   // 4 matrices of 8 rows x 8 halves per warp would be 256 halves; one warp only
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot;
  // 1. every thread writes value = lane_id * 256 + column into columns 0..15 of its own lane
  const uint32_t L = threadIdx.x;
  uint32_t v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = L * 256 + j;
  const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  // read back as 16x256b.x2 fragments: lanes [32w, 32w+16) then [32w+16, 32w+32)
#pragma unroll
  for (int hl = 0; hl < 2; ++hl) {
    uint32_t r[8];
    const uint32_t ta = base + ((uint32_t)(warp * 32 + hl * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(ta)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 8; ++k) frag_out[((warp * 2 + hl) * 32 + lane) * 8 + k] = r[k];
    // 2. warp 0, first half only: treat r[0..3] as four b16x2 fragments and store them transposed
    if (warp == 0 && hl == 0) {
      // element tags: (register index k, thread T, half h) -> 16-bit value k * 4096 + T * 2 + h
      uint32_t f[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = (uint32_t)(k * 4096 + lane * 2) | ((uint32_t)(k * 4096 + lane * 2 + 1) << 16);
      // row address provided by thread T: matrix T / 8, row T % 8 -> tile + (T / 8) * 64 + (T % 8) * 8 halves
      const uint32_t addr = smem_u32(tile + (lane >> 3) * 64 + (lane & 7) * 8);
      asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(f[0]), "r"(f[1]),
                   "r"(f[2]), "r"(f[3])
                   : "memory");
      __syncwarp();
      for (int i = lane; i < 256; i += 32) smem_out[i] = tile[i];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32));
}

int main() {
  uint32_t* d_frag;
  uint16_t* d_smem;
  cudaMalloc(&d_frag, 4 * 2 * 32 * 8 * 4);
  cudaMalloc(&d_smem, 256 * 2);
  probe<<<1, 128>>>(d_frag, d_smem);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("probe failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  static uint32_t frag[4 * 2 * 32 * 8];
  static uint16_t sm[256];
  cudaMemcpy(frag, d_frag, sizeof(frag), cudaMemcpyDeviceToHost);
  cudaMemcpy(sm, d_smem, sizeof(sm), cudaMemcpyDeviceToHost);
  printf("== tcgen05.ld.16x256b.x2 after tcgen05.st.32x32b.x16: thread T register k <- (lane, column)\n");
  for (int w = 0; w < 2; ++w)
    for (int hl = 0; hl < 2; ++hl)
      for (int T = 0; T < 32; T += (T < 8 ? 1 : 8)) {
        printf("warp %d half %d T %2d:", w, hl, T);
        for (int k = 0; k < 8; ++k) {
          const uint32_t v = frag[((w * 2 + hl) * 32 + T) * 8 + k];
          printf(" r%d=(%3u,%2u)", k, v >> 8, v & 255);
        }
        printf("\n");
      }
  // check the conjectured rule on everything: r[2a+b + 4c] of thread T <- lane base + T/4 + 8a, column 8c + 2(T%4) + b
  int bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int hl = 0; hl < 2; ++hl)
      for (int T = 0; T < 32; ++T)
        for (int k = 0; k < 8; ++k) {
          const uint32_t v = frag[((w * 2 + hl) * 32 + T) * 8 + k];
          const int c = k >> 2, a = (k >> 1) & 1, b = k & 1;
          const uint32_t want = (uint32_t)(w * 32 + hl * 16 + T / 4 + 8 * a) * 256 + (8 * c + 2 * (T % 4) + b);
          bad += v != want;
        }
  printf("rule r[4c+2a+b](T) = (lane0 + T/4 + 8a, 8c + 2(T%%4) + b): %s (%d mismatches)\n", bad ? "WRONG" : "holds", bad);
  printf("== stmatrix.x4.trans: smem matrix m, row i, column j <- (register, thread, half)\n");
  int bad2 = 0;
  for (int m = 0; m < 4; ++m)
    for (int i = 0; i < 8; ++i) {
      if (m == 0) printf("m0 row %d:", i);
      for (int j = 0; j < 8; ++j) {
        const uint16_t v = sm[m * 64 + i * 8 + j];
        const int k = v >> 12, T = (v & 4095) >> 1, h = v & 1;
        if (m == 0) printf(" (r%d,T%2d,h%d)", k, T, h);
        // conjecture (.trans): memory[m][i][j] comes from register m of thread T = 4 * j + i / 2, half i % 2
        bad2 += !(k == m && T == 4 * j + i / 2 && h == (i & 1));
      }
      if (m == 0) printf("\n");
    }
  printf("rule mem[m][i][j] = reg m of thread 4j + i/2, half i%%2: %s (%d mismatches)\n", bad2 ? "WRONG" : "holds", bad2);
  return 0;
}
