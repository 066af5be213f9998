// tcgen05 building blocks shared by the tensor-core kernels (sm_100a).
//
// Every dense product on the hot path has the shape  D^T[c][r] = sum_k W[c][k] * X[r][k]  with
//   c = output channel (<= 128 per CTA), r = a row of edge / node state, k = input channel (K = H).
// It is issued TRANSPOSED on purpose: W is the MMA "A" operand and lives in TENSOR MEMORY for the whole
// life of the persistent CTA (tcgen05.mma with A in TMEM), the row tile X is the "B" operand in shared
// memory, and the accumulator comes out with TMEM lane = output channel, TMEM column = row.  The
// epilogue thread that owns lane c therefore holds ONE channel of MANY consecutive rows -- exactly the
// layout the per-destination aggregation wants (register accumulators, coalesced 128-byte gathers of
// node-table rows across the 32 lanes of a warp).
//
// Precision: fp32 inputs are split on the fly into fp16 (hi, lo) pairs, x*2^-4 = hi + lo (22 mantissa
// bits), weights likewise (W*2^4), and three MMAs  Whi*Xhi + Whi*Xlo + Wlo*Xhi  accumulate in fp32:
// measured against the fp64 oracle this is indistinguishable from the fp32 FFMA path, whereas single
// pass tf32/bf16/fp16 and bf16x3 all miss the 1e-4 edge-probability tolerance.
#pragma once

#include <cuda_fp16.h>

#include "gnb_common.cuh"

namespace gnb {
namespace tc {

constexpr int kM = 128;            // MMA M = output channels per CTA (TMEM lanes)
constexpr float kXScale = 0.0625f; // activations are multiplied by 2^-4 before the fp16 split ...
constexpr float kWScale = 16.0f;   // ... and weights by 2^4, so the product needs no rescale

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Has the phase with this parity completed?  NOT a cheap probe: try_wait suspends the warp until the phase completes
// or a system time limit expires -- about 7.5k cycles on this part, and a suspendTimeHint does not shorten it
// (tools/microbench/trywait_probe.cu, profiles/r02h_trywait_probe.txt); a waiter wakes ~130 cycles after the arrive.
// That makes it the right instruction for waiting on ONE barrier (a wait of a pipeline stage costs one or two polls)
// and the wrong one for polling several: use mbar_poll there.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// returns at once
__device__ __forceinline__ bool mbar_poll(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// ---------------------------------------------------------------------------------------------
// Spin watchdog.  Every device-side wait of the tensor-core kernels is bounded: a wait that lasts longer than
// Watch::timeout_ns (wall clock, %globaltimer) writes WHO waited for WHAT into a record in host-mapped memory and
// traps, so that a lost barrier phase surfaces as a CUDA error with a diagnosis (gnb_hang_report) within seconds
// instead of a silent spin.  The clock is read once every 256 polls of a wait that has not succeeded at once.
// ---------------------------------------------------------------------------------------------
struct Watch {
  unsigned long long* rec;         // kWatchWords words in cudaHostAllocMapped memory, or nullptr
  unsigned long long timeout_ns;   // 0 = wait for ever
};
constexpr int kWatchWords = 8;
constexpr unsigned long long kWatchMagic = 0x474e42484e47ull;   // "GNBHNG"
enum WatchKernel : uint32_t { kWkEdge2 = 1, kWkLinear2 = 2, kWkScore2 = 3 };
enum WatchRole : uint32_t { kWrProducer = 1, kWrMma = 2, kWrStore = 3, kWrEpilogue = 4 };
enum WatchBar : uint32_t { kWbFull = 1, kWbEmpty = 2, kWbDFull = 3, kWbDEmpty = 4, kWbOFull = 5, kWbOEmpty = 6, kWbIFull = 7 };
__host__ __device__ constexpr uint32_t watch_tag(uint32_t kernel, uint32_t role, uint32_t bar) {
  return (kernel << 16) | (role << 8) | bar;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

static __device__ __noinline__ void watch_fail(Watch w, uint32_t tag, uint32_t index, uint32_t parity, long long iter,
                                        unsigned long long waited_ns) {
  if (w.rec != nullptr) {
    volatile unsigned long long* r = w.rec;
    r[1] = tag;
    r[2] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;
    r[3] = ((unsigned long long)index << 32) | parity;
    r[4] = (unsigned long long)iter;
    r[5] = waited_ns;
    __threadfence_system();
    r[0] = kWatchMagic;
    __threadfence_system();
  }
  __trap();
}

// State of one bounded wait (a few registers, live only while a wait is unsuccessful).
struct SpinGuard {
  uint32_t polls = 0;
  unsigned long long t0 = 0;
  __device__ __forceinline__ void poll(const Watch& w, uint32_t tag, uint32_t index, uint32_t parity, long long iter) {
    if (((++polls) & 0xffu) == 0 && w.timeout_ns != 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > w.timeout_ns) watch_fail(w, tag, index, parity, iter, now - t0);
    }
  }
};

// Wait for the phase with this parity; `ns` > 0 backs off between polls so that a warp that is expected to wait
// for a whole pipeline stage does not take issue slots from the warps doing the work on the same scheduler.
// tag / index / iter only describe the wait for the watchdog record.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t ns, const Watch& w, uint32_t tag,
                                          uint32_t index, long long iter) {
  if (mbar_test(bar, parity)) return;
  SpinGuard g;
  for (;;) {
    if (ns != 0) __nanosleep(ns);
    if (mbar_test(bar, parity)) return;
    g.poll(w, tag, index, parity, iter);
  }
}
// One lane of a fully converged warp.  Unlike `lane == 0` the compiler knows that exactly one thread is active in
// the guarded region, so operands of the warp-level instructions issued there (tcgen05.mma, TMA) move to uniform
// registers directly instead of through a broadcast loop per instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address (lane 0, first column) to *slot in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem descriptor]; one thread issues for the CTA
__device__ __forceinline__ void mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// the same arrival delivered to the mbarrier at this offset in every CTA of the cluster named in cta_mask
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32-byte global load of 8 consecutive floats (one full sector per lane).  Coherent path on purpose: the
// edge kernel updates the same array in place (rows it has already consumed).
__device__ __forceinline__ void ldg_f32x8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// Descriptors
// ---------------------------------------------------------------------------------------------
// Instruction descriptor, kind::f16: D fp32, A/B fp16, both K-major, M x N.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) /* D = f32 */ | (0u << 7) /* A = f16 */ | (0u << 10) /* B = f16 */ | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor, K-major, no swizzle: 8x(16 byte) core matrices; LBO = byte step between
// core matrices along K, SBO = byte step between 8-row groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, layout type 0 = SWIZZLE_NONE
}

// ---------------------------------------------------------------------------------------------
// Tile geometry.  K = input channels (H).  A row tile is NT rows; in shared memory it is two fp16
// operand images (hi, lo), each [NT/8][K/8] core matrices of 8 rows x 8 k (128 bytes, row-contiguous).
// ---------------------------------------------------------------------------------------------
template <int K, int NT_>
struct Tile {
  static constexpr int NT = NT_;
  static constexpr int LBO = 128;                 // next core matrix along K
  static constexpr int SBO = (K / 8) * 128;       // next 8-row group
  static constexpr int IMG_BYTES = NT * K * 2;    // one operand image
  static constexpr int BUF_BYTES = 2 * IMG_BYTES; // hi + lo
  static constexpr int W_COLS = K / 2;            // TMEM columns of one weight image (2 fp16 per column)
  static constexpr int KSTEPS = K / 16;
  static_assert(K % 16 == 0 && NT % 16 == 0 && NT >= 16 && NT <= 256, "tile shape");
};

__host__ __device__ constexpr uint32_t pow2_cols(uint32_t c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// x[0..7] (fp32, already scaled) -> 8 fp16 hi + 8 fp16 lo, packed for one 16-byte core-matrix row each
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - back.x, x[2 * i + 1] - back.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Producer: rows [row0, row0 + NT) of the fp32 matrix X[rows][K] -> (hi, lo) operand images at `buf`.
// Called by NPW producer warps (warp index pw in [0, NPW)).  A warp step covers 8 rows x 32 floats:
// lane = (row % 8) + 8 * (k-chunk % 4) so that a quarter warp writes one whole 128-byte core matrix
// (conflict free) and a warp reads four full 32-byte sectors of each of 8 rows.
template <int K, int NT, int NPW>
__device__ __forceinline__ void produce_tile(const float* __restrict__ X, int64_t rows, int64_t row0, uint8_t* buf,
                                             int pw, int lane) {
  using T = Tile<K, NT>;
  constexpr int RG = NT / 8, CG = K / 32;       // row groups x chunk groups of a tile
  constexpr int STEPS = RG * CG / NPW;          // warp steps per producer warp
  static_assert((RG * CG) % NPW == 0, "producer split");
  constexpr int BATCH = STEPS < 8 ? STEPS : 8;  // loads in flight per thread (8 x 32 B)
  static_assert(STEPS % BATCH == 0, "producer batch");
  const int rin = lane & 7, cin = lane >> 3;
#pragma unroll 1
  for (int s0 = 0; s0 < STEPS; s0 += BATCH) {
    float v[BATCH][8];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int step = (s0 + b) * NPW + pw;
      const int rg = step / CG, cg = step % CG;
      const int64_t r = row0 + rg * 8 + rin;
      if (r < rows) {
        ldg_f32x8(X + r * K + cg * 32 + cin * 8, v[b]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[b][i] = 0.f;
      }
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int step = (s0 + b) * NPW + pw;
      const int rg = step / CG, cg = step % CG;
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = v[b][i] * kXScale;
      uint4 hi, lo;
      split8(x, hi, lo);
      uint8_t* dst = buf + rg * T::SBO + (cg * 4 + cin) * T::LBO + rin * 16;
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + T::IMG_BYTES) = lo;
    }
  }
}

// One thread: the three split-precision products of a tile into D (TMEM column d_col .. d_col + NT).
// tmem_w = first column of the W_hi image (W_lo follows at + W_COLS); b_addr = shared address of the tile.
template <int K, int NT>
__device__ __forceinline__ void issue_tile_mma(uint32_t tmem_w, uint32_t tmem_d, uint32_t b_addr) {
  using T = Tile<K, NT>;
  constexpr uint32_t idesc = make_idesc(kM, NT);
  uint32_t acc = 0;
#pragma unroll
  for (int term = 0; term < 3; ++term) {  // small terms first: Wlo*Xhi, Whi*Xlo, Whi*Xhi
    const uint32_t a0 = tmem_w + (term == 0 ? T::W_COLS : 0);
    const uint32_t b0 = b_addr + (term == 1 ? T::IMG_BYTES : 0);
#pragma unroll
    for (int ks = 0; ks < T::KSTEPS; ++ks) {
      mma_ts_f16(tmem_d, a0 + ks * 8, make_smem_desc(b0 + ks * 2 * T::LBO, T::LBO, T::SBO), idesc, acc);
      acc = 1;
    }
  }
}

// Load this CTA's 128 x K weight block (pre-split fp16 images, [2][128][K] halves, see gnb_pack_linear_tc)
// into TMEM columns [0, 2 * W_COLS).  Called by four warps that cover the four lane quarters.
template <int K>
__device__ __forceinline__ void load_weights_to_tmem(const __half* __restrict__ Wp_block, uint32_t tmem_base,
                                                     int quarter, int lane) {
  constexpr int W_COLS = K / 2;
  const int row = quarter * 32 + lane;
#pragma unroll 1
  for (int img = 0; img < 2; ++img) {
    const uint4* src = reinterpret_cast<const uint4*>(Wp_block + ((size_t)img * kM + row) * K);
#pragma unroll 1
    for (int c0 = 0; c0 < W_COLS; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 q = __ldg(src + c0 / 4 + i);
        v[4 * i + 0] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
      }
      tmem_st16(tmem_base + ((uint32_t)(quarter * 32) << 16) + img * W_COLS + c0, v);
    }
  }
  tmem_st_wait();
}

}  // namespace tc
}  // namespace gnb
