// TMA (cp.async.bulk.tensor) + 128-byte-swizzled operand tiles for the second-generation tensor-core kernels.
//
// State layout in HBM ("split16"): a matrix X[rows][K] of fp32 values is held as fp16 pairs, row r =
//   [ hi[r][0..K) | lo[r][0..K) ],  hi = fp16(x / 16),  lo = fp16(x / 16 - hi)     (x ~ 16 * (hi + lo), 22 significant bits)
// i.e. one row-major fp16 matrix [rows][2K]: the same 4 bytes per element as fp32 and a row is still one contiguous
// piece of memory (the reverse aggregation gathers whole rows), but the two halves ARE the tcgen05 operands: a
// tile of NT rows is brought into shared memory by the TMA engine (no conversion instructions, no registers in
// flight), one [NT rows][64 halves = 128 bytes] box per 64-channel block and half, written in the SWIZZLE_128B
// pattern the UMMA shared-memory descriptor expects, and e' / h' go back with TMA stores from the same tile.
#pragma once

#include <cuda.h>

#include "gnb_tc.cuh"

namespace gnb {
namespace tc {

constexpr int kKB = 64;                       // halves per 128-byte swizzle row (one "k block")

// Bytes of one k block of an NT-row tile, and of one whole stage (both images, K / 64 k blocks each).
template <int K, int NT>
struct Tile2 {
  static constexpr int KBLOCKS = K / kKB;
  static constexpr int KB_BYTES = NT * 128;
  static constexpr int IMG_BYTES = KBLOCKS * KB_BYTES;
  static constexpr int BUF_BYTES = 2 * IMG_BYTES;
  static constexpr int W_COLS = K / 2;
  static constexpr int KSTEPS = K / 16;
  static_assert(K % kKB == 0 && NT % 8 == 0, "tile shape");
};

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// 2-D tensor map over a split16 matrix [rows][2K] fp16 (hi columns [0, K), lo columns [K, 2K)), box =
// [box_rows][64 halves], SWIZZLE_128B; out-of-bounds rows read as zeros (loads) / are clipped (stores).
// Returns 0 or a negative GNB_E_* code.
int make_state_map(CUtensorMap* map, const void* base, int64_t rows, int K, int box_rows);

// The watchdog argument of a tensor-core kernel launch: record in host-mapped memory (allocated on first use) and the
// per-wait timeout (GNB_SPIN_TIMEOUT_MS or gnb_set_spin_timeout_ms; default 10 s, 0 = none).  gnb_tc2_common.cu.
Watch watch_get();

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// global -> shared, completion counted on an mbarrier (x = element index along K, y = row)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
// L2 eviction policy for traffic that is touched once (the e stream of the aggregation kernel): such lines are the first
// to go, so that they do not push out the node-table rows the gathers come back to.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc_hint(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar,
                                                    uint16_t cta_mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5, %6;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)), "h"(cta_mask), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, const void* smem_src, int x, int y,
                                                  uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(x), "r"(y), "l"(policy)
               : "memory");
}
// Same, delivered to the same shared-memory offset (and signalling the mbarrier at the same offset) of every CTA of
// the cluster named in cta_mask: the tile is read from L2 / HBM once for the whole cluster.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared -> global, tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of every committed store has been read (the tile may be reused)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Shared-memory matrix descriptor of a K-major, SWIZZLE_128B operand: rows are 128 bytes apart, 8-row swizzle
// atoms 1024 bytes apart (SBO); LBO is not used by swizzled K-major layouts (set to 1).  `addr` must lie in a
// 1024-byte aligned tile; advancing K by 16 halves = +32 bytes on the start address.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

// Byte offset of element (row r, channel k) inside one image of a stage laid out by tma_load_2d.
template <int NT>
__device__ __forceinline__ uint32_t sw128_offset(int r, int k) {
  const int kb = k >> 6, kk = k & 63;
  return (uint32_t)(kb * (NT * 128) + r * 128 + ((((kk >> 3) ^ (r & 7)) << 4) | ((kk & 7) << 1)));
}

// The low word of that descriptor (start address field + LBO = 1) and the descriptor of an operand `byte_off` bytes
// further into the same stage: the address field is (addr >> 4) mod 2^14 and a stage never crosses the 256 KB the
// field spans, so advancing it is ONE 32-bit add.  The MMA-issuing thread computes the base once per tile; building
// every descriptor from its byte address cost five uniform-datapath instructions per tcgen05.mma, and that thread --
// one warp sharing a scheduler with four epilogue warps -- was busy 60-70 % of the time in the aggregation kernel,
// so that the accumulators of a tile arrived ~3k cycles after its epilogue group had asked for them (profiles/r02h).
__device__ __forceinline__ uint32_t sw128_desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t sw128_desc_at(uint32_t base_lo, uint32_t byte_off) {
  constexpr uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO, version 1, SWIZZLE_128B
  return ((uint64_t)hi << 32) | (uint64_t)(base_lo + (byte_off >> 4));
}

// The three split-precision products of one tile: D (+)= Whi*Xhi + Whi*Xlo + Wlo*Xhi (small terms first).
// tmem_w: first column of the W_hi image (W_lo at + K/2); b_addr: shared address of the stage.
template <int K, int NT>
__device__ __forceinline__ void issue_tile_mma_sw128(uint32_t tmem_w, uint32_t tmem_d, uint32_t b_addr) {
  using T = Tile2<K, NT>;
  constexpr uint32_t idesc = make_idesc(kM, NT);
  const uint32_t lo0 = sw128_desc_lo(b_addr);
  uint32_t acc = 0;
#pragma unroll
  for (int term = 0; term < 3; ++term) {  // Wlo*Xhi, Whi*Xlo, Whi*Xhi
    const uint32_t a0 = tmem_w + (term == 0 ? T::W_COLS : 0);
#pragma unroll
    for (int ks = 0; ks < T::KSTEPS; ++ks) {
      const uint32_t off = (term == 1 ? T::IMG_BYTES : 0) + (ks >> 2) * T::KB_BYTES + (ks & 3) * 32;
      mma_ts_f16(tmem_d, a0 + ks * 8, sw128_desc_at(lo0, off), idesc, acc);
      acc = 1;
    }
  }
}

// x (fp32) -> the two fp16 halves of the split16 format
__device__ __forceinline__ void split1(float x, __half& hi, __half& lo) {
  const float xs = x * kXScale;
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));
}
__device__ __forceinline__ float merge1(__half hi, __half lo) {
  return (__half2float(hi) + __half2float(lo)) * kWScale;   // kWScale == 1 / kXScale == 16
}

}  // namespace tc
}  // namespace gnb
