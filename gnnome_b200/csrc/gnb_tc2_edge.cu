// Fused edge pass of one (Sym)GatedGCN layer: TMA-fed tcgen05 with the edge state held in the split16 format
// (gnb_tma.cuh).  Reference layers/gated_gcn_full.py:97,104-114:
//   z_p  = B1h[src_p] + B2h[dst_p] + e_p * W_B3^T         (E x H x H product on tcgen05, fp16 hi/lo split, fp32 in TMEM)
//   e'_p = relu(z_p * scale + shift) (+ e_p)               written over e in place
//   F_i  = sum_{p: dst_p = i} sigmoid(e'_p) * A2h[src_p] / (sum_p sigmoid(e'_p) + 1e-6)
// The eval-mode norm affine is folded into the operands by the caller: the rows of W_B3 and B_1 are multiplied by
// scale, B2h' = scale * B2h + shift, so that the epilogue computes  relu(D + B1h'[src] + B2h'[dst]) / 16 + e/16  in
// the scaled domain the split16 state is stored in.  The residual e/16 = hi + lo is read from TENSOR MEMORY too: a
// second, identity product (A = I from shared memory, B = the stage's own k blocks of this CTA's channels) puts it
// there in the accumulator layout, exactly (one product per output), instead of two 16-bit shared loads, two
// conversions and two floating-point operations per element.
//
// Persistent CTAs, one per SM; a CTA owns HC = min(H, 128) output channels (H = 256: the two channel halves of a
// tile run on the two CTAs of a cluster that share the tile through TMA multicast) with its W_B3 block resident in
// TMEM, and walks 32-edge tiles of the dst-sorted edge array.  Tile i of a CTA belongs to epilogue group i % 4.
//
//   warp 0      : producer.  TMA loads of the tile's operand images into input stage i % NB (the loads run NB tiles
//                 ahead of the tensor core), the tile's (src, dst) indices into the group's index slot; optionally
//                 (EdgeMode, off by default) an L2 prefetch of the node rows the epilogue will gather.
//   warps 1, 3  : MMA issue (one lane each; warp 1 the even tiles of the CTA, warp 3 the odd ones): 3 K/16 split products
//                 into the group's accumulator set + 2 HC/16 identity products for the residual, descriptors advanced
//                 from one base per tile; the commit RELEASES THE INPUT STAGE -- nothing but the tensor core
//                 reads a stage, so the load pipeline is as deep as the stage ring (with the in-place stage of
//                 round 1 / 2a, held from load to store, four of five stages sat under the epilogue groups and a
//                 third of the epilogue's time was spent waiting for the next tile: profiles/r02b).
//   warps 4..19 : epilogue, 4 groups x 4 TMEM lane quarters.  A thread owns ONE channel and walks the tile's 32
//                 edges: gathers of the (B1h', A2h) node rows are coalesced across the warp, per-destination sums
//                 are register accumulators closed at warp-uniform segment boundaries (no atomics, fixed summation
//                 order).  Segments that straddle a tile leave partial sums in carry[tile][4][H], resolved by
//                 gnb_node_update2.  e' goes back through TENSOR MEMORY: the thread writes its fp32 values over the
//                 residual columns of the accumulator set (its own lane), and after the tile the warp reads its 32
//                 lanes back in the matrix-fragment layout (tcgen05.ld.16x256b: a thread then holds pairs of
//                 consecutive edges of one channel), splits the pairs into fp16 (hi, lo) with packed conversions and
//                 stores 8 x 8 blocks TRANSPOSED into the group's OUTPUT BUFFER with stmatrix -- 16-byte rows of 8
//                 channels in the 128-byte swizzle the TMA store expects.
//   warp 2      : TMA stores, lane 0 for all four groups, polling their hand-over barriers with test_wait.
//
// Barriers.  Every barrier has ONE producer side and ONE consumer that sees each of its phases in sequence, and every
// ring has back-pressure, so a parity wait can neither be satisfied by the phase before last nor miss a phase (the
// round-1 dead-lock was a group completing two phases of an un-back-pressured hand-over barrier; profiles/r02a):
//   full[NB]      TMA bytes                       -> MMA warp
//   empty[NB]     MMA commit (both CTAs in a cluster: either multicast writes into the stage) -> producer
//   ifull[4][2]   producer (indices published)    -> the group's warps; a slot is rewritten 8 tiles later, which the
//                 chain  load(i+8) <- MMA(i+8-NB) issued after MMA(i+4) <- dempty: group done with tile i  allows (NB <= 4)
//   dfull[4]      MMA commit                      -> the group's warps
//   dempty[4]     the group's warps (last TMEM read of the tile) -> MMA warp
//   ofull[4]      the group's warps (e' in the output buffer)    -> store thread
//   oempty[4]     store thread (buffer read by the TMA engine)    -> the group's warps
// Every wait is bounded by the spin watchdog (gnb_tc.cuh): a lost phase traps with a record instead of spinning.
#include <cstdlib>
#include <type_traits>

#include "gnb_tma.cuh"

namespace gnb {
namespace tc {

constexpr int kE2NT = 32;        // edges per tile (MMA N) = edges per epilogue warp = carry granularity
constexpr int kE2Chunk = kE2NT;
constexpr int kE2Groups = 4;     // epilogue groups of four warps (one per TMEM lane quarter); group g takes tiles g, g+G, ...
                                 // (five groups were measured slower, r01h)
constexpr int kE2DCols = 2 * kE2NT;   // TMEM columns of one accumulator set: z (32 edges) | residual e/16, then e'/16
constexpr int kE2FirstEpiWarp = 4;
constexpr int kE2Threads = 32 * (kE2FirstEpiWarp + 4 * kE2Groups);
constexpr int kE2IdxInts = 2 * kE2NT + 4;   // src[32], dst[32], prev_dst, next_dst (+ pad)
constexpr int kE2IdxSlots = 2;              // index slots per group (see ifull above)
// L2 management of the launch (gnb_debug_edge_mode; the default is what measured best, profiles/r02i):
enum EdgeMode : int {
  kEmLoadEvictFirst = 1,    // TMA loads of the e tiles carry an evict-first policy
  kEmStoreEvictFirst = 2,   // TMA stores of the e' tiles carry an evict-first policy
  kEmPrefetchLate = 4,      // the node rows of a tile are prefetched into L2 when its stage is free, not before
  kEmNoPrefetch = 8,        // no L2 prefetch of the node rows
  kEmDefault = kEmNoPrefetch,
};

template <int H>
struct Edge2Cfg {
  static constexpr int HC = H < kM ? H : kM;   // live channels per CTA
  static constexpr int NH = H / HC;            // channel halves (CTAs per tile)
  // H = 256: the two channel halves of a tile form a 2-CTA cluster; each loads half of the tile's boxes and
  // multicasts them to both, so the e tile is read from L2 / HBM once.  Both read whole rows of e and each overwrites
  // its own channels in global memory; no flag is needed: a CTA stores tile t only after its own epilogue, i.e. after
  // its own stage of tile t was complete, which means every box of the tile had been read from global memory on
  // behalf of both CTAs.
  static constexpr bool MC = NH == 2;
  using T = Tile2<H, kE2NT>;
  static constexpr int NB = 4;                           // input stages (NB <= 4: see ifull)
  static_assert(NB <= 4, "index slot reuse distance");
  // the identity block [128 rows (TMEM lanes)][HC input channels] as a K-major SWIZZLE_128B A operand
  static constexpr int ID_KB_BYTES = kM * 128;
  static constexpr int ID_BYTES = (HC / kKB) * ID_KB_BYTES;
  // output buffer of a group: this CTA's channels of a tile, (hi | lo) x k blocks x 32 rows x 128 bytes
  static constexpr int OKB = HC / kKB;
  static constexpr int OIMG_BYTES = OKB * T::KB_BYTES;
  static constexpr int OBUF_BYTES = 2 * OIMG_BYTES;
  // epilogue warps of a group that own live channels (the others idle: they must not feed the barriers)
  static constexpr int LIVE_WARPS = HC / 32;
  // accumulator sets in TMEM: one per epilogue group (H = 256: the weight images take half of the 512 columns)
  static constexpr uint32_t TMEM_COLS = pow2_cols(2 * T::W_COLS + kE2Groups * kE2DCols);
  static constexpr uint32_t D_COL0 = 2 * T::W_COLS;
  static_assert(2 * T::W_COLS + kE2Groups * kE2DCols <= 512, "tensor memory budget");
  static constexpr int NBARS = 2 * NB + kE2Groups * (kE2IdxSlots + 4);
  // no alignment slack: the dynamic shared memory window starts 1024-byte aligned (the kernel traps if it does not)
  static constexpr size_t SMEM = (size_t)NB * T::BUF_BYTES + ID_BYTES + (size_t)kE2Groups * OBUF_BYTES +
                                 (size_t)kE2Groups * kE2IdxSlots * kE2IdxInts * 4 + NBARS * 8 + 16;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// D[tmem] (+)= A[smem descriptor] * B[smem descriptor]; one thread issues for the CTA
__device__ __forceinline__ void mma_ss_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// The residual product: D[c][edge] = sum_k I[c][k] * (Xhi + Xlo)[edge][k] over the KI input channels this CTA owns
// (b_addr = the stage's first k block of those channels; the lo image follows img_bytes later).
template <int KI, int NT>
__device__ __forceinline__ void issue_ident_mma_sw128(uint32_t tmem_d, uint32_t ident_addr, uint32_t b_addr,
                                                      uint32_t img_bytes, uint32_t kb_bytes) {
  constexpr uint32_t idesc = make_idesc(kM, NT);
  const uint32_t a_lo = sw128_desc_lo(ident_addr), b_lo = sw128_desc_lo(b_addr);
  uint32_t acc = 0;
#pragma unroll
  for (int img = 0; img < 2; ++img) {
#pragma unroll
    for (int ks = 0; ks < KI / 16; ++ks) {
      mma_ss_f16(tmem_d, sw128_desc_at(a_lo, (ks >> 2) * (kM * 128) + (ks & 3) * 32),
                 sw128_desc_at(b_lo, img * img_bytes + (ks >> 2) * kb_bytes + (ks & 3) * 32), idesc, acc);
      acc = 1;
    }
  }
}

// sigmoid(16 x): the state is held in the scaled domain x = e / 16
__device__ __forceinline__ float sigmoid16f_fast(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-16.0f * 1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
  return r;
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
               : "memory");
}
// 16 TMEM lanes x 32 columns in the matrix-fragment layout (measured, tools/microbench/tmem_probe.cu): register
// r[4c + 2a + b] of thread T holds (lane0 + T / 4 + 8a, column 8c + 2 (T % 4) + b)
__device__ __forceinline__ void tmem_ld_frag16x32(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// four 8 x 8 b16 matrices, transposed on the way: memory[m][i][j] = half (i % 2) of register m of thread 4j + i / 2;
// thread T supplies the address of row T % 8 of matrix T / 8 (16 bytes)
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r0), "r"(r1),
               "r"(r2), "r"(r3)
               : "memory");
}

template <int H, bool kResidual, bool kTiming>
__global__ void __launch_bounds__(kE2Threads, 1)
edge_forward_tc2_kernel(const __grid_constant__ CUtensorMap map_e, gnb_graph_t g, const float* __restrict__ P, int64_t ldP, const __half* __restrict__ Wp,
                        float* __restrict__ F, float* __restrict__ carry, int flags, int workers,
                        unsigned long long* timing, const Watch watch, int store_delay_ns, int mode) {
  using C = Edge2Cfg<H>;
  using T = typename C::T;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (smem_u32(smem_raw) & 1023u) __trap();   // SWIZZLE_128B tiles need 1024-byte alignment; the budget has no slack for it
  uint8_t* bufs = smem_raw;                                                        // [NB] input stages
  uint8_t* ident = bufs + (size_t)C::NB * T::BUF_BYTES;
  uint8_t* obufs = ident + C::ID_BYTES;                                            // [G] output buffers
  int* idx_area = reinterpret_cast<int*>(obufs + (size_t)kE2Groups * C::OBUF_BYTES);   // [G][2] index slots
  uint64_t* bars = reinterpret_cast<uint64_t*>(idx_area + kE2Groups * kE2IdxSlots * kE2IdxInts);
  uint64_t* full = bars;                                   // [NB]
  uint64_t* empty = full + C::NB;                          // [NB]
  uint64_t* ifull = empty + C::NB;                         // [G][2]
  uint64_t* dfull = ifull + kE2Groups * kE2IdxSlots;       // [G]
  uint64_t* dempty = dfull + kE2Groups;                    // [G]
  uint64_t* ofull = dempty + kE2Groups;                    // [G]
  uint64_t* oempty = ofull + kE2Groups;                    // [G]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(oempty + kE2Groups);

  const int half = blockIdx.x % C::NH, worker = blockIdx.x / C::NH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t E = g.num_edges;
  const int64_t num_tiles = (E + kE2NT - 1) / kE2NT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NB; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], C::MC ? 2 : 1);   // multicast: both CTAs of the cluster write into a stage
    }
    for (int i = 0; i < kE2Groups * kE2IdxSlots; ++i) mbar_init(&ifull[i], 32);
    for (int i = 0; i < kE2Groups; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], C::LIVE_WARPS);
      mbar_init(&ofull[i], C::LIVE_WARPS);
      mbar_init(&oempty[i], 1);
    }
    fence_barrier_init();
    prefetch_tensormap(&map_e);
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if (C::MC) cluster_sync_all();   // the peer's multicast / commit must find initialised barriers
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= kE2FirstEpiWarp && warp < kE2FirstEpiWarp + 4)
    load_weights_to_tmem<H>(Wp + (size_t)half * 2 * kM * H, tmem_base, warp & 3, lane);
  if (kResidual && threadIdx.x < kM) {
    // row r of the identity block: 1.0 at input channel r (if this CTA has that many), in the 128-byte swizzle
    const int r = threadIdx.x;
#pragma unroll
    for (int kb = 0; kb < C::HC / kKB; ++kb) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (r < C::HC && (r >> 6) == kb && ((r & 63) >> 3) == j) w[(r & 7) >> 1] = (r & 1) ? 0x3C000000u : 0x00003C00u;
        *reinterpret_cast<uint4*>(ident + kb * C::ID_KB_BYTES + r * 128 + ((j ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: TMA + indices
    auto load_idx = [&](int64_t t, int (&r)[3]) {
      const int64_t p0 = t * kE2NT + lane;
      r[0] = (p0 < E) ? g.in_src[p0] : 0;
      r[1] = (p0 < E) ? g.in_dst[p0] : 0;    // rows past a ragged end: any valid node (every use is guarded by n)
      r[2] = -1;
      if (lane == 0 && t > 0) r[2] = g.in_dst[t * kE2NT - 1];
      if (lane == 1 && (t + 1) * kE2NT < E) r[2] = g.in_dst[(t + 1) * kE2NT];
    };
    // Every node row is touched for the first time by SOME gather of the epilogue, and that one waits for HBM; the
    // producer knows the tile's endpoints NB tile periods before the epilogue needs them and can pull this CTA's slices
    // of the (B1h, A2h)[src] and B2h[dst] rows into L2 ahead of time.  OFF by default (kEmNoPrefetch): at config 3 a
    // line lives ~12 us in the L2, about the distance between this prefetch and its use, so rows were read from DRAM
    // twice (150 GB per launch instead of 94 GB, 45.2 ms instead of 40.2 ms: profiles/r02i); the three quads of
    // gathers each epilogue warp keeps in flight cover the first-touch latency instead.
    auto prefetch_rows = [&](const int (&r)[3]) {
      const char* a = reinterpret_cast<const char*>(P + (int64_t)r[0] * ldP + 2 * half * C::HC);
#pragma unroll
      for (int l = 0; l < C::HC * 8 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + l * 128));
      const char* b = reinterpret_cast<const char*>(P + (int64_t)r[1] * ldP + 2 * H + half * C::HC);
#pragma unroll
      for (int l = 0; l < C::HC * 4 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + l * 128));
    };
    // The endpoints are read kIdxAhead tiles ahead of their use.  The ring is indexed STATICALLY (the tile loop is
    // unrolled by its length): a ring that is shifted by register moves makes every iteration wait for the load
    // issued in the iteration before it -- one DRAM round trip per tile, about a tile period, so the producer never
    // got ahead of the tensor core, the stages ran empty and the epilogue groups waited for their indices 10 % of
    // the time (profiles/r02g: the producer warp spent 60 % of its samples on these loads and 4 % on `empty`).
    constexpr int kIdxAhead = 4;
    static_assert(kIdxAhead % C::NB == 0 && kIdxAhead % kE2Groups == 0, "stage / group of ring entry k are static");
    int ring[kIdxAhead][3];
#pragma unroll
    for (int k = 0; k < kIdxAhead; ++k) {
      ring[k][0] = ring[k][1] = 0;
      ring[k][2] = -1;
      if (worker + (int64_t)k * workers < num_tiles) load_idx(worker + (int64_t)k * workers, ring[k]);
    }
    const int64_t ahead = (int64_t)kIdxAhead * workers;
    const uint64_t l2_stream = l2_policy_evict_first();
    bool more = worker < num_tiles;
    unsigned long long ptm[2] = {0, 0};   // kTiming: cycles waiting for an empty stage, tiles
    for (int i0 = 0; more; i0 += kIdxAhead) {
#pragma unroll
      for (int k = 0; k < kIdxAhead; ++k) {
        const int i = i0 + k;
        const int64_t t = worker + (int64_t)i * workers;
        if (t >= num_tiles) { more = false; break; }
        const int s = k % C::NB;
        if (!(mode & (kEmPrefetchLate | kEmNoPrefetch))) prefetch_rows(ring[k]);
        const long long p0 = kTiming ? clock64() : 0;
        mbar_wait(&empty[s], ((i / C::NB) & 1) ^ 1, 64, watch, watch_tag(kWkEdge2, kWrProducer, kWbEmpty), s, i);
        if (kTiming) { ptm[0] += clock64() - p0; ptm[1] += 1; }
        if ((mode & kEmPrefetchLate) && !(mode & kEmNoPrefetch)) prefetch_rows(ring[k]);
        if (elect_one()) {
          uint8_t* stage = bufs + (size_t)s * T::BUF_BYTES;
          mbar_arrive_expect_tx(&full[s], T::BUF_BYTES);
          if (mode & kEmLoadEvictFirst) {
#pragma unroll
            for (int kb = 0; kb < T::KBLOCKS; ++kb) {
              if (C::MC) {
                tma_load_2d_mc_hint(stage + half * T::IMG_BYTES + kb * T::KB_BYTES, &map_e, half * H + kb * kKB,
                                    (int)(t * kE2NT), &full[s], (uint16_t)3, l2_stream);
              } else {
                tma_load_2d_hint(stage + kb * T::KB_BYTES, &map_e, kb * kKB, (int)(t * kE2NT), &full[s], l2_stream);
                tma_load_2d_hint(stage + T::IMG_BYTES + kb * T::KB_BYTES, &map_e, H + kb * kKB, (int)(t * kE2NT), &full[s],
                                 l2_stream);
              }
            }
          } else {
#pragma unroll
            for (int kb = 0; kb < T::KBLOCKS; ++kb) {
              if (C::MC) {   // rank 0 brings the hi halves, rank 1 the lo halves, for both CTAs
                tma_load_2d_mc(stage + half * T::IMG_BYTES + kb * T::KB_BYTES, &map_e, half * H + kb * kKB, (int)(t * kE2NT),
                               &full[s], (uint16_t)3);
              } else {
                tma_load_2d(stage + kb * T::KB_BYTES, &map_e, kb * kKB, (int)(t * kE2NT), &full[s]);
                tma_load_2d(stage + T::IMG_BYTES + kb * T::KB_BYTES, &map_e, H + kb * kKB, (int)(t * kE2NT), &full[s]);
              }
            }
          }
        }
        const int grp = k % kE2Groups, slot = (i / kE2Groups) % kE2IdxSlots;
        int* ia = idx_area + (grp * kE2IdxSlots + slot) * kE2IdxInts;
        ia[lane] = ring[k][0];
        ia[kE2NT + lane] = ring[k][1];
        if (lane < 2) ia[2 * kE2NT + lane] = ring[k][2];
        mbar_arrive(&ifull[grp * kE2IdxSlots + slot]);
        if (t + ahead < num_tiles) load_idx(t + ahead, ring[k]);
      }
    }
    if (kTiming && timing != nullptr && lane == 0) {
      timing[((size_t)blockIdx.x * 32 + warp) * 5 + 0] = ptm[0];
      timing[((size_t)blockIdx.x * 32 + warp) * 5 + 4] = ptm[1];
    }
  } else if (warp == 1 || warp == 3) {
    // ---------------------------------------------------------------- MMA issue
    // Two issuing warps on two schedulers, warp 1 for the even tiles of the CTA and warp 3 for the odd ones (stage
    // and accumulator set follow the tile index, so each barrier keeps its one consumer).  One warp was busy 2.4k of
    // the 3.2k cycles of a tile period just ISSUING (64 tcgen05.mma + their descriptors at one instruction every ~9
    // cycles next to four epilogue warps on the same scheduler), and the accumulators of a tile arrived 1.5-3k cycles
    // after its group had asked for them (profiles/r02h).
    int i = warp >> 1;
    unsigned long long mtm[4] = {0, 0, 0, 0};   // kTiming: cycles waiting for the stage, for the accumulator set, issuing; tiles
    for (int64_t t = worker + (int64_t)i * workers; t < num_tiles; t += 2 * (int64_t)workers, i += 2) {
      const int s = i % C::NB, d = i % kE2Groups;
      const long long m0 = kTiming ? clock64() : 0;
      mbar_wait(&full[s], (i / C::NB) & 1, 32, watch, watch_tag(kWkEdge2, kWrMma, kWbFull), s, i);
      const long long m1 = kTiming ? clock64() : 0;
      mbar_wait(&dempty[d], ((i / kE2Groups) & 1) ^ 1, 32, watch, watch_tag(kWkEdge2, kWrMma, kWbDEmpty), d, i);
      const long long m2 = kTiming ? clock64() : 0;
      tc_fence_after();
      if (elect_one()) {
        const uint32_t stage = smem_u32(bufs + (size_t)s * T::BUF_BYTES);
        issue_tile_mma_sw128<H, kE2NT>(tmem_base, tmem_base + C::D_COL0 + d * kE2DCols, stage);
        if (kResidual)
          issue_ident_mma_sw128<C::HC, kE2NT>(tmem_base + C::D_COL0 + d * kE2DCols + kE2NT, smem_u32(ident),
                                              stage + half * (C::HC / kKB) * T::KB_BYTES, T::IMG_BYTES, T::KB_BYTES);
        // the products have read the stage: hand it back to the producer(s) ...
        if (C::MC) mma_commit_mc(&empty[s], (uint16_t)3);
        else mma_commit(&empty[s]);
        mma_commit(&dfull[d]);   // ... and the accumulators to the group
      }
      __syncwarp();
      if (kTiming) { mtm[0] += m1 - m0; mtm[1] += m2 - m1; mtm[2] += clock64() - m2; mtm[3] += 1; }
    }
    if (kTiming && timing != nullptr && lane == 0) {
      timing[((size_t)blockIdx.x * 32 + warp) * 5 + 0] = mtm[0];
      timing[((size_t)blockIdx.x * 32 + warp) * 5 + 1] = mtm[1];
      timing[((size_t)blockIdx.x * 32 + warp) * 5 + 2] = mtm[2];
      timing[((size_t)blockIdx.x * 32 + warp) * 5 + 4] = mtm[3];
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------------- TMA stores
    // Lane 0 stores the tiles of all four groups.  It POLLS them with test_wait, which returns at once (try_wait would
    // park the thread on one group's barrier for microseconds while another group's tile is ready; their tiles finish
    // out of order).  A group cannot refill its output buffer before this thread has released it (oempty), so ofull
    // never runs ahead of its one consumer.
    if (lane == 0) {
      constexpr int kPer = kE2Groups;
      const int g0 = 0;
      const uint64_t l2_stream = l2_policy_evict_first();
      int it[kPer];                                        // next tile iteration of each served group
#pragma unroll
      for (int k = 0; k < kPer; ++k) it[k] = g0 + k;
      auto remaining = [&](int k) { return worker + (int64_t)it[k] * workers < num_tiles; };
      auto any_remaining = [&]() {
        bool r = false;
#pragma unroll
        for (int k = 0; k < kPer; ++k) r = r || remaining(k);
        return r;
      };
      SpinGuard guard;
      while (any_remaining()) {
        bool progressed = false;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          if (!remaining(k)) continue;
          const int i = it[k], grp = g0 + k;
          if (!mbar_poll(&ofull[grp], (i / kE2Groups) & 1)) continue;
          const int64_t t = worker + (int64_t)i * workers;
          const uint8_t* ob = obufs + (size_t)grp * C::OBUF_BYTES;
          if (mode & kEmStoreEvictFirst) {
#pragma unroll
            for (int kbl = 0; kbl < C::OKB; ++kbl) {
              const int kb = half * C::OKB + kbl;
              tma_store_2d_hint(&map_e, ob + kbl * T::KB_BYTES, kb * kKB, (int)(t * kE2NT), l2_stream);
              tma_store_2d_hint(&map_e, ob + C::OIMG_BYTES + kbl * T::KB_BYTES, H + kb * kKB, (int)(t * kE2NT), l2_stream);
            }
          } else {
#pragma unroll
            for (int kbl = 0; kbl < C::OKB; ++kbl) {
              const int kb = half * C::OKB + kbl;
              tma_store_2d(&map_e, ob + kbl * T::KB_BYTES, kb * kKB, (int)(t * kE2NT));
              tma_store_2d(&map_e, ob + C::OIMG_BYTES + kbl * T::KB_BYTES, H + kb * kKB, (int)(t * kE2NT));
            }
          }
          tma_store_commit();
          tma_store_wait_read();
          if (store_delay_ns > 0) __nanosleep((unsigned)store_delay_ns);   // fault injection (gnb_debug_store_delay_ns)
          if (store_delay_ns >= 0) mbar_arrive(&oempty[grp]);   // < 0: never released -- the watchdog test's dead-lock
          it[k] += kE2Groups;
          progressed = true;
        }
        if (progressed) {
          guard = SpinGuard();
        } else {
          __nanosleep(64);
          guard.poll(watch, watch_tag(kWkEdge2, kWrStore, kWbOFull), (uint32_t)g0, (uint32_t)((it[0] / kE2Groups) & 1), it[0]);
        }
      }
      tma_store_wait_all();
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - kE2FirstEpiWarp;
    const int grp = ew >> 2, q = warp & 3;
    const int cl = q * 32 + lane;            // TMEM lane = channel within the CTA's block
    const bool ch_ok = cl < C::HC;           // warp-uniform (HC is a multiple of 32)
    const int c = half * C::HC + (ch_ok ? cl : 0);
    const int64_t my_tiles = ch_ok ? num_tiles : 0;   // warps without live channels (H = 64) sit the loop out
    const char* Pc = reinterpret_cast<const char*>(P + 2 * c);        // (B1h'[c], A2h[c]) interleaved
    const char* Pb2 = reinterpret_cast<const char*>(P + 2 * H + c);   // B2h'[c]
    const int ldPb = (int)(ldP * (int64_t)sizeof(float));            // row pitch in bytes (< 2^31, checked by the host)
    constexpr unsigned kFull = 0xffffffffu;
    // The warp's 32 channels q * 32 .. + 31 of the CTA's block inside an output image: element (row, ch) sits at
    //   (ch / 64) * KB_BYTES + row * 128 + ((((ch % 64) / 8) ^ (row % 8)) * 16) + (ch % 8) * 2
    const uint32_t ob_row = smem_u32(obufs + (size_t)grp * C::OBUF_BYTES) + (uint32_t)(((q * 32) >> 6) * T::KB_BYTES) +
                            (uint32_t)lane * 128u;
    const int cx0 = ((q * 32) & 63) >> 3;    // first of the warp's four 16-byte chunks of a row (0 or 4)
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::D_COL0 + grp * kE2DCols;
    // optional cycle accounting (gnb_debug_edge_timing): [index wait, accumulator wait, batches, pack + hand-off, tiles]
    unsigned long long tm[5] = {0, 0, 0, 0, 0};
    int jj = 0;                              // tiles of this group so far
    for (int64_t t = worker + (int64_t)grp * workers; t < my_tiles; t += (int64_t)kE2Groups * workers, ++jj) {
      const int i = grp + kE2Groups * jj;
      const int64_t cs = t * kE2NT;
      const int n = (int)((E - cs < kE2Chunk) ? (E - cs) : kE2Chunk);
      const long long t0 = kTiming ? clock64() : 0;
      const int slot = jj % kE2IdxSlots;
      mbar_wait(&ifull[grp * kE2IdxSlots + slot], (jj / kE2IdxSlots) & 1, 32, watch,
                watch_tag(kWkEdge2, kWrEpilogue, kWbIFull), grp * kE2IdxSlots + slot, i);
      const long long t1 = kTiming ? clock64() : 0;
      const int* ia = idx_area + (grp * kE2IdxSlots + slot) * kE2IdxInts;
      const int my_dst = ia[kE2NT + lane];
      const int prev_dst = ia[2 * kE2NT];
      const int next_dst = ia[2 * kE2NT + 1];
      int head_dst = -1, tail_dst = -1;
      const int up = __shfl_up_sync(kFull, my_dst, 1);
      // bit j: edge j of the chunk opens a new destination segment
      const unsigned segmask = __ballot_sync(kFull, lane == 0 || (lane < n && my_dst != up));
      const int first_dst = __shfl_sync(kFull, my_dst, 0);
      const int last_dst = __shfl_sync(kFull, my_dst, n - 1);
      if (prev_dst == first_dst) head_dst = first_dst;
      if (next_dst == last_dst) tail_dst = last_dst;

      const int64_t chunk = cs / kE2Chunk;
      int cur = -1;
      float num = 0.f, den = 0.f;

      // Software pipeline over the tile's eight quads of edges with three quads of gathers in flight: xa = (B1h', A2h)[src]
      // (one coalesced 8-byte row segment per warp and edge) and xb = B2h'[dst] of quad k + 3 are requested before quad
      // k is computed, into four statically indexed register buffers (the q0 loop is unrolled by the ring length).
      // With two batches of eight (r02c) a gather had one batch of arithmetic to arrive in and 10 % of all warp
      // samples sat on its first use (profiles/r02g).
      constexpr int kQ = 4, kNQ = kE2Chunk / kQ, kRing = 4;
      const int* ia_src = ia;                            // endpoints of the chunk's 32 edges: warp-uniform reads
      const int* ia_dst = ia + kE2NT;
      auto fetch = [&](int qd, float2 (&xa)[kQ], float (&xb)[kQ]) {
        int sj[kQ], dj[kQ];
        *reinterpret_cast<int4*>(&sj[0]) = *reinterpret_cast<const int4*>(ia_src + qd * kQ);
        *reinterpret_cast<int4*>(&dj[0]) = *reinterpret_cast<const int4*>(ia_dst + qd * kQ);
#pragma unroll
        for (int u = 0; u < kQ; ++u) xa[u] = __ldg(reinterpret_cast<const float2*>(Pc + (int64_t)sj[u] * ldPb));
        // B2h'[dst]: the four edges of a quad share it unless a destination segment opens at its 2nd..4th edge.  The
        // shared value stays in xb[0] and compute() selects it: copying it into xb[1..3] here would make this warp
        // wait for the load inside the fetch, i.e. before the arithmetic that is meant to cover its latency.
        xb[0] = __ldg(reinterpret_cast<const float*>(Pb2 + (int64_t)dj[0] * ldPb));
        if ((segmask >> (qd * kQ)) & 0xeu) {   // warp-uniform
#pragma unroll
          for (int w = 1; w < kQ; ++w) xb[w] = __ldg(reinterpret_cast<const float*>(Pb2 + (int64_t)dj[w] * ldPb));
        }
      };
      auto compute = [&](auto full_tag, int qd, const float2 (&xa)[kQ], const float (&xb)[kQ]) {
        constexpr bool kFullChunk = decltype(full_tag)::value;   // all 32 edges exist: no per-edge validity tests
        uint32_t zr[4], er[4] = {0u, 0u, 0u, 0u};
        tmem_ld4(taddr + qd * kQ, zr);
        if (kResidual) tmem_ld4(taddr + kE2NT + qd * kQ, er);   // e / 16 = hi + lo, exact
        tmem_ld_wait();
        const unsigned mb = (segmask >> (qd * kQ)) & 0xfu;
        const bool shared_b2 = (mb & 0xeu) == 0;                 // warp-uniform
        float v[4], sg[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float b2 = (w == 0 || shared_b2) ? xb[0] : xb[w];
          // scaled domain: v = e' / 16 (the norm affine is folded into D, B1h' and B2h' by the caller)
          v[w] = fmaf(fmaxf(__uint_as_float(zr[w]) + xa[w].x + b2, 0.f), kXScale, __uint_as_float(er[w]));
          sg[w] = (kFullChunk || qd * kQ + w < n) ? sigmoid16f_fast(v[w]) : 0.f;
        }
        tmem_st4(taddr + kE2NT + qd * kQ, v);   // e' / 16 over the residual it was computed from (same lane)
        if (mb == 0) {                 // warp-uniform: the four edges continue the running segment
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            num = fmaf(sg[w], xa[w].y, num);
            den += sg[w];
          }
        } else {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            if (mb & (1u << w)) {      // warp-uniform: close the running segment, open the next
              if (cur >= 0) {
                if (cur == head_dst) {
                  carry[(chunk * 4 + 0) * H + c] = num;
                  carry[(chunk * 4 + 1) * H + c] = den;
                } else {
                  F[(int64_t)cur * H + c] = gate_div(num, den);
                }
              }
              cur = ia_dst[qd * kQ + w];
              num = 0.f;
              den = 0.f;
            }
            num = fmaf(sg[w], xa[w].y, num);
            den += sg[w];
          }
        }
      };
      float2 fa[kRing][kQ];
      float fb[kRing][kQ];
#pragma unroll
      for (int k = 0; k < kRing; ++k)
#pragma unroll
        for (int w = 0; w < kQ; ++w) fb[k][w] = 0.f;   // xb[1..3] of a quad with one destination are never loaded
#pragma unroll
      for (int k = 0; k + 1 < kRing; ++k) fetch(k, fa[k], fb[k]);   // in flight while the tile's products complete
      const long long t2 = kTiming ? clock64() : 0;
      mbar_wait(&dfull[grp], jj & 1, 32, watch, watch_tag(kWkEdge2, kWrEpilogue, kWbDFull), grp, i);
      tc_fence_after();
      const long long t3 = kTiming ? clock64() : 0;
      auto run_chunk = [&](auto full_tag) {
#pragma unroll 1
        for (int q0 = 0; q0 < kNQ; q0 += kRing) {
#pragma unroll
          for (int k = 0; k < kRing; ++k) {
            if (q0 + k + kRing - 1 < kNQ) fetch(q0 + k + kRing - 1, fa[(k + kRing - 1) % kRing], fb[(k + kRing - 1) % kRing]);
            compute(full_tag, q0 + k, fa[k], fb[k]);
          }
        }
      };
      if (n == kE2Chunk) run_chunk(std::true_type{});
      else run_chunk(std::false_type{});
      const long long t4 = kTiming ? clock64() : 0;
      // close the segment that is still open at the end of the chunk
      if (cur == tail_dst) {
        carry[(chunk * 4 + 2) * H + c] = num;
        carry[(chunk * 4 + 3) * H + c] = den;
      } else if (cur == head_dst) {
        carry[(chunk * 4 + 0) * H + c] = num;
        carry[(chunk * 4 + 1) * H + c] = den;
      } else {
        F[(int64_t)cur * H + c] = gate_div(num, den);
      }
      // e' (fp32, scaled) sits in this warp's 32 TMEM lanes x 32 columns: read it back as fragments, split, and store
      // transposed into the group's output buffer: edge row T, channels 16 hl + 8 a + [0, 8) = one 16-byte swizzle chunk.
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      // Both 16-lane halves are read before anything else happens, and that is the LAST read of this accumulator
      // set: it goes back to the MMA warp here, so that the products of the group's next tile (about 1.1k cycles the
      // group cannot hide: tensor memory holds one set per group) run under the conversion and the hand-off below.
      uint32_t fr[2][16];
      tmem_ld_frag16x32(taddr + kE2NT, fr[0]);
      tmem_ld_frag16x32(taddr + kE2NT + ((uint32_t)16 << 16), fr[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[grp]);
      // the buffer still holds the group's previous tile until the store thread has handed it back
      mbar_wait(&oempty[grp], (jj & 1) ^ 1, 32, watch, watch_tag(kWkEdge2, kWrEpilogue, kWbOEmpty), grp, i);
#pragma unroll
      for (int hl = 0; hl < 2; ++hl) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          uint32_t fh[4], fl[4];
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {   // edge block cb: edges 8 cb + 2 (T % 4), + 1 of channel T / 4 + 8 a
            const float x0 = __uint_as_float(fr[hl][4 * cb + 2 * a]), x1 = __uint_as_float(fr[hl][4 * cb + 2 * a + 1]);
            const __half2 hh = __floats2half2_rn(x0, x1);
            const float2 back = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(x0 - back.x, x1 - back.y);
            fh[cb] = *reinterpret_cast<const uint32_t*>(&hh);
            fl[cb] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          const uint32_t addr = ob_row + ((((uint32_t)(cx0 + 2 * hl + a)) ^ ((uint32_t)lane & 7u)) << 4);
          stmatrix_x4_trans(addr, fh[0], fh[1], fh[2], fh[3]);
          stmatrix_x4_trans(addr + C::OIMG_BYTES, fl[0], fl[1], fl[2], fl[3]);
        }
      }
      // e' is in the output buffer: make it visible to the async proxy, then hand it to the store thread
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ofull[grp]);
      if (kTiming) {
        const long long t5 = clock64();
        tm[0] += t1 - t0; tm[1] += t3 - t2; tm[2] += (t4 - t3) + (t2 - t1); tm[3] += t5 - t4; tm[4] += 1;
      }
    }
    if (kTiming && timing != nullptr && lane == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) timing[((size_t)blockIdx.x * 32 + warp) * 5 + k] = tm[k];
    }
  }
  tc_fence_before();
  if (C::MC) cluster_sync_all();   // the peer may still multicast into this CTA's stages / arrive on its barriers
  else __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// debugging aid: when set, every epilogue warp leaves its cycle accounting in timing[blockIdx][warp][5]
static unsigned long long* g_edge_timing = nullptr;
// fault injection: stall the store thread this long after every tile (the schedule must tolerate a slow store thread)
static int g_store_delay_ns = 0;
// L2 management (EdgeMode bits): gnb_debug_edge_mode, or GNB_EDGE_MODE in the environment at the first launch
static int g_edge_mode = -1;
static int edge_mode() {
  if (g_edge_mode < 0) {
    const char* env = getenv("GNB_EDGE_MODE");
    g_edge_mode = env ? atoi(env) : kEmDefault;
  }
  return g_edge_mode;
}

template <int H>
static int edge_forward_tc2_impl(const gnb_graph_t* g, const float* P, int64_t ldP, const void* Wp,
                                 void* e16, float* F, float* carry, int flags, cudaStream_t stream) {
  using C = Edge2Cfg<H>;
  const bool res = flags & GNB_F_RESIDUAL;
  auto kern = g_edge_timing ? (res ? edge_forward_tc2_kernel<H, true, true> : edge_forward_tc2_kernel<H, false, true>)
                            : (res ? edge_forward_tc2_kernel<H, true, false> : edge_forward_tc2_kernel<H, false, false>);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (err != cudaSuccess) {
    set_error("gnb_edge_forward_tc2: cudaFuncSetAttribute(%zu): %s", C::SMEM, cudaGetErrorString(err));
    return (int)err;
  }
  const int64_t E = g->num_edges;
  CUtensorMap map_e;
  int rc = make_state_map(&map_e, e16, E, H, kE2NT);
  if (rc) return rc;
  const int64_t num_tiles = (E + kE2NT - 1) / kE2NT;
  int workers = sm_count() / C::NH;
  if (workers > num_tiles) workers = (int)num_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(workers * C::NH));
  cfg.blockDim = dim3(kE2Threads);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C::MC ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  err = cudaLaunchKernelEx(&cfg, kern, map_e, *g, P, ldP, (const __half*)Wp, F, carry, flags, workers,
                           g_edge_timing, watch_get(), g_store_delay_ns, edge_mode());
  if (err != cudaSuccess) {
    set_error("gnb_edge_forward_tc2: launch failed: %s", cudaGetErrorString(err));
    return (int)err;
  }
  return check_launch("gnb_edge_forward_tc2");
}

}  // namespace tc
}  // namespace gnb

using namespace gnb;

extern "C" int gnb_edge_tile_tc2(int H) { return (H == 64 || H == 128 || H == 256) ? tc::kE2NT : GNB_E_INVALID; }

extern "C" int gnb_edge_chunk_tc2(int H) { return (H == 64 || H == 128 || H == 256) ? tc::kE2Chunk : GNB_E_INVALID; }

extern "C" void gnb_debug_edge_timing(void* buf) { tc::g_edge_timing = (unsigned long long*)buf; }

extern "C" void gnb_debug_store_delay_ns(int ns) { tc::g_store_delay_ns = ns; }

extern "C" void gnb_debug_edge_mode(int mode) { tc::g_edge_mode = mode < 0 ? tc::kEmDefault : mode; }

extern "C" int gnb_edge_forward_tc2(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const void* Wp,
                                    void* e16, float* F, float* carry, int flags, void* stream) {
  GNB_REQUIRE(g != nullptr && g->num_edges >= 0 && g->in_ptr != nullptr, "graph not staged");
  if (g->num_edges == 0) return 0;
  GNB_REQUIRE(g->in_src && g->in_dst, "graph not staged");
  GNB_REQUIRE(P && Wp && e16 && F && carry, "null pointer");
  GNB_REQUIRE(ldP >= ((flags & GNB_F_SYMMETRIC) ? 5 : 4) * (int64_t)H && ldP % 2 == 0 && ldP < ((int64_t)1 << 29),
              "ldP=%lld out of range", (long long)ldP);
  GNB_REQUIRE(((uintptr_t)P % 8 == 0) && ((uintptr_t)e16 % 16 == 0) && ((uintptr_t)Wp % 16 == 0),
              "gnb_edge_forward_tc2: e16 must be 16-byte aligned, P 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  switch (H) {
    case 64: return tc::edge_forward_tc2_impl<64>(g, P, ldP, Wp, e16, F, carry, flags, s);
    case 128: return tc::edge_forward_tc2_impl<128>(g, P, ldP, Wp, e16, F, carry, flags, s);
    case 256: return tc::edge_forward_tc2_impl<256>(g, P, ldP, Wp, e16, F, carry, flags, s);
  }
  set_error("gnb_edge_forward_tc2: hidden_features=%d unsupported (64, 128, 256)", H);
  return GNB_E_INVALID;
}
