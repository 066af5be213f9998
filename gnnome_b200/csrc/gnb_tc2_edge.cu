// Fused edge pass of one (Sym)GatedGCN layer, second generation: TMA-fed tcgen05 with the edge state held in the
// split16 format (gnb_tma.cuh).  Reference layers/gated_gcn_full.py:97,104-114:
//   z_p  = B1h[src_p] + B2h[dst_p] + e_p * W_B3^T         (E x H x H product on tcgen05, fp16 hi/lo split, fp32 in TMEM)
//   e'_p = relu(z_p * scale + shift) (+ e_p)               written over e in place
//   F_i  = sum_{p: dst_p = i} sigmoid(e'_p) * A2h[src_p] / (sum_p sigmoid(e'_p) + 1e-6)
//
// Persistent CTAs, one per SM; a CTA owns HC = min(H, 128) output channels (H = 256: the two channel halves of a
// tile run on neighbouring CTAs) with its W_B3 block resident in TMEM, and walks 32-edge tiles of the dst-sorted
// edge array (six to eight stages in flight: one per epilogue group plus the tiles being loaded / multiplied ahead).
// One shared-memory stage serves a tile through its whole life:
//   TMA load (hi, lo images, 128B-swizzled)  ->  MMA B operand  ->  residual source for the epilogue  ->
//   e' written over it in place (same thread, same address)  ->  TMA store back to HBM.
//   warp 0      : producer: TMA loads (lane 0) + the tile's (src, dst) indices into the stage's index area
//   warp 1      : MMA issue (one lane)
//   warps 2, 3  : TMA stores, one thread for the even stages, one for the odd ones, polling (store whichever finished stage
//                 group is ready, release the stage).  The hand-over barrier sfull is PER STAGE and each stage has ONE
//                 store thread: the barrier's phases are consumed in sequence by a single waiter, and a stage
//                 cannot be armed again before that waiter has stored it and released it for reloading.
//   Every wait is bounded by the spin watchdog (gnb_tc.cuh): a lost phase traps with a record instead of spinning.
//   warps 4..19 : epilogue, 4 groups x 4 TMEM lane quarters; group g takes tiles g, g+4, ... (one 32-edge chunk).
//                 A thread owns ONE channel and walks its 32 consecutive edges: gathers of the (B1h, A2h) node
//                 rows are coalesced across the warp, per-destination sums are register accumulators closed at
//                 warp-uniform segment boundaries (no atomics, fixed summation order).  Segments that straddle
//                 a 32-edge chunk leave partial sums in carry[chunk][4][H], resolved by gnb_node_update2.
#include <type_traits>

#include "gnb_tma.cuh"

namespace gnb {
namespace tc {

constexpr int kE2NT = 32;        // edges per tile (MMA N) = edges per epilogue warp = carry granularity
constexpr int kE2Chunk = kE2NT;
constexpr int kE2Groups = 4;     // epilogue groups of four warps (one per TMEM lane quarter); group g takes tiles g, g+G, ...
                                 // (five groups were measured slower: the schedulers are issue-bound, r01h)
constexpr int kE2DBufs = 8;      // accumulator buffers (two per group: the MMA of a group's next tile overlaps its epilogue)
constexpr int kE2FirstEpiWarp = 4;
constexpr int kE2Threads = 32 * (kE2FirstEpiWarp + 4 * kE2Groups);
constexpr int kE2IdxInts = 2 * kE2NT + 4;   // src[32], dst[32], prev_dst, next_dst (+ pad: stages stay 16-byte aligned)

template <int H>
struct Edge2Cfg {
  static constexpr int HC = H < kM ? H : kM;   // live channels per CTA
  static constexpr int NH = H / HC;            // channel halves (CTAs per tile)
  // H = 256: the two channel halves of a tile form a 2-CTA cluster; each loads half of the tile's boxes and
  // multicasts them to both, so the e tile is read from L2 / HBM once, and no inter-CTA flag is needed for the
  // in-place update (a CTA stores its channels only after ITS stage is complete, i.e. after every box of the tile
  // has been read from global memory on behalf of both)
  static constexpr bool MC = NH == 2;
  using T = Tile2<H, kE2NT>;
  // Stages: one per group in its epilogue plus two being loaded / multiplied ahead.
  static constexpr int NB = (H >= 256) ? 6 : 8;
  // epilogue warps of a group that own live channels (the others idle: they must not feed the barriers, or they
  // would run ahead of the live ones and complete a phase early)
  static constexpr int LIVE_WARPS = HC / 32;
  static constexpr uint32_t TMEM_COLS = pow2_cols(2 * T::W_COLS + kE2DBufs * kE2NT);
  static constexpr uint32_t D_COL0 = 2 * T::W_COLS;
  static constexpr size_t SMEM = (size_t)NB * T::BUF_BYTES + 1024 + (size_t)NB * kE2IdxInts * 4 + 512;
};

__device__ __forceinline__ uint16_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

template <int H, bool kResidual, bool kTiming>
__global__ void __launch_bounds__(kE2Threads, 1)
edge_forward_tc2_kernel(const __grid_constant__ CUtensorMap map_e, gnb_graph_t g, const float* __restrict__ P, int64_t ldP, const __half* __restrict__ Wp,
                        const float* __restrict__ scale_e, const float* __restrict__ shift_e,
                        float* __restrict__ F, float* __restrict__ carry, int flags, int workers,
                        unsigned long long* timing, const Watch watch, int store_delay_ns) {
  using C = Edge2Cfg<H>;
  using T = typename C::T;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the shared array (an integer round-trip would demote every later
  // access through these pointers to generic loads)
  uint8_t* bufs = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  int* idx_area = reinterpret_cast<int*>(bufs + (size_t)C::NB * T::BUF_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(idx_area + C::NB * kE2IdxInts);
  uint64_t* full = bars;                    // [NB] producer (TMA bytes + 32 index lanes) -> MMA, epilogue
  uint64_t* empty = full + C::NB;           // [NB] store warp -> producer
  uint64_t* dfull = empty + C::NB;          // [D]  MMA -> epilogue
  uint64_t* dempty = dfull + kE2DBufs;      // [D]  epilogue -> MMA
  uint64_t* sfull = dempty + kE2DBufs;      // [NB] epilogue (e' written into the stage) -> store warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sfull + C::NB);
  static_assert((3 * C::NB + 2 * kE2DBufs) * 8 + 4 <= 512, "barrier area");

  const int half = blockIdx.x % C::NH, worker = blockIdx.x / C::NH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t E = g.num_edges;
  const int64_t num_tiles = (E + kE2NT - 1) / kE2NT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NB; ++i) {
      mbar_init(&full[i], 33);
      mbar_init(&empty[i], C::MC ? 2 : 1);   // multicast: both CTAs of the cluster write into a stage
      mbar_init(&sfull[i], C::LIVE_WARPS);
    }
    for (int i = 0; i < kE2DBufs; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], C::LIVE_WARPS);
    }
    fence_barrier_init();
    prefetch_tensormap(&map_e);
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if (C::MC) cluster_sync_all();   // the peer's multicast must find initialised barriers
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= kE2FirstEpiWarp && warp < kE2FirstEpiWarp + 4)
    load_weights_to_tmem<H>(Wp + (size_t)half * 2 * kM * H, tmem_base, warp & 3, lane);
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
    // Every node row is touched for the first time by SOME gather of the epilogue, and that one would wait for
    // HBM; the producer knows the tile's endpoints two to three tile periods before the epilogue needs them, so
    // it pulls this CTA's slices of the (B1h, A2h)[src] and B2h[dst] rows into L2 ahead of time.
    auto prefetch_rows = [&](const int (&r)[3]) {
      {
        const int sj = r[0], dj = r[1];
        {   // rows past a ragged end carry node 0: a harmless prefetch
          const char* a = reinterpret_cast<const char*>(P + (int64_t)sj * ldP + 2 * half * C::HC);
#pragma unroll
          for (int l = 0; l < C::HC * 8 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + l * 128));
          const char* b = reinterpret_cast<const char*>(P + (int64_t)dj * ldP + 2 * H + half * C::HC);
#pragma unroll
          for (int l = 0; l < C::HC * 4 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + l * 128));
        }
      }
    };
    int cur[3], nxt[3];
    if (worker < num_tiles) load_idx(worker, cur);
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB;
      prefetch_rows(cur);
      mbar_wait(&empty[s], ((i / C::NB) & 1) ^ 1, 64, watch, watch_tag(kWkEdge2, kWrProducer, kWbEmpty), s, i);
      if (elect_one()) {
        uint8_t* stage = bufs + (size_t)s * T::BUF_BYTES;
        mbar_arrive_expect_tx(&full[s], T::BUF_BYTES);
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
      if (t + workers < num_tiles) load_idx(t + workers, nxt);   // in flight while this tile's indices are published
      int* ia = idx_area + s * kE2IdxInts;
      ia[lane] = cur[0];
      ia[kE2NT + lane] = cur[1];
      if (lane < 2) ia[2 * kE2NT + lane] = cur[2];
      mbar_arrive(&full[s]);
#pragma unroll
      for (int k = 0; k < 3; ++k) cur[k] = nxt[k];
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issue
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB, d = i % kE2DBufs;
      mbar_wait(&full[s], (i / C::NB) & 1, 32, watch, watch_tag(kWkEdge2, kWrMma, kWbFull), s, i);
      mbar_wait(&dempty[d], ((i / kE2DBufs) & 1) ^ 1, 32, watch, watch_tag(kWkEdge2, kWrMma, kWbDEmpty), d, i);
      tc_fence_after();
      if (elect_one()) {
        issue_tile_mma_sw128<H, kE2NT>(tmem_base, tmem_base + C::D_COL0 + d * kE2NT,
                                       smem_u32(bufs + (size_t)s * T::BUF_BYTES));
        mma_commit(&dfull[d]);
      }
      __syncwarp();
    }
  } else if (warp < kE2FirstEpiWarp) {
    // ---------------------------------------------------------------- TMA store of one epilogue group
    // Lane 0 of warp 2 stores the tiles that pass through the even stages, lane 0 of warp 3 those of the odd stages.
    // A stage's hand-over barrier thus has ONE consumer that sees every phase of it in sequence: it tests phase k
    // only after it has stored phase k - 1, and phase k + 1 cannot complete before it releases the stage -- the
    // parity test can neither alias forwards nor be satisfied by the phase before last.  (Round 1 assigned the
    // stores by epilogue GROUP with one barrier per group: a group could complete two phases while the thread was
    // busy with its other group -- the cfg3 dead-lock.  One barrier per stage polled by the group's thread is wrong
    // the other way round: with NB = 6 a stage alternates between two groups, so each thread saw every OTHER phase
    // and a fresh barrier satisfied its parity test before the tile had even been processed.)
    // The thread POLLS its stages without blocking on any of them (tiles of different groups finish out of order).
    if (lane == 0) {
      const int l0 = warp - 2;                              // stages l0, l0 + 2, ...
      constexpr int kMaxSt = (C::NB + 1) / 2;
      int it[kMaxSt];                                       // next tile iteration that passes through each served stage
#pragma unroll
      for (int k = 0; k < kMaxSt; ++k) it[k] = l0 + 2 * k;
      auto remaining = [&](int k) { return l0 + 2 * k < C::NB && worker + (int64_t)it[k] * workers < num_tiles; };
      auto any_remaining = [&]() {
        bool r = false;
#pragma unroll
        for (int k = 0; k < kMaxSt; ++k) r = r || remaining(k);
        return r;
      };
      SpinGuard guard;
      while (any_remaining()) {
        bool progressed = false;
#pragma unroll
        for (int k = 0; k < kMaxSt; ++k) {
          if (!remaining(k)) continue;
          const int i = it[k];
          const int s = l0 + 2 * k;                         // == i % NB
          if (!mbar_test(&sfull[s], (i / C::NB) & 1)) continue;
          const int64_t t = worker + (int64_t)i * workers;
          const uint8_t* stage = bufs + (size_t)s * T::BUF_BYTES;
#pragma unroll
          for (int kbl = 0; kbl < C::HC / kKB; ++kbl) {
            const int kb = half * (C::HC / kKB) + kbl;
            tma_store_2d(&map_e, stage + kb * T::KB_BYTES, kb * kKB, (int)(t * kE2NT));
            tma_store_2d(&map_e, stage + T::IMG_BYTES + kb * T::KB_BYTES, H + kb * kKB, (int)(t * kE2NT));
          }
          tma_store_commit();
          tma_store_wait_read();
          if (store_delay_ns > 0) __nanosleep((unsigned)store_delay_ns);   // fault injection (gnb_debug_store_delay_ns)
          if (store_delay_ns >= 0) {   // < 0: the stage is never released -- the watchdog test's dead-lock
            mbar_arrive(&empty[s]);
            if (C::MC) mbar_arrive_cluster(&empty[s], (uint32_t)(half ^ 1));   // the peer also writes into this stage
          }
          it[k] += C::NB;
          progressed = true;
        }
        if (progressed) {
          guard = SpinGuard();
        } else {
          __nanosleep(32);
          guard.poll(watch, watch_tag(kWkEdge2, kWrStore, kWbSFull), (uint32_t)l0, (uint32_t)((it[0] / C::NB) & 1), it[0]);
        }
      }
      tma_store_wait_all();
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - kE2FirstEpiWarp;
    const int grp = ew >> 2, q = warp & 3;
    constexpr int sub = 0;                   // one 32-edge chunk per tile
    const int cl = q * 32 + lane;            // TMEM lane = channel within the CTA's block
    const bool ch_ok = cl < C::HC;           // warp-uniform (HC is a multiple of 32)
    const int c = half * C::HC + (ch_ok ? cl : 0);
    const int64_t my_tiles = ch_ok ? num_tiles : 0;   // warps without live channels (H = 64) sit the loop out
    const float sc = scale_e[c], sh = shift_e[c];
    const char* Pc = reinterpret_cast<const char*>(P + 2 * c);        // (B1h[c], A2h[c]) interleaved
    const char* Pb2 = reinterpret_cast<const char*>(P + 2 * H + c);   // B2h[c]
    const int ldPb = (int)(ldP * (int64_t)sizeof(float));            // row pitch in bytes (< 2^31, checked by the host)
    constexpr unsigned kFull = 0xffffffffu;
    // this thread's column of the stage: element (row, c) of an image sits at
    //   (c / 64) * KB_BYTES + row * 128 + ((((c % 64) / 8) ^ (row % 8)) * 16) + (c % 8) * 2
    const uint32_t col_base = (uint32_t)((c >> 6) * T::KB_BYTES + sub * kE2Chunk * 128 + ((c & 7) << 1));
    const uint32_t col_x = (uint32_t)((c & 63) >> 3);
    // Hand a finished stage to the store warp: every live warp of the group arrives on the STAGE's sfull.  All
    // arrivals of the stage's previous use precede its store, its reload and hence the `full` phase this warp has
    // waited for, so a fast warp can never complete a phase on behalf of a slower one.
    // optional cycle accounting (gnb_debug_edge_timing): [full wait, dfull wait, batches, flush + hand-off, tiles]
    unsigned long long tm[5] = {0, 0, 0, 0, 0};
    auto stage_done = [&](int s) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull[s]);
    };
    int i = 0;
    for (int64_t t = worker; t < my_tiles; t += workers, ++i) {
      if (i % kE2Groups != grp) continue;
      const int s = i % C::NB, d = i % kE2DBufs;
      const uint32_t dpar = (i / kE2DBufs) & 1;
      const int64_t cs = t * kE2NT + sub * kE2Chunk;
      const bool live = cs < E;              // warp-uniform; false only for the second half of a ragged last tile
      const int n = live ? (int)((E - cs < kE2Chunk) ? (E - cs) : kE2Chunk) : 0;
      const long long t0 = kTiming ? clock64() : 0;
      // indices published (and the operand tile has landed)
      mbar_wait(&full[s], (i / C::NB) & 1, 32, watch, watch_tag(kWkEdge2, kWrEpilogue, kWbFull), s, i);
      const long long t1 = kTiming ? clock64() : 0;
      if (!live) {  // nothing to compute, but the barriers still have to be fed
        mbar_wait(&dfull[d], dpar, 64, watch, watch_tag(kWkEdge2, kWrEpilogue, kWbDFull), d, i);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dempty[d]);
        stage_done(s);
        continue;
      }
      const int* ia = idx_area + s * kE2IdxInts;
      const int my_dst = ia[kE2NT + sub * kE2Chunk + lane];
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
      const uint32_t st_hi = smem_u32(bufs + (size_t)s * T::BUF_BYTES) + col_base;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::D_COL0 + d * kE2NT + sub * kE2Chunk;
      int cur = -1;
      float num = 0.f, den = 0.f;

      // Software pipeline over four batches of eight edges: the gathers of batch b+1 -- xa = (B1h, A2h)[src] and
      // xb = B2h[dst], one coalesced row segment per warp each -- are in flight while batch b is computed.
      constexpr int kEB = 8;
      const int* ia_src = ia + sub * kE2Chunk;            // endpoints of the chunk's 32 edges: warp-uniform reads
      const int* ia_dst = ia + kE2NT + sub * kE2Chunk;
      auto fetch = [&](int b, float2 (&xa)[kEB], float (&xb)[kEB]) {
        int sj[kEB], dj[kEB];
        *reinterpret_cast<int4*>(&sj[0]) = *reinterpret_cast<const int4*>(ia_src + b * kEB);
        *reinterpret_cast<int4*>(&sj[4]) = *reinterpret_cast<const int4*>(ia_src + b * kEB + 4);
        *reinterpret_cast<int4*>(&dj[0]) = *reinterpret_cast<const int4*>(ia_dst + b * kEB);
        *reinterpret_cast<int4*>(&dj[4]) = *reinterpret_cast<const int4*>(ia_dst + b * kEB + 4);
#pragma unroll
        for (int u = 0; u < kEB; ++u) xa[u] = __ldg(reinterpret_cast<const float2*>(Pc + (int64_t)sj[u] * ldPb));
        // B2h[dst]: the four edges of a quad share it unless a destination segment opens at its 2nd..4th edge
        const unsigned m8 = segmask >> (b * kEB);
#pragma unroll
        for (int hq = 0; hq < kEB; hq += 4) {
          if (((m8 >> hq) & 0xeu) == 0) {   // warp-uniform
            const float v = __ldg(reinterpret_cast<const float*>(Pb2 + (int64_t)dj[hq] * ldPb));
            xb[hq] = v; xb[hq + 1] = v; xb[hq + 2] = v; xb[hq + 3] = v;
          } else {
#pragma unroll
            for (int w = 0; w < 4; ++w)
              xb[hq + w] = __ldg(reinterpret_cast<const float*>(Pb2 + (int64_t)dj[hq + w] * ldPb));
          }
        }
      };
      // Stage address of row r of the chunk for this thread's channel: st_hi + r * 128 + (((c%64)/8 ^ r%8) << 4).
      // rowx[w] covers rows = w (mod 4) of the current half-batch; it advances by 4 rows per half-batch, which
      // flips bit 2 of the swizzle term (xor 64 bytes) on top of the 512-byte step.
      uint32_t rowx[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) rowx[w] = st_hi + (uint32_t)(w * 128) + ((col_x ^ (uint32_t)w) << 4);
      auto compute = [&](auto full_tag, int b, const float2 (&xa)[kEB], const float (&xb)[kEB]) {
        constexpr bool kFullChunk = decltype(full_tag)::value;   // all 32 edges exist: no per-edge validity tests
        // Two half-batches of four edges.  The shared-memory accesses are volatile asm statements, which the
        // compiler keeps in program order: the loads of a half-batch come first and its stores last, so that the
        // four per-edge dependency chains in between interleave.
#pragma unroll
        for (int hb = 0; hb < kEB; hb += 4) {
          uint32_t zr[4];
          tmem_ld4(taddr + b * kEB + hb, zr);
          float ein[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            ein[w] = 0.f;
            if (kResidual) {
              const __half_raw hr{lds_u16(rowx[w])}, lr{lds_u16(rowx[w] + T::IMG_BYTES)};
              ein[w] = (__half2float(__half(hr)) + __half2float(__half(lr))) * kWScale;
            }
          }
          tmem_ld_wait();
          if (b == kE2Chunk / kEB - 1 && hb == 4) {   // last read of this accumulator buffer: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dempty[d]);
          }
          float v[4], sg[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int u = hb + w;
            v[w] = fmaxf(fmaf(__uint_as_float(zr[w]) + xa[u].x + xb[u], sc, sh), 0.f) + ein[w];
            sg[w] = (kFullChunk || b * kEB + u < n) ? sigmoidf_fast(v[w]) : 0.f;
          }
          // split fp16: two edges per packed conversion
          uint32_t ph[2], pl[2];
#pragma unroll
          for (int w2 = 0; w2 < 2; ++w2) {
            const float x0 = v[2 * w2] * kXScale, x1 = v[2 * w2 + 1] * kXScale;
            const __half2 hh = __floats2half2_rn(x0, x1);
            const float2 back = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(x0 - back.x, x1 - back.y);
            ph[w2] = *reinterpret_cast<const uint32_t*>(&hh);
            pl[w2] = *reinterpret_cast<const uint32_t*>(&ll);
          }
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            if (kFullChunk || b * kEB + hb + w < n) {
              sts_u16(rowx[w], (uint16_t)((w & 1) ? (ph[w >> 1] >> 16) : (ph[w >> 1] & 0xffffu)));
              sts_u16(rowx[w] + T::IMG_BYTES, (uint16_t)((w & 1) ? (pl[w >> 1] >> 16) : (pl[w >> 1] & 0xffffu)));
            }
            rowx[w] = (rowx[w] + 512u) ^ 64u;
          }
          const unsigned mb = (segmask >> (b * kEB + hb)) & 0xfu;
          if (mb == 0) {                 // warp-uniform: the four edges continue the running segment
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              num = fmaf(sg[w], xa[hb + w].y, num);
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
                cur = ia_dst[b * kEB + hb + w];
                num = 0.f;
                den = 0.f;
              }
              num = fmaf(sg[w], xa[hb + w].y, num);
              den += sg[w];
            }
          }
        }
      };
      float2 fa0[kEB], fa1[kEB];
      float fb0[kEB], fb1[kEB];
      fetch(0, fa0, fb0);
      const long long t2 = kTiming ? clock64() : 0;
      mbar_wait(&dfull[d], dpar, 32, watch, watch_tag(kWkEdge2, kWrEpilogue, kWbDFull), d, i);
      tc_fence_after();
      const long long t3 = kTiming ? clock64() : 0;
      auto run_chunk = [&](auto full_tag) {
#pragma unroll 1
        for (int b = 0; b < kE2Chunk / kEB; b += 2) {
          fetch(b + 1, fa1, fb1);
          compute(full_tag, b, fa0, fb0);
          if (b + 2 < kE2Chunk / kEB) fetch(b + 2, fa0, fb0);
          compute(full_tag, b + 1, fa1, fb1);
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
      // e' is in the stage: make it visible to the async proxy, then hand the stage to the store warp
      fence_proxy_async();
      stage_done(s);
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

template <int H>
static int edge_forward_tc2_impl(const gnb_graph_t* g, const float* P, int64_t ldP, const void* Wp,
                                 const float* scale_e, const float* shift_e, void* e16, float* F, float* carry,
                                 int flags, cudaStream_t stream) {
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
  err = cudaLaunchKernelEx(&cfg, kern, map_e, *g, P, ldP, (const __half*)Wp, scale_e, shift_e, F, carry, flags, workers,
                           g_edge_timing, watch_get(), g_store_delay_ns);
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

extern "C" int gnb_edge_forward_tc2(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const void* Wp,
                                    const float* scale_e, const float* shift_e, void* e16, float* F, float* carry,
                                    int flags, void* stream) {
  GNB_REQUIRE(g != nullptr && g->num_edges >= 0 && g->in_ptr != nullptr, "graph not staged");
  if (g->num_edges == 0) return 0;
  GNB_REQUIRE(g->in_src && g->in_dst, "graph not staged");
  GNB_REQUIRE(P && Wp && scale_e && shift_e && e16 && F && carry, "null pointer");
  GNB_REQUIRE(ldP >= ((flags & GNB_F_SYMMETRIC) ? 5 : 4) * (int64_t)H && ldP % 2 == 0 && ldP < ((int64_t)1 << 29),
              "ldP=%lld out of range", (long long)ldP);
  GNB_REQUIRE(((uintptr_t)P % 8 == 0) && ((uintptr_t)e16 % 16 == 0) && ((uintptr_t)Wp % 16 == 0),
              "gnb_edge_forward_tc2: e16 must be 16-byte aligned, P 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  switch (H) {
    case 64: return tc::edge_forward_tc2_impl<64>(g, P, ldP, Wp, scale_e, shift_e, e16, F, carry, flags, s);
    case 128: return tc::edge_forward_tc2_impl<128>(g, P, ldP, Wp, scale_e, shift_e, e16, F, carry, flags, s);
    case 256: return tc::edge_forward_tc2_impl<256>(g, P, ldP, Wp, scale_e, shift_e, e16, F, carry, flags, s);
  }
  set_error("gnb_edge_forward_tc2: hidden_features=%d unsupported (64, 128, 256)", H);
  return GNB_E_INVALID;
}
