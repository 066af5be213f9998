// Fused edge pass of one (Sym)GatedGCN layer on the tensor cores (reference layers/gated_gcn_full.py:97,104-114):
//   z_p  = B1h[src_p] + B2h[dst_p] + e_p * W_B3^T         (the E x H x H product runs on tcgen05)
//   e'_p = relu(z_p * scale + shift) (+ e_p)               written over e in place
//   F_i  = sum_{p: dst_p = i} sigmoid(e'_p) * A2h[src_p] / (sum_p sigmoid(e'_p) + 1e-6)
//
// Persistent CTAs (one per SM).  A CTA owns HC = min(H, 128) output channels (for H = 256 the two channel
// halves of a tile run on neighbouring CTAs) with its W_B3 block resident in TMEM, and walks 64-edge tiles
// of the dst-sorted edge array:
//   warp 0      : MMA issue (D^T[channel][edge] = W * e^T, fp16 hi/lo split, fp32 accumulate in TMEM)
//   warps 1..8  : producers: e rows (fp32, coalesced 32-byte lane loads) -> fp16 (hi, lo) operand images
//   warps 9..24 : epilogue: 2 accumulator buffers x 4 TMEM lane quarters x 2 chunks of 32 edges.
//                 A thread owns ONE channel and walks its 32 consecutive edges: gathers of the
//                 (B1h, A2h) node rows are 128/256-byte coalesced across the warp, the per-destination
//                 sums are register accumulators closed at warp-uniform segment boundaries (no atomics,
//                 fixed summation order).  Segments that straddle a 32-edge chunk leave partial sums in
//                 carry[chunk][4][H], resolved by gnb_node_update.
#include <type_traits>

#include "gnb_tc.cuh"

namespace gnb {
namespace tc {

constexpr int kEdgeNT = 64;      // edges per tile (MMA N)
constexpr int kEdgeChunk = 32;   // edges per epilogue warp = carry granularity
constexpr int kEdgeProducerWarps = 8;
constexpr int kEdgeEpiWarps = 16;
constexpr int kEdgeFirstEpiWarp = 1 + kEdgeProducerWarps;   // must be 1 (mod 4): 4 consecutive warps cover the 4 TMEM lane quarters
constexpr int kEdgeThreads = 32 * (kEdgeFirstEpiWarp + kEdgeEpiWarps);
static_assert(kEdgeFirstEpiWarp % 4 == 1, "epilogue warp numbering");

template <int H>
struct EdgeTcCfg {
  static constexpr int HC = H < kM ? H : kM;   // live channels per CTA
  static constexpr int NH = H / HC;            // channel halves (CTAs per tile)
  using T = Tile<H, kEdgeNT>;
  static constexpr int NB = (H >= 256) ? 3 : 4;
  static constexpr uint32_t TMEM_COLS = pow2_cols(2 * T::W_COLS + 2 * kEdgeNT);
  static constexpr uint32_t D_COL0 = 2 * T::W_COLS;
  static constexpr size_t SMEM = (size_t)NB * T::BUF_BYTES + 256;
};

__device__ __forceinline__ void red_release_add(int32_t* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int H>
__global__ void __launch_bounds__(kEdgeThreads, 1)
edge_forward_tc_kernel(gnb_graph_t g, const float* __restrict__ P, int64_t ldP, const __half* __restrict__ Wp,
                       const float* __restrict__ scale_e, const float* __restrict__ shift_e, float* e,
                       float* __restrict__ F, float* __restrict__ carry, int32_t* tile_flags, int epoch, int flags,
                       int workers) {
  using C = EdgeTcCfg<H>;
  using T = typename C::T;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* bufs = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)C::NB * T::BUF_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::NB;
  uint64_t* dfull = empty + C::NB;
  uint64_t* dempty = dfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

  const int half = blockIdx.x % C::NH, worker = blockIdx.x / C::NH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t E = g.num_edges;
  const int64_t num_tiles = (E + kEdgeNT - 1) / kEdgeNT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NB; ++i) {
      mbar_init(&full[i], kEdgeProducerWarps * 32);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], kEdgeEpiWarps / 2);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= kEdgeFirstEpiWarp && warp < kEdgeFirstEpiWarp + 4) load_weights_to_tmem<H>(Wp + (size_t)half * 2 * kM * H, tmem_base, warp & 3, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ---------------------------------------------------------------- MMA issue
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB, d = i & 1;
      mbar_wait(&full[s], (i / C::NB) & 1);
      // both channel halves read whole rows of e and overwrite their own half in place: tell the other
      // half that this CTA's copy of tile t has left global memory
      if (C::NH > 1 && lane == 0) red_release_add(tile_flags + t, 1);
      mbar_wait(&dempty[d], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        issue_tile_mma<H, kEdgeNT>(tmem_base, tmem_base + C::D_COL0 + d * kEdgeNT,
                                   smem_u32(bufs + (size_t)s * T::BUF_BYTES));
        mma_commit(&empty[s]);
        mma_commit(&dfull[d]);
      }
      __syncwarp();
    }
  } else if (warp <= kEdgeProducerWarps) {
    // ---------------------------------------------------------------- producers
    const int pw = warp - 1;
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB;
      mbar_wait(&empty[s], ((i / C::NB) & 1) ^ 1);
      produce_tile<H, kEdgeNT, kEdgeProducerWarps>(e, E, t * kEdgeNT, bufs + (size_t)s * T::BUF_BYTES, pw, lane);
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - kEdgeFirstEpiWarp;
    const int grp = ew >> 3, sub = (ew >> 2) & 1, q = warp & 3;
    const int cl = q * 32 + lane;            // TMEM lane = channel within the CTA's block
    const bool ch_ok = cl < C::HC;           // warp-uniform (HC is a multiple of 32)
    const int c = half * C::HC + (ch_ok ? cl : 0);
    const float sc = scale_e[c], sh = shift_e[c];
    const bool residual = flags & GNB_F_RESIDUAL;
    const float* Pc = P + 2 * c;             // (B1h[c], A2h[c]) interleaved
    const float* Pb2 = P + 2 * H + c;        // B2h[c]
    constexpr unsigned kFull = 0xffffffffu;
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      if ((i & 1) != grp) continue;
      const int64_t cs = t * kEdgeNT + sub * kEdgeChunk;
      const bool live = cs < E && ch_ok;     // warp-uniform
      const int n = live ? (int)((E - cs < kEdgeChunk) ? (E - cs) : kEdgeChunk) : 0;
      int my_src = 0, my_dst = -1, head_dst = -1, tail_dst = -1;
      unsigned segmask = 0;                  // bit j: edge j of the chunk opens a new destination segment
      if (live) {
        const int64_t pl = cs + (lane < n ? lane : n - 1);   // tail lanes repeat the last edge (loads only)
        my_src = g.in_src[pl];
        my_dst = g.in_dst[pl];
        const int prev_dst = (cs > 0) ? g.in_dst[cs - 1] : -1;
        const int next_dst = (cs + n < E) ? g.in_dst[cs + n] : -1;
        const int up = __shfl_up_sync(kFull, my_dst, 1);
        segmask = __ballot_sync(kFull, lane == 0 || (lane < n && my_dst != up));
        const int first_dst = __shfl_sync(kFull, my_dst, 0);
        const int last_dst = __shfl_sync(kFull, my_dst, n - 1);
        if (prev_dst == first_dst) head_dst = first_dst;
        if (next_dst == last_dst) tail_dst = last_dst;
      }
      mbar_wait(&dfull[grp], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::D_COL0 + grp * kEdgeNT + sub * kEdgeChunk;
      if (!live) {  // nothing to compute, but the accumulator buffer still has to be handed back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dempty[grp]);
        continue;
      }
      if (C::NH > 1) {  // the other half must have read tile t before we overwrite our channels of it
        if (lane == 0) {
          while (ld_acquire(tile_flags + t) < C::NH * epoch) __nanosleep(32);
        }
        __syncwarp();
      }
      const int64_t chunk = cs / kEdgeChunk;
      float* erow = e + cs * H + c;          // this thread's channel of the chunk's first edge
      int cur = -1;
      float num = 0.f, den = 0.f, b2s = 0.f;

      // Software pipeline over eight batches of four edges: the gathers of batch b+1 are in flight while
      // batch b is computed.  fa = (B1h, A2h)[src], fe = layer input (residual), fb = B2h[dst] (fetched only
      // where a destination segment opens).  kTail = the one ragged chunk at the end of the edge array.
      constexpr int kEB = 4;   // edges per batch
      float2 ba[2][kEB];
      float ein[2][kEB], b2v[2][kEB];
      auto fetch = [&](auto tail, int b, float2 (&fa)[kEB], float (&fe)[kEB], float (&fb)[kEB]) {
        constexpr bool kTail = decltype(tail)::value;
        const unsigned mb = segmask >> (b * kEB);
        const float* eb = erow + (int64_t)b * kEB * H;
#pragma unroll
        for (int u = 0; u < kEB; ++u) {
          const int sj = __shfl_sync(kFull, my_src, b * kEB + u);
          fa[u] = __ldg(reinterpret_cast<const float2*>(Pc + (int64_t)sj * ldP));
          if (kTail) fe[u] = (b * kEB + u < n) ? eb[u * H] : 0.f;
          else fe[u] = eb[u * H];
          fb[u] = 0.f;
          if (mb & (1u << u)) {   // warp-uniform
            const int dj = __shfl_sync(kFull, my_dst, b * kEB + u);
            fb[u] = __ldg(Pb2 + (int64_t)dj * ldP);
          }
        }
      };
      auto compute = [&](auto tail, int b, const float2 (&fa)[kEB], const float (&fe)[kEB], const float (&fb)[kEB]) {
        constexpr bool kTail = decltype(tail)::value;
        uint32_t zr[kEB];
        tmem_ld4(taddr + b * kEB, zr);
        tmem_ld_wait();
        if (b == kEdgeChunk / kEB - 1) {   // last read of this accumulator buffer: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&dempty[grp]);
        }
        const unsigned mb = segmask >> (b * kEB);
        float* eb = erow + (int64_t)b * kEB * H;
#pragma unroll
        for (int u = 0; u < kEB; ++u) {
          if (mb & (1u << u)) {          // warp-uniform: close the running segment, open the next
            if (cur >= 0) {
              if (cur == head_dst) {
                carry[(chunk * 4 + 0) * H + c] = num;
                carry[(chunk * 4 + 1) * H + c] = den;
              } else {
                F[(int64_t)cur * H + c] = gate_div(num, den);
              }
            }
            cur = __shfl_sync(kFull, my_dst, b * kEB + u);
            num = 0.f;
            den = 0.f;
            b2s = fmaf(fb[u], sc, sh);
          }
          float v = fmaf(__uint_as_float(zr[u]) + fa[u].x, sc, b2s);
          v = fmaxf(v, 0.f);
          if (residual) v += fe[u];
          if (!kTail || b * kEB + u < n) {
            eb[u * H] = v;
            const float sg = sigmoidf_fast(v);
            num = fmaf(sg, fa[u].y, num);
            den += sg;
          }
        }
      };
      auto run_chunk = [&](auto tail) {
        fetch(tail, 0, ba[0], ein[0], b2v[0]);
#pragma unroll 1
        for (int b = 0; b < kEdgeChunk / kEB; b += 2) {
          fetch(tail, b + 1, ba[1], ein[1], b2v[1]);
          compute(tail, b, ba[0], ein[0], b2v[0]);
          if (b + 2 < kEdgeChunk / kEB) fetch(tail, b + 2, ba[0], ein[0], b2v[0]);
          compute(tail, b + 1, ba[1], ein[1], b2v[1]);
        }
      };
      if (n == kEdgeChunk) run_chunk(std::false_type{});
      else run_chunk(std::true_type{});
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int H>
static int edge_forward_tc_impl(const gnb_graph_t* g, const float* P, int64_t ldP, const void* Wp,
                                const float* scale_e, const float* shift_e, float* e, float* F, float* carry,
                                int32_t* tile_flags, int epoch, int flags, cudaStream_t stream) {
  using C = EdgeTcCfg<H>;
  cudaError_t err = cudaFuncSetAttribute(edge_forward_tc_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM);
  if (err != cudaSuccess) {
    set_error("gnb_edge_forward_tc: cudaFuncSetAttribute(%zu): %s", C::SMEM, cudaGetErrorString(err));
    return (int)err;
  }
  if (C::NH > 1) GNB_REQUIRE(tile_flags != nullptr && epoch > 0, "gnb_edge_forward_tc: H=%d needs tile_flags and epoch >= 1", H);
  const int64_t num_tiles = (g->num_edges + kEdgeNT - 1) / kEdgeNT;
  int workers = sm_count() / C::NH;
  if (workers > num_tiles) workers = (int)num_tiles;
  // every CTA must be resident at the same time (the channel halves wait on each other's flags)
  edge_forward_tc_kernel<H><<<workers * C::NH, kEdgeThreads, C::SMEM, stream>>>(
      *g, P, ldP, (const __half*)Wp, scale_e, shift_e, e, F, carry, tile_flags, epoch, flags, workers);
  return check_launch("gnb_edge_forward_tc");
}

}  // namespace tc
}  // namespace gnb

using namespace gnb;

extern "C" int gnb_edge_chunk_tc(int H) { return supported_h(H) ? tc::kEdgeChunk : GNB_E_INVALID; }

extern "C" int gnb_edge_tile_tc(int H) { return supported_h(H) ? tc::kEdgeNT : GNB_E_INVALID; }

extern "C" int gnb_edge_forward_tc(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const void* Wp,
                                   const float* scale_e, const float* shift_e, float* e, float* F, float* carry,
                                   int32_t* tile_flags, int epoch, int flags, void* stream) {
  GNB_REQUIRE(g != nullptr && g->num_edges >= 0 && g->in_ptr != nullptr, "graph not staged");
  if (g->num_edges == 0) return 0;
  GNB_REQUIRE(g->in_src && g->in_dst, "graph not staged");
  GNB_REQUIRE(P && Wp && scale_e && shift_e && e && F && carry, "null pointer");
  GNB_REQUIRE(ldP >= ((flags & GNB_F_SYMMETRIC) ? 5 : 4) * (int64_t)H && ldP % 2 == 0, "ldP=%lld too small",
              (long long)ldP);
  GNB_REQUIRE(((uintptr_t)P % 8 == 0) && ((uintptr_t)e % 32 == 0) && ((uintptr_t)Wp % 16 == 0),
              "gnb_edge_forward_tc: e must be 32-byte aligned, P 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  switch (H) {
    case 32: return tc::edge_forward_tc_impl<32>(g, P, ldP, Wp, scale_e, shift_e, e, F, carry, tile_flags, epoch, flags, s);
    case 64: return tc::edge_forward_tc_impl<64>(g, P, ldP, Wp, scale_e, shift_e, e, F, carry, tile_flags, epoch, flags, s);
    case 128: return tc::edge_forward_tc_impl<128>(g, P, ldP, Wp, scale_e, shift_e, e, F, carry, tile_flags, epoch, flags, s);
    case 256: return tc::edge_forward_tc_impl<256>(g, P, ldP, Wp, scale_e, shift_e, e, F, carry, tile_flags, epoch, flags, s);
  }
  set_error("hidden_features=%d unsupported (32, 64, 128, 256)", H);
  return GNB_E_INVALID;
}
