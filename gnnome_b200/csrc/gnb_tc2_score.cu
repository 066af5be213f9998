// ScorePredictor (reference layers/score_predictor.py:12-24) on the split16 path.
//   W1 = [W1s | W1d | W1e]  =>  W1 . cat(x[src], x[dst], e) = S[src][0:hs] + S[dst][hs:2hs] + e . W1e^T
// with S[n] = [x_n W1s^T | x_n W1d^T + b1] projected once per node (gnb_node_linear_tc2).  Per 64-edge tile:
//   e . W1e^T          tcgen05 (W1e, zero-padded to 128 rows, resident in TMEM; e tile by TMA, fp16 hi/lo split)
//   phase A (epilogue) thread = hidden unit k: t[edge][k] = relu(D[k][edge] + S[src][k] + S[dst][hs + k]) -> smem
//   phase B            lane = unit m, 8 edges per warp: u = relu(W2 t + b2), score = W3 . u + b3 -> scores[in_eid[p]]
//   warp 0 : producer (TMA + src/dst/eid of the tile)   warp 1 : MMA issue   warps 4..19 : two epilogue groups
#include "gnb_tma.cuh"

namespace gnb {
namespace tc {

constexpr int kS2NT = 64;
constexpr int kS2Groups = 2;
constexpr int kS2FirstEpiWarp = 4;
constexpr int kS2Threads = 32 * (kS2FirstEpiWarp + 8 * kS2Groups);
constexpr int kS2IdxInts = 3 * kS2NT;   // src, dst, eid

template <int H, int HS>
struct Score2Cfg {
  using T = Tile2<H, kS2NT>;
  static constexpr int NB = (H >= 256) ? 2 : 4;
  static constexpr int TS = HS + 4;                      // stride of the hidden tile: rows stay 16-byte aligned
  static constexpr uint32_t TMEM_COLS = pow2_cols(2 * T::W_COLS + kS2Groups * kS2NT);
  static constexpr uint32_t D_COL0 = 2 * T::W_COLS;
  static constexpr size_t T_FLOATS = (size_t)kS2Groups * kS2NT * TS;
  static constexpr size_t SMEM = (size_t)NB * T::BUF_BYTES + 1024 + (size_t)NB * kS2IdxInts * 4 +
                                 (T_FLOATS + (size_t)HS * 32 + 64) * 4 + 256;
};

template <int H, int HS>
__global__ void __launch_bounds__(kS2Threads, 1)
score_forward_tc2_kernel(const __grid_constant__ CUtensorMap map_e, gnb_graph_t g, const float* __restrict__ S, const __half* __restrict__ Wp,
                         const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ W3,
                         const float* __restrict__ b3, float* __restrict__ scores, const Watch watch) {
  using C = Score2Cfg<H, HS>;
  using T = typename C::T;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the shared array (an integer round-trip would demote every later
  // access through these pointers to generic loads)
  uint8_t* bufs = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  int* idx_area = reinterpret_cast<int*>(bufs + (size_t)C::NB * T::BUF_BYTES);
  float* t_s = reinterpret_cast<float*>(idx_area + C::NB * kS2IdxInts);   // [G][64][TS]
  float* w2t_s = t_s + C::T_FLOATS;                                       // [HS][32]  (W2 transposed)
  float* b2_s = w2t_s + HS * 32;                                          // [32]
  float* w3_s = b2_s + 32;                                                // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(w3_s + 32);
  uint64_t* full = bars;                   // [NB] producer -> MMA, epilogue
  uint64_t* empty = full + C::NB;          // [NB] MMA -> producer
  uint64_t* dfull = empty + C::NB;         // [G]  MMA -> epilogue
  uint64_t* dempty = dfull + kS2Groups;    // [G]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + kS2Groups);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int worker = blockIdx.x, workers = gridDim.x;
  const int64_t E = g.num_edges;
  const int64_t num_tiles = (E + kS2NT - 1) / kS2NT;

  for (int i = threadIdx.x; i < 32 * HS; i += kS2Threads) w2t_s[(i % HS) * 32 + (i / HS)] = W2[i];   // W2 is [32][HS]
  if (threadIdx.x < 32) {
    b2_s[threadIdx.x] = b2[threadIdx.x];
    w3_s[threadIdx.x] = W3[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NB; ++i) {
      mbar_init(&full[i], 33);
      mbar_init(&empty[i], 1 + 8 * kS2Groups / kS2Groups);   // MMA commit + the 8 warps of the tile's group (indices)
    }
    for (int i = 0; i < kS2Groups; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], 8);
    }
    fence_barrier_init();
    prefetch_tensormap(&map_e);
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= kS2FirstEpiWarp && warp < kS2FirstEpiWarp + 4) load_weights_to_tmem<H>(Wp, tmem_base, warp & 3, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: TMA + indices
    auto load_idx = [&](int64_t t, int (&r)[6]) {
      const int64_t p0 = t * kS2NT + lane, p1 = p0 + 32;
      r[0] = (p0 < E) ? g.in_src[p0] : 0;
      r[1] = (p1 < E) ? g.in_src[p1] : 0;
      r[2] = (p0 < E) ? g.in_dst[p0] : 0;
      r[3] = (p1 < E) ? g.in_dst[p1] : 0;
      r[4] = (p0 < E) ? g.in_eid[p0] : -1;
      r[5] = (p1 < E) ? g.in_eid[p1] : -1;
    };
    // The projected node rows S[src][0:hs], S[dst][hs:2hs] are gathered by the epilogue; whoever touches a row
    // first would wait for HBM, so the producer pulls them into L2 a tile or two ahead (it has the endpoints).
    auto prefetch_rows = [&](const int (&r)[6]) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const char* a = reinterpret_cast<const char*>(S + (int64_t)r[k] * 2 * HS);
        const char* b = reinterpret_cast<const char*>(S + (int64_t)r[2 + k] * 2 * HS + HS);
#pragma unroll
        for (int l = 0; l < HS * 4 / 128; ++l) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a + l * 128));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(b + l * 128));
        }
      }
    };
    int cur[6], nxt[6];
    if (worker < num_tiles) load_idx(worker, cur);
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB;
      prefetch_rows(cur);
      mbar_wait(&empty[s], ((i / C::NB) & 1) ^ 1, 64, watch, watch_tag(kWkScore2, kWrProducer, kWbEmpty), s, i);
      if (elect_one()) {
        uint8_t* stage = bufs + (size_t)s * T::BUF_BYTES;
        mbar_arrive_expect_tx(&full[s], T::BUF_BYTES);
#pragma unroll
        for (int kb = 0; kb < T::KBLOCKS; ++kb) {
          tma_load_2d(stage + kb * T::KB_BYTES, &map_e, kb * kKB, (int)(t * kS2NT), &full[s]);
          tma_load_2d(stage + T::IMG_BYTES + kb * T::KB_BYTES, &map_e, H + kb * kKB, (int)(t * kS2NT), &full[s]);
        }
      }
      if (t + workers < num_tiles) load_idx(t + workers, nxt);
      int* ia = idx_area + s * kS2IdxInts;
      ia[lane] = cur[0];
      ia[32 + lane] = cur[1];
      ia[kS2NT + lane] = cur[2];
      ia[kS2NT + 32 + lane] = cur[3];
      ia[2 * kS2NT + lane] = cur[4];
      ia[2 * kS2NT + 32 + lane] = cur[5];
      mbar_arrive(&full[s]);
#pragma unroll
      for (int k = 0; k < 6; ++k) cur[k] = nxt[k];
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issue
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB, d = i % kS2Groups;
      mbar_wait(&full[s], (i / C::NB) & 1, 32, watch, watch_tag(kWkScore2, kWrMma, kWbFull), s, i);
      mbar_wait(&dempty[d], ((i / kS2Groups) & 1) ^ 1, 32, watch, watch_tag(kWkScore2, kWrMma, kWbDEmpty), d, i);
      tc_fence_after();
      if (elect_one()) {
        issue_tile_mma_sw128<H, kS2NT>(tmem_base, tmem_base + C::D_COL0 + d * kS2NT,
                                       smem_u32(bufs + (size_t)s * T::BUF_BYTES));
        mma_commit(&empty[s]);
        mma_commit(&dfull[d]);
      }
      __syncwarp();
    }
  } else if (warp >= kS2FirstEpiWarp) {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - kS2FirstEpiWarp;
    const int grp = ew >> 3, sub = (ew >> 2) & 1, q = warp & 3;
    const int w8 = ew & 7;                    // warp inside the group: phase B takes edges [8 * w8, 8 * w8 + 8)
    const int my_edge = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);   // see the butterfly below
    const int k = q * 32 + lane;              // hidden unit (TMEM lane)
    const bool unit_ok = k < HS;              // warp-uniform
    const float bias3 = b3[0];
    float* tg = t_s + (size_t)grp * kS2NT * C::TS;
    constexpr unsigned kFull = 0xffffffffu;
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      if (i % kS2Groups != grp) continue;
      const int s = i % C::NB;
      mbar_wait(&full[s], (i / C::NB) & 1, 32, watch, watch_tag(kWkScore2, kWrEpilogue, kWbFull), s, i);
      const int* ia = idx_area + s * kS2IdxInts;
      // copy what phase A / B need out of the stage's index area, then release our share of the stage
      const int my_src = ia[sub * 32 + lane];
      const int my_dst = ia[kS2NT + sub * 32 + lane];
      const int b_eid = ia[2 * kS2NT + w8 * 8 + my_edge];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      mbar_wait(&dfull[grp], (i / kS2Groups) & 1, 32, watch, watch_tag(kWkScore2, kWrEpilogue, kWbDFull), grp, i);
      tc_fence_after();
      // ---- phase A: hidden layer 1 for this warp's 32 edges x 32 units ---------------------------------
      if (unit_ok) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::D_COL0 + grp * kS2NT + sub * 32;
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          uint32_t zr[8];
          tmem_ld8(taddr + b * 8, zr);
          float sv[8], dv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int sj = __shfl_sync(kFull, my_src, b * 8 + u), dj = __shfl_sync(kFull, my_dst, b * 8 + u);
            sv[u] = __ldg(S + (int64_t)sj * 2 * HS + k);
            dv[u] = __ldg(S + (int64_t)dj * 2 * HS + HS + k);
          }
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 8; ++u)
            tg[(sub * 32 + b * 8 + u) * C::TS + k] = fmaxf(__uint_as_float(zr[u]) + sv[u] + dv[u], 0.f);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[grp]);
      named_bar_sync(1 + grp, 256);
      // ---- phase B: lane = second-layer unit m, each warp takes 8 of the tile's 64 edges ------------------------
      // The W2 row of the unit sits in registers (32 columns at a time) and the hidden tile is read with 16-byte
      // broadcast loads: 24 shared-memory wavefronts per edge instead of 72 with four threads per edge.
      {
        const float* trow = tg + (size_t)(w8 * 8) * C::TS;
        float acc[8];
        const float bias2 = b2_s[lane];
#pragma unroll
        for (int ed = 0; ed < 8; ++ed) acc[ed] = bias2;
#pragma unroll 1
        for (int c0 = 0; c0 < HS; c0 += 32) {
          float w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j] = w2t_s[(c0 + j) * 32 + lane];
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            float4 tv[8];
#pragma unroll
            for (int ed = 0; ed < 8; ++ed) tv[ed] = *reinterpret_cast<const float4*>(trow + ed * C::TS + c0 + k4 * 4);
#pragma unroll
            for (int ed = 0; ed < 8; ++ed) acc[ed] = fmaf(w[k4 * 4 + 0], tv[ed].x, acc[ed]);
#pragma unroll
            for (int ed = 0; ed < 8; ++ed) acc[ed] = fmaf(w[k4 * 4 + 1], tv[ed].y, acc[ed]);
#pragma unroll
            for (int ed = 0; ed < 8; ++ed) acc[ed] = fmaf(w[k4 * 4 + 2], tv[ed].z, acc[ed]);
#pragma unroll
            for (int ed = 0; ed < 8; ++ed) acc[ed] = fmaf(w[k4 * 4 + 3], tv[ed].w, acc[ed]);
          }
        }
        // score = W3 . relu(u) + b3: sum over the 32 lanes for 8 edges at once (transposed butterfly, 9 shuffles);
        // afterwards lane 4 * j holds edge j of this warp
        const float w3v = w3_s[lane];
        float v[8];
#pragma unroll
        for (int ed = 0; ed < 8; ++ed) v[ed] = w3v * fmaxf(acc[ed], 0.f);
        const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
        float r4[4], r2[2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float keep = h16 ? v[j + 4] : v[j], send = h16 ? v[j] : v[j + 4];
          r4[j] = keep + __shfl_xor_sync(kFull, send, 16);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float keep = h8 ? r4[j + 2] : r4[j], send = h8 ? r4[j] : r4[j + 2];
          r2[j] = keep + __shfl_xor_sync(kFull, send, 8);
        }
        float r1 = (h4 ? r2[1] : r2[0]) + __shfl_xor_sync(kFull, h4 ? r2[0] : r2[1], 4);
        r1 += __shfl_xor_sync(kFull, r1, 2);
        r1 += __shfl_xor_sync(kFull, r1, 1);
        if ((lane & 3) == 0 && b_eid >= 0) scores[b_eid] = r1 + bias3;
      }
      named_bar_sync(1 + grp, 256);   // t_s of this group is free again
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int H, int HS>
static int score_forward_tc2_impl(const gnb_graph_t* g, const float* S, const void* Wp, const float* W2,
                                  const float* b2, const float* W3, const float* b3, const void* e16, float* scores,
                                  cudaStream_t stream) {
  using C = Score2Cfg<H, HS>;
  cudaError_t err = cudaFuncSetAttribute(score_forward_tc2_kernel<H, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::SMEM);
  if (err != cudaSuccess) {
    set_error("gnb_score_forward_tc2: cudaFuncSetAttribute(%zu): %s", C::SMEM, cudaGetErrorString(err));
    return (int)err;
  }
  const int64_t E = g->num_edges;
  CUtensorMap map_e;
  int rc = make_state_map(&map_e, e16, E, H, kS2NT);
  if (rc) return rc;
  const int64_t num_tiles = (E + kS2NT - 1) / kS2NT;
  int64_t grid = sm_count();
  if (grid > num_tiles) grid = num_tiles;
  score_forward_tc2_kernel<H, HS><<<(unsigned)grid, kS2Threads, C::SMEM, stream>>>(map_e, *g, S, (const __half*)Wp,
                                                                                  W2, b2, W3, b3, scores, watch_get());
  return check_launch("gnb_score_forward_tc2");
}

template <int H>
static int score_forward_tc2_hs(int hs, const gnb_graph_t* g, const float* S, const void* Wp, const float* W2,
                                const float* b2, const float* W3, const float* b3, const void* e16, float* scores,
                                cudaStream_t stream) {
  switch (hs) {
    case 32: return score_forward_tc2_impl<H, 32>(g, S, Wp, W2, b2, W3, b3, e16, scores, stream);
    case 64: return score_forward_tc2_impl<H, 64>(g, S, Wp, W2, b2, W3, b3, e16, scores, stream);
    case 128: return score_forward_tc2_impl<H, 128>(g, S, Wp, W2, b2, W3, b3, e16, scores, stream);
  }
  set_error("hidden_edge_scores=%d unsupported (32, 64, 128)", hs);
  return GNB_E_INVALID;
}

}  // namespace tc
}  // namespace gnb

using namespace gnb;

extern "C" int gnb_score_forward_tc2(const gnb_graph_t* g, int H, int hs, const float* S, const void* Wp,
                                     const float* W2, const float* b2, const float* W3, const float* b3,
                                     const void* e16, float* scores, void* stream) {
  GNB_REQUIRE(g != nullptr && g->num_edges >= 0 && g->in_ptr != nullptr, "graph not staged");
  if (g->num_edges == 0) return 0;
  GNB_REQUIRE(g->in_src && g->in_dst && g->in_eid, "graph not staged");
  GNB_REQUIRE(S && Wp && W2 && b2 && W3 && b3 && e16 && scores, "null pointer");
  GNB_REQUIRE(((uintptr_t)e16 % 16 == 0) && ((uintptr_t)Wp % 16 == 0), "pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  switch (H) {
    case 64: return tc::score_forward_tc2_hs<64>(hs, g, S, Wp, W2, b2, W3, b3, e16, scores, s);
    case 128: return tc::score_forward_tc2_hs<128>(hs, g, S, Wp, W2, b2, W3, b3, e16, scores, s);
    case 256: return tc::score_forward_tc2_hs<256>(hs, g, S, Wp, W2, b2, W3, b3, e16, scores, s);
  }
  set_error("gnb_score_forward_tc2: hidden_features=%d unsupported (64, 128, 256)", H);
  return GNB_E_INVALID;
}
