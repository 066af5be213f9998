// Node-side linear maps on tcgen05 with TMA-fed operands:  out[r][c] = sum_k X[r][k] * W[c][k] + bias[c].
//
// Second generation of gnb_tc_linear.cu (same transposed formulation: the 128-channel weight block of a CTA is
// the MMA A operand and stays in TENSOR MEMORY; a 64-row tile of X is the B operand), but X arrives in the
// split16 format (gnb_tma.cuh): the TMA engine drops the two fp16 images of a row tile straight into
// 128-byte-swizzled shared memory, so there are no producer warps, no conversion instructions and no loads
// held in registers.  Replaces the five node nn.Linear calls of a (Sym)GatedGCN layer
// (reference layers/gated_gcn_full.py:91-96) and the node halves of ScorePredictor.W1 (score_predictor.py:13-14).
//   warp 0      : TMA producer (one lane)
//   warp 1      : MMA issue (one lane): 3 x K/16 tcgen05.mma per tile (Wlo*Xhi + Whi*Xlo + Whi*Xhi, fp32 in TMEM)
//   warps 4..11 : epilogue, two groups of four warps (one per TMEM lane quarter); group g drains accumulator g:
//                 + bias, 128-byte coalesced fp32 stores
#include "gnb_tma.cuh"

namespace gnb {
namespace tc {

constexpr int kLin2NT = 64;
constexpr int kLin2FirstEpiWarp = 4;
constexpr int kLin2Threads = 32 * (kLin2FirstEpiWarp + 8);

template <int K>
struct Lin2Cfg {
  using T = Tile2<K, kLin2NT>;
  static constexpr int NB = (K >= 256) ? 3 : (K >= 128 ? 4 : 6);
  static constexpr uint32_t TMEM_COLS = pow2_cols(2 * T::W_COLS + 2 * kLin2NT);
  static constexpr uint32_t D_COL0 = 2 * T::W_COLS;
  static constexpr size_t SMEM = (size_t)NB * T::BUF_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
};

// kMC: two neighbouring channel blocks of a worker form a 2-CTA cluster; each CTA issues half of the X tile's TMA boxes
// with .multicast::cluster (rank 0 the hi halves, rank 1 the lo halves), so a tile is pulled out of L2 once per pair
// instead of once per channel block -- the kernel is bound by that L2 -> SM traffic (X is re-read by all 5H / 128
// channel blocks).  A stage is written by both CTAs, so its `empty` barrier collects the MMA commits of both.
template <int K, bool kMC>
__global__ void __launch_bounds__(kLin2Threads, 1)
node_linear_tc2_kernel(const __grid_constant__ CUtensorMap map_x, int64_t rows, const __half* __restrict__ Wp, const float* __restrict__ bias, int M,
                       float* __restrict__ out, int64_t ld_out, int nblk, int workers, const Watch watch,
                       const float* __restrict__ out_scale) {
  using C = Lin2Cfg<K>;
  using T = typename C::T;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the shared array (an integer round-trip would demote every later
  // access through these pointers to generic loads)
  uint8_t* bufs = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bufs + (size_t)C::NB * T::BUF_BYTES);
  uint64_t* full = bars;             // [NB] TMA -> MMA
  uint64_t* empty = bars + C::NB;    // [NB] MMA -> TMA
  uint64_t* dfull = empty + C::NB;   // [2]  MMA -> epilogue
  uint64_t* dempty = dfull + 2;      // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

  const int cb = blockIdx.x % nblk, worker = blockIdx.x / nblk;
  const int rank = kMC ? (int)(blockIdx.x & 1) : 0;   // position in the cluster (nblk is even when kMC)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (rows + kLin2NT - 1) / kLin2NT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NB; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kMC ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], 4);
    }
    fence_barrier_init();
    prefetch_tensormap(&map_x);
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if (kMC) cluster_sync_all();   // the peer's multicast must find initialised barriers
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= kLin2FirstEpiWarp && warp < kLin2FirstEpiWarp + 4)
    load_weights_to_tmem<K>(Wp + (size_t)cb * 2 * kM * K, tmem_base, warp & 3, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB;
      mbar_wait(&empty[s], ((i / C::NB) & 1) ^ 1, 64, watch, watch_tag(kWkLinear2, kWrProducer, kWbEmpty), s, i);
      if (elect_one()) {
        uint8_t* stage = bufs + (size_t)s * T::BUF_BYTES;
        mbar_arrive_expect_tx(&full[s], T::BUF_BYTES);
#pragma unroll
        for (int kb = 0; kb < T::KBLOCKS; ++kb) {
          if (kMC) {
            tma_load_2d_mc(stage + rank * T::IMG_BYTES + kb * T::KB_BYTES, &map_x, rank * K + kb * kKB, (int)(t * kLin2NT),
                           &full[s], (uint16_t)3);
          } else {
            tma_load_2d(stage + kb * T::KB_BYTES, &map_x, kb * kKB, (int)(t * kLin2NT), &full[s]);
            tma_load_2d(stage + T::IMG_BYTES + kb * T::KB_BYTES, &map_x, K + kb * kKB, (int)(t * kLin2NT), &full[s]);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issue
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB, d = i & 1;
      mbar_wait(&full[s], (i / C::NB) & 1, 32, watch, watch_tag(kWkLinear2, kWrMma, kWbFull), s, i);
      mbar_wait(&dempty[d], ((i >> 1) & 1) ^ 1, 32, watch, watch_tag(kWkLinear2, kWrMma, kWbDEmpty), d, i);
      tc_fence_after();
      if (elect_one()) {
        issue_tile_mma_sw128<K, kLin2NT>(tmem_base, tmem_base + C::D_COL0 + d * kLin2NT,
                                         smem_u32(bufs + (size_t)s * T::BUF_BYTES));
        if (kMC) mma_commit_mc(&empty[s], (uint16_t)3);   // both CTAs of the pair write into each other's stage
        else mma_commit(&empty[s]);
        mma_commit(&dfull[d]);
      }
      __syncwarp();
    }
  } else if (warp >= kLin2FirstEpiWarp) {
    // ---------------------------------------------------------------- epilogue
    const int g = (warp - kLin2FirstEpiWarp) >> 2, q = warp & 3;
    const int ch = cb * kM + q * 32 + lane;
    const bool ch_ok = ch < M;
    const float b = ch_ok ? bias[ch] : 0.f;
    const float os = out_scale ? *out_scale : 1.f;   // optional per-tensor post-scale (undoes a pre-scale of X)
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      if ((i & 1) != g) continue;
      mbar_wait(&dfull[g], (i >> 1) & 1, 32, watch, watch_tag(kWkLinear2, kWrEpilogue, kWbDFull), g, i);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::D_COL0 + g * kLin2NT;
      tmem_ld32(taddr, v0);
      tmem_ld32(taddr + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[g]);
      const int64_t r0 = t * kLin2NT;
      if (ch_ok) {
        float* o = out + r0 * ld_out + ch;
        if (r0 + kLin2NT <= rows) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j * ld_out] = fmaf(__uint_as_float(v0[j]), os, b);
          o += 32 * ld_out;
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j * ld_out] = fmaf(__uint_as_float(v1[j]), os, b);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (r0 + j < rows) o[j * ld_out] = fmaf(__uint_as_float(v0[j]), os, b);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (r0 + 32 + j < rows) o[(32 + j) * ld_out] = fmaf(__uint_as_float(v1[j]), os, b);
        }
      }
    }
  }
  tc_fence_before();
  if (kMC) cluster_sync_all();   // the peer may still multicast into this CTA's stages / arrive on its barriers
  else __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int K>
static int node_linear_tc2_impl(const void* X16, int64_t rows, const void* Wp, const float* bias, int M, float* out,
                                int64_t ld_out, const float* out_scale, cudaStream_t stream) {
  using C = Lin2Cfg<K>;
  const int nblk_ = (M + kM - 1) / kM;
  const bool mc = nblk_ % 2 == 0;
  auto kern = mc ? node_linear_tc2_kernel<K, true> : node_linear_tc2_kernel<K, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (e != cudaSuccess) {
    set_error("gnb_node_linear_tc2: cudaFuncSetAttribute(%zu): %s", C::SMEM, cudaGetErrorString(e));
    return (int)e;
  }
  CUtensorMap map_x;
  int rc = make_state_map(&map_x, X16, rows, K, kLin2NT);
  if (rc) return rc;
  const int nblk = (M + kM - 1) / kM;
  const int sms = sm_count();
  GNB_REQUIRE(nblk <= sms, "gnb_node_linear_tc2: M=%d needs more channel blocks than SMs", M);
  const int64_t num_tiles = (rows + kLin2NT - 1) / kLin2NT;
  int workers = sms / nblk;
  if (workers > num_tiles) workers = (int)num_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(workers * nblk));
  cfg.blockDim = dim3(kLin2Threads);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = mc ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, map_x, rows, (const __half*)Wp, bias, M, out, ld_out, nblk, workers, watch_get(),
                         out_scale);
  if (e != cudaSuccess) {
    set_error("gnb_node_linear_tc2: launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return check_launch("gnb_node_linear_tc2");
}

}  // namespace tc
}  // namespace gnb

using namespace gnb;

extern "C" int gnb_node_linear_tc2(const void* X16, int64_t rows, int K, const void* Wp, const float* bias, int M,
                                   float* out, int64_t ld_out, const float* out_scale, void* stream) {
  GNB_REQUIRE(M > 0 && ld_out >= M, "gnb_node_linear_tc2: bad output shape (M=%d ld=%lld)", M, (long long)ld_out);
  if (rows == 0) return 0;
  GNB_REQUIRE(X16 && Wp && bias && out, "null pointer");
  GNB_REQUIRE(((uintptr_t)X16 % 16 == 0) && ((uintptr_t)Wp % 16 == 0), "gnb_node_linear_tc2: X16 / Wp must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  switch (K) {
    case 64: return tc::node_linear_tc2_impl<64>(X16, rows, Wp, bias, M, out, ld_out, out_scale, s);
    case 128: return tc::node_linear_tc2_impl<128>(X16, rows, Wp, bias, M, out, ld_out, out_scale, s);
    case 256: return tc::node_linear_tc2_impl<256>(X16, rows, Wp, bias, M, out, ld_out, out_scale, s);
  }
  set_error("gnb_node_linear_tc2: K=%d unsupported (64, 128, 256)", K);
  return GNB_E_INVALID;
}
