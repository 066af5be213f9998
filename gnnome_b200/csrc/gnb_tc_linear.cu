// Tensor-core (tcgen05) node-side linear maps:  out[r][c] = sum_k X[r][k] * W[c][k] + bias[c].
//
// Replaces the five node nn.Linear calls of a (Sym)GatedGCN layer (reference layers/gated_gcn_full.py:91-96,
// one concatenated weight) and the node halves of ScorePredictor.W1 (layers/score_predictor.py:13-14).
// See gnb_tc.cuh for the transposed TMEM-resident-weight formulation and the fp16 (hi, lo) split.
//
// Persistent CTAs, one per SM.  CTA b owns output-channel block cb = b % nblk (128 channels, its W block
// stays in TMEM) and walks row tiles worker, worker + workers, ...  All channel blocks visit the same row
// tile at about the same time, so X is read from HBM once and from L2 nblk times.
//   warp 0        : TMEM allocation, MMA issue (one lane)
//   warps 1..8    : producers  (global fp32 rows -> fp16 hi/lo operand images in shared memory)
//   warps 9..16   : epilogue, two groups of four warps (one warp per TMEM lane quarter), group g drains
//                   accumulator buffer g: +bias, coalesced 128-byte stores
#include "gnb_tc.cuh"

namespace gnb {
namespace tc {

constexpr int kLinNT = 64;
constexpr int kLinProducerWarps = 8;
constexpr int kLinFirstEpiWarp = 1 + kLinProducerWarps;  // 1 (mod 4): 4 consecutive warps cover the 4 TMEM lane quarters
constexpr int kLinThreads = 32 * (kLinFirstEpiWarp + 8);
static_assert(kLinFirstEpiWarp % 4 == 1, "epilogue warp numbering");

template <int K>
struct LinCfg {
  using T = Tile<K, kLinNT>;
  static constexpr int NB = (K >= 256) ? 3 : 4;
  static constexpr uint32_t TMEM_COLS = pow2_cols(2 * T::W_COLS + 2 * kLinNT);
  static constexpr uint32_t D_COL0 = 2 * T::W_COLS;
  static constexpr size_t SMEM = (size_t)NB * T::BUF_BYTES + 256;
};

template <int K>
__global__ void __launch_bounds__(kLinThreads, 1)
node_linear_tc_kernel(const float* __restrict__ X, int64_t rows, const __half* __restrict__ Wp,
                      const float* __restrict__ bias, int M, float* __restrict__ out, int64_t ld_out, int nblk,
                      int workers) {
  using C = LinCfg<K>;
  using T = typename C::T;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* bufs = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)C::NB * T::BUF_BYTES);
  uint64_t* full = bars;             // [NB] producers -> MMA
  uint64_t* empty = bars + C::NB;    // [NB] MMA -> producers
  uint64_t* dfull = empty + C::NB;   // [2]  MMA -> epilogue
  uint64_t* dempty = dfull + 2;      // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

  const int cb = blockIdx.x % nblk, worker = blockIdx.x / nblk;
  if (worker >= workers) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (rows + kLinNT - 1) / kLinNT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NB; ++i) {
      mbar_init(&full[i], kLinProducerWarps * 32);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kLinFirstEpiWarp && warp < kLinFirstEpiWarp + 4) {  // first epilogue group doubles as the weight loader
    load_weights_to_tmem<K>(Wp + (size_t)cb * 2 * kM * K, tmem_base, warp & 3, lane);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ---------------------------------------------------------------- MMA issue
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB, d = i & 1;
      mbar_wait(&full[s], (i / C::NB) & 1);
      mbar_wait(&dempty[d], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        issue_tile_mma<K, kLinNT>(tmem_base, tmem_base + C::D_COL0 + d * kLinNT,
                                  smem_u32(bufs + (size_t)s * T::BUF_BYTES));
        mma_commit(&empty[s]);
        mma_commit(&dfull[d]);
      }
      __syncwarp();
    }
  } else if (warp <= kLinProducerWarps) {
    // ---------------------------------------------------------------- producers
    const int pw = warp - 1;
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      const int s = i % C::NB;
      mbar_wait(&empty[s], ((i / C::NB) & 1) ^ 1);
      produce_tile<K, kLinNT, kLinProducerWarps>(X, rows, t * kLinNT, bufs + (size_t)s * T::BUF_BYTES, pw, lane);
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int g = (warp - kLinFirstEpiWarp) >> 2, q = warp & 3;
    const int ch = cb * kM + q * 32 + lane;
    const bool ch_ok = ch < M;
    const float b = ch_ok ? bias[ch] : 0.f;
    int i = 0;
    for (int64_t t = worker; t < num_tiles; t += workers, ++i) {
      if ((i & 1) != g) continue;
      mbar_wait(&dfull[g], (i >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::D_COL0 + g * kLinNT;
      tmem_ld32(taddr, v0);
      tmem_ld32(taddr + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty[g]);
      const int64_t r0 = t * kLinNT;
      if (ch_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (r0 + j < rows) out[(r0 + j) * ld_out + ch] = __uint_as_float(v0[j]) + b;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (r0 + 32 + j < rows) out[(r0 + 32 + j) * ld_out + ch] = __uint_as_float(v1[j]) + b;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// W[M][K] fp32 (nn.Linear layout) -> Wp[ceil(M/128)][2][128][K] fp16: (hi, lo) images of 16 * W, zero padded
__global__ void pack_linear_tc_kernel(const float* __restrict__ W, int M, int K, __half* __restrict__ Wp, int nblk) {
  const int64_t total = (int64_t)nblk * kM * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int64_t row = i / K;
    const int blk = (int)(row / kM), r = (int)(row % kM);
    const float w = (row < M) ? W[row * K + k] * kWScale : 0.f;
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    Wp[((size_t)(blk * 2 + 0) * kM + r) * K + k] = hi;
    Wp[((size_t)(blk * 2 + 1) * kM + r) * K + k] = lo;
  }
}

template <int K>
static int node_linear_tc_impl(const float* X, int64_t rows, const void* Wp, const float* bias, int M, float* out,
                               int64_t ld_out, cudaStream_t stream) {
  using C = LinCfg<K>;
  cudaError_t e = cudaFuncSetAttribute(node_linear_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::SMEM);
  if (e != cudaSuccess) {
    set_error("gnb_node_linear_tc: cudaFuncSetAttribute(%zu): %s", C::SMEM, cudaGetErrorString(e));
    return (int)e;
  }
  const int nblk = (M + kM - 1) / kM;
  const int sms = sm_count();
  GNB_REQUIRE(nblk <= sms, "gnb_node_linear_tc: M=%d needs more channel blocks than SMs", M);
  const int64_t num_tiles = (rows + kLinNT - 1) / kLinNT;
  int workers = sms / nblk;
  if (workers > num_tiles) workers = (int)num_tiles;
  node_linear_tc_kernel<K><<<workers * nblk, kLinThreads, C::SMEM, stream>>>(X, rows, (const __half*)Wp, bias, M, out,
                                                                             ld_out, nblk, workers);
  return check_launch("gnb_node_linear_tc");
}

}  // namespace tc
}  // namespace gnb

using namespace gnb;

extern "C" size_t gnb_packed_linear_bytes(int M, int K) {
  if (M <= 0 || K <= 0) return 0;
  return (size_t)((M + tc::kM - 1) / tc::kM) * 2 * tc::kM * K * sizeof(__half);
}

extern "C" int gnb_pack_linear_tc(const float* W, int M, int K, void* Wp, void* stream) {
  GNB_REQUIRE(W && Wp && M > 0 && K > 0 && K % 16 == 0, "gnb_pack_linear_tc: bad arguments (M=%d K=%d)", M, K);
  const int nblk = (M + tc::kM - 1) / tc::kM;
  const int64_t total = (int64_t)nblk * tc::kM * K;
  unsigned blocks = (unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  tc::pack_linear_tc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, M, K, (__half*)Wp, nblk);
  return check_launch("gnb_pack_linear_tc");
}

extern "C" int gnb_node_linear_tc(const float* X, int64_t rows, int K, const void* Wp, const float* bias, int M,
                                  float* out, int64_t ld_out, void* stream) {
  GNB_REQUIRE(M > 0 && ld_out >= M, "gnb_node_linear_tc: bad output shape (M=%d ld=%lld)", M, (long long)ld_out);
  if (rows == 0) return 0;
  GNB_REQUIRE(X && Wp && bias && out, "null pointer");
  GNB_REQUIRE(((uintptr_t)X % 32 == 0) && ((uintptr_t)Wp % 16 == 0), "gnb_node_linear_tc: X must be 32-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  switch (K) {
    case 32: return tc::node_linear_tc_impl<32>(X, rows, Wp, bias, M, out, ld_out, s);
    case 64: return tc::node_linear_tc_impl<64>(X, rows, Wp, bias, M, out, ld_out, s);
    case 128: return tc::node_linear_tc_impl<128>(X, rows, Wp, bias, M, out, ld_out, s);
    case 256: return tc::node_linear_tc_impl<256>(X, rows, Wp, bias, M, out, ld_out, s);
  }
  set_error("gnb_node_linear_tc: K=%d unsupported (32, 64, 128, 256)", K);
  return GNB_E_INVALID;
}
