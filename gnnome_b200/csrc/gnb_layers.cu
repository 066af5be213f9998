// GatedGCN layer kernels, CUDA-core (fp32 FFMA) edition.
//
// Layout idea shared by every edge kernel here: ONE THREAD PER CHANNEL, walking edge positions
// sequentially.  A group of H threads owns a chunk of consecutive dst-sorted positions; because all
// threads of the group see the same (src, dst) stream, per-destination sums are plain register
// accumulators and segment boundaries are group-uniform branches -- no atomics, no shuffles, and a
// bit-reproducible summation order (position order).  Gathers of node-table rows are 128-byte
// coalesced per warp (32 consecutive channels of one row).
#include <cuda_fp16.h>

#include "gnb_common.cuh"

namespace gnb {

// ------------------------------------------------------------------------------------------------
// Encoders: out[r] = W2 * relu(W1 * in[idx[r]] + b1) + b2     (models/full_graph.py:26-27)
// ------------------------------------------------------------------------------------------------
constexpr int kEncRows = 64;

__global__ void __launch_bounds__(kThreads)
encode_kernel(const float* __restrict__ in, const int32_t* __restrict__ idx, int64_t rows, int in_f,
              int hid, int H, const float* __restrict__ W1, const float* __restrict__ b1,
              const float* __restrict__ W2t, const float* __restrict__ b2, float* __restrict__ out) {
  extern __shared__ float smem[];
  float* w2 = smem;                    // [hid][H]
  float* w1 = w2 + hid * H;            // [hid][in_f]
  float* bb1 = w1 + hid * in_f;        // [hid]
  float* hdn = bb1 + hid;              // [kEncRows][hid]
  for (int i = threadIdx.x; i < hid * H; i += kThreads) w2[i] = W2t[i];
  for (int i = threadIdx.x; i < hid * in_f; i += kThreads) w1[i] = W1[i];
  for (int i = threadIdx.x; i < hid; i += kThreads) bb1[i] = b1[i];
  __syncthreads();
  const int64_t num_tiles = (rows + kEncRows - 1) / kEncRows;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kEncRows;
    for (int i = threadIdx.x; i < kEncRows * hid; i += kThreads) {
      int r = i / hid, j = i - r * hid;
      float acc = 0.f;
      if (r0 + r < rows) {
        int64_t row = idx ? (int64_t)idx[r0 + r] : (r0 + r);
        acc = bb1[j];
        for (int f = 0; f < in_f; ++f) acc = fmaf(in[row * in_f + f], w1[j * in_f + f], acc);
        acc = fmaxf(acc, 0.f);
      }
      hdn[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kEncRows * H; i += kThreads) {
      int r = i / H, c = i - r * H;
      if (r0 + r < rows) {
        float acc = b2[c];
        for (int j = 0; j < hid; ++j) acc = fmaf(hdn[r * hid + j], w2[j * H + c], acc);
        out[(r0 + r) * H + c] = acc;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Node projections: out[rows][M] = A[rows][K] * Wt[K][M] + bias   (gated_gcn_full.py:91-96)
// 128x128x16 tiles, 8x8 register micro-tiles, fp32 FFMA.
// ------------------------------------------------------------------------------------------------
constexpr int kBM = 128, kBN = 128, kBK = 16;

__global__ void __launch_bounds__(kThreads)
node_linear_kernel(const float* __restrict__ A, int64_t rows, int K, const float* __restrict__ Wt,
                   const float* __restrict__ bias, int M, float* __restrict__ out, int64_t ld_out) {
  __shared__ __align__(16) float As[kBK][kBM + 4];
  __shared__ __align__(16) float Bs[kBK][kBN];
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * kBM;
  const int col0 = blockIdx.y * kBN;
  const int ty = tid / 16, tx = tid % 16;  // 16 x 16 threads, each 8 rows x 8 cols (cols split 4+4)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // loader mapping: A tile 128 rows x 16 k = 512 float4 (2 per thread); B tile 16 k x 128 cols = 512 float4
  for (int k0 = 0; k0 < K; k0 += kBK) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int f = tid + it * kThreads;       // 0..511
      int r = f / 4, kq = (f % 4) * 4;   // row in tile, k offset
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < rows) v = *reinterpret_cast<const float4*>(A + (row0 + r) * K + k0 + kq);
      As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
      int kk = f / 32, cq = (f % 32) * 4;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col0 + cq < M) w = *reinterpret_cast<const float4*>(Wt + (int64_t)(k0 + kk) * M + col0 + cq);
      *reinterpret_cast<float4*>(&Bs[kk][cq]) = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float a[8], b[8];
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t r = row0 + ty * 8 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      int c = col0 + half * 64 + tx * 4;
      if (c < M) {
        float4 bv = *reinterpret_cast<const float4*>(bias + c);
        float4 o = make_float4(acc[i][half * 4 + 0] + bv.x, acc[i][half * 4 + 1] + bv.y,
                               acc[i][half * 4 + 2] + bv.z, acc[i][half * 4 + 3] + bv.w);
        *reinterpret_cast<float4*>(out + r * ld_out + c) = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tile GEMM used inside the edge kernels: z_s[R][NOUT] = a_s[R][K] * Wt[K][NOUT]  (Wt in global/L2)
// warp w owns rows [w*R/8, (w+1)*R/8), lane l owns columns {l + 32 i}.  All 256 threads call it.
// ------------------------------------------------------------------------------------------------
template <int R, int K, int NOUT>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ a_s, const float* __restrict__ Wt,
                                          float* __restrict__ w_s /* [16][NOUT] */,
                                          float* __restrict__ z_s) {
  constexpr int RPW = R / 8, CPL = NOUT / 32, KC = 16;
  static_assert(R % 8 == 0 && NOUT % 32 == 0 && K % KC == 0, "tile_gemm shape");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[RPW][CPL];
#pragma unroll
  for (int i = 0; i < RPW; ++i)
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KC) {
    __syncthreads();  // previous chunk fully consumed (and a_s / z_s hazards of the caller)
    for (int i = threadIdx.x; i < KC * NOUT / 4; i += kThreads)
      reinterpret_cast<float4*>(w_s)[i] = reinterpret_cast<const float4*>(Wt + (size_t)k0 * NOUT)[i];
    __syncthreads();
#pragma unroll
    for (int kq = 0; kq < KC; kq += 4) {
      float b[4][CPL];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int j = 0; j < CPL; ++j) b[kk][j] = w_s[(kq + kk) * NOUT + lane + 32 * j];
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(a_s + (warp * RPW + i) * K + k0 + kq);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          acc[i][j] = fmaf(a.x, b[0][j], acc[i][j]);
          acc[i][j] = fmaf(a.y, b[1][j], acc[i][j]);
          acc[i][j] = fmaf(a.z, b[2][j], acc[i][j]);
          acc[i][j] = fmaf(a.w, b[3][j], acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i)
#pragma unroll
    for (int j = 0; j < CPL; ++j) z_s[(warp * RPW + i) * NOUT + lane + 32 * j] = acc[i][j];
}

// ------------------------------------------------------------------------------------------------
// K2: fused edge pass over the dst-CSR (gated_gcn_full.py:97,104-114)
// ------------------------------------------------------------------------------------------------
template <int H>
struct EdgeCfg {
  static constexpr int G = kThreads / H;     // channel groups per CTA (each owns its own chunk)
  static constexpr int R = kTileRows * G;    // edge rows per CTA tile
  static constexpr size_t smem_floats = (size_t)2 * R * H + 16 * H;
  static constexpr size_t smem_bytes = smem_floats * 4 + (size_t)2 * R * 4;
};

template <int H>
__global__ void __launch_bounds__(kThreads)
edge_forward_kernel(gnb_graph_t g, const float* __restrict__ P, int64_t ldP,
                    const float* __restrict__ We_t, const float* __restrict__ scale_e,
                    const float* __restrict__ shift_e, float* __restrict__ e, float* __restrict__ F,
                    float* __restrict__ carry, int flags) {
  using C = EdgeCfg<H>;
  extern __shared__ __align__(16) float smem[];
  float* e_s = smem;                      // [R][H] layer input rows (residual + GEMM operand)
  float* z_s = e_s + C::R * H;            // [R][H] e * We_t
  float* w_s = z_s + C::R * H;            // [16][H] weight chunk
  int* src_s = reinterpret_cast<int*>(w_s + 16 * H);  // [R]
  int* dst_s = src_s + C::R;                          // [R]

  const int64_t E = g.num_edges;
  const int64_t num_chunks = (E + kChunk - 1) / kChunk;
  const int64_t num_items = (num_chunks + C::G - 1) / C::G;
  const int grp = threadIdx.x / H, c = threadIdx.x % H;
  const float sc = scale_e[c], sh = shift_e[c];
  const bool residual = flags & GNB_F_RESIDUAL;

  for (int64_t item = blockIdx.x; item < num_items; item += gridDim.x) {
    const int64_t chunk = item * C::G + grp;                 // this group's chunk
    const int64_t cs = chunk * kChunk;
    const int64_t ce = (cs + kChunk < E) ? cs + kChunk : E;  // may be <= cs when the chunk is void
    const bool live = cs < E;
    int head_dst = -1, tail_dst = -1;
    if (live) {
      if (cs > 0 && g.in_dst[cs - 1] == g.in_dst[cs]) head_dst = g.in_dst[cs];
      if (ce < E && g.in_dst[ce] == g.in_dst[ce - 1]) tail_dst = g.in_dst[ce - 1];
    }
    int cur = -1;
    float num = 0.f, den = 0.f, b2 = 0.f;

    for (int t = 0; t < kTilesPerChunk; ++t) {
      __syncthreads();  // previous tile's phase B is done with e_s / z_s / src_s
      // ---- stage the tile: each group brings kTileRows rows of its own chunk -------------------
      for (int i = threadIdx.x; i < C::R * (H / 4); i += kThreads) {
        int row = i / (H / 4), q4 = i - row * (H / 4);
        int gg = row / kTileRows;
        int64_t p = (item * C::G + gg) * (int64_t)kChunk + t * kTileRows + (row - gg * kTileRows);
        int64_t pe = ((item * C::G + gg) * (int64_t)kChunk + kChunk < E)
                         ? (item * C::G + gg) * (int64_t)kChunk + kChunk : E;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < pe) v = *reinterpret_cast<const float4*>(e + p * H + q4 * 4);
        *reinterpret_cast<float4*>(e_s + row * H + q4 * 4) = v;
      }
      for (int row = threadIdx.x; row < C::R; row += kThreads) {
        int gg = row / kTileRows;
        int64_t base = (item * C::G + gg) * (int64_t)kChunk;
        int64_t p = base + t * kTileRows + (row - gg * kTileRows);
        int64_t pe = (base + kChunk < E) ? base + kChunk : E;
        src_s[row] = (p < pe) ? g.in_src[p] : -1;
        dst_s[row] = (p < pe) ? g.in_dst[p] : -1;
      }
      // ---- z = e * We_t (first __syncthreads inside makes the staging visible) -----------------
      tile_gemm<C::R, H, H>(e_s, We_t, w_s, z_s);
      __syncthreads();
      // ---- phase B: one thread per channel walks its group's rows -----------------------------
      const int64_t p0 = cs + (int64_t)t * kTileRows;
      if (live && p0 < ce) {
        const int nrows = (ce - p0 < kTileRows) ? (int)(ce - p0) : kTileRows;
        const int rbase = grp * kTileRows;
        for (int j0 = 0; j0 < nrows; j0 += 8) {
          float2 ba[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            int s = (j0 + u < nrows) ? src_s[rbase + j0 + u] : -1;
            ba[u] = (s >= 0) ? __ldg(reinterpret_cast<const float2*>(P + (int64_t)s * ldP) + c)
                             : make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = j0 + u;
            if (j < nrows) {
              const int d = dst_s[rbase + j];
              if (d != cur) {
                if (cur >= 0) {
                  if (cur == head_dst) {
                    carry[(chunk * 4 + 0) * H + c] = num;
                    carry[(chunk * 4 + 1) * H + c] = den;
                  } else {
                    F[(int64_t)cur * H + c] = num / (den + kGateEps);
                  }
                }
                cur = d;
                num = 0.f;
                den = 0.f;
                b2 = __ldg(P + (int64_t)d * ldP + 2 * H + c);
              }
              const float ein = e_s[(rbase + j) * H + c];
              float v = fmaf(z_s[(rbase + j) * H + c] + ba[u].x + b2, sc, sh);
              v = fmaxf(v, 0.f);
              if (residual) v += ein;
              e[(p0 + j) * H + c] = v;
              const float sg = sigmoidf_fast(v);
              num = fmaf(sg, ba[u].y, num);
              den += sg;
            }
          }
        }
      }
    }
    // ---- close the segment that is open at the end of the chunk ----------------------------------
    if (live && cur >= 0) {
      if (cur == tail_dst) {
        carry[(chunk * 4 + 2) * H + c] = num;
        carry[(chunk * 4 + 3) * H + c] = den;
      } else if (cur == head_dst) {
        carry[(chunk * 4 + 0) * H + c] = num;
        carry[(chunk * 4 + 1) * H + c] = den;
      } else {
        F[(int64_t)cur * H + c] = num / (den + kGateEps);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3: reverse aggregation over the src-CSR + node update (gated_gcn_full.py:124-137)
// H/4 threads per node, float4 per thread.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_sigmoid(float4 a) {
  return make_float4(sigmoidf_fast(a.x), sigmoidf_fast(a.y), sigmoidf_fast(a.z), sigmoidf_fast(a.w));
}
__device__ __forceinline__ float4 f4_gate_div(float4 n, float4 d) {
  return make_float4(n.x / (d.x + kGateEps), n.y / (d.y + kGateEps), n.z / (d.z + kGateEps),
                     n.w / (d.w + kGateEps));
}

template <int H>
__global__ void __launch_bounds__(kThreads)
node_update_kernel(gnb_graph_t g, const float* __restrict__ P, int64_t ldP, const float* __restrict__ e,
                   const float* __restrict__ F, const float* __restrict__ carry,
                   const float* __restrict__ h_in, const float* __restrict__ scale_h,
                   const float* __restrict__ shift_h, float* __restrict__ h_out, int flags, int chunk,
                   int64_t node_begin, int64_t node_end, const int32_t* __restrict__ xp_ptr,
                   const int32_t* __restrict__ xp_row, const float* __restrict__ xp_buf,
                   float* __restrict__ partial_out) {
  constexpr int TPN = H / 4;               // threads per node
  constexpr int NPB = kThreads / TPN;      // nodes in flight per CTA
  const int t = threadIdx.x % TPN, slot = threadIdx.x / TPN;
  const bool sym = flags & GNB_F_SYMMETRIC, residual = flags & GNB_F_RESIDUAL;
  const int a1_off = sym ? 4 * H : 3 * H;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 sc = partial_out ? zero : reinterpret_cast<const float4*>(scale_h)[t];
  const float4 sh = partial_out ? zero : reinterpret_cast<const float4*>(shift_h)[t];
  for (int64_t i = node_begin + (int64_t)blockIdx.x * NPB + slot; i < node_end; i += (int64_t)gridDim.x * NPB) {
    // ---- Bk: gate-normalised sum over out-edges ------------------------------------------------
    float4 bk = zero;
    if (sym || partial_out) {
      float4 num = zero, den = zero;
      const int qa = g.out_ptr[i], qb = g.out_ptr[i + 1];
      for (int q0 = qa; q0 < qb; q0 += 4) {
        float4 ev[4], av[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (q0 + u < qb) {
            const int64_t p = g.out_pos[q0 + u];
            const int64_t d = g.out_dst[q0 + u];
            ev[u] = reinterpret_cast<const float4*>(e + p * H)[t];
            av[u] = __ldg(reinterpret_cast<const float4*>(P + d * ldP + 3 * H) + t);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (q0 + u < qb) {
            const float4 sg = f4_sigmoid(ev[u]);
            num = f4_fma(sg, av[u], num);
            den = f4_add(den, sg);
          }
        }
      }
      if (partial_out) {  // multi-GPU: un-normalised partial sums of a halo source node, for its owner
        reinterpret_cast<float4*>(partial_out + (i - node_begin) * 2 * H)[t] = num;
        reinterpret_cast<float4*>(partial_out + (i - node_begin) * 2 * H + H)[t] = den;
        continue;
      }
      if (xp_ptr) {       // multi-GPU: partial sums other ranks computed for this node, in rank order
        for (int r = xp_ptr[i], re = xp_ptr[i + 1]; r < re; ++r) {
          const float* row = xp_buf + (int64_t)xp_row[r] * 2 * H;
          num = f4_add(num, reinterpret_cast<const float4*>(row)[t]);
          den = f4_add(den, reinterpret_cast<const float4*>(row + H)[t]);
        }
      }
      bk = f4_gate_div(num, den);
    }
    // ---- F: from K2, resolving chunk-straddling segments ------------------------------------------
    float4 f = zero;
    const int pa = g.in_ptr[i], pb = g.in_ptr[i + 1];
    if (pb > pa) {
      const int c0 = pa / chunk, c1 = (pb - 1) / chunk;
      if (c0 == c1) {
        f = reinterpret_cast<const float4*>(F + i * H)[t];
      } else {
        float4 num = zero, den = zero;
        for (int k = c0; k < c1; ++k) {
          num = f4_add(num, reinterpret_cast<const float4*>(carry + ((int64_t)k * 4 + 2) * H)[t]);
          den = f4_add(den, reinterpret_cast<const float4*>(carry + ((int64_t)k * 4 + 3) * H)[t]);
        }
        num = f4_add(num, reinterpret_cast<const float4*>(carry + ((int64_t)c1 * 4 + 0) * H)[t]);
        den = f4_add(den, reinterpret_cast<const float4*>(carry + ((int64_t)c1 * 4 + 1) * H)[t]);
        f = f4_gate_div(num, den);
      }
    }
    // ---- h' = relu(bn_h(A1h + F + Bk)) + h ---------------------------------------------------------
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(P + i * ldP + a1_off) + t);
    float4 u = f4_add(f4_add(a1, f), bk);
    u = f4_fma(u, sc, sh);
    u = make_float4(fmaxf(u.x, 0.f), fmaxf(u.y, 0.f), fmaxf(u.z, 0.f), fmaxf(u.w, 0.f));
    if (residual) u = f4_add(u, reinterpret_cast<const float4*>(h_in + i * H)[t]);
    reinterpret_cast<float4*>(h_out + i * H)[t] = u;
  }
}

// ------------------------------------------------------------------------------------------------
// Score predictor (score_predictor.py:12-24) with the W1 split
// ------------------------------------------------------------------------------------------------
template <int H, int HS>
struct ScoreCfg {
  static constexpr int R = 8192 / H;  // rows per tile (same 32 KB e tile as K2)
  static constexpr int TS = HS + 1;   // padded stride of the hidden tile
  static constexpr size_t smem_bytes =
      ((size_t)R * H + 16 * HS + (size_t)R * HS + (size_t)R * TS + 32 * TS + 32 + 32) * 4 + (size_t)3 * R * 4;
};

// kSplit16: e points at the split16 edge state (fp16 rows [hi 0..H | lo 0..H]; x = 16 * (hi + lo))
template <int H, int HS, bool kSplit16>
__global__ void __launch_bounds__(kThreads)
score_forward_kernel(gnb_graph_t g, const float* __restrict__ S, const float* __restrict__ W1e_t,
                     const float* __restrict__ W2, const float* __restrict__ b2,
                     const float* __restrict__ W3, const float* __restrict__ b3,
                     const float* __restrict__ e, float* __restrict__ scores) {
  using C = ScoreCfg<H, HS>;
  extern __shared__ __align__(16) float smem[];
  float* e_s = smem;                       // [R][H]
  float* w_s = e_s + C::R * H;             // [16][HS]
  float* z_s = w_s + 16 * HS;              // [R][HS]   e * W1e_t
  float* t_s = z_s + C::R * HS;            // [R][TS]   relu(hidden)
  float* w2_s = t_s + C::R * C::TS;        // [32][TS]
  float* b2_s = w2_s + 32 * C::TS;         // [32]
  float* w3_s = b2_s + 32;                 // [32]
  int* src_s = reinterpret_cast<int*>(w3_s + 32);
  int* dst_s = src_s + C::R;
  int* eid_s = dst_s + C::R;

  for (int i = threadIdx.x; i < 32 * HS; i += kThreads) w2_s[(i / HS) * C::TS + (i % HS)] = W2[i];
  if (threadIdx.x < 32) {
    b2_s[threadIdx.x] = b2[threadIdx.x];
    w3_s[threadIdx.x] = W3[threadIdx.x];
  }
  const float bias3 = b3[0];
  const int64_t E = g.num_edges;
  const int64_t num_tiles = (E + C::R - 1) / C::R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t p0 = tile * C::R;
    __syncthreads();
    for (int i = threadIdx.x; i < C::R * (H / 4); i += kThreads) {
      int row = i / (H / 4), q4 = i - row * (H / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p0 + row < E) {
        if (kSplit16) {
          const __half* e16 = reinterpret_cast<const __half*>(e);
          const uint2 hh = *reinterpret_cast<const uint2*>(e16 + (p0 + row) * 2 * H + q4 * 4);
          const uint2 ll = *reinterpret_cast<const uint2*>(e16 + (p0 + row) * 2 * H + H + q4 * 4);
          const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&hh.x));
          const float2 h1 = __half22float2(*reinterpret_cast<const __half2*>(&hh.y));
          const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&ll.x));
          const float2 l1 = __half22float2(*reinterpret_cast<const __half2*>(&ll.y));
          v = make_float4((h0.x + l0.x) * 16.f, (h0.y + l0.y) * 16.f, (h1.x + l1.x) * 16.f, (h1.y + l1.y) * 16.f);
        } else {
          v = *reinterpret_cast<const float4*>(e + (p0 + row) * H + q4 * 4);
        }
      }
      *reinterpret_cast<float4*>(e_s + row * H + q4 * 4) = v;
    }
    for (int row = threadIdx.x; row < C::R; row += kThreads) {
      bool ok = p0 + row < E;
      src_s[row] = ok ? g.in_src[p0 + row] : -1;
      dst_s[row] = ok ? g.in_dst[p0 + row] : -1;
      eid_s[row] = ok ? g.in_eid[p0 + row] : -1;
    }
    tile_gemm<C::R, H, HS>(e_s, W1e_t, w_s, z_s);
    __syncthreads();
    // hidden = relu(S[src][0:HS] + S[dst][HS:2HS] + z): one warp per row, coalesced
    for (int row = warp; row < C::R; row += kThreads / 32) {
      const int s = src_s[row], d = dst_s[row];
      if (s < 0) continue;
#pragma unroll
      for (int k = lane; k < HS; k += 32) {
        float v = z_s[row * HS + k] + __ldg(S + (int64_t)s * 2 * HS + k) + __ldg(S + (int64_t)d * 2 * HS + HS + k);
        t_s[row * C::TS + k] = fmaxf(v, 0.f);
      }
    }
    __syncthreads();
    // 8 threads per row, 4 of the 32 second-layer units each
    for (int rb = 0; rb < C::R; rb += kThreads / 8) {
      const int row = rb + threadIdx.x / 8, sub = threadIdx.x % 8;
      float part = 0.f;
      if (row < C::R && src_s[row] >= 0) {
        float u0 = b2_s[sub], u1 = b2_s[sub + 8], u2 = b2_s[sub + 16], u3 = b2_s[sub + 24];
#pragma unroll 8
        for (int k = 0; k < HS; ++k) {
          const float tv = t_s[row * C::TS + k];
          u0 = fmaf(w2_s[(sub)*C::TS + k], tv, u0);
          u1 = fmaf(w2_s[(sub + 8) * C::TS + k], tv, u1);
          u2 = fmaf(w2_s[(sub + 16) * C::TS + k], tv, u2);
          u3 = fmaf(w2_s[(sub + 24) * C::TS + k], tv, u3);
        }
        part = w3_s[sub] * fmaxf(u0, 0.f) + w3_s[sub + 8] * fmaxf(u1, 0.f) +
               w3_s[sub + 16] * fmaxf(u2, 0.f) + w3_s[sub + 24] * fmaxf(u3, 0.f);
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (sub == 0 && row < C::R && src_s[row] >= 0) scores[eid_s[row]] = part + bias3;
    }
  }
}

template <typename Kern>
static int launch_cfg(Kern kern, size_t smem, int* blocks_per_sm) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(e));
    return (int)e;
  }
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, kThreads, smem);
  if (e != cudaSuccess || *blocks_per_sm < 1) {
    set_error("kernel does not fit on an SM (smem %zu): %s", smem, cudaGetErrorString(e));
    return e != cudaSuccess ? (int)e : GNB_E_INVALID;
  }
  return 0;
}

static unsigned grid_for(int64_t items, int blocks_per_sm) {
  int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  return (unsigned)(items < cap ? (items > 0 ? items : 1) : cap);
}

template <int H>
static int edge_forward_impl(const gnb_graph_t* g, const float* P, int64_t ldP, const float* We_t,
                             const float* scale_e, const float* shift_e, float* e, float* F,
                             float* carry, int flags, cudaStream_t stream) {
  using C = EdgeCfg<H>;
  int bps = 0;
  int rc = launch_cfg(edge_forward_kernel<H>, C::smem_bytes, &bps);
  if (rc) return rc;
  int64_t chunks = (g->num_edges + kChunk - 1) / kChunk;
  int64_t items = (chunks + C::G - 1) / C::G;
  edge_forward_kernel<H><<<grid_for(items, bps), kThreads, C::smem_bytes, stream>>>(
      *g, P, ldP, We_t, scale_e, shift_e, e, F, carry, flags);
  return check_launch("gnb_edge_forward");
}

template <int H>
static int node_update_impl(const gnb_graph_t* g, const float* P, int64_t ldP, const float* e,
                            const float* F, const float* carry, const float* h_in,
                            const float* scale_h, const float* shift_h, float* h_out, int flags, int chunk,
                            int64_t node_begin, int64_t node_end, const int32_t* xp_ptr, const int32_t* xp_row,
                            const float* xp_buf, float* partial_out, cudaStream_t stream) {
  int bps = 0;
  int rc = launch_cfg(node_update_kernel<H>, 0, &bps);
  if (rc) return rc;
  constexpr int NPB = kThreads / (H / 4);
  int64_t items = (node_end - node_begin + NPB - 1) / NPB;
  node_update_kernel<H><<<grid_for(items, bps * 4), kThreads, 0, stream>>>(
      *g, P, ldP, e, F, carry, h_in, scale_h, shift_h, h_out, flags, chunk, node_begin, node_end, xp_ptr, xp_row,
      xp_buf, partial_out);
  return check_launch(partial_out ? "gnb_reverse_partial" : "gnb_node_update");
}

template <int H, int HS, bool kSplit16>
static int score_forward_impl(const gnb_graph_t* g, const float* S, const float* W1e_t, const float* W2,
                              const float* b2, const float* W3, const float* b3, const float* e,
                              float* scores, cudaStream_t stream) {
  using C = ScoreCfg<H, HS>;
  int bps = 0;
  int rc = launch_cfg(score_forward_kernel<H, HS, kSplit16>, C::smem_bytes, &bps);
  if (rc) return rc;
  int64_t tiles = (g->num_edges + C::R - 1) / C::R;
  score_forward_kernel<H, HS, kSplit16><<<grid_for(tiles, bps), kThreads, C::smem_bytes, stream>>>(
      *g, S, W1e_t, W2, b2, W3, b3, e, scores);
  return check_launch("gnb_score_forward");
}

template <int H, bool kSplit16 = false>
static int score_forward_hs(int hs, const gnb_graph_t* g, const float* S, const float* W1e_t,
                            const float* W2, const float* b2, const float* W3, const float* b3,
                            const float* e, float* scores, cudaStream_t stream) {
  switch (hs) {
    case 32: return score_forward_impl<H, 32, kSplit16>(g, S, W1e_t, W2, b2, W3, b3, e, scores, stream);
    case 64: return score_forward_impl<H, 64, kSplit16>(g, S, W1e_t, W2, b2, W3, b3, e, scores, stream);
    case 128: return score_forward_impl<H, 128, kSplit16>(g, S, W1e_t, W2, b2, W3, b3, e, scores, stream);
  }
  set_error("hidden_edge_scores=%d unsupported (32, 64, 128)", hs);
  return GNB_E_INVALID;
}

}  // namespace gnb

using namespace gnb;

#define GNB_DISPATCH_H(H, CALL)                                 \
  switch (H) {                                                  \
    case 32: { constexpr int kH = 32; return CALL; }            \
    case 64: { constexpr int kH = 64; return CALL; }            \
    case 128: { constexpr int kH = 128; return CALL; }          \
    case 256: { constexpr int kH = 256; return CALL; }          \
    default: break;                                             \
  }                                                             \
  set_error("hidden_features=%d unsupported (32, 64, 128, 256)", H); \
  return GNB_E_INVALID

extern "C" int gnb_edge_chunk(int H) { return supported_h(H) ? kChunk : GNB_E_INVALID; }

extern "C" int gnb_encode(const float* in, const int32_t* idx, int64_t rows, int in_f, int hid, int H,
                          const float* W1, const float* b1, const float* W2t, const float* b2,
                          float* out, void* stream) {
  GNB_REQUIRE(in_f > 0 && hid > 0 && H > 0, "bad encoder shape");
  if (rows == 0) return 0;
  GNB_REQUIRE(in && W1 && b1 && W2t && b2 && out, "null pointer");
  size_t smem = ((size_t)hid * H + (size_t)hid * in_f + hid + (size_t)kEncRows * hid) * 4;
  GNB_REQUIRE(smem <= 200 * 1024, "encoder weights do not fit in shared memory (hid=%d H=%d)", hid, H);
  int bps = 0;
  int rc = launch_cfg(encode_kernel, smem, &bps);
  if (rc) return rc;
  int64_t tiles = (rows + kEncRows - 1) / kEncRows;
  encode_kernel<<<grid_for(tiles, bps), kThreads, smem, (cudaStream_t)stream>>>(in, idx, rows, in_f, hid, H, W1,
                                                                                 b1, W2t, b2, out);
  return check_launch("gnb_encode");
}

extern "C" int gnb_node_linear(const float* A, int64_t rows, int K, const float* Wt, const float* bias,
                               int M, float* out, int64_t ld_out, void* stream) {
  GNB_REQUIRE(K > 0 && K % 16 == 0 && M > 0 && M % 4 == 0 && ld_out >= M && ld_out % 4 == 0,
              "gnb_node_linear: need K %% 16 == 0, M %% 4 == 0, ld_out %% 4 == 0 (K=%d M=%d ld=%lld)", K, M,
              (long long)ld_out);
  if (rows == 0) return 0;
  GNB_REQUIRE(A && Wt && bias && out, "null pointer");
  GNB_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)Wt % 16 == 0) && ((uintptr_t)bias % 16 == 0) &&
                  ((uintptr_t)out % 16 == 0), "gnb_node_linear: pointers must be 16-byte aligned");
  dim3 grid((unsigned)((rows + kBM - 1) / kBM), (unsigned)((M + kBN - 1) / kBN));
  node_linear_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(A, rows, K, Wt, bias, M, out, ld_out);
  return check_launch("gnb_node_linear");
}

static int check_graph(const gnb_graph_t* g) {
  GNB_REQUIRE(g != nullptr, "graph is null");
  GNB_REQUIRE(g->num_edges >= 0 && g->num_nodes >= 0, "negative graph size");
  GNB_REQUIRE(g->in_ptr && g->out_ptr, "graph not staged");
  if (g->num_edges > 0)
    GNB_REQUIRE(g->in_src && g->in_dst && g->in_eid && g->out_pos && g->out_dst, "graph not staged");
  return 0;
}

extern "C" int gnb_edge_forward(const gnb_graph_t* g, int H, const float* P, int64_t ldP,
                                const float* We_t, const float* scale_e, const float* shift_e, float* e,
                                float* F, float* carry, int flags, void* stream) {
  int rc = check_graph(g);
  if (rc) return rc;
  if (g->num_edges == 0) return 0;
  GNB_REQUIRE(P && We_t && scale_e && shift_e && e && F && carry, "null pointer");
  GNB_REQUIRE(ldP >= ((flags & GNB_F_SYMMETRIC) ? 5 : 4) * (int64_t)H && ldP % 4 == 0, "ldP=%lld too small",
              (long long)ldP);
  GNB_REQUIRE(((uintptr_t)P % 16 == 0) && ((uintptr_t)e % 16 == 0) && ((uintptr_t)We_t % 16 == 0),
              "pointers must be 16-byte aligned");
  GNB_DISPATCH_H(H, (edge_forward_impl<kH>(g, P, ldP, We_t, scale_e, shift_e, e, F, carry, flags,
                                           (cudaStream_t)stream)));
}

extern "C" int gnb_node_update(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const float* e,
                               const float* F, const float* carry, const float* h_in,
                               const float* scale_h, const float* shift_h, float* h_out, int flags,
                               int chunk, int64_t node_begin, int64_t node_end, const int32_t* xp_ptr,
                               const int32_t* xp_row, const float* xp_buf, void* stream) {
  GNB_REQUIRE(chunk > 0, "gnb_node_update: chunk must be the carry granularity of the edge pass that filled F/carry");
  int rc = check_graph(g);
  if (rc) return rc;
  GNB_REQUIRE(0 <= node_begin && node_begin <= node_end && node_end <= g->num_nodes, "gnb_node_update: bad node range");
  if (node_end == node_begin) return 0;
  GNB_REQUIRE(P && h_in && scale_h && shift_h && h_out, "null pointer");
  GNB_REQUIRE((xp_ptr == nullptr) == (xp_row == nullptr) && (xp_ptr == nullptr) == (xp_buf == nullptr),
              "gnb_node_update: xp_ptr / xp_row / xp_buf go together");
  GNB_REQUIRE((uintptr_t)xp_buf % 16 == 0, "gnb_node_update: xp_buf must be 16-byte aligned");
  if (g->num_edges > 0) GNB_REQUIRE(e && F && carry, "null pointer");
  GNB_REQUIRE(ldP >= ((flags & GNB_F_SYMMETRIC) ? 5 : 4) * (int64_t)H && ldP % 4 == 0, "ldP=%lld too small",
              (long long)ldP);
  GNB_REQUIRE(((uintptr_t)P % 16 == 0) && ((uintptr_t)e % 16 == 0) && ((uintptr_t)F % 16 == 0) &&
                  ((uintptr_t)carry % 16 == 0) && ((uintptr_t)h_in % 16 == 0) && ((uintptr_t)h_out % 16 == 0) &&
                  ((uintptr_t)scale_h % 16 == 0) && ((uintptr_t)shift_h % 16 == 0),
              "pointers must be 16-byte aligned");
  GNB_DISPATCH_H(H, (node_update_impl<kH>(g, P, ldP, e, F, carry, h_in, scale_h, shift_h, h_out, flags, chunk,
                                          node_begin, node_end, xp_ptr, xp_row, xp_buf, nullptr,
                                          (cudaStream_t)stream)));
}

extern "C" int gnb_reverse_partial(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const float* e,
                                   int64_t node_begin, int64_t node_end, float* out, void* stream) {
  int rc = check_graph(g);
  if (rc) return rc;
  GNB_REQUIRE(0 <= node_begin && node_begin <= node_end && node_end <= g->num_nodes, "gnb_reverse_partial: bad node range");
  if (node_end == node_begin) return 0;
  GNB_REQUIRE(P && out && (g->num_edges == 0 || e), "null pointer");
  GNB_REQUIRE(ldP >= 4 * (int64_t)H && ldP % 4 == 0, "ldP=%lld too small", (long long)ldP);
  GNB_REQUIRE(((uintptr_t)P % 16 == 0) && ((uintptr_t)e % 16 == 0) && ((uintptr_t)out % 16 == 0),
              "pointers must be 16-byte aligned");
  GNB_DISPATCH_H(H, (node_update_impl<kH>(g, P, ldP, e, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                          GNB_F_SYMMETRIC, 1, node_begin, node_end, nullptr, nullptr, nullptr, out,
                                          (cudaStream_t)stream)));
}

extern "C" int gnb_score_forward(const gnb_graph_t* g, int H, int hs, const float* S, const float* W1e_t,
                                 const float* W2, const float* b2, const float* W3, const float* b3,
                                 const float* e, float* scores, void* stream) {
  int rc = check_graph(g);
  if (rc) return rc;
  if (g->num_edges == 0) return 0;
  GNB_REQUIRE(S && W1e_t && W2 && b2 && W3 && b3 && e && scores, "null pointer");
  GNB_REQUIRE(((uintptr_t)e % 16 == 0) && ((uintptr_t)W1e_t % 16 == 0), "pointers must be 16-byte aligned");
  GNB_DISPATCH_H(H, (score_forward_hs<kH>(hs, g, S, W1e_t, W2, b2, W3, b3, e, scores, (cudaStream_t)stream)));
}

extern "C" int gnb_score_forward2(const gnb_graph_t* g, int H, int hs, const float* S, const float* W1e_t,
                                  const float* W2, const float* b2, const float* W3, const float* b3,
                                  const void* e16, float* scores, void* stream) {
  int rc = check_graph(g);
  if (rc) return rc;
  if (g->num_edges == 0) return 0;
  GNB_REQUIRE(S && W1e_t && W2 && b2 && W3 && b3 && e16 && scores, "null pointer");
  GNB_REQUIRE(((uintptr_t)e16 % 16 == 0) && ((uintptr_t)W1e_t % 16 == 0), "pointers must be 16-byte aligned");
  GNB_DISPATCH_H(H, (score_forward_hs<kH, true>(hs, g, S, W1e_t, W2, b2, W3, b3, (const float*)e16, scores,
                                               (cudaStream_t)stream)));
}
