// Training-mode primitives (fp32 state, edge rows in dst-sorted position order).
//
// train.py runs the model under autograd (train.py:141-183, 329, 346).  The training path is built from a small
// set of graph primitives, each with its adjoint, that gnnome_b200/autograd.py composes into the layer exactly as
// layers/gated_gcn_full.py:82-142 does; torch's autograd engine chains them, the dense nn.Linear products stay
// plain library GEMMs.  Every sum runs over a CSR range in a fixed order (no atomics), so a training step is
// bit-reproducible.  Correctness first: these kernels stream fp32 rows and are not fused like the inference path.
//
//   gather_add3   z_p = A[src_p] + B[dst_p] (+ C_p)                      apply_edges(u_add_v)          :104-105
//   seg_sum       out_i = sum of X_p over the in- / out-edges of i        adjoint of the gathers above
//   agg_fwd       out_i = sum sigma_p * A[nbr_p] / (sum sigma_p + 1e-6)   update_all(u_mul_e, sum) / (copy_e, sum), :112-114, :125-127
//   agg_bwd_edge / agg_bwd_node                                            its adjoints w.r.t. sigma and A
//                 (mode | 2: the UN-NORMALISED sums (num, den) and their adjoints -- the sharded training step adds the
//                 partial sums of several ranks before the division)
//   gate_fwd/bwd  e' = relu(ehat) (+ e), sigma = sigmoid(e')             :107-111
//   col_stats, affine2                                                     BatchNorm1d batch statistics / normalise / backward, :106,:132
#include "gnb_common.cuh"

namespace gnb {
namespace train {

constexpr int kT = 256;
constexpr int kB = 4;     // edge rows in flight per thread in the per-node loops

static unsigned blocks_for(int64_t items) {
  int64_t b = (items + kT - 1) / kT, cap = (int64_t)sm_count() * 16;
  return (unsigned)(b < 1 ? 1 : (b < cap ? b : cap));
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 fma4(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + __expf(-x)); }

// z[p] = A[src_p] + B[dst_p] (+ C[p]); A, B node tables with row pitch ldA / ldB
__global__ void gather_add3_kernel(gnb_graph_t g, int H, const float* __restrict__ A, int64_t ldA,
                                   const float* __restrict__ B, int64_t ldB, const float* __restrict__ C,
                                   float* __restrict__ out) {
  const int h4 = H / 4;
  const int64_t total = g.num_edges * h4;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t p = i / h4;
    const int c = (int)(i - p * h4) * 4;
    float4 v = add4(ld4(A + (int64_t)g.in_src[p] * ldA + c), ld4(B + (int64_t)g.in_dst[p] * ldB + c));
    if (C) v = add4(v, ld4(C + p * H + c));
    st4(out + p * H + c, v);
  }
}

// out[i] = sum over the in-edges (mode 0) / out-edges (mode 1) of node i of X[p]; H/4 threads per node
__global__ void seg_sum_kernel(gnb_graph_t g, int H, const float* __restrict__ X, int mode, float* __restrict__ out,
                               int64_t ldo) {
  const int h4 = H / 4;
  const int64_t total = g.num_nodes * h4;
  for (int64_t k = (int64_t)blockIdx.x * kT + threadIdx.x; k < total; k += (int64_t)gridDim.x * kT) {
    const int64_t i = k / h4;
    const int c = (int)(k - i * h4) * 4;
    // kB edge rows in flight per thread (a loop with one dependent load per trip ran at ~1 TB/s); same summation order
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int a0 = mode == 0 ? g.in_ptr[i] : g.out_ptr[i], a1 = mode == 0 ? g.in_ptr[i + 1] : g.out_ptr[i + 1];
    for (int p0 = a0; p0 < a1; p0 += kB) {
      int64_t row[kB];
      float4 v[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) row[u] = (p0 + u < a1) ? (mode == 0 ? (int64_t)(p0 + u) : (int64_t)g.out_pos[p0 + u]) : -1;
#pragma unroll
      for (int u = 0; u < kB; ++u) v[u] = row[u] >= 0 ? ld4(X + row[u] * H + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kB; ++u)
        if (row[u] >= 0) acc = add4(acc, v[u]);
    }
    st4(out + i * ldo + c, acc);
  }
}

// Gate-normalised aggregation.  mode 0: i = dst, neighbour = src (in-edges); mode 1: i = src, neighbour = dst.
__global__ void agg_fwd_kernel(gnb_graph_t g, int H, const float* __restrict__ A, int64_t ldA,
                               const float* __restrict__ sigma, int mode, float* __restrict__ den,
                               float* __restrict__ out) {
  const bool raw = mode & 2;
  mode &= 1;
  const int h4 = H / 4;
  const int64_t total = g.num_nodes * h4;
  for (int64_t k = (int64_t)blockIdx.x * kT + threadIdx.x; k < total; k += (int64_t)gridDim.x * kT) {
    const int64_t i = k / h4;
    const int c = (int)(k - i * h4) * 4;
    float4 num = make_float4(0.f, 0.f, 0.f, 0.f), dn = num;
    const int a0 = mode == 0 ? g.in_ptr[i] : g.out_ptr[i], a1 = mode == 0 ? g.in_ptr[i + 1] : g.out_ptr[i + 1];
    for (int p0 = a0; p0 < a1; p0 += kB) {   // indices, then rows, then arithmetic: kB rows in flight, same order
      int64_t row[kB], nb[kB];
      float4 sv[kB], av[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const bool ok = p0 + u < a1;
        row[u] = ok ? (mode == 0 ? (int64_t)(p0 + u) : (int64_t)g.out_pos[p0 + u]) : -1;
        nb[u] = ok ? (mode == 0 ? (int64_t)g.in_src[p0 + u] : (int64_t)g.out_dst[p0 + u]) : 0;
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        if (row[u] >= 0) {
          sv[u] = ld4(sigma + row[u] * H + c);
          av[u] = ld4(A + nb[u] * ldA + c);
        }
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        if (row[u] >= 0) {
          num = fma4(sv[u], av[u], num);
          dn = add4(dn, sv[u]);
        }
      }
    }
    st4(den + i * H + c, dn);
    if (raw) st4(out + i * H + c, num);
    else st4(out + i * H + c, make_float4(num.x / (dn.x + kGateEps), num.y / (dn.y + kGateEps), num.z / (dn.z + kGateEps),
                                          num.w / (dn.w + kGateEps)));
  }
}

// gsigma[p] (+)= gnum[i_p] * A[nbr_p] + gden[i_p] with gnum = gout / (den + eps), gden = -gout * out / (den + eps)
// (mode | 2: gout IS gnum and out IS gden, den unused)
__global__ void agg_bwd_edge_kernel(gnb_graph_t g, int H, const float* __restrict__ gout, const float* __restrict__ out,
                                    const float* __restrict__ den, const float* __restrict__ A, int64_t ldA, int mode,
                                    float* __restrict__ gsigma, int accumulate) {
  const bool raw = mode & 2;
  mode &= 1;
  const int h4 = H / 4;
  const int64_t total = g.num_edges * h4;
  for (int64_t k = (int64_t)blockIdx.x * kT + threadIdx.x; k < total; k += (int64_t)gridDim.x * kT) {
    const int64_t p = k / h4;
    const int c = (int)(k - p * h4) * 4;
    const int64_t i = mode == 0 ? g.in_dst[p] : g.in_src[p];
    const int64_t nb = mode == 0 ? g.in_src[p] : g.in_dst[p];
    const float4 go = ld4(gout + i * H + c), o = ld4(out + i * H + c);
    const float4 a = ld4(A + nb * ldA + c);
    float4 r;
    if (raw) {
      r = fma4(go, a, o);
    } else {
      const float4 d = ld4(den + i * H + c);
      r.x = go.x / (d.x + kGateEps) * (a.x - o.x);
      r.y = go.y / (d.y + kGateEps) * (a.y - o.y);
      r.z = go.z / (d.z + kGateEps) * (a.z - o.z);
      r.w = go.w / (d.w + kGateEps) * (a.w - o.w);
    }
    if (accumulate) r = add4(r, ld4(gsigma + p * H + c));
    st4(gsigma + p * H + c, r);
  }
}

// gA[n] = sum over the edges whose NEIGHBOUR end is n of gnum[i_p] * sigma[p]   (walks the other CSR view)
__global__ void agg_bwd_node_kernel(gnb_graph_t g, int H, const float* __restrict__ gout, const float* __restrict__ den,
                                    const float* __restrict__ sigma, int mode, float* __restrict__ gA, int64_t ldg) {
  const bool raw = mode & 2;   // gout is gnum: no division by (den + eps)
  mode &= 1;
  const int h4 = H / 4;
  const int64_t total = g.num_nodes * h4;
  for (int64_t k = (int64_t)blockIdx.x * kT + threadIdx.x; k < total; k += (int64_t)gridDim.x * kT) {
    const int64_t n = k / h4;
    const int c = (int)(k - n * h4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // mode 0: the forward aggregated over in-edges with neighbour = src, so n is the SOURCE of the edges walked here
    // (src-CSR: row = out_pos[q], owner i = out_dst[q]); mode 1: n is the DESTINATION (dst-CSR: row = p, i = in_src[p]).
    const int a0 = mode == 0 ? g.out_ptr[n] : g.in_ptr[n], a1 = mode == 0 ? g.out_ptr[n + 1] : g.in_ptr[n + 1];
    for (int p0 = a0; p0 < a1; p0 += kB) {   // kB rows in flight, same summation order
      int64_t row[kB], own[kB];
      float4 gv[kB], sv[kB], dv[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        const bool ok = p0 + u < a1;
        row[u] = ok ? (mode == 0 ? (int64_t)g.out_pos[p0 + u] : (int64_t)(p0 + u)) : -1;
        own[u] = ok ? (mode == 0 ? (int64_t)g.out_dst[p0 + u] : (int64_t)g.in_src[p0 + u]) : 0;
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        if (row[u] >= 0) {
          gv[u] = ld4(gout + own[u] * H + c);
          sv[u] = ld4(sigma + row[u] * H + c);
          if (!raw) dv[u] = ld4(den + own[u] * H + c);
        }
      }
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        if (row[u] < 0) continue;
        if (raw) {
          acc = fma4(gv[u], sv[u], acc);
        } else {
          acc.x = fmaf(gv[u].x / (dv[u].x + kGateEps), sv[u].x, acc.x);
          acc.y = fmaf(gv[u].y / (dv[u].y + kGateEps), sv[u].y, acc.y);
          acc.z = fmaf(gv[u].z / (dv[u].z + kGateEps), sv[u].z, acc.z);
          acc.w = fmaf(gv[u].w / (dv[u].w + kGateEps), sv[u].w, acc.w);
        }
      }
    }
    st4(gA + n * ldg + c, acc);
  }
}

// e' = relu(ehat) (+ e_in), sigma = sigmoid(e')
__global__ void gate_fwd_kernel(const float* __restrict__ ehat, const float* __restrict__ e_in, int64_t n4,
                                float* __restrict__ e_out, float* __restrict__ sigma) {
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kT) {
    float4 v = ld4(ehat + 4 * i);
    v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    if (e_in) v = add4(v, ld4(e_in + 4 * i));
    st4(e_out + 4 * i, v);
    st4(sigma + 4 * i, make_float4(sigmoid_exact(v.x), sigmoid_exact(v.y), sigmoid_exact(v.z), sigmoid_exact(v.w)));
  }
}

// t = g_e + g_sigma * sigma * (1 - sigma);  g_ehat = t * [ehat > 0];  g_ein = t
__global__ void gate_bwd_kernel(const float* __restrict__ g_e, const float* __restrict__ g_sigma,
                                const float* __restrict__ ehat, const float* __restrict__ sigma, int64_t n4,
                                float* __restrict__ g_ehat, float* __restrict__ g_ein) {
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kT) {
    const float4 ge = ld4(g_e + 4 * i), gs = ld4(g_sigma + 4 * i), s = ld4(sigma + 4 * i), eh = ld4(ehat + 4 * i);
    float4 t;
    t.x = fmaf(gs.x * s.x, 1.f - s.x, ge.x);
    t.y = fmaf(gs.y * s.y, 1.f - s.y, ge.y);
    t.z = fmaf(gs.z * s.z, 1.f - s.z, ge.z);
    t.w = fmaf(gs.w * s.w, 1.f - s.w, ge.w);
    if (g_ein) st4(g_ein + 4 * i, t);
    st4(g_ehat + 4 * i, make_float4(eh.x > 0.f ? t.x : 0.f, eh.y > 0.f ? t.y : 0.f, eh.z > 0.f ? t.z : 0.f,
                                    eh.w > 0.f ? t.w : 0.f));
  }
}

// out = a * (x - sx) (+ b * (y - sy)) + c with per-channel a, b, c, sx, sy: BatchNorm normalise (y = null) and its input
// gradient.  The shifts (null = 0) are the column means: a channel whose mean is large against its spread would lose
// its digits in a * x + (c - a * mean).
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

__global__ void affine2_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ a,
                               const float* __restrict__ b, const float* __restrict__ cc, const float* __restrict__ sx,
                               const float* __restrict__ sy, int64_t rows, int H, float* __restrict__ out) {
  const int h4 = H / 4;
  const int64_t total = rows * h4;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c = (int)(i % h4) * 4;
    float4 xv = ld4(x + 4 * i);
    if (sx) xv = sub4(xv, ld4(sx + c));
    float4 v = fma4(ld4(a + c), xv, ld4(cc + c));
    if (y) {
      float4 yv = ld4(y + 4 * i);
      if (sy) yv = sub4(yv, ld4(sy + c));
      v = fma4(ld4(b + c), yv, v);
    }
    st4(out + 4 * i, v);
  }
}

// Column statistics in two deterministic steps: per-CTA partial sums of a' and a'*b' (b = null: a'*a') over a contiguous
// block of rows in fp32 (<= kStatRows rows each), combined in fp64 in CTA order; a' = a - shift_a, b' = b - shift_b
// per channel (the variance is taken about the mean of a first pass: no cancellation).
constexpr int kStatRows = 1024;

__global__ void col_stats_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                         const float* __restrict__ shift_a, const float* __restrict__ shift_b,
                                         int64_t rows, int H, double* __restrict__ part /* [blocks][2][H] */) {
  const int64_t r0 = (int64_t)blockIdx.x * kStatRows;
  const int64_t r1 = (r0 + kStatRows < rows) ? r0 + kStatRows : rows;
  for (int c = threadIdx.x; c < H; c += kT) {
    float s1 = 0.f, s2 = 0.f;
    const float sa = shift_a ? shift_a[c] : 0.f, sb = shift_b ? shift_b[c] : 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      const float av = a[r * H + c] - sa;
      const float bv = b ? b[r * H + c] - sb : av;
      s1 += av;
      s2 = fmaf(av, bv, s2);
    }
    part[((int64_t)blockIdx.x * 2 + 0) * H + c] = (double)s1;
    part[((int64_t)blockIdx.x * 2 + 1) * H + c] = (double)s2;
  }
}

__global__ void col_stats_combine_kernel(const double* __restrict__ part, int64_t blocks, int H, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t k = 0; k < blocks; ++k) {
    s1 += part[(k * 2 + 0) * H + c];
    s2 += part[(k * 2 + 1) * H + c];
  }
  out[c] = s1;
  out[H + c] = s2;
}

// LayerNorm over the channels of each row (nn.LayerNorm, gated_gcn_full.py:39-42): one warp per row.
// fwd: xhat = (x - mean_row) * rstd_row, y = gamma * xhat + beta; saves xhat and rstd.
__global__ void layer_norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int64_t rows, int W, float eps,
                                      float* __restrict__ y, float* __restrict__ xhat, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * kT + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * kT) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float* xr = x + r * W;
    float s = 0.f;
    for (int c = lane; c < W; c += 32) s += xr[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)W;
    float v = 0.f;
    for (int c = lane; c < W; c += 32) {
      const float d = xr[c] - mean;
      v = fmaf(d, d, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rs = rsqrtf(v / (float)W + eps);
    if (lane == 0) rstd[r] = rs;
    for (int c = lane; c < W; c += 32) {
      const float xh = (xr[c] - mean) * rs;
      xhat[r * W + c] = xh;
      y[r * W + c] = fmaf(gamma[c], xh, beta[c]);
    }
  }
}

// bwd: gxhat = g * gamma; gx = rstd * (gxhat - mean_row(gxhat) - xhat * mean_row(gxhat * xhat))
__global__ void layer_norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ xhat,
                                      const float* __restrict__ rstd, const float* __restrict__ gamma, int64_t rows,
                                      int W, float* __restrict__ gx) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * kT + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * kT) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < W; c += 32) {
      const float gh = g[r * W + c] * gamma[c];
      s1 += gh;
      s2 = fmaf(gh, xhat[r * W + c], s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / (float)W, m2 = s2 / (float)W, rs = rstd[r];
    for (int c = lane; c < W; c += 32) gx[r * W + c] = rs * (g[r * W + c] * gamma[c] - m1 - xhat[r * W + c] * m2);
  }
}

}  // namespace train
}  // namespace gnb

using namespace gnb;
using namespace gnb::train;

extern "C" int gnb_t_layer_norm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int W, float eps,
                                    float* y, float* xhat, float* rstd, void* stream) {
  GNB_REQUIRE(W > 0, "gnb_t_layer_norm_fwd: bad width %d", W);
  if (rows == 0) return 0;
  GNB_REQUIRE(x && gamma && beta && y && xhat && rstd, "null pointer");
  layer_norm_fwd_kernel<<<blocks_for(rows * 32), kT, 0, (cudaStream_t)stream>>>(x, gamma, beta, rows, W, eps, y, xhat, rstd);
  return check_launch("gnb_t_layer_norm_fwd");
}

extern "C" int gnb_t_layer_norm_bwd(const float* g, const float* xhat, const float* rstd, const float* gamma, int64_t rows,
                                    int W, float* gx, void* stream) {
  GNB_REQUIRE(W > 0, "gnb_t_layer_norm_bwd: bad width %d", W);
  if (rows == 0) return 0;
  GNB_REQUIRE(g && xhat && rstd && gamma && gx, "null pointer");
  layer_norm_bwd_kernel<<<blocks_for(rows * 32), kT, 0, (cudaStream_t)stream>>>(g, xhat, rstd, gamma, rows, W, gx);
  return check_launch("gnb_t_layer_norm_bwd");
}

static int check_train_graph(const gnb_graph_t* g, int H) {
  GNB_REQUIRE(g != nullptr && g->in_ptr && g->out_ptr, "graph not staged");
  GNB_REQUIRE(H > 0 && H % 4 == 0, "hidden_features=%d must be a positive multiple of 4", H);
  if (g->num_edges > 0) GNB_REQUIRE(g->in_src && g->in_dst && g->out_pos && g->out_dst, "graph not staged");
  return 0;
}

extern "C" int gnb_t_gather_add3(const gnb_graph_t* g, int H, const float* A, int64_t ldA, const float* B, int64_t ldB,
                                 const float* C, float* out, void* stream) {
  int rc = check_train_graph(g, H);
  if (rc) return rc;
  if (g->num_edges == 0) return 0;
  GNB_REQUIRE(A && B && out && ldA % 4 == 0 && ldB % 4 == 0, "gnb_t_gather_add3: bad arguments");
  gather_add3_kernel<<<blocks_for(g->num_edges * (H / 4)), kT, 0, (cudaStream_t)stream>>>(*g, H, A, ldA, B, ldB, C, out);
  return check_launch("gnb_t_gather_add3");
}

extern "C" int gnb_t_seg_sum(const gnb_graph_t* g, int H, const float* X, int mode, float* out, int64_t ldo, void* stream) {
  int rc = check_train_graph(g, H);
  if (rc) return rc;
  if (g->num_nodes == 0) return 0;
  GNB_REQUIRE(out && (X || g->num_edges == 0) && ldo % 4 == 0 && (mode == 0 || mode == 1), "gnb_t_seg_sum: bad arguments");
  seg_sum_kernel<<<blocks_for(g->num_nodes * (H / 4)), kT, 0, (cudaStream_t)stream>>>(*g, H, X, mode, out, ldo);
  return check_launch("gnb_t_seg_sum");
}

extern "C" int gnb_t_agg_fwd(const gnb_graph_t* g, int H, const float* A, int64_t ldA, const float* sigma, int mode,
                             float* den, float* out, void* stream) {
  int rc = check_train_graph(g, H);
  if (rc) return rc;
  if (g->num_nodes == 0) return 0;
  GNB_REQUIRE(A && den && out && (sigma || g->num_edges == 0) && ldA % 4 == 0 && mode >= 0 && mode <= 3,
              "gnb_t_agg_fwd: bad arguments");
  agg_fwd_kernel<<<blocks_for(g->num_nodes * (H / 4)), kT, 0, (cudaStream_t)stream>>>(*g, H, A, ldA, sigma, mode, den, out);
  return check_launch("gnb_t_agg_fwd");
}

extern "C" int gnb_t_agg_bwd_edge(const gnb_graph_t* g, int H, const float* gout, const float* out, const float* den,
                                  const float* A, int64_t ldA, int mode, float* gsigma, int accumulate, void* stream) {
  int rc = check_train_graph(g, H);
  if (rc) return rc;
  if (g->num_edges == 0) return 0;
  GNB_REQUIRE(gout && out && (den || (mode & 2)) && A && gsigma && ldA % 4 == 0 && mode >= 0 && mode <= 3,
              "gnb_t_agg_bwd_edge: bad arguments");
  agg_bwd_edge_kernel<<<blocks_for(g->num_edges * (H / 4)), kT, 0, (cudaStream_t)stream>>>(*g, H, gout, out, den, A, ldA,
                                                                                         mode, gsigma, accumulate);
  return check_launch("gnb_t_agg_bwd_edge");
}

extern "C" int gnb_t_agg_bwd_node(const gnb_graph_t* g, int H, const float* gout, const float* den, const float* sigma,
                                  int mode, float* gA, int64_t ldg, void* stream) {
  int rc = check_train_graph(g, H);
  if (rc) return rc;
  if (g->num_nodes == 0) return 0;
  GNB_REQUIRE(gout && (den || (mode & 2)) && gA && (sigma || g->num_edges == 0) && ldg % 4 == 0 && mode >= 0 && mode <= 3,
              "gnb_t_agg_bwd_node: bad arguments");
  agg_bwd_node_kernel<<<blocks_for(g->num_nodes * (H / 4)), kT, 0, (cudaStream_t)stream>>>(*g, H, gout, den, sigma, mode, gA, ldg);
  return check_launch("gnb_t_agg_bwd_node");
}

extern "C" int gnb_t_gate_fwd(const float* ehat, const float* e_in, int64_t rows, int H, float* e_out, float* sigma,
                              void* stream) {
  GNB_REQUIRE(H > 0 && H % 4 == 0, "hidden_features=%d must be a positive multiple of 4", H);
  if (rows == 0) return 0;
  GNB_REQUIRE(ehat && e_out && sigma, "null pointer");
  gate_fwd_kernel<<<blocks_for(rows * (H / 4)), kT, 0, (cudaStream_t)stream>>>(ehat, e_in, rows * (H / 4), e_out, sigma);
  return check_launch("gnb_t_gate_fwd");
}

extern "C" int gnb_t_gate_bwd(const float* g_e, const float* g_sigma, const float* ehat, const float* sigma, int64_t rows,
                              int H, float* g_ehat, float* g_ein, void* stream) {
  GNB_REQUIRE(H > 0 && H % 4 == 0, "hidden_features=%d must be a positive multiple of 4", H);
  if (rows == 0) return 0;
  GNB_REQUIRE(g_e && g_sigma && ehat && sigma && g_ehat, "null pointer");
  gate_bwd_kernel<<<blocks_for(rows * (H / 4)), kT, 0, (cudaStream_t)stream>>>(g_e, g_sigma, ehat, sigma, rows * (H / 4),
                                                                              g_ehat, g_ein);
  return check_launch("gnb_t_gate_bwd");
}

extern "C" int gnb_t_affine2(const float* x, const float* y, const float* a, const float* b, const float* c,
                             const float* shift_x, const float* shift_y, int64_t rows, int H, float* out, void* stream) {
  GNB_REQUIRE(H > 0 && H % 4 == 0, "hidden_features=%d must be a positive multiple of 4", H);
  if (rows == 0) return 0;
  GNB_REQUIRE(x && a && c && out && (y == nullptr || b != nullptr), "null pointer");
  affine2_kernel<<<blocks_for(rows * (H / 4)), kT, 0, (cudaStream_t)stream>>>(x, y, a, b, c, shift_x, shift_y, rows, H, out);
  return check_launch("gnb_t_affine2");
}

extern "C" size_t gnb_t_col_stats_workspace(int64_t rows, int H) {
  if (rows <= 0 || H <= 0) return 0;
  return (size_t)((rows + kStatRows - 1) / kStatRows) * 2 * (size_t)H * sizeof(double);
}

extern "C" int gnb_t_col_stats(const float* a, const float* b, const float* shift_a, const float* shift_b, int64_t rows,
                               int H, double* out, void* workspace, void* stream) {
  GNB_REQUIRE(H > 0 && rows >= 0 && out, "gnb_t_col_stats: bad arguments");
  if (rows == 0) {
    GNB_CUDA(cudaMemsetAsync(out, 0, (size_t)2 * H * sizeof(double), (cudaStream_t)stream));
    return 0;
  }
  GNB_REQUIRE(a && workspace, "null pointer");
  const int64_t blocks = (rows + kStatRows - 1) / kStatRows;
  col_stats_partial_kernel<<<(unsigned)blocks, kT, 0, (cudaStream_t)stream>>>(a, b, shift_a, shift_b, rows, H, (double*)workspace);
  col_stats_combine_kernel<<<(H + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const double*)workspace, blocks, H, out);
  return check_launch("gnb_t_col_stats");
}
