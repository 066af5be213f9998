// CUDA-core kernels of the split16 path (gnb_tma.cuh): the input encoders and the reverse aggregation + node update.
// They stream: every byte is touched once, so the work is organised for coalescing and loads in flight, not flops.
#include "gnb_tma.cuh"

namespace gnb {
namespace tc {

// ------------------------------------------------------------------------------------------------
// Encoders: y[r] = W2 * relu(W1 * in[idx ? idx[r] : r] + b1) + b2   (models/full_graph.py:26-27)
// written as split16 images (out16) and / or fp32 rows (out32).
// A thread owns 8 consecutive channels of R = 4 rows: per hidden unit j it reads its W2 slice once (two 16-byte
// shared loads) and applies it to the four rows; H / 8 lanes cover a row, so stores are row-contiguous.
// ------------------------------------------------------------------------------------------------
constexpr int kEncR = 4;
constexpr int kEncThreads = 256;

// HID = 16 (hidden_ne_features of every shipped configuration): the hidden layer of a row is computed ONCE, its 16
// units spread over the H / 8 lanes that share the row, and broadcast by shuffle inside the fully unrolled j loop;
// HID = 0: any hidden width, every lane recomputes the hidden units it needs.
template <int H, int HID>
__global__ void __launch_bounds__(kEncThreads)
encode2_kernel(const float* __restrict__ in, const int32_t* __restrict__ idx, int64_t rows, int in_f, int hid,
               const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2t,
               const float* __restrict__ b2, __half* __restrict__ out16, float* __restrict__ out32) {
  constexpr int LPR = H / 8;                       // lanes per row
  constexpr int RPB = (kEncThreads / LPR) * kEncR; // rows per CTA iteration
  extern __shared__ __align__(16) float smem_enc[];
  float* w2 = smem_enc;                 // [hid][H]
  float* w1 = w2 + hid * H;             // [hid][in_f]
  float* bb1 = w1 + hid * in_f;         // [hid]
  for (int i = threadIdx.x; i < hid * H; i += kEncThreads) w2[i] = W2t[i];
  for (int i = threadIdx.x; i < hid * in_f; i += kEncThreads) w1[i] = W1[i];
  for (int i = threadIdx.x; i < hid; i += kEncThreads) bb1[i] = b1[i];
  __syncthreads();
  const int lr = threadIdx.x % LPR, rg = threadIdx.x / LPR;
  const int c0 = lr * 8;
  float bias[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bias[k] = b2[c0 + k];
  const int64_t num_it = (rows + RPB - 1) / RPB;
  for (int64_t it = blockIdx.x; it < num_it; it += gridDim.x) {
    const int64_t r0 = it * RPB + (int64_t)rg * kEncR;
    float x[kEncR][4];   // in_f <= 4
#pragma unroll
    for (int r = 0; r < kEncR; ++r) {
      const int64_t row = r0 + r;
#pragma unroll
      for (int f = 0; f < 4; ++f) x[r][f] = 0.f;
      if (row < rows) {
        const int64_t src = idx ? (int64_t)idx[row] : row;
#pragma unroll
        for (int f = 0; f < 4; ++f)
          if (f < in_f) x[r][f] = in[src * in_f + f];
      }
    }
    float acc[kEncR][8];
#pragma unroll
    for (int r = 0; r < kEncR; ++r)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[r][k] = bias[k];
    if (HID > 0) {
      constexpr int UPL = HID > 0 ? (HID + LPR - 1) / LPR : 1;   // hidden units per lane
      const unsigned gmask = LPR >= 32 ? 0xffffffffu : (((1u << LPR) - 1u) << (((threadIdx.x & 31) / LPR) * LPR));
      float hv[kEncR][UPL];
#pragma unroll
      for (int r = 0; r < kEncR; ++r)
#pragma unroll
        for (int q = 0; q < UPL; ++q) {
          const int j = lr + q * LPR;
          float hsum = 0.f;
          if (j < HID) {
            hsum = bb1[j];
#pragma unroll
            for (int f = 0; f < 4; ++f)
              if (f < in_f) hsum = fmaf(x[r][f], w1[j * in_f + f], hsum);
            hsum = fmaxf(hsum, 0.f);
          }
          hv[r][q] = hsum;
        }
#pragma unroll
      for (int j = 0; j < HID; ++j) {
        const float4 wa = *reinterpret_cast<const float4*>(w2 + j * H + c0);
        const float4 wb = *reinterpret_cast<const float4*>(w2 + j * H + c0 + 4);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int r = 0; r < kEncR; ++r) {
          const float hsum = __shfl_sync(gmask, hv[r][j / LPR], j % LPR, LPR < 32 ? LPR : 32);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[r][k] = fmaf(hsum, w[k], acc[r][k]);
        }
      }
    } else
    for (int j = 0; j < hid; ++j) {
      const float4 wa = *reinterpret_cast<const float4*>(w2 + j * H + c0);
      const float4 wb = *reinterpret_cast<const float4*>(w2 + j * H + c0 + 4);
      const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int r = 0; r < kEncR; ++r) {
        float hsum = bb1[j];
#pragma unroll
        for (int f = 0; f < 4; ++f)
          if (f < in_f) hsum = fmaf(x[r][f], w1[j * in_f + f], hsum);
        hsum = fmaxf(hsum, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r][k] = fmaf(hsum, w[k], acc[r][k]);
      }
    }
#pragma unroll
    for (int r = 0; r < kEncR; ++r) {
      const int64_t row = r0 + r;
      if (row >= rows) continue;
      if (out32) {
        *reinterpret_cast<float4*>(out32 + row * H + c0) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        *reinterpret_cast<float4*>(out32 + row * H + c0 + 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
      }
      if (out16) {
        float xs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) xs[k] = acc[r][k] * kXScale;
        uint4 h, l;
        split8(xs, h, l);
        *reinterpret_cast<uint4*>(out16 + row * 2 * H + c0) = h;
        *reinterpret_cast<uint4*>(out16 + row * 2 * H + H + c0) = l;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Reverse aggregation over the src-CSR + node update (gated_gcn_full.py:124-137), e' read from its split16
// images.  H / 8 threads per node, 8 channels per thread (16-byte hi + 16-byte lo + 32-byte A3h loads per edge).
// ------------------------------------------------------------------------------------------------
struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 f8_zero() {
  F8 r;
#pragma unroll
  for (int k = 0; k < 8; ++k) r.v[k] = 0.f;
  return r;
}
__device__ __forceinline__ F8 f8_load(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  F8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ F8 f8_ldg(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  F8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void f8_store(float* p, const F8& x) {
  *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
// 8 halves hi + 8 halves lo -> 8 floats (x = 16 * (hi + lo))
__device__ __forceinline__ F8 f8_merge(const uint4& h, const uint4& l) {
  const __half2* hp = reinterpret_cast<const __half2*>(&h);
  const __half2* lp = reinterpret_cast<const __half2*>(&l);
  F8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 hf = __half22float2(hp[j]), lf = __half22float2(lp[j]);
    r.v[2 * j] = (hf.x + lf.x) * kWScale;
    r.v[2 * j + 1] = (hf.y + lf.y) * kWScale;
  }
  return r;
}

constexpr int kNu2Threads = 256;
constexpr int kNu2Batch = 4;   // out-edges in flight per thread (64 bytes each)

// H / 8 threads ("group") per node, 8 channels per thread: per out-edge a 16-byte hi + 16-byte lo load of the e'
// row (one contiguous 4H-byte row per group) and a 32-byte load of A3h[dst].  Latency is what bounds this kernel
// (CSR pointers -> edge indices -> rows are dependent loads), so: the pointers of the group's NEXT node are
// fetched one iteration ahead, everything that depends only on the node id is issued before the edge loop, the
// group loads the indices of up to H / 8 edges with one instruction and broadcasts them by shuffle, and four
// edge rows are in flight per thread.
template <int H>
__global__ void __launch_bounds__(kNu2Threads, 2)
node_update2_kernel(gnb_graph_t g, const float* __restrict__ P, int64_t ldP, const __half* __restrict__ e16,
                    const float* __restrict__ F, const float* __restrict__ carry, const float* __restrict__ h_in,
                    const float* __restrict__ scale_h, const float* __restrict__ shift_h, float* __restrict__ h_out,
                    __half* __restrict__ h16_out, int flags, int chunk, int64_t node_begin, int64_t node_end,
                    const int32_t* __restrict__ xp_ptr, const int32_t* __restrict__ xp_row,
                    const float* __restrict__ xp_buf, float* __restrict__ partial_out) {
  constexpr int TPN = H / 8;               // threads per node
  constexpr int NPB = kNu2Threads / TPN;   // nodes in flight per CTA
  const int t = threadIdx.x % TPN, slot = threadIdx.x / TPN;
  const int c0 = t * 8;
  const bool sym = flags & GNB_F_SYMMETRIC, residual = flags & GNB_F_RESIDUAL;
  const bool agg = sym || partial_out;
  const int a1_off = sym ? 4 * H : 3 * H;
  const unsigned gmask = TPN >= 32 ? 0xffffffffu : (((1u << TPN) - 1u) << (((threadIdx.x & 31) / TPN) * TPN));
  // Node ids and counts fit 32 bits (the graph's indices are int32); the loop state is kept narrow because this kernel
  // sits at its register cap.
  const int stride = (int)gridDim.x * NPB;
  const int count = (int)(node_end - node_begin);
  auto node_at = [&](int k) -> int { return (int)node_begin + k; };
  int k = (int)blockIdx.x * NPB + slot;
  int i32 = k < count ? node_at(k) : 0;
  int qa = 0, qb = 0, pa = 0, pb = 0;
  if (k < count) {
    if (agg) { qa = g.out_ptr[i32]; qb = g.out_ptr[i32 + 1]; }
    if (!partial_out) { pa = g.in_ptr[i32]; pb = g.in_ptr[i32 + 1]; }
  }
  for (; k < count; k += stride) {
    const int64_t i = i32;
    int nqa = 0, nqb = 0, npa = 0, npb = 0;   // CSR pointers of this group's next node: one iteration ahead
    int inext = 0;
    if (k + stride < count) {
      inext = node_at(k + stride);
      if (agg) { nqa = g.out_ptr[inext]; nqb = g.out_ptr[inext + 1]; }
      if (!partial_out) { npa = g.in_ptr[inext]; npb = g.in_ptr[inext + 1]; }
    }
    // the list of partial sums other ranks computed for the node (multi-GPU): requested here, ahead of the edge
    // loop -- read after it, the two dependent loads added a DRAM round trip to every node of a latency-bound
    // kernel (the sharded launches ran 30 % longer per node than the single-GPU ones, profiles/r02f)
    int xa = 0, xb = 0;
    if (xp_ptr) { xa = xp_ptr[i]; xb = xp_ptr[i + 1]; }
    // ---- everything that depends on the node id only ------------------------------------------------
    F8 a1 = f8_zero(), hin = f8_zero(), f = f8_zero();
    const bool f_direct = pb > pa && pa / chunk == (pb - 1) / chunk;   // the whole in-segment sits in one chunk
    if (!partial_out) {
      a1 = f8_ldg(P + i * ldP + a1_off + c0);
      if (residual) hin = f8_load(h_in + i * H + c0);
      if (f_direct) f = f8_load(F + i * H + c0);
    }
    // ---- Bk: gate-normalised sum over out-edges ------------------------------------------------
    F8 bk = f8_zero();
    if (agg) {
      F8 num = f8_zero(), den = f8_zero();
      for (int q0 = qa; q0 < qb; q0 += TPN) {
        const int cnt = (qb - q0 < TPN) ? (qb - q0) : TPN;   // group-uniform
        int myp = 0, myd = 0;
        if (t < cnt) {
          myp = g.out_pos[q0 + t];
          myd = g.out_dst[q0 + t];
        }
        for (int u0 = 0; u0 < cnt; u0 += kNu2Batch) {
          uint4 eh[kNu2Batch], el[kNu2Batch];
          F8 av[kNu2Batch];
#pragma unroll
          for (int u = 0; u < kNu2Batch; ++u) {
            const int lane_u = (u0 + u < cnt) ? (u0 + u) : (cnt - 1);   // tail slots repeat the last edge (loads only)
            const int64_t p = __shfl_sync(gmask, myp, lane_u, TPN);
            const int64_t d = __shfl_sync(gmask, myd, lane_u, TPN);
            eh[u] = *reinterpret_cast<const uint4*>(e16 + p * 2 * H + c0);
            el[u] = *reinterpret_cast<const uint4*>(e16 + p * 2 * H + H + c0);
            av[u] = f8_ldg(P + d * ldP + 3 * H + c0);
          }
#pragma unroll
          for (int u = 0; u < kNu2Batch; ++u) {
            if (u0 + u < cnt) {
              const F8 ev = f8_merge(eh[u], el[u]);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float sg = sigmoidf_fast(ev.v[k]);
                num.v[k] = fmaf(sg, av[u].v[k], num.v[k]);
                den.v[k] += sg;
              }
            }
          }
        }
      }
      if (partial_out) {  // multi-GPU: un-normalised partial sums of a halo source node, for its owner
        f8_store(partial_out + (i - node_begin) * 2 * H + c0, num);
        f8_store(partial_out + (i - node_begin) * 2 * H + H + c0, den);
        qa = nqa; qb = nqb; i32 = inext;
        continue;
      }
      if (xp_ptr) {       // multi-GPU: partial sums other ranks computed for this node, in rank order
        for (int r = xa; r < xb; ++r) {
          const float* row = xp_buf + (int64_t)xp_row[r] * 2 * H;
          const F8 xn = f8_load(row + c0), xd = f8_load(row + H + c0);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            num.v[k] += xn.v[k];
            den.v[k] += xd.v[k];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) bk.v[k] = gate_div(num.v[k], den.v[k]);
    }
    // ---- F: from the edge pass, resolving chunk-straddling segments ---------------------------------
    if (pb > pa && !f_direct) {
      const int k0 = pa / chunk, k1 = (pb - 1) / chunk;
      F8 num = f8_zero(), den = f8_zero();
      for (int k = k0; k < k1; ++k) {
        const F8 a = f8_load(carry + ((int64_t)k * 4 + 2) * H + c0), b = f8_load(carry + ((int64_t)k * 4 + 3) * H + c0);
#pragma unroll
        for (int jx = 0; jx < 8; ++jx) {
          num.v[jx] += a.v[jx];
          den.v[jx] += b.v[jx];
        }
      }
      const F8 a = f8_load(carry + ((int64_t)k1 * 4 + 0) * H + c0), b = f8_load(carry + ((int64_t)k1 * 4 + 1) * H + c0);
#pragma unroll
      for (int jx = 0; jx < 8; ++jx) f.v[jx] = gate_div(num.v[jx] + a.v[jx], den.v[jx] + b.v[jx]);
    }
    // ---- h' = relu(bn_h(A1h + F + Bk)) + h ---------------------------------------------------------
    const F8 sc = f8_ldg(scale_h + c0), sh = f8_ldg(shift_h + c0);   // L1-resident; not worth 16 live registers
    F8 u;
#pragma unroll
    for (int k = 0; k < 8; ++k) u.v[k] = fmaxf(fmaf(a1.v[k] + f.v[k] + bk.v[k], sc.v[k], sh.v[k]), 0.f) + hin.v[k];
    f8_store(h_out + i * H + c0, u);
    if (h16_out) {
      float xs[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) xs[k] = u.v[k] * kXScale;
      uint4 hh, ll;
      split8(xs, hh, ll);
      *reinterpret_cast<uint4*>(h16_out + i * 2 * H + c0) = hh;     // absolute rows, like h_out
      *reinterpret_cast<uint4*>(h16_out + i * 2 * H + H + c0) = ll;
    }
    qa = nqa; qb = nqb; pa = npa; pb = npb; i32 = inext;
  }
}

template <int H>
static int encode2_impl(const float* in, const int32_t* idx, int64_t rows, int in_f, int hid, const float* W1,
                        const float* b1, const float* W2t, const float* b2, void* out16, float* out32,
                        cudaStream_t stream) {
  const size_t smem = ((size_t)hid * H + (size_t)hid * in_f + hid) * sizeof(float);
  auto kern = hid == 16 ? encode2_kernel<H, 16> : encode2_kernel<H, 0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("gnb_encode2: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return (int)e;
  }
  constexpr int RPB = (kEncThreads / (H / 8)) * kEncR;
  const int64_t its = (rows + RPB - 1) / RPB;
  const int64_t cap = (int64_t)sm_count() * 4;
  kern<<<(unsigned)(its < cap ? its : cap), kEncThreads, smem, stream>>>(
      in, idx, rows, in_f, hid, W1, b1, W2t, b2, (__half*)out16, out32);
  return check_launch("gnb_encode2");
}

template <int H>
static int node_update2_impl(const gnb_graph_t* g, const float* P, int64_t ldP, const void* e16, const float* F,
                             const float* carry, const float* h_in, const float* scale_h, const float* shift_h,
                             float* h_out, void* h16_out, int flags, int chunk, int64_t node_begin, int64_t node_end,
                             const int32_t* xp_ptr, const int32_t* xp_row, const float* xp_buf, float* partial_out,
                             cudaStream_t stream) {
  constexpr int NPB = kNu2Threads / (H / 8);
  const int64_t items = (node_end - node_begin + NPB - 1) / NPB;
  const int64_t cap = (int64_t)sm_count() * 2;   // persistent: two resident CTAs per SM, grid-stride over the nodes
  node_update2_kernel<H><<<(unsigned)(items < cap ? items : cap), kNu2Threads, 0, stream>>>(
      *g, P, ldP, (const __half*)e16, F, carry, h_in, scale_h, shift_h, h_out, (__half*)h16_out, flags, chunk,
      node_begin, node_end, xp_ptr, xp_row, xp_buf, partial_out);
  return check_launch(partial_out ? "gnb_reverse_partial2" : "gnb_node_update2");
}

}  // namespace tc
}  // namespace gnb

using namespace gnb;

#define GNB_DISPATCH_H2(H, CALL)                                \
  switch (H) {                                                  \
    case 64: { constexpr int kH = 64; return CALL; }            \
    case 128: { constexpr int kH = 128; return CALL; }          \
    case 256: { constexpr int kH = 256; return CALL; }          \
    default: break;                                             \
  }                                                             \
  set_error("hidden_features=%d unsupported by the split16 path (64, 128, 256)", H); \
  return GNB_E_INVALID

extern "C" int gnb_encode2(const float* in, const int32_t* idx, int64_t rows, int in_f, int hid, int H,
                           const float* W1, const float* b1, const float* W2t, const float* b2, void* out16,
                           float* out32, void* stream) {
  GNB_REQUIRE(in_f > 0 && in_f <= 4 && hid > 0 && hid <= 64, "gnb_encode2: need in_f <= 4, hid <= 64 (in_f=%d hid=%d)", in_f, hid);
  if (rows == 0) return 0;
  GNB_REQUIRE(in && W1 && b1 && W2t && b2 && (out16 || out32), "null pointer");
  GNB_REQUIRE(((uintptr_t)out16 % 16 == 0) && ((uintptr_t)out32 % 16 == 0) && ((uintptr_t)W2t % 16 == 0),
              "gnb_encode2: outputs / W2t must be 16-byte aligned");
  GNB_DISPATCH_H2(H, (tc::encode2_impl<kH>(in, idx, rows, in_f, hid, W1, b1, W2t, b2, out16, out32, (cudaStream_t)stream)));
}

static int check_graph2(const gnb_graph_t* g) {
  GNB_REQUIRE(g != nullptr, "graph is null");
  GNB_REQUIRE(g->num_edges >= 0 && g->num_nodes >= 0, "negative graph size");
  GNB_REQUIRE(g->in_ptr && g->out_ptr, "graph not staged");
  if (g->num_edges > 0)
    GNB_REQUIRE(g->in_src && g->in_dst && g->in_eid && g->out_pos && g->out_dst, "graph not staged");
  return 0;
}

extern "C" int gnb_node_update2(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const void* e16,
                                const float* F, const float* carry, const float* h_in, const float* scale_h,
                                const float* shift_h, float* h_out, void* h16_out, int flags, int chunk,
                                int64_t node_begin, int64_t node_end, const int32_t* xp_ptr, const int32_t* xp_row,
                                const float* xp_buf, void* stream) {
  GNB_REQUIRE(chunk > 0, "gnb_node_update2: chunk must be the carry granularity of the edge pass that filled F/carry");
  int rc = check_graph2(g);
  if (rc) return rc;
  GNB_REQUIRE(0 <= node_begin && node_begin <= node_end && node_end <= g->num_nodes, "gnb_node_update2: bad node range");
  if (node_end == node_begin) return 0;
  GNB_REQUIRE(P && h_in && scale_h && shift_h && h_out, "null pointer");
  if (g->num_edges > 0) GNB_REQUIRE(e16 && F && carry, "null pointer");
  GNB_REQUIRE((xp_ptr == nullptr) == (xp_row == nullptr) && (xp_ptr == nullptr) == (xp_buf == nullptr),
              "gnb_node_update2: xp_ptr / xp_row / xp_buf go together");
  GNB_REQUIRE(ldP >= ((flags & GNB_F_SYMMETRIC) ? 5 : 4) * (int64_t)H && ldP % 4 == 0, "ldP=%lld too small", (long long)ldP);
  GNB_REQUIRE(((uintptr_t)P % 16 == 0) && ((uintptr_t)e16 % 16 == 0) && ((uintptr_t)F % 16 == 0) &&
                  ((uintptr_t)carry % 16 == 0) && ((uintptr_t)h_in % 16 == 0) && ((uintptr_t)h_out % 16 == 0) &&
                  ((uintptr_t)h16_out % 16 == 0) && ((uintptr_t)scale_h % 16 == 0) && ((uintptr_t)shift_h % 16 == 0) &&
                  ((uintptr_t)xp_buf % 16 == 0),
              "pointers must be 16-byte aligned");
  GNB_DISPATCH_H2(H, (tc::node_update2_impl<kH>(g, P, ldP, e16, F, carry, h_in, scale_h, shift_h, h_out, h16_out, flags,
                                                chunk, node_begin, node_end, xp_ptr, xp_row, xp_buf, nullptr,
                                                (cudaStream_t)stream)));
}

extern "C" int gnb_reverse_partial2(const gnb_graph_t* g, int H, const float* P, int64_t ldP, const void* e16,
                                    int64_t node_begin, int64_t node_end, float* out, void* stream) {
  int rc = check_graph2(g);
  if (rc) return rc;
  GNB_REQUIRE(0 <= node_begin && node_begin <= node_end && node_end <= g->num_nodes, "gnb_reverse_partial2: bad node range");
  if (node_end == node_begin) return 0;
  GNB_REQUIRE(P && out && (g->num_edges == 0 || e16), "null pointer");
  GNB_REQUIRE(ldP >= 4 * (int64_t)H && ldP % 4 == 0, "ldP=%lld too small", (long long)ldP);
  GNB_REQUIRE(((uintptr_t)P % 16 == 0) && ((uintptr_t)e16 % 16 == 0) && ((uintptr_t)out % 16 == 0),
              "pointers must be 16-byte aligned");
  GNB_DISPATCH_H2(H, (tc::node_update2_impl<kH>(g, P, ldP, e16, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                nullptr, GNB_F_SYMMETRIC, 1, node_begin, node_end, nullptr, nullptr,
                                                nullptr, out, (cudaStream_t)stream)));
}
