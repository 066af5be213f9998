// Node-induced subgraph on the device: dgl.node_subgraph(g, keep, store_ids=True) as used by the reference's strand-wise
// masking (train.py:91-100) and by its mini-batches (sub_g.ndata['_ID'] / sub_g.edata['_ID'], train.py:125-135).
//   kept nodes are renumbered in increasing order of their original id, induced edges (both endpoints kept) keep the
//   relative order of their original ids; node_id / edge_id are the '_ID' maps back to the parent graph.
// Integer work, bit-exact: flag -> block counts -> scan of the block counts -> in-block scan + scatter.  Two passes over
// the edge list (8 + 1 bytes per edge read twice, 12 bytes per induced edge written): HBM-bound streaming.
#include "gnb_common.cuh"

namespace gnb {
namespace {

constexpr int kSgThreads = 256;
constexpr int kSgItems = 4;                          // consecutive items per thread
constexpr int kSgBlock = kSgThreads * kSgItems;      // items per block

inline int64_t sg_blocks(int64_t n) { return (n + kSgBlock - 1) / kSgBlock; }

struct SubgraphWs {
  int32_t* new_id;   // [N]  rank of node i among the kept nodes (meaningful where keep[i])
  int32_t* boff_n;   // [blocks(N)] kept nodes before block b
  int32_t* boff_e;   // [blocks(E)] induced edges before block b
};

inline size_t ws_layout(int64_t N, int64_t E, void* base, SubgraphWs* w) {
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = 0;
  char* p = static_cast<char*>(base);
  if (w) w->new_id = reinterpret_cast<int32_t*>(p + off);
  off += up((size_t)(N > 0 ? N : 1) * 4);
  if (w) w->boff_n = reinterpret_cast<int32_t*>(p + off);
  off += up((size_t)(sg_blocks(N) + 1) * 4);
  if (w) w->boff_e = reinterpret_cast<int32_t*>(p + off);
  off += up((size_t)(sg_blocks(E) + 1) * 4);
  return off;
}

// the thread's kSgItems consecutive items; one 16-byte (4-byte for the flags) load when they all exist
__device__ __forceinline__ void load_ids(const int32_t* __restrict__ p, int64_t base, int64_t n, int (&v)[kSgItems]) {
  if (base + kSgItems <= n) {
    const int4 t = *reinterpret_cast<const int4*>(p + base);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  } else {
#pragma unroll
    for (int k = 0; k < kSgItems; ++k) v[k] = (base + k < n) ? p[base + k] : -1;
  }
}

__device__ __forceinline__ void load_keep(const uint8_t* __restrict__ keep, int64_t base, int64_t n, int (&f)[kSgItems]) {
  if (base + kSgItems <= n) {
    const uchar4 t = *reinterpret_cast<const uchar4*>(keep + base);
    f[0] = t.x != 0, f[1] = t.y != 0, f[2] = t.z != 0, f[3] = t.w != 0;
  } else {
#pragma unroll
    for (int k = 0; k < kSgItems; ++k) f[k] = (base + k < n) ? (keep[base + k] != 0) : 0;
  }
}

// flags of the thread's items: kept nodes, or edges with both endpoints kept (s / d: their endpoints)
template <bool kEdges>
__device__ __forceinline__ int sg_flags(const uint8_t* __restrict__ keep, const int32_t* __restrict__ src,
                                        const int32_t* __restrict__ dst, int64_t base, int64_t n, int (&f)[kSgItems],
                                        int (&s)[kSgItems], int (&d)[kSgItems]) {
  static_assert(kSgItems == 4, "vector loads assume four items per thread");
  int mine = 0;
  if (kEdges) {
    load_ids(src, base, n, s);
    load_ids(dst, base, n, d);
#pragma unroll
    for (int k = 0; k < kSgItems; ++k) {
      f[k] = (s[k] >= 0) && (keep[s[k]] != 0) && (keep[d[k]] != 0);
      mine += f[k];
    }
  } else {
    load_keep(keep, base, n, f);
#pragma unroll
    for (int k = 0; k < kSgItems; ++k) mine += f[k];
  }
  return mine;
}

// exclusive prefix of `mine` over the block (thread order) and the block total
__device__ __forceinline__ int block_exclusive(int mine, int* total) {
  __shared__ int warp_s[kSgThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_s[warp] = incl;
  __syncthreads();
  int before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < kSgThreads / 32; ++w) {
    const int v = warp_s[w];
    if (w < warp) before += v;
    all += v;
  }
  __syncthreads();
  *total = all;
  return before + incl - mine;
}

template <bool kEdges>
__global__ void __launch_bounds__(kSgThreads) sg_count_kernel(const uint8_t* __restrict__ keep,
                                                               const int32_t* __restrict__ src,
                                                               const int32_t* __restrict__ dst, int64_t n,
                                                               int32_t* __restrict__ bcount) {
  const int64_t base = (int64_t)blockIdx.x * kSgBlock + (int64_t)threadIdx.x * kSgItems;
  int f[kSgItems], s[kSgItems], d[kSgItems];
  const int mine = sg_flags<kEdges>(keep, src, dst, base, n, f, s, d);
  int total;
  block_exclusive(mine, &total);
  if (threadIdx.x == 0) bcount[blockIdx.x] = total;
}

// in place: counts -> exclusive offsets; total[0] = sum.  One block walks the array (<= 60 K entries for 60 M edges).
__global__ void __launch_bounds__(kSgThreads) sg_scan_blocks_kernel(int32_t* __restrict__ b, int64_t nb,
                                                                    int64_t* __restrict__ total) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += kSgThreads) {
    const int64_t i = base + threadIdx.x;
    const int mine = (i < nb) ? b[i] : 0;
    int all;
    const int excl = block_exclusive(mine, &all);
    const int carry = carry_s;
    if (i < nb) b[i] = carry + excl;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + all;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = carry_s;
}

__global__ void __launch_bounds__(kSgThreads) sg_rank_nodes_kernel(const uint8_t* __restrict__ keep, int64_t n,
                                                                    const int32_t* __restrict__ boff,
                                                                    int32_t* __restrict__ new_id) {
  const int64_t base = (int64_t)blockIdx.x * kSgBlock + (int64_t)threadIdx.x * kSgItems;
  int f[kSgItems], s[kSgItems], d[kSgItems];
  const int mine = sg_flags<false>(keep, nullptr, nullptr, base, n, f, s, d);
  int total;
  int r = boff[blockIdx.x] + block_exclusive(mine, &total);
#pragma unroll
  for (int k = 0; k < kSgItems; ++k)
    if (base + k < n) {
      new_id[base + k] = r;
      r += f[k];
    }
}

__global__ void __launch_bounds__(kSgThreads) sg_fill_nodes_kernel(const uint8_t* __restrict__ keep, int64_t n,
                                                                    const int32_t* __restrict__ new_id,
                                                                    int32_t* __restrict__ node_id) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (keep[i]) node_id[new_id[i]] = (int32_t)i;
}

__global__ void __launch_bounds__(kSgThreads) sg_fill_edges_kernel(const uint8_t* __restrict__ keep,
                                                                    const int32_t* __restrict__ src,
                                                                    const int32_t* __restrict__ dst, int64_t n,
                                                                    const int32_t* __restrict__ boff,
                                                                    const int32_t* __restrict__ new_id,
                                                                    int32_t* __restrict__ edge_id,
                                                                    int32_t* __restrict__ sub_src,
                                                                    int32_t* __restrict__ sub_dst) {
  const int64_t base = (int64_t)blockIdx.x * kSgBlock + (int64_t)threadIdx.x * kSgItems;
  int f[kSgItems], s[kSgItems], d[kSgItems];
  const int mine = sg_flags<true>(keep, src, dst, base, n, f, s, d);
  int total;
  int r = boff[blockIdx.x] + block_exclusive(mine, &total);
#pragma unroll
  for (int k = 0; k < kSgItems; ++k)
    if (f[k]) {
      edge_id[r] = (int32_t)(base + k);
      sub_src[r] = new_id[s[k]];
      sub_dst[r] = new_id[d[k]];
      ++r;
    }
}

}  // namespace
}  // namespace gnb

using namespace gnb;

extern "C" int gnb_subgraph_workspace(int64_t num_nodes, int64_t num_edges, size_t* bytes) {
  GNB_REQUIRE(bytes != nullptr, "gnb_subgraph_workspace: null pointer");
  GNB_REQUIRE(num_nodes >= 0 && num_edges >= 0 && num_nodes < (1ll << 31) && num_edges < (1ll << 31),
              "gnb_subgraph_workspace: N=%lld / E=%lld too large for int32 ids", (long long)num_nodes, (long long)num_edges);
  *bytes = ws_layout(num_nodes, num_edges, nullptr, nullptr);
  return 0;
}

extern "C" int gnb_subgraph_count(const uint8_t* keep, const int32_t* src, const int32_t* dst, int64_t num_nodes,
                                  int64_t num_edges, void* workspace, size_t workspace_bytes, int64_t* counts,
                                  void* stream) {
  size_t need = 0;
  int rc = gnb_subgraph_workspace(num_nodes, num_edges, &need);
  if (rc) return rc;
  GNB_REQUIRE(counts != nullptr && workspace != nullptr, "gnb_subgraph_count: null pointer");
  if (workspace_bytes < need) {
    set_error("gnb_subgraph_count: workspace %zu < %zu bytes", workspace_bytes, need);
    return GNB_E_WORKSPACE;
  }
  GNB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gnb_subgraph_count: workspace must be 256-byte aligned");
  GNB_REQUIRE((keep || num_nodes == 0) && ((src && dst) || num_edges == 0), "gnb_subgraph_count: null pointer");
  GNB_REQUIRE((reinterpret_cast<uintptr_t>(keep) & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
              "gnb_subgraph_count: keep must be 4-byte, src / dst 16-byte aligned");
  SubgraphWs w;
  ws_layout(num_nodes, num_edges, workspace, &w);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nbn = sg_blocks(num_nodes), nbe = sg_blocks(num_edges);
  if (nbn > 0) sg_count_kernel<false><<<(unsigned)nbn, kSgThreads, 0, st>>>(keep, nullptr, nullptr, num_nodes, w.boff_n);
  sg_scan_blocks_kernel<<<1, kSgThreads, 0, st>>>(w.boff_n, nbn, counts);
  if (nbn > 0) sg_rank_nodes_kernel<<<(unsigned)nbn, kSgThreads, 0, st>>>(keep, num_nodes, w.boff_n, w.new_id);
  if (nbe > 0) sg_count_kernel<true><<<(unsigned)nbe, kSgThreads, 0, st>>>(keep, src, dst, num_edges, w.boff_e);
  sg_scan_blocks_kernel<<<1, kSgThreads, 0, st>>>(w.boff_e, nbe, counts + 1);
  return check_launch("gnb_subgraph_count");
}

extern "C" int gnb_subgraph_fill(const uint8_t* keep, const int32_t* src, const int32_t* dst, int64_t num_nodes,
                                 int64_t num_edges, void* workspace, int32_t* node_id, int32_t* edge_id,
                                 int32_t* sub_src, int32_t* sub_dst, void* stream) {
  size_t need = 0;
  int rc = gnb_subgraph_workspace(num_nodes, num_edges, &need);
  if (rc) return rc;
  GNB_REQUIRE(workspace != nullptr, "gnb_subgraph_fill: null pointer");
  // the outputs may be NULL when pass 1 counted no kept node / no induced edge (nothing is written then)
  GNB_REQUIRE(keep || num_nodes == 0, "gnb_subgraph_fill: null pointer");
  GNB_REQUIRE((src && dst) || num_edges == 0, "gnb_subgraph_fill: null pointer");
  GNB_REQUIRE((reinterpret_cast<uintptr_t>(keep) & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
              "gnb_subgraph_fill: keep must be 4-byte, src / dst 16-byte aligned");
  SubgraphWs w;
  ws_layout(num_nodes, num_edges, workspace, &w);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nbe = sg_blocks(num_edges);
  if (num_nodes > 0) {
    const int64_t want = (num_nodes + kSgThreads - 1) / kSgThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    sg_fill_nodes_kernel<<<(unsigned)(want < cap ? want : cap), kSgThreads, 0, st>>>(keep, num_nodes, w.new_id, node_id);
  }
  if (nbe > 0)
    sg_fill_edges_kernel<<<(unsigned)nbe, kSgThreads, 0, st>>>(keep, src, dst, num_edges, w.boff_e, w.new_id, edge_id,
                                                             sub_src, sub_dst);
  return check_launch("gnb_subgraph_fill");
}
