// Graph staging: COO edge list -> dst-CSR + src-CSR views (gnb_graph_t), plus row gather/scatter.
// Replaces the index structures DGL builds for apply_edges / update_all / dgl.reverse
// (reference layers/gated_gcn_full.py:99,104,112,125).  Device-side, two stable radix sorts.
#include <cub/device/device_radix_sort.cuh>
#include <stdarg.h>

#include <string>

#include "gnb_common.cuh"

namespace gnb {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}

__global__ void iota_kernel(int32_t* out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)i;
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ in, const int32_t* __restrict__ idx,
                                  int32_t* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

// keys sorted ascending, values in [0, num_nodes): ptr[v] = first position whose key >= v
__global__ void row_ptr_kernel(const int32_t* __restrict__ keys, int64_t n, int64_t num_nodes,
                               int32_t* __restrict__ ptr) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  int64_t prev = (p == 0) ? -1 : keys[p - 1];
  int64_t cur = (p == n) ? num_nodes : keys[p];
  for (int64_t v = prev + 1; v <= cur; ++v) ptr[v] = (int32_t)p;
}

// ld_in / ld_out: row strides in float4 units
template <bool kScatter>
__global__ void move_rows_kernel(const float4* __restrict__ in, int64_t ld_in, const int32_t* __restrict__ idx,
                                 int64_t rows, int w4, float4* __restrict__ out, int64_t ld_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = rows * w4;
  for (; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / w4;
    int c = (int)(i - r * w4);
    int64_t other = idx[r];
    if (kScatter) out[other * ld_out + c] = in[r * ld_in + c];
    else out[r * ld_out + c] = in[other * ld_in + c];
  }
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int key_bits(int64_t num_nodes) {
  int b = 1;
  while (((int64_t)1 << b) < num_nodes) ++b;
  return b;
}

static size_t cub_sort_bytes(int64_t E, int bits) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)E, 0, bits);
  return bytes;
}

}  // namespace gnb

using namespace gnb;

extern "C" int gnb_abi_version(void) { return GNB_ABI_VERSION; }

extern "C" const char* gnb_last_error(void) { return g_error.c_str(); }

extern "C" int gnb_graph_stage_workspace(int64_t E, int64_t N, size_t* bytes) {
  GNB_REQUIRE(bytes != nullptr, "bytes is null");
  GNB_REQUIRE(E >= 0 && N >= 0 && E < (int64_t)2147483647 && N < (int64_t)2147483647,
              "graph too large for int32 indices (E=%lld N=%lld)", (long long)E, (long long)N);
  size_t e_bytes = align_up((size_t)(E > 0 ? E : 1) * sizeof(int32_t), 256);
  *bytes = 2 * e_bytes + align_up(cub_sort_bytes(E > 0 ? E : 1, key_bits(N > 1 ? N : 2)), 256);
  return 0;
}

extern "C" int gnb_graph_stage(const int32_t* src, const int32_t* dst, gnb_graph_t* g,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  GNB_REQUIRE(g != nullptr, "graph is null");
  const int64_t E = g->num_edges, N = g->num_nodes;
  size_t need = 0;
  int rc = gnb_graph_stage_workspace(E, N, &need);
  if (rc) return rc;
  if (workspace_bytes < need) {
    set_error("gnb_graph_stage: workspace %zu < %zu bytes", workspace_bytes, need);
    return GNB_E_WORKSPACE;
  }
  GNB_REQUIRE(g->in_ptr && g->out_ptr, "row pointer arrays are null");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int T = 256;
  if (E == 0) {
    GNB_CUDA(cudaMemsetAsync(g->in_ptr, 0, (size_t)(N + 1) * sizeof(int32_t), stream));
    GNB_CUDA(cudaMemsetAsync(g->out_ptr, 0, (size_t)(N + 1) * sizeof(int32_t), stream));
    return 0;
  }
  GNB_REQUIRE(src && dst && workspace && g->in_src && g->in_dst && g->in_eid && g->out_pos && g->out_dst,
              "null array");
  const size_t e_bytes = align_up((size_t)E * sizeof(int32_t), 256);
  int32_t* iota = (int32_t*)workspace;
  int32_t* keys_sorted = (int32_t*)((char*)workspace + e_bytes);
  void* cub_tmp = (char*)workspace + 2 * e_bytes;
  size_t cub_bytes = workspace_bytes - 2 * e_bytes;
  const int bits = key_bits(N > 1 ? N : 2);
  const unsigned blocksE = (unsigned)((E + T - 1) / T), blocksE1 = (unsigned)((E + T) / T);

  iota_kernel<<<blocksE, T, 0, stream>>>(iota, E);
  // view 1: stable sort by dst -> position order p
  GNB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, dst, g->in_dst, (const int32_t*)iota,
                                           g->in_eid, (int)E, 0, bits, stream));
  gather_i32_kernel<<<blocksE, T, 0, stream>>>(src, g->in_eid, g->in_src, E);
  row_ptr_kernel<<<blocksE1, T, 0, stream>>>(g->in_dst, E, N, g->in_ptr);
  // view 2: stable sort of the positions by src
  GNB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, (const int32_t*)g->in_src, keys_sorted,
                                           (const int32_t*)iota, g->out_pos, (int)E, 0, bits, stream));
  gather_i32_kernel<<<blocksE, T, 0, stream>>>(g->in_dst, g->out_pos, g->out_dst, E);
  row_ptr_kernel<<<blocksE1, T, 0, stream>>>(keys_sorted, E, N, g->out_ptr);
  return check_launch("gnb_graph_stage");
}

static int move_rows(bool scatter, const float* in, int64_t ld_in, const int32_t* idx, int64_t rows, int W,
                     float* out, int64_t ld_out, void* stream_) {
  GNB_REQUIRE(W > 0 && W % 4 == 0, "row width %d must be a positive multiple of 4", W);
  GNB_REQUIRE(ld_in >= W && ld_out >= W && ld_in % 4 == 0 && ld_out % 4 == 0, "row strides must be multiples of 4, >= W");
  GNB_REQUIRE(((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0), "rows must be 16-byte aligned");
  if (rows == 0) return 0;
  GNB_REQUIRE(in && idx && out, "null pointer");
  int64_t total = rows * (W / 4);
  unsigned blocks = (unsigned)((total + 255) / 256 < (int64_t)sm_count() * 16 ? (total + 255) / 256
                                                                              : (int64_t)sm_count() * 16);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (scatter)
    move_rows_kernel<true><<<blocks, 256, 0, stream>>>((const float4*)in, ld_in / 4, idx, rows, W / 4, (float4*)out, ld_out / 4);
  else
    move_rows_kernel<false><<<blocks, 256, 0, stream>>>((const float4*)in, ld_in / 4, idx, rows, W / 4, (float4*)out, ld_out / 4);
  return check_launch("gnb_move_rows");
}

extern "C" int gnb_gather_rows(const float* in, const int32_t* idx, int64_t rows, int W, float* out,
                               void* stream) {
  return move_rows(false, in, W, idx, rows, W, out, W, stream);
}

extern "C" int gnb_gather_rows_ld(const float* in, int64_t ld_in, const int32_t* idx, int64_t rows, int W,
                                  float* out, int64_t ld_out, void* stream) {
  return move_rows(false, in, ld_in, idx, rows, W, out, ld_out, stream);
}

extern "C" int gnb_scatter_rows(const float* in, const int32_t* idx, int64_t rows, int W, float* out,
                                void* stream) {
  return move_rows(true, in, W, idx, rows, W, out, W, stream);
}
