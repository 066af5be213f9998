// Input features of the scoring pass, built on the device from the staged graph and the raw overlap attributes
// (what the reference's callers do with torch on the host before model(g, x, e)):
//   x = [z(in_deg), z(out_deg)]        inference.py:413-420, train.py:112-120, utils/data_utils.py:50-51
//   e = [z(overlap_length), overlap_similarity]   utils/data_utils.py:31-41
// with z(v) = (v - v.mean()) / v.std() and torch's unbiased std.  8 bytes per node / edge: HBM-bound streaming passes,
// fp64 accumulation, deterministic (fixed grid, per-block partials summed in a fixed order).
#include "gnb_common.cuh"

namespace gnb {
namespace {

constexpr int kMaxCols = 4;
constexpr int kZBlocks = 512;   // fixed grid: the workspace layout and the summation order do not depend on the device
constexpr int kZThreads = 256;

__global__ void degree_rows_kernel(const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ out_ptr, int64_t N,
                                   int swap, float2* __restrict__ x) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const float din = (float)(in_ptr[i + 1] - in_ptr[i]);
    const float dout = (float)(out_ptr[i + 1] - out_ptr[i]);
    x[i] = swap ? make_float2(dout, din) : make_float2(din, dout);
  }
}

// total[c] = sum over the kZBlocks partials of column c, in block order (every consumer block repeats it: 2 K adds)
__device__ __forceinline__ void load_totals(const double* __restrict__ part, int cols, double* tot_s) {
  if (threadIdx.x < cols) {
    double t = 0.0;
    for (int b = 0; b < kZBlocks; ++b) t += part[threadIdx.x * kZBlocks + b];
    tot_s[threadIdx.x] = t;
  }
  __syncthreads();
}

// part_out[c][block] = sum over this block's rows of d (kPow == 1) or d * d (kPow == 2), d = x[r][c] - mean_c;
// mean_c = 0 for the first pass, sum_c / rows (from the first pass' partials) for the second.
template <int kPow>
__global__ void __launch_bounds__(kZThreads) col_moment_kernel(const float* __restrict__ x, int64_t rows, int cols,
                                                               const double* __restrict__ part_in,
                                                               double* __restrict__ part_out) {
  __shared__ double tot_s[kMaxCols];
  __shared__ double red_s[kZThreads / 32][kMaxCols];
  double mean[kMaxCols], acc[kMaxCols];
#pragma unroll
  for (int c = 0; c < kMaxCols; ++c) mean[c] = 0.0, acc[c] = 0.0;
  if (kPow == 2) {
    load_totals(part_in, cols, tot_s);
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c)
      if (c < cols) mean[c] = tot_s[c] / (double)rows;
  }
  for (int64_t r = (int64_t)blockIdx.x * kZThreads + threadIdx.x; r < rows; r += (int64_t)kZBlocks * kZThreads) {
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c)
      if (c < cols) {
        const double d = (double)x[r * cols + c] - mean[c];
        acc[c] += kPow == 1 ? d : d * d;
      }
  }
#pragma unroll
  for (int c = 0; c < kMaxCols; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], o);
    if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5][c] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x < cols) {
    double t = 0.0;
    for (int w = 0; w < kZThreads / 32; ++w) t += red_s[w][threadIdx.x];
    part_out[threadIdx.x * kZBlocks + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kZThreads) zscore_apply_kernel(float* __restrict__ x, int64_t rows, int cols, int col_mask,
                                                                const double* __restrict__ part1,
                                                                const double* __restrict__ part2,
                                                                double* __restrict__ stats) {
  __shared__ double sum_s[kMaxCols], ssq_s[kMaxCols];
  load_totals(part1, cols, sum_s);
  load_totals(part2, cols, ssq_s);
  float mean[kMaxCols], sd[kMaxCols];
#pragma unroll
  for (int c = 0; c < kMaxCols; ++c)
    if (c < cols) {
      const double m = sum_s[c] / (double)rows;
      const double s = sqrt(ssq_s[c] / (double)(rows - 1));  // rows == 1: 0 / 0 = nan, like torch
      mean[c] = (float)m, sd[c] = (float)s;
      if (stats && blockIdx.x == 0 && threadIdx.x == 0) stats[2 * c] = m, stats[2 * c + 1] = s;
    }
  for (int64_t r = (int64_t)blockIdx.x * kZThreads + threadIdx.x; r < rows; r += (int64_t)kZBlocks * kZThreads) {
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c)
      if (c < cols && ((col_mask >> c) & 1)) x[r * cols + c] = (x[r * cols + c] - mean[c]) / sd[c];
  }
}

}  // namespace
}  // namespace gnb

using namespace gnb;

extern "C" int gnb_degree_rows(const gnb_graph_t* g, int swap, float* x, void* stream) {
  GNB_REQUIRE(g != nullptr && g->in_ptr && g->out_ptr, "gnb_degree_rows: graph not staged");
  if (g->num_nodes == 0) return 0;
  GNB_REQUIRE(x != nullptr && (reinterpret_cast<uintptr_t>(x) & 7) == 0, "gnb_degree_rows: x must be 8-byte aligned");
  const int64_t want = (g->num_nodes + kZThreads - 1) / kZThreads;
  const int blocks = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
  degree_rows_kernel<<<blocks, kZThreads, 0, (cudaStream_t)stream>>>(g->in_ptr, g->out_ptr, g->num_nodes, swap,
                                                                     reinterpret_cast<float2*>(x));
  return check_launch("gnb_degree_rows");
}

extern "C" size_t gnb_zscore_workspace(void) { return (size_t)2 * kMaxCols * kZBlocks * sizeof(double); }

extern "C" int gnb_zscore_cols(float* x, int64_t rows, int cols, int col_mask, double* stats, void* workspace,
                               void* stream) {
  GNB_REQUIRE(cols >= 1 && cols <= kMaxCols, "gnb_zscore_cols: cols=%d not in 1..%d", cols, kMaxCols);
  GNB_REQUIRE(rows >= 0 && col_mask >= 0 && col_mask < (1 << cols), "gnb_zscore_cols: bad rows / col_mask");
  if (rows == 0) return 0;
  GNB_REQUIRE(x != nullptr && workspace != nullptr, "gnb_zscore_cols: null pointer");
  GNB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "gnb_zscore_cols: workspace must be 8-byte aligned");
  double* part1 = static_cast<double*>(workspace);
  double* part2 = part1 + kMaxCols * kZBlocks;
  cudaStream_t st = (cudaStream_t)stream;
  col_moment_kernel<1><<<kZBlocks, kZThreads, 0, st>>>(x, rows, cols, nullptr, part1);
  col_moment_kernel<2><<<kZBlocks, kZThreads, 0, st>>>(x, rows, cols, part1, part2);
  zscore_apply_kernel<<<kZBlocks, kZThreads, 0, st>>>(x, rows, cols, col_mask, part1, part2, stats);
  return check_launch("gnb_zscore_cols");
}
