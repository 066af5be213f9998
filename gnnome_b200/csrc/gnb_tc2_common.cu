// Shared host / device pieces of the tensor-core kernels: the spin watchdog record, tensor maps, weight packing and
// the split16 state format (fp32 <-> (hi, lo) fp16 image conversion, see gnb_tma.cuh).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "gnb_tma.cuh"

namespace gnb {
namespace tc {

// ---------------------------------------------------------------------------------------------
// spin watchdog (device side: gnb_tc.cuh)
// ---------------------------------------------------------------------------------------------
static std::mutex g_watch_mutex;
static unsigned long long* g_watch_rec = nullptr;     // host-mapped, kWatchWords words
static bool g_watch_tried = false;
static long long g_watch_timeout_ms = -1;             // -1: not initialised (GNB_SPIN_TIMEOUT_MS or 10 s)

Watch watch_get() {
  std::lock_guard<std::mutex> lock(g_watch_mutex);
  if (g_watch_timeout_ms < 0) {
    const char* env = getenv("GNB_SPIN_TIMEOUT_MS");
    g_watch_timeout_ms = (env != nullptr && *env) ? atoll(env) : 10000;
    if (g_watch_timeout_ms < 0) g_watch_timeout_ms = 0;
  }
  if (!g_watch_tried) {
    g_watch_tried = true;
    void* p = nullptr;
    // survives the sticky error a trap leaves behind: the host reads it through its own pointer
    if (cudaHostAlloc(&p, kWatchWords * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) ==
        cudaSuccess) {
      memset(p, 0, kWatchWords * sizeof(unsigned long long));
      g_watch_rec = (unsigned long long*)p;
    } else {
      (void)cudaGetLastError();
    }
  }
  Watch w;
  w.rec = g_watch_rec;   // unified addressing: the host pointer of mapped memory is valid on the device
  w.timeout_ns = (unsigned long long)g_watch_timeout_ms * 1000000ull;
  return w;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {  // resolved through the runtime: no link-time dependency on libcuda
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_state_map(CUtensorMap* map, const void* base, int64_t rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return GNB_E_ARCH;
  }
  if (((uintptr_t)base & 15) != 0 || K % kKB != 0 || rows <= 0 || box_rows <= 0 || box_rows > 256) {
    set_error("make_state_map: bad matrix (base %p rows %lld K %d box %d)", base, (long long)rows, K, box_rows);
    return GNB_E_INVALID;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)2 * K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)2 * K * sizeof(__half)};
  const cuuint32_t box[2] = {(cuuint32_t)kKB, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d; rows %lld K %d)", (int)r, (long long)rows, K);
    return GNB_E_INVALID;
  }
  return 0;
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

// out16[r][:] = split(in[idx ? idx[r] : r][:]); one thread per 8 consecutive channels
__global__ void split_rows_kernel(const float* __restrict__ in, const int32_t* __restrict__ idx, int64_t rows, int K,
                                  __half* __restrict__ out, const float* __restrict__ scale) {
  const float xs = kXScale * (scale ? *scale : 1.f);   // optional per-tensor pre-scale (a power of two: exact)
  const int k8 = K / 8;
  const int64_t total = rows * k8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k8;
    const int c = (int)(i - r * k8) * 8;
    const int64_t rs = idx ? (int64_t)idx[r] : r;
    const float4 a = *reinterpret_cast<const float4*>(in + rs * K + c);
    const float4 b = *reinterpret_cast<const float4*>(in + rs * K + c + 4);
    float x[8] = {a.x * xs, a.y * xs, a.z * xs, a.w * xs, b.x * xs, b.y * xs, b.z * xs, b.w * xs};
    uint4 h, l;
    split8(x, h, l);
    *reinterpret_cast<uint4*>(out + r * 2 * K + c) = h;
    *reinterpret_cast<uint4*>(out + r * 2 * K + K + c) = l;
  }
}

// out[idx ? idx[r] : r][:] = merge(in16[r][:])
__global__ void merge_rows_kernel(const __half* __restrict__ in, const int32_t* __restrict__ idx, int64_t rows, int K,
                                  float* __restrict__ out) {
  const int k8 = K / 8;
  const int64_t total = rows * k8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k8;
    const int c = (int)(i - r * k8) * 8;
    const int64_t rd = idx ? (int64_t)idx[r] : r;
    const uint4 h = *reinterpret_cast<const uint4*>(in + r * 2 * K + c);
    const uint4 l = *reinterpret_cast<const uint4*>(in + r * 2 * K + K + c);
    const __half2* hp = reinterpret_cast<const __half2*>(&h);
    const __half2* lp = reinterpret_cast<const __half2*>(&l);
    float o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 hf = __half22float2(hp[j]), lf = __half22float2(lp[j]);
      o[2 * j] = (hf.x + lf.x) * kWScale;
      o[2 * j + 1] = (hf.y + lf.y) * kWScale;
    }
    *reinterpret_cast<float4*>(out + rd * K + c) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(out + rd * K + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

}  // namespace tc
}  // namespace gnb

using namespace gnb;

extern "C" void gnb_set_spin_timeout_ms(long long ms) {
  std::lock_guard<std::mutex> lock(tc::g_watch_mutex);
  tc::g_watch_timeout_ms = ms < 0 ? 0 : ms;
}

extern "C" int gnb_hang_report(char* buf, size_t cap) {
  static const char* kKernel[] = {"?", "gnb_edge_forward_tc2", "gnb_node_linear_tc2", "gnb_score_forward_tc2"};
  static const char* kRole[] = {"?", "producer", "mma", "store", "epilogue"};
  static const char* kBar[] = {"?", "full", "empty", "dfull", "dempty", "ofull", "oempty", "ifull"};
  const volatile unsigned long long* r = tc::g_watch_rec;
  if (r == nullptr || r[0] != tc::kWatchMagic) {
    if (buf != nullptr && cap > 0) buf[0] = 0;
    return 0;
  }
  const unsigned tag = (unsigned)r[1];
  const unsigned k = (tag >> 16) & 0xff, role = (tag >> 8) & 0xff, bar = tag & 0xff;
  if (buf != nullptr && cap > 0)
    snprintf(buf, cap,
             "device spin timed out: kernel %s, block %u, thread %u (warp %u, role %s) waited %.1f ms for barrier %s[%u] "
             "parity %u at tile iteration %lld",
             kKernel[k < 4 ? k : 0], (unsigned)(r[2] >> 32), (unsigned)(r[2] & 0xffffffffu),
             (unsigned)(r[2] & 0xffffffffu) / 32, kRole[role < 5 ? role : 0], (double)r[5] * 1e-6, kBar[bar < 8 ? bar : 0],
             (unsigned)(r[3] >> 32), (unsigned)(r[3] & 0xffffffffu), (long long)r[4]);
  return 1;
}

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

extern "C" size_t gnb_split16_bytes(int64_t rows, int K) {
  if (rows < 0 || K <= 0) return 0;
  return (size_t)2 * (size_t)rows * (size_t)K * sizeof(__half);
}

static unsigned stream_blocks(int64_t total) {
  int64_t b = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  return (unsigned)(b < 1 ? 1 : (b < cap ? b : cap));
}

extern "C" int gnb_split_rows(const float* in, const int32_t* idx, int64_t rows, int K, void* out16, const float* scale,
                              void* stream) {
  GNB_REQUIRE(K > 0 && K % 8 == 0, "gnb_split_rows: K=%d must be a positive multiple of 8", K);
  if (rows == 0) return 0;
  GNB_REQUIRE(in && out16, "null pointer");
  GNB_REQUIRE(((uintptr_t)in % 16 == 0) && ((uintptr_t)out16 % 16 == 0), "gnb_split_rows: pointers must be 16-byte aligned");
  tc::split_rows_kernel<<<stream_blocks(rows * (K / 8)), 256, 0, (cudaStream_t)stream>>>(in, idx, rows, K, (__half*)out16, scale);
  return check_launch("gnb_split_rows");
}

extern "C" int gnb_merge_rows(const void* in16, const int32_t* idx, int64_t rows, int K, float* out, void* stream) {
  GNB_REQUIRE(K > 0 && K % 8 == 0, "gnb_merge_rows: K=%d must be a positive multiple of 8", K);
  if (rows == 0) return 0;
  GNB_REQUIRE(in16 && out, "null pointer");
  GNB_REQUIRE(((uintptr_t)in16 % 16 == 0) && ((uintptr_t)out % 16 == 0), "gnb_merge_rows: pointers must be 16-byte aligned");
  tc::merge_rows_kernel<<<stream_blocks(rows * (K / 8)), 256, 0, (cudaStream_t)stream>>>((const __half*)in16, idx, rows, K, out);
  return check_launch("gnb_merge_rows");
}
