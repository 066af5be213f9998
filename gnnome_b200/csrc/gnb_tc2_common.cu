// split16 state format helpers: tensor maps, fp32 <-> (hi, lo) fp16 image conversion (see gnb_tma.cuh).
#include <cuda_runtime.h>

#include "gnb_tma.cuh"

namespace gnb {
namespace tc {

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

// out16[r][:] = split(in[idx ? idx[r] : r][:]); one thread per 8 consecutive channels
__global__ void split_rows_kernel(const float* __restrict__ in, const int32_t* __restrict__ idx, int64_t rows, int K,
                                  __half* __restrict__ out) {
  const int k8 = K / 8;
  const int64_t total = rows * k8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k8;
    const int c = (int)(i - r * k8) * 8;
    const int64_t rs = idx ? (int64_t)idx[r] : r;
    const float4 a = *reinterpret_cast<const float4*>(in + rs * K + c);
    const float4 b = *reinterpret_cast<const float4*>(in + rs * K + c + 4);
    float x[8] = {a.x * kXScale, a.y * kXScale, a.z * kXScale, a.w * kXScale,
                  b.x * kXScale, b.y * kXScale, b.z * kXScale, b.w * kXScale};
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

extern "C" size_t gnb_split16_bytes(int64_t rows, int K) {
  if (rows < 0 || K <= 0) return 0;
  return (size_t)2 * (size_t)rows * (size_t)K * sizeof(__half);
}

static unsigned stream_blocks(int64_t total) {
  int64_t b = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  return (unsigned)(b < 1 ? 1 : (b < cap ? b : cap));
}

extern "C" int gnb_split_rows(const float* in, const int32_t* idx, int64_t rows, int K, void* out16, void* stream) {
  GNB_REQUIRE(K > 0 && K % 8 == 0, "gnb_split_rows: K=%d must be a positive multiple of 8", K);
  if (rows == 0) return 0;
  GNB_REQUIRE(in && out16, "null pointer");
  GNB_REQUIRE(((uintptr_t)in % 16 == 0) && ((uintptr_t)out16 % 16 == 0), "gnb_split_rows: pointers must be 16-byte aligned");
  tc::split_rows_kernel<<<stream_blocks(rows * (K / 8)), 256, 0, (cudaStream_t)stream>>>(in, idx, rows, K, (__half*)out16);
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
