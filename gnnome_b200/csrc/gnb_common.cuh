// Shared helpers for the gnnome_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gnnome_b200.h"

namespace gnb {

constexpr int kThreads = 256;      // every edge/node kernel runs 256-thread CTAs
constexpr int kTileRows = 32;      // edge rows one channel-group consumes per tile
constexpr int kTilesPerChunk = 16; // chunk = 512 consecutive positions; carries exist per chunk
constexpr int kChunk = kTileRows * kTilesPerChunk;
constexpr float kGateEps = 1e-6f;  // gated_gcn_full.py:114,127

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

inline bool supported_h(int H) { return H == 32 || H == 64 || H == 128 || H == 256; }

__device__ __forceinline__ float sigmoidf_fast(float x) {
  // 1 / (1 + 2^(-x log2 e)): ex2.approx + rcp.approx, ~2 ulp; saturates cleanly at +-inf
  return __frcp_rn(1.0f + exp2f(-1.4426950408889634f * x));
}

}  // namespace gnb

#define GNB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      gnb::set_error(__VA_ARGS__);        \
      return GNB_E_INVALID;               \
    }                                     \
  } while (0)

#define GNB_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      gnb::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));               \
      return (int)_e;                                                               \
    }                                                                               \
  } while (0)
