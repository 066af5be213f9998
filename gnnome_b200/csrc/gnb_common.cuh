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
  // 1 / (1 + 2^(-x log2 e)) with the two MUFU approximations (ex2: 2 ulp, rcp: 1 ulp) -- four instructions,
  // ~3e-7 absolute error; saturates cleanly (ex2 -> inf -> rcp -> 0, ex2 -> 0 -> 1).
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
  return r;
}

__device__ __forceinline__ float gate_div(float num, float den) {
  // num / (den + 1e-6): den >= 0, so the approximate reciprocal (1 ulp) is safe
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den + kGateEps));
  return num * r;
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
