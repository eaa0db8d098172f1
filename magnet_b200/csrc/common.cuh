// Shared device helpers for the MaGNet matching kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/magnet_b200.h"

namespace magnet {

// Kernel-side view of magnet_cost_args; k values travel in the launch parameters
// (constant bank, uniform loads) so that no __constant__ symbol / extra copy is needed.
struct CostParams {
  int B, V, D, H, W, HW;
  int k_sorted;       // 1 when k[0..D) is non-decreasing (enables the analytic cell walk)
  float kappa;
  float inv_v_exact;  // 1/V when V is a power of two (exact), else 0 -> use IEEE division
  float vf;           // float(V)
  const float* __restrict__ ref_feat;
  const float* __restrict__ src_feat;
  const float* __restrict__ src_gmm;
  const float* __restrict__ rays;
  const magnet_camera* __restrict__ cams;
  const float* __restrict__ d_volume;
  const float* __restrict__ ref_gmm;
  float* __restrict__ out;
  float k[MAGNET_MAX_PLANES];
};

// Kernel-side arguments of the F-volume backward (magnet_cost_f_bwd_args + the geometry of its forward call).
struct BwdParams {
  int B, V, D, C, H, W, HW;
  int softmax;
  float vf;
  const float* __restrict__ ref_feat;   // (B,C,H,W)
  const float* __restrict__ src_feat;   // (V*B,C,H,W) NCHW
  const float* __restrict__ rays;
  const magnet_camera* __restrict__ cams;
  const float* __restrict__ prob;       // (B,D,H,W) forward output (softmax == 1) or unused
  const float* __restrict__ grad_out;   // (B,D,H,W)
  float* __restrict__ g_score;          // (B,D,H,W) workspace
  float* __restrict__ grad_ref;         // (B,C,H,W), written
  float* __restrict__ grad_src;         // (V*B,C,H,W), accumulated with atomics (caller zeroes it)
  float k[MAGNET_MAX_PLANES];
};

// 1/x to <1 ulp: MUFU.RCP + one Newton step.  x == 0 -> NaN/inf, which the callers'
// coordinate clamp turns into "out of bounds" (same outcome as the reference's +-10 clamp).
__device__ __forceinline__ float rcp_nr(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return __fmaf_rn(r, __fmaf_rn(-x, r, 1.0f), r);
}

__device__ __forceinline__ float ldg_f(const float* p) { return __ldg(p); }

}  // namespace magnet
