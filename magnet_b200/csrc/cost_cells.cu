// MAGNET_VARIANT_CELLS — tap-sharing fused warp + sample + consistency + view-fusion kernel.
//
// Replaces homography.py:79-161 (and :10-75 with CW == false) without materialising any
// D x C x H x W tensor.  Key identity (SURVEY §7 hard part 1):
//     sum_c ref_c * (sum_t w_t * src_{t,c})  ==  sum_t w_t * <ref, src_t>
// The D hypotheses of one reference pixel project onto a short epipolar segment of the source
// view, so they fall into only a handful of bilinear cells (about 6 at 640x480 / D=64).  Each
// distinct cell needs 4 channel dot products <ref[p], src[tap]> (plus 4 taps of source mu and
// sigma); every hypothesis inside the cell is then a 3-FMA bilinear polynomial per quantity.
//
// Mapping: one thread per reference pixel (a warp = 32 consecutive pixels of a row, so source
// gathers of neighbouring lanes hit neighbouring addresses), loop over views inside the thread.
// Per (pixel, view) the work is split into three phases that every lane of a warp executes in
// LOCKSTEP — a run-per-cell loop would diverge 2.2x on the bench workload (measured by
// simulation):
//   A  walk the hypotheses j in order, record every change of bilinear cell -> per-lane cell list
//      (headers in shared memory, at most NCELL per round);
//   B  for cell i = 0..warp-max: 4 x C-channel dot products + mu/sigma taps -> 12 polynomial
//      coefficients per cell, stored as 3 float4 per lane in shared memory;
//   C  walk the hypotheses again; on leaving the current cell reload the next record (3 LDS.128,
//      the only divergent code), evaluate cost / mu~ / sigma~, apply the consistency test and
//      accumulate over views in a shared-memory column owned by the lane.
// If a lane needs more than NCELL cells (incoherent depth, e.g. random test inputs) the warp
// processes the hypotheses in several rounds [j_lo, j_end), j_end = warp-min of the first
// hypothesis a lane could not record — always correct, no separate slow path.
//
// Numerics (DESIGN.md "parity"): same formulas as the reference, but (i) 1/Zp via MUFU.RCP + one
// Newton step and ix = P0/Zp - 0.5 folded into one FMA instead of the normalise / clamp /
// unnormalise round trip, (ii) channel sums re-associated (dot-then-blend), (iii) bilinear
// polynomial instead of 4 explicit weights, (iv) fp32 view accumulation.  Each changes results at
// the 1e-6 relative level; the hard consistency threshold can flip for elements within ~1e-6 of it.
#include "common.cuh"

namespace magnet {

constexpr int NT = 128;    // threads per CTA = reference pixels per CTA
constexpr int NCELL = 8;   // cell records per lane per round

// dynamic shared memory: rec[NCELL][3][NT] float4 | hdr[NCELL][NT] float2 | acc[D][NT] float
__host__ __device__ inline size_t cells_smem_bytes(int D) {
  return (size_t)NCELL * 3 * NT * 16 + (size_t)NCELL * NT * 8 + (size_t)D * NT * 4;
}

template <int MODE>
struct DepthSrc {
  float mu, sg;
  const float* dv;   // d_volume + b*D*HW + n
  int HW;
};

// Projection of hypothesis j: continuous source-image sample position (ix, iy) = projected
// pixel - 0.5, and z = depth in the source camera.  Used by phases A and C;
// written with explicit intrinsics so both phases get bit-identical results.
template <int MODE>
__device__ __forceinline__ void project(const CostParams& p, const DepthSrc<MODE>& ds, int j,
                                        float a0, float a1, float a2, float q0, float q1, float q2,
                                        float xmax, float ymax, float& ix, float& iy, float& z) {
  float d;
  if (MODE == MAGNET_DEPTH_VOLUME) d = ldg_f(ds.dv + (size_t)j * ds.HW);
  else if (MODE == MAGNET_DEPTH_GAUSS) d = __fadd_rn(ds.mu, __fmul_rn(ds.sg, p.k[j]));
  else d = p.k[j];
  const float P0 = __fmaf_rn(q0, d, a0);
  const float P1 = __fmaf_rn(q1, d, a1);
  z = __fadd_rn(a2, __fmul_rn(q2, d));            // exactly the reference's z (mul, then add)
  const float r = rcp_nr(__fadd_rn(z, 1e-10f));
  ix = __fmaf_rn(P0, r, -0.5f);
  iy = __fmaf_rn(P1, r, -0.5f);
  // Anything left of -1 / right of W (above / below likewise) has all four taps out of bounds;
  // clamp so that cell coordinates stay small and NaN (fmaxf drops it) maps to "out of bounds".
  ix = fminf(fmaxf(ix, -2.0f), xmax);
  iy = fminf(fmaxf(iy, -2.0f), ymax);
}

template <int C, int LAYOUT>
__device__ __forceinline__ float tap_dot(const float* __restrict__ src_img, const float (&ref)[C], int off,
                                         int HW) {
  float s0 = 0.0f, s1 = 0.0f;
  if (LAYOUT == MAGNET_SRC_C4HW4) {
    const float4* s = reinterpret_cast<const float4*>(src_img) + off;
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 t = __ldg(s + (size_t)c4 * HW);
      s0 = __fmaf_rn(ref[4 * c4 + 0], t.x, s0);
      s1 = __fmaf_rn(ref[4 * c4 + 1], t.y, s1);
      s0 = __fmaf_rn(ref[4 * c4 + 2], t.z, s0);
      s1 = __fmaf_rn(ref[4 * c4 + 3], t.w, s1);
    }
  } else {
    const float* s = src_img + off;
#pragma unroll
    for (int c = 0; c < C; c += 2) {
      s0 = __fmaf_rn(ref[c], ldg_f(s + (size_t)c * HW), s0);
      s1 = __fmaf_rn(ref[c + 1], ldg_f(s + (size_t)(c + 1) * HW), s1);
    }
  }
  return s0 + s1;
}

template <int C, int MODE, int LAYOUT, bool CW, bool SOFTMAX>
__global__ void __launch_bounds__(NT)
cost_cells_kernel(const __grid_constant__ CostParams p) {
  extern __shared__ float4 smem4[];
  float4* rec = smem4;                                                   // [NCELL][3][NT]
  float2* hdr = reinterpret_cast<float2*>(smem4 + NCELL * 3 * NT);       // [NCELL][NT]
  float* acc = reinterpret_cast<float*>(hdr + NCELL * NT);               // [D][NT]

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int H = p.H, W = p.W, HW = p.HW, D = p.D;
  const int n_raw = blockIdx.x * NT + tid;
  const bool live = n_raw < HW;
  const int n = live ? n_raw : HW - 1;          // dead lanes shadow the last pixel, never store
  const unsigned FULL = 0xffffffffu;

  float ref[C];
  {
    const float* rp = p.ref_feat + (size_t)b * C * HW + n;
#pragma unroll
    for (int c = 0; c < C; ++c) ref[c] = ldg_f(rp + (size_t)c * HW);
  }
  const float r0 = ldg_f(p.rays + ((size_t)b * 3 + 0) * HW + n);
  const float r1 = ldg_f(p.rays + ((size_t)b * 3 + 1) * HW + n);
  const float r2 = ldg_f(p.rays + ((size_t)b * 3 + 2) * HW + n);

  DepthSrc<MODE> ds;
  ds.HW = HW;
  ds.dv = nullptr;
  ds.mu = ds.sg = 0.0f;
  if (MODE == MAGNET_DEPTH_VOLUME) ds.dv = p.d_volume + (size_t)b * D * HW + n;
  if (MODE == MAGNET_DEPTH_GAUSS) {
    ds.mu = ldg_f(p.ref_gmm + ((size_t)b * 2 + 0) * HW + n);
    ds.sg = ldg_f(p.ref_gmm + ((size_t)b * 2 + 1) * HW + n);
  }
  for (int j = 0; j < D; ++j) acc[j * NT + tid] = 0.0f;

  const float xmax = (float)W + 1.0f, ymax = (float)H + 1.0f;

  for (int v = 0; v < p.V; ++v) {
    const magnet_camera* cam = p.cams + (b * p.V + v);
    if (cam->valid != 1.0f) continue;                                    // CTA-uniform
    const float a0 = cam->a[0], a1 = cam->a[1], a2 = cam->a[2];
    const float q0 = __fmaf_rn(cam->A[2], r2, __fmaf_rn(cam->A[1], r1, __fmul_rn(cam->A[0], r0)));
    const float q1 = __fmaf_rn(cam->A[5], r2, __fmaf_rn(cam->A[4], r1, __fmul_rn(cam->A[3], r0)));
    const float q2 = __fmaf_rn(cam->A[8], r2, __fmaf_rn(cam->A[7], r1, __fmul_rn(cam->A[6], r0)));
    const int vb = v * p.B + b;
    const float* src_img = p.src_feat + (size_t)vb * C * HW;
    const float* gm = CW ? p.src_gmm + (size_t)vb * 2 * HW : nullptr;

    int j_lo = 0;
    while (j_lo < D) {                                                   // rounds; warp-uniform
      // ---------------- phase A: cell list --------------------------------------------------
      int ncell = 0, j_stop = D;
      {
        float cx = -1e30f, cy = -1e30f;
        for (int j = j_lo; j < D; ++j) {
          float ix, iy, z;
          project<MODE>(p, ds, j, a0, a1, a2, q0, q1, q2, xmax, ymax, ix, iy, z);
          const float fx = ix - cx, fy = iy - cy;
          if (!(fx >= 0.0f && fx < 1.0f && fy >= 0.0f && fy < 1.0f)) {
            if (ncell == NCELL) { j_stop = j; break; }
            cx = floorf(ix);
            cy = floorf(iy);
            hdr[ncell * NT + tid] = make_float2(cx, cy);
            ++ncell;
          }
        }
      }
      const int j_end = __reduce_min_sync(FULL, j_stop);
      const int nmax = __reduce_max_sync(FULL, ncell);

      // ---------------- phase B: per-cell records -------------------------------------------
      for (int i = 0; i < nmax; ++i) {
        if (i < ncell) {
          const float2 h = hdr[i * NT + tid];
          const int x0 = (int)h.x, y0 = (int)h.y, x1 = x0 + 1, y1 = y0 + 1;
          const bool xin0 = x0 >= 0 && x0 < W, xin1 = x1 >= 0 && x1 < W;
          const bool yin0 = y0 >= 0 && y0 < H, yin1 = y1 >= 0 && y1 < H;
          const int o00 = y0 * W + x0;
          float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
          if (xin0 && yin0) v00 = tap_dot<C, LAYOUT>(src_img, ref, o00, HW);
          if (xin1 && yin0) v01 = tap_dot<C, LAYOUT>(src_img, ref, o00 + 1, HW);
          if (xin0 && yin1) v10 = tap_dot<C, LAYOUT>(src_img, ref, o00 + W, HW);
          if (xin1 && yin1) v11 = tap_dot<C, LAYOUT>(src_img, ref, o00 + W + 1, HW);
          // bilinear polynomial  v(fx,fy) = c0 + fx*cx + fy*(cy + fx*cxy)
          rec[(i * 3 + 0) * NT + tid] = make_float4(v00, v01 - v00, v10 - v00, (v00 - v01) - (v10 - v11));
          if (CW) {
            float m00 = 0.f, m01 = 0.f, m10 = 0.f, m11 = 0.f, s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
            if (xin0 && yin0) { m00 = ldg_f(gm + o00); s00 = ldg_f(gm + HW + o00); }
            if (xin1 && yin0) { m01 = ldg_f(gm + o00 + 1); s01 = ldg_f(gm + HW + o00 + 1); }
            if (xin0 && yin1) { m10 = ldg_f(gm + o00 + W); s10 = ldg_f(gm + HW + o00 + W); }
            if (xin1 && yin1) { m11 = ldg_f(gm + o00 + W + 1); s11 = ldg_f(gm + HW + o00 + W + 1); }
            rec[(i * 3 + 1) * NT + tid] = make_float4(m00, m01 - m00, m10 - m00, (m00 - m01) - (m10 - m11));
            rec[(i * 3 + 2) * NT + tid] = make_float4(s00, s01 - s00, s10 - s00, (s00 - s01) - (s10 - s11));
          }
        }
      }

      // ---------------- phase C: evaluate hypotheses [j_lo, j_end) ---------------------------
      {
        int i = -1;
        float cx = -1e30f, cy = -1e30f;
        float4 rd = make_float4(0.f, 0.f, 0.f, 0.f), rm = rd, rs = rd;
#pragma unroll 2
        for (int j = j_lo; j < j_end; ++j) {
          float ix, iy, z;
          project<MODE>(p, ds, j, a0, a1, a2, q0, q1, q2, xmax, ymax, ix, iy, z);
          float fx = ix - cx, fy = iy - cy;
          if (!(fx >= 0.0f && fx < 1.0f && fy >= 0.0f && fy < 1.0f)) {
            ++i;
            const float2 h = hdr[i * NT + tid];
            cx = h.x;
            cy = h.y;
            rd = rec[(i * 3 + 0) * NT + tid];
            if (CW) {
              rm = rec[(i * 3 + 1) * NT + tid];
              rs = rec[(i * 3 + 2) * NT + tid];
            }
            fx = ix - cx;
            fy = iy - cy;
          }
          const float cost = __fmaf_rn(fy, __fmaf_rn(fx, rd.w, rd.z), __fmaf_rn(fx, rd.y, rd.x));
          float val = cost;
          if (CW) {
            const float mu = __fmaf_rn(fy, __fmaf_rn(fx, rm.w, rm.z), __fmaf_rn(fx, rm.y, rm.x));
            const float sg = __fmaf_rn(fy, __fmaf_rn(fx, rs.w, rs.z), __fmaf_rn(fx, rs.y, rs.x));
            // homography.py:157-158: |z - mu~| < sigma~ * kappa, strict
            val = (fabsf(__fsub_rn(z, mu)) < __fmul_rn(sg, p.kappa)) ? cost : 0.0f;
          }
          acc[j * NT + tid] += val;
        }
      }
      j_lo = j_end;
    }
  }

  // ---------------- epilogue: 1/V mean over ALL views (homography.py:120), optional softmax -----
  float* outp = p.out + (size_t)b * D * HW + n;
  if (!SOFTMAX) {
    if (live) {
      if (p.inv_v_exact != 0.0f) {
        for (int j = 0; j < D; ++j) outp[(size_t)j * HW] = acc[j * NT + tid] * p.inv_v_exact;
      } else {
        for (int j = 0; j < D; ++j) outp[(size_t)j * HW] = __fdiv_rn(acc[j * NT + tid], p.vf);
      }
    }
  } else {
    float m = -INFINITY;
    for (int j = 0; j < D; ++j) {
      const float x = __fdiv_rn(acc[j * NT + tid], p.vf);
      acc[j * NT + tid] = x;
      m = fmaxf(m, x);
    }
    float s = 0.0f;
    for (int j = 0; j < D; ++j) {
      const float e = expf(acc[j * NT + tid] - m);
      acc[j * NT + tid] = e;
      s += e;
    }
    if (live)
      for (int j = 0; j < D; ++j) outp[(size_t)j * HW] = __fdiv_rn(acc[j * NT + tid], s);
  }
}

template <int C, int MODE, int LAYOUT>
static cudaError_t launch_cml(const CostParams& p, bool cw, bool softmax, cudaStream_t st) {
  const size_t smem = cells_smem_bytes(p.D);
  dim3 grid((p.HW + NT - 1) / NT, p.B), block(NT);
#define MAGNET_LAUNCH(CWv, SMv)                                                                         \
  do {                                                                                                  \
    auto kern = cost_cells_kernel<C, MODE, LAYOUT, CWv, SMv>;                                           \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                     \
    kern<<<grid, block, smem, st>>>(p);                                                                 \
    return cudaGetLastError();                                                                          \
  } while (0)
  if (cw && !softmax) MAGNET_LAUNCH(true, false);
  if (!cw && softmax) MAGNET_LAUNCH(false, true);
  if (!cw && !softmax) MAGNET_LAUNCH(false, false);
#undef MAGNET_LAUNCH
  return cudaErrorInvalidValue;   // cw && softmax is not a reference configuration
}

template <int C>
static cudaError_t launch_c(const CostParams& p, int mode, int layout, bool cw, bool softmax, cudaStream_t st) {
  if (layout == MAGNET_SRC_C4HW4) {
    if (mode == MAGNET_DEPTH_VOLUME) return launch_cml<C, MAGNET_DEPTH_VOLUME, MAGNET_SRC_C4HW4>(p, cw, softmax, st);
    if (mode == MAGNET_DEPTH_GAUSS) return launch_cml<C, MAGNET_DEPTH_GAUSS, MAGNET_SRC_C4HW4>(p, cw, softmax, st);
    return launch_cml<C, MAGNET_DEPTH_PLANES, MAGNET_SRC_C4HW4>(p, cw, softmax, st);
  }
  if (mode == MAGNET_DEPTH_VOLUME) return launch_cml<C, MAGNET_DEPTH_VOLUME, MAGNET_SRC_NCHW>(p, cw, softmax, st);
  if (mode == MAGNET_DEPTH_GAUSS) return launch_cml<C, MAGNET_DEPTH_GAUSS, MAGNET_SRC_NCHW>(p, cw, softmax, st);
  return launch_cml<C, MAGNET_DEPTH_PLANES, MAGNET_SRC_NCHW>(p, cw, softmax, st);
}

bool cells_supports(int C, int D) {
  return (C == 16 || C == 32 || C == 64) && cells_smem_bytes(D) <= 220 * 1024;
}

void cells_launch_info(int B, int HW, int D, int* grid, int* block, int* smem) {
  *grid = ((HW + NT - 1) / NT) * B;
  *block = NT;
  *smem = (int)cells_smem_bytes(D);
}

cudaError_t launch_cost_cells(const CostParams& p, int mode, int layout, int C, bool cw, bool softmax,
                              cudaStream_t st, int* launches) {
  *launches = 1;
  switch (C) {
    case 16: return launch_c<16>(p, mode, layout, cw, softmax, st);
    case 32: return launch_c<32>(p, mode, layout, cw, softmax, st);
    case 64: return launch_c<64>(p, mode, layout, cw, softmax, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace magnet
