// MAGNET_VARIANT_WINDOW — tap-sharing kernel with the CTA's source window staged in shared memory.
//
// Same algorithm as cost_cells.cu (per-lane bilinear-cell lists, register tap reuse, polynomial records,
// lockstep phases A / B / C — see there), different memory plan.  ncu on cost_cells showed 70 % of the warp
// stalls on the tap gathers: every tap is 16 LDG.128 that miss L1 (hit rate 1-20 %: 16 warps x ~20 KB per
// cell iteration do not fit beside the shared memory), so 3.5 GB per launch cross L2 -> SM for 0.16 GB of
// distinct source data.  Here a CTA of 256 threads owns a 16x16 pixel tile and ONE chunk of 32 hypotheses;
// per source view it
//   1. builds the cell lists (phase A) and reduces their bounding box over the CTA,
//   2. copies that window of the TILED32 source image (all C channels, + the source Gaussian) into shared
//      memory with cp.async (16-byte LDGSTS, coalesced 512-byte global segments) — each source byte now
//      crosses L2 -> SM once per (CTA, view) instead of ~7 times,
//   3. runs phase B against the window: a 64-channel dot product is 16 LDS.128 at immediate offsets
//      (pixel stride 17 float4 = 272 B makes the quarter-warp accesses bank-conflict free) + 32 FFMA2,
//   4. runs phase C with the 32 accumulators in registers (fully unrolled, no shared-memory column).
// Two CTAs of 128 threads (16x8 tiles) per SM, up to 255 registers per thread; the window gets the shared
// memory that the records leave (~300 pixels per CTA).  A window that does not fit falls back to global
// gathers for that round.  Measured (cfg2): 0.54 ms vs 0.35 ms for the global-gather kernel — it moves 2.5x
// fewer bytes L2 -> SM but with 8 warps per SM it cannot hide the shared-memory / barrier latencies; it is
// kept as MAGNET_VARIANT_WINDOW for the next round.
#include <cstdlib>

#include "cells_common.cuh"

namespace magnet {

#ifndef MAGNET_WNT
#define MAGNET_WNT 128
#endif
#ifndef MAGNET_WCTAS
#define MAGNET_WCTAS 2
#endif
#ifndef MAGNET_WNCELL
#define MAGNET_WNCELL 5
#endif
constexpr int WNT = MAGNET_WNT;       // threads per CTA
constexpr int WTW = 16, WTH = WNT / 16;  // CTA tile (pixels)
constexpr int WCTAS = MAGNET_WCTAS;   // resident CTAs per SM the shared memory is budgeted for
constexpr int WNCELL = MAGNET_WNCELL; // cell records per lane per round
constexpr int WCHUNK = 32;            // hypotheses per CTA, accumulated in registers
constexpr int WSMEM_MAX = (227 * 1024) / WCTAS - (WCTAS > 1 ? 1024 : 0);  // dynamic smem per CTA (sm_100: 227 KB/SM)

__host__ __device__ constexpr size_t window_fixed_bytes() {
  return (size_t)WNCELL * 3 * WNT * 16 + (size_t)WNCELL * WNT * 8 + WCHUNK * 4 + 2 * 8 * 4;
}
template <int C>
__host__ __device__ constexpr int window_cap_px() {
  return (int)((WSMEM_MAX - window_fixed_bytes()) / ((C / 4 + 1) * 16));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// <ref, window[tap]>: C/4 LDS.128 at immediate offsets + packed FMAs; slot C/4 of the pixel holds (mu, sigma).
template <int C, bool CW>
__device__ __forceinline__ Tap load_tap_win(const float4* __restrict__ win, const float2 (&ref2)[C / 2], int x, int y,
                                            int W, int H, int wx0, int wy0, int ww) {
  Tap t;
  t.f = t.m = t.s = 0.0f;
  if (x >= 0 && x < W && y >= 0 && y < H) {
    const float4* s = win + ((y - wy0) * ww + (x - wx0)) * (C / 4 + 1);
    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 q = s[c4];
      s0 = __ffma2_rn(ref2[2 * c4 + 0], make_float2(q.x, q.y), s0);
      s1 = __ffma2_rn(ref2[2 * c4 + 1], make_float2(q.z, q.w), s1);
    }
    t.f = (s0.x + s0.y) + (s1.x + s1.y);
    if (CW) {
      const float4 ms = s[C / 4];
      t.m = ms.x;
      t.s = ms.y;
    }
  }
  return t;
}

template <int C, int MODE, bool CW>
__global__ void __launch_bounds__(WNT, WCTAS)
cost_window_kernel(const __grid_constant__ CostParams p) {
  constexpr int QN = C / 4, PF4 = QN + 1;
  extern __shared__ float4 smem4[];
  float4* rec = smem4;                                                   // [WNCELL][3][WNT]
  float2* hdr = reinterpret_cast<float2*>(smem4 + WNCELL * 3 * WNT);     // [WNCELL][WNT]
  float* ks = reinterpret_cast<float*>(hdr + WNCELL * WNT);              // [WCHUNK]
  int* scr = reinterpret_cast<int*>(ks + WCHUNK);                        // [2][8] CTA reductions, double buffered
  float4* win = reinterpret_cast<float4*>(scr + 16);                     // [cap][PF4]
  constexpr int CAP = window_cap_px<C>();

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int H = p.H, W = p.W, HW = p.HW, D = p.D;
  const int XB = (W + 31) >> 5;
  const int nchunks = (D + WCHUNK - 1) / WCHUNK;
  const int tiles_x = (W + WTW - 1) / WTW;
  const int tile = blockIdx.x / nchunks;                                 // chunks of a tile are adjacent CTAs (L2)
  const int jc = (blockIdx.x % nchunks) * WCHUNK;
  const int jc_end = min(jc + WCHUNK, D);
  const int px = (tile % tiles_x) * WTW + tid % WTW;
  const int py = (tile / tiles_x) * WTH + tid / WTW;
  const bool live = px < W && py < H;
  const int n = live ? py * W + px : HW - 1;    // dead lanes shadow the last pixel, never store
  const unsigned FULL = 0xffffffffu;

  float2 ref2[C / 2];
  {
    const float* rp = p.ref_feat + (size_t)b * C * HW + n;
#pragma unroll
    for (int c = 0; c < C / 2; ++c) ref2[c] = make_float2(ldg_f(rp + (size_t)(2 * c) * HW), ldg_f(rp + (size_t)(2 * c + 1) * HW));
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
  const size_t img_stride4 = (size_t)H * XB * QN * 32;                   // float4 per source image

  if (tid < WCHUNK) ks[tid] = (MODE != MAGNET_DEPTH_VOLUME && jc + tid < jc_end) ? p.k[jc + tid] : 0.0f;
  if (tid < 16) scr[tid] = (tid & 7) == 0 || (tid & 7) == 2 || (tid & 7) == 4 ? (1 << 30) : -(1 << 30);
  __syncthreads();
  // scr[par][0..4] = min x_lo, max x_hi, min y_lo, max y_hi, min j_stop   (slots 0,2,4 are minima)

  float accr[WCHUNK];
#pragma unroll
  for (int j = 0; j < WCHUNK; ++j) accr[j] = 0.0f;
  int par = 0;

  for (int v = 0; v < p.V; ++v) {
    const magnet_camera* cam = p.cams + (b * p.V + v);
    if (cam->valid != 1.0f) continue;                                    // CTA-uniform
    const float a0 = cam->a[0], a1 = cam->a[1], a2 = cam->a[2];
    const float q0 = __fmaf_rn(cam->A[2], r2, __fmaf_rn(cam->A[1], r1, __fmul_rn(cam->A[0], r0)));
    const float q1 = __fmaf_rn(cam->A[5], r2, __fmaf_rn(cam->A[4], r1, __fmul_rn(cam->A[3], r0)));
    const float q2 = __fmaf_rn(cam->A[8], r2, __fmaf_rn(cam->A[7], r1, __fmul_rn(cam->A[6], r0)));
    const int vb = v * p.B + b;
    const float4* src_img = reinterpret_cast<const float4*>(p.src_feat) + (size_t)vb * img_stride4;
    const float* gm = CW ? p.src_gmm + (size_t)vb * 2 * HW : nullptr;

    const bool walk = __all_sync(FULL, walk_ok<MODE>(p, ds, jc, jc_end, a2, q2));   // warp-uniform choice
    // direction of travel of the sample as depth grows: sign of d(ix)/dd = (q0 a2 - a0 q2) / z^2
    const float gx = __fmaf_rn(q0, a2, -__fmul_rn(a0, q2)), gy = __fmaf_rn(q1, a2, -__fmul_rn(a1, q2));
    const int sx = gx > 0.0f ? 1 : (gx < 0.0f ? -1 : 0), sy = gy > 0.0f ? 1 : (gy < 0.0f ? -1 : 0);

    int px0 = -1000000, py0 = -1000000;          // previous cell of this lane, taps kept for reuse
    Tap p00, p01, p10, p11;
    p00.f = p00.m = p00.s = 0.f;
    p01 = p10 = p11 = p00;

    int j_lo = jc;
    while (j_lo < jc_end) {                                              // rounds; CTA-uniform
      // ---------------- phase A: cell list + CTA bounding box ------------------------------------
      int ncell, j_stop;
      CellBox box;
      unsigned startmask;
      cell_list<MODE, WNCELL, WNT>(p, ds, ks, hdr + tid, walk, jc, j_lo, jc_end, a0, a1, a2, q0, q1, q2, sx, sy, W, H,
                                   ncell, j_stop, box, startmask);
      {
        const int xl = __reduce_min_sync(FULL, box.x_lo), xh = __reduce_max_sync(FULL, box.x_hi);
        const int yl = __reduce_min_sync(FULL, box.y_lo), yh = __reduce_max_sync(FULL, box.y_hi);
        const int js = __reduce_min_sync(FULL, j_stop);
        if (lane == 0) {
          int* s = scr + par * 8;
          atomicMin(s + 0, xl); atomicMax(s + 1, xh); atomicMin(s + 2, yl); atomicMax(s + 3, yh); atomicMin(s + 4, js);
        }
      }
      __syncthreads();                                                   // B1
      const int* sr = scr + par * 8;
      const int wx0 = max(sr[0], 0), wx1 = min(sr[1] + 1, W - 1);
      const int wy0 = max(sr[2], 0), wy1 = min(sr[3] + 1, H - 1);
      const int j_end = sr[4];
      const int ww = wx1 - wx0 + 1, wh = wy1 - wy0 + 1;
      const bool has_win = ww > 0 && wh > 0;
      const bool staged = has_win && ww * wh <= CAP;
      if (tid < 8) scr[(par ^ 1) * 8 + tid] = (tid == 0 || tid == 2 || tid == 4) ? (1 << 30) : -(1 << 30);
      par ^= 1;

      // ---------------- window: TILED32 global -> shared, pixel-major [wh][ww][QN + 1] float4 -----
      if (staged) {
        for (int r = warp; r < wh; r += WNT / 32) {
          const int y = wy0 + r;
          const float4* grow = src_img + (size_t)y * XB * QN * 32;
          float4* srow = win + (size_t)r * ww * PF4;
          for (int idx = lane; idx < ww * QN; idx += 32) {
            const int q = idx & (QN - 1), pxi = idx / QN, x = wx0 + pxi;
            cp_async16(srow + pxi * PF4 + q, grow + ((x >> 5) * QN + q) * 32 + (x & 31));
          }
          if (CW) {
            for (int pxi = lane; pxi < ww; pxi += 32) {
              const int o = y * W + wx0 + pxi;
              srow[pxi * PF4 + QN] = make_float4(ldg_f(gm + o), ldg_f(gm + HW + o), 0.f, 0.f);
            }
          }
        }
        cp_async_wait_all();
      }
      __syncthreads();                                                   // B2: window visible to all warps

      // ---------------- phase B: per-cell records --------------------------------------------------
      const int nmax = __reduce_max_sync(FULL, ncell);
      for (int i = 0; i < nmax; ++i) {
        if (i < ncell) {
          const float2 h = hdr[i * WNT + tid];
          const int x0 = (int)h.x, y0 = (int)h.y;
          const int dx = x0 - px0, dy = y0 - py0;
          const bool mvx = dy == 0 && (dx == 1 || dx == -1);
          const bool mvy = dx == 0 && (dy == 1 || dy == -1);
          // two taps every lane computes: the new column (x move), the new row (y move), or the top row
          int ax = x0, ay = y0, bx = x0 + 1, by = y0;
          if (mvx) { ax = bx = (dx == 1) ? x0 + 1 : x0; by = y0 + 1; }
          if (mvy) { ay = by = (dy == 1) ? y0 + 1 : y0; }
          Tap tA, tB, tC, tD;
          tC.f = tC.m = tC.s = 0.f;
          tD = tC;
          const bool all4 = !(mvx || mvy);                               // first cell / diagonal / jump
          if (staged) {
            tA = load_tap_win<C, CW>(win, ref2, ax, ay, W, H, wx0, wy0, ww);
            tB = load_tap_win<C, CW>(win, ref2, bx, by, W, H, wx0, wy0, ww);
            if (all4) {
              tC = load_tap_win<C, CW>(win, ref2, x0, y0 + 1, W, H, wx0, wy0, ww);
              tD = load_tap_win<C, CW>(win, ref2, x0 + 1, y0 + 1, W, H, wx0, wy0, ww);
            }
          } else {
            tA = load_tap<C, CW>(src_img, gm, ref2, ax, ay, W, H, XB, HW);
            tB = load_tap<C, CW>(src_img, gm, ref2, bx, by, W, H, XB, HW);
            if (all4) {
              tC = load_tap<C, CW>(src_img, gm, ref2, x0, y0 + 1, W, H, XB, HW);
              tD = load_tap<C, CW>(src_img, gm, ref2, x0 + 1, y0 + 1, W, H, XB, HW);
            }
          }
          Tap n00, n01, n10, n11;
          if (mvx) {
            if (dx == 1) { n00 = p01; n10 = p11; n01 = tA; n11 = tB; }
            else         { n01 = p00; n11 = p10; n00 = tA; n10 = tB; }
          } else if (mvy) {
            if (dy == 1) { n00 = p10; n01 = p11; n10 = tA; n11 = tB; }
            else         { n10 = p00; n11 = p01; n00 = tA; n01 = tB; }
          } else {
            n00 = tA; n01 = tB; n10 = tC; n11 = tD;
          }
          p00 = n00; p01 = n01; p10 = n10; p11 = n11;
          px0 = x0; py0 = y0;
          rec[(i * 3 + 0) * WNT + tid] = bilinear_poly(n00.f, n01.f, n10.f, n11.f);
          if (CW) {
            rec[(i * 3 + 1) * WNT + tid] = bilinear_poly(n00.m, n01.m, n10.m, n11.m);
            rec[(i * 3 + 2) * WNT + tid] = bilinear_poly(n00.s, n01.s, n10.s, n11.s);
          }
        }
      }

      // ---------------- phase C: evaluate hypotheses [j_lo, j_end), accumulators in registers -------
      // Branch-free and chain-free: the index of the lane's cell at hypothesis j is the number of cell
      // starts up to j (a popcount of the start mask from phase A), so every unrolled step is independent of
      // the others; the record is simply (re)loaded each step — a lane enters a new cell on ~10 % of its
      // steps, but some lane of the warp does on ~97 % of them, so a branch would be taken almost always.
      auto step = [&](const int jj) {
        float ix, iy, z;
        project(depth_of<MODE>(p, ds, jc + jj), a0, a1, a2, q0, q1, q2, ix, iy, z);
        const int ci = __popc(startmask & ((2u << jj) - 1u)) - 1;        // >= 0: bit (j_lo - jc) is always set
        const int ro = ci * WNT + tid;
        const float2 hc = hdr[ro];
        const float4 rd = rec[3 * ro - 2 * tid];                         // rec[(3*ci + 0) * WNT + tid]
        const float fx = ix - hc.x, fy = iy - hc.y;
        float cost = __fmaf_rn(fy, __fmaf_rn(fx, rd.w, rd.z), __fmaf_rn(fx, rd.y, rd.x));
        if (!(fabsf(cost) < 3.0e38f)) cost = 0.0f;                       // all-zero record x non-finite position
        float val = cost;
        if (CW) {
          const float4 rm = rec[3 * ro - 2 * tid + WNT], rs = rec[3 * ro - 2 * tid + 2 * WNT];
          const float mu = __fmaf_rn(fy, __fmaf_rn(fx, rm.w, rm.z), __fmaf_rn(fx, rm.y, rm.x));
          const float sg = __fmaf_rn(fy, __fmaf_rn(fx, rs.w, rs.z), __fmaf_rn(fx, rs.y, rs.x));
          // homography.py:157-158: |z - mu~| < sigma~ * kappa, strict
          val = (fabsf(__fsub_rn(z, mu)) < __fmul_rn(sg, p.kappa)) ? cost : 0.0f;
        }
        accr[jj] += val;
      };
      if (j_lo == jc && j_end == jc + WCHUNK) {                          // the usual case: one round, full chunk
#pragma unroll
        for (int jj = 0; jj < WCHUNK; ++jj) step(jj);
      } else {
#pragma unroll
        for (int jj = 0; jj < WCHUNK; ++jj)
          if (jc + jj >= j_lo && jc + jj < j_end) step(jj);              // CTA-uniform guard
      }
      j_lo = j_end;
    }
  }

  // -------- epilogue: 1/V mean over ALL views (homography.py:120) ---------------------------------
  if (live) {
    float* outp = p.out + ((size_t)b * D + jc) * HW + n;
#pragma unroll
    for (int jj = 0; jj < WCHUNK; ++jj) {
      if (jc + jj < jc_end)
        outp[(size_t)jj * HW] = p.inv_v_exact != 0.0f ? accr[jj] * p.inv_v_exact : __fdiv_rn(accr[jj], p.vf);
    }
  }
}

template <int C, int MODE>
static cudaError_t launch_wm(const CostParams& p, bool cw, cudaStream_t st) {
  const int nchunks = (p.D + WCHUNK - 1) / WCHUNK;
  const int tiles = ((p.W + WTW - 1) / WTW) * ((p.H + WTH - 1) / WTH);
  dim3 grid(tiles * nchunks, p.B), block(WNT);
  const size_t smem = window_fixed_bytes() + (size_t)window_cap_px<C>() * (C / 4 + 1) * 16;
#define MAGNET_LAUNCH(CWv)                                                                              \
  do {                                                                                                  \
    auto kern = cost_window_kernel<C, MODE, CWv>;                                                       \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                     \
    kern<<<grid, block, smem, st>>>(p);                                                                 \
    return cudaGetLastError();                                                                          \
  } while (0)
  if (cw) MAGNET_LAUNCH(true);
  MAGNET_LAUNCH(false);
#undef MAGNET_LAUNCH
}

template <int C>
static cudaError_t launch_w(const CostParams& p, int mode, bool cw, cudaStream_t st) {
  if (mode == MAGNET_DEPTH_VOLUME) return launch_wm<C, MAGNET_DEPTH_VOLUME>(p, cw, st);
  if (mode == MAGNET_DEPTH_GAUSS) return launch_wm<C, MAGNET_DEPTH_GAUSS>(p, cw, st);
  return launch_wm<C, MAGNET_DEPTH_PLANES>(p, cw, st);
}

bool window_supports(int C, int D, int layout) {
  return (C == 16 || C == 32 || C == 64) && layout == MAGNET_SRC_TILED32 && D >= 1;
}

void window_launch_info(int B, int H, int W, int D, int C, int* grid, int* block, int* smem) {
  const int nchunks = (D + WCHUNK - 1) / WCHUNK;
  *grid = ((W + WTW - 1) / WTW) * ((H + WTH - 1) / WTH) * nchunks * B;
  *block = WNT;
  const int pf4 = C / 4 + 1;
  *smem = (int)(window_fixed_bytes() + (size_t)((WSMEM_MAX - window_fixed_bytes()) / (pf4 * 16)) * pf4 * 16);
}

cudaError_t launch_cost_window(const CostParams& p, int mode, int C, bool cw, cudaStream_t st) {
  switch (C) {
    case 16: return launch_w<16>(p, mode, cw, st);
    case 32: return launch_w<32>(p, mode, cw, st);
    case 64: return launch_w<64>(p, mode, cw, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace magnet
