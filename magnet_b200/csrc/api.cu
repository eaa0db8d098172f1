// C-ABI entry points declared in include/magnet_b200.h: argument validation + dispatch only.
#include <atomic>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace magnet {
cudaError_t launch_cost_direct(const CostParams& p, int depth_mode, int src_layout, int C, bool cw,
                               bool softmax, cudaStream_t st, int* launches);
cudaError_t launch_cost_cells(const CostParams& p, int mode, int C, bool cw, bool reuse, cudaStream_t st);
cudaError_t launch_softmax_planes(float* vol, int B, int D, int HW, cudaStream_t st);
bool cells_supports(int C, int D, int layout);
cudaError_t launch_cost_tma(const CostParams& p, int mode, int C, bool cw, cudaStream_t st);
bool tma_supports(int C, int D, int V, int layout);
void tma_launch_info(int B, int H, int W, int D, int* grid, int* block, int* smem);
cudaError_t launch_cost_mma(const CostParams& p, int mode, bool cw, cudaStream_t st);
bool mma_supports(int C, int D, int V, int layout);
void mma_launch_info(int B, int H, int W, int D, int* grid, int* block, int* smem);
size_t split16_buffer_bytes(int N, int H, int W);
cudaError_t launch_repack_split16(const float* src, const float* gmm, void* dst, int N, int C, int H, int W,
                                  cudaStream_t st, int* launches);
#ifdef MAGNET_MMA_DEBUG
void mma_set_debug_buffer(float* p);
#endif
cudaError_t launch_repack_pixc(const float* src, const float* gmm, float* dst, int N, int C, int H, int W,
                               cudaStream_t st);
void cells_launch_info(int B, int H, int W, int D, int* grid, int* block, int* smem);
cudaError_t launch_pack_cameras(const float* intM, const float* R, int64_t r_sb, int64_t r_sv, int64_t r_si,
                                int64_t r_sj, const float* t, int64_t t_sb, int64_t t_sv, int64_t t_si,
                                const int32_t* is_valid, int B, int V, magnet_camera* out, cudaStream_t st);
cudaError_t launch_repack(const float* src, float* dst, int N, int C, int H, int W, cudaStream_t st);
cudaError_t launch_sample(const float* gmm, const float* k_host, int B, int D, int HW, float* dvol,
                          cudaStream_t st);
cudaError_t launch_update_fwd(const float* dout, const float* gmm0, int B, int HW, float* out, cudaStream_t st);
cudaError_t launch_update_bwd(const float* gout, const float* dout, const float* gmm0, int B, int HW, float* gin,
                              cudaStream_t st);
cudaError_t launch_relative_poses(const float* ext_ref, const float* ext_nghbr, int B, int V, float* poses,
                                  int32_t* valid, cudaStream_t st);
cudaError_t launch_camera_rays(const double* raw, int B, int H, int W, float* intM, float* rays, cudaStream_t st);
cudaError_t launch_upsample_fwd(const float* depth, const float* mask, int B, int CH, int H, int W, int k, float* out,
                                cudaStream_t st);
cudaError_t launch_upsample_bwd(const float* gout, const float* depth, const float* mask, int B, int CH, int H, int W,
                                int k, float* gdepth, float* gmask, cudaStream_t st);
cudaError_t launch_cost_f_bwd(const BwdParams& p, cudaStream_t st, int* launches);
cudaError_t launch_upsample_nll_fwd(const float* depth, const float* mask, const float* gt, const uint8_t* gtm, int B,
                                    int H, int W, int k, float* partial, cudaStream_t st);
cudaError_t launch_upsample_nll_bwd(const float* depth, const float* mask, const float* gt, const uint8_t* gtm, float scale,
                                    int B, int H, int W, int k, float* gdepth, float* gmask, cudaStream_t st);
}  // namespace magnet

namespace {
std::atomic<uint64_t> g_launches{0};
thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return MAGNET_ERR_CUDA;
}

int validate_cost(const magnet_cost_args* a) {
  if (!a) return MAGNET_ERR_NULL;
  if (a->B <= 0 || a->V <= 0 || a->D <= 0 || a->C <= 0 || a->H <= 0 || a->W <= 0) return MAGNET_ERR_SHAPE;
  if (a->D > MAGNET_MAX_PLANES) return MAGNET_ERR_UNSUPPORTED;
  if ((int64_t)a->H * a->W > (1 << 26)) return MAGNET_ERR_SHAPE;
  if (!a->ref_feat || !a->src_feat || !a->rays || !a->cams || !a->out) return MAGNET_ERR_NULL;
  if (a->consistency && !a->src_gmm && a->src_layout != MAGNET_SRC_PIXC && a->src_layout != MAGNET_SRC_SPLIT16)
    return MAGNET_ERR_NULL;
  if (a->consistency && a->softmax) return MAGNET_ERR_UNSUPPORTED;
  switch (a->depth_mode) {
    case MAGNET_DEPTH_VOLUME: if (!a->d_volume) return MAGNET_ERR_NULL; break;
    case MAGNET_DEPTH_GAUSS: if (!a->ref_gmm || !a->k_host) return MAGNET_ERR_NULL; break;
    case MAGNET_DEPTH_PLANES: if (!a->k_host) return MAGNET_ERR_NULL; break;
    default: return MAGNET_ERR_UNSUPPORTED;
  }
  if (a->src_layout == MAGNET_SRC_TILED32) {
    if (a->C % 4 != 0) return MAGNET_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(a->src_feat) % 16 != 0) return MAGNET_ERR_ALIGN;
  } else if (a->src_layout == MAGNET_SRC_PIXC) {
    if (!magnet::tma_supports(a->C, a->D, a->V, a->src_layout)) return MAGNET_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(a->src_feat) % 16 != 0) return MAGNET_ERR_ALIGN;
    if (a->variant != MAGNET_VARIANT_AUTO && a->variant != MAGNET_VARIANT_TMA) return MAGNET_ERR_UNSUPPORTED;
  } else if (a->src_layout == MAGNET_SRC_SPLIT16) {
    if (!magnet::mma_supports(a->C, a->D, a->V, a->src_layout)) return MAGNET_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(a->src_feat) % 16 != 0 || reinterpret_cast<uintptr_t>(a->ref_feat) % 16 != 0)
      return MAGNET_ERR_ALIGN;
    if (a->variant != MAGNET_VARIANT_AUTO && a->variant != MAGNET_VARIANT_MMA) return MAGNET_ERR_UNSUPPORTED;
  } else if (a->src_layout != MAGNET_SRC_NCHW) {
    return MAGNET_ERR_UNSUPPORTED;
  }
  if (a->variant < MAGNET_VARIANT_AUTO || a->variant > MAGNET_VARIANT_MMA) return MAGNET_ERR_UNSUPPORTED;
  if (a->variant == MAGNET_VARIANT_MMA && !magnet::mma_supports(a->C, a->D, a->V, a->src_layout))
    return MAGNET_ERR_UNSUPPORTED;
  if (a->variant == MAGNET_VARIANT_TMA && !magnet::tma_supports(a->C, a->D, a->V, a->src_layout))
    return MAGNET_ERR_UNSUPPORTED;
  if ((a->variant == MAGNET_VARIANT_CELLS || a->variant == MAGNET_VARIANT_CELLS_NOREUSE) &&
      !magnet::cells_supports(a->C, a->D, a->src_layout))
    return MAGNET_ERR_UNSUPPORTED;
  if (a->variant == MAGNET_VARIANT_CELLS_NOREUSE && a->depth_mode != MAGNET_DEPTH_GAUSS) return MAGNET_ERR_UNSUPPORTED;
  return MAGNET_OK;
}

bool use_cells(const magnet_cost_args* a) {
  if (a->variant == MAGNET_VARIANT_DIRECT || a->variant == MAGNET_VARIANT_TMA || a->variant == MAGNET_VARIANT_MMA)
    return false;
  return magnet::cells_supports(a->C, a->D, a->src_layout);
}

bool use_tma(const magnet_cost_args* a) {                  // the PIXC layout is served by the TMA kernel only
  return a->src_layout == MAGNET_SRC_PIXC && magnet::tma_supports(a->C, a->D, a->V, a->src_layout);
}

bool use_mma(const magnet_cost_args* a) {                  // the SPLIT16 layout is served by the tensor-core kernel only
  return a->src_layout == MAGNET_SRC_SPLIT16 && magnet::mma_supports(a->C, a->D, a->V, a->src_layout);
}
}  // namespace

extern "C" {

int magnet_abi_version(void) { return MAGNET_ABI_VERSION; }

const char* magnet_strerror(int status) {
  switch (status) {
    case MAGNET_OK: return "ok";
    case MAGNET_ERR_NULL: return "required pointer is NULL";
    case MAGNET_ERR_SHAPE: return "bad or inconsistent dimension";
    case MAGNET_ERR_UNSUPPORTED: return "unsupported C / D / layout / mode / variant";
    case MAGNET_ERR_CUDA: return "CUDA runtime error (see magnet_last_cuda_error)";
    case MAGNET_ERR_ALIGN: return "pointer not 16-byte aligned";
    default: return "unknown status";
  }
}

const char* magnet_last_cuda_error(void) { return g_cuda_err; }

uint64_t magnet_launch_count(void) { return g_launches.load(); }

int magnet_cost_launch_info(const magnet_cost_args* a, int* grid_ctas, int* block_threads, int* smem_bytes) {
  const int st = validate_cost(a);
  if (st != MAGNET_OK) return st;
  if (!grid_ctas || !block_threads || !smem_bytes) return MAGNET_ERR_NULL;
  if (use_mma(a)) {
    magnet::mma_launch_info(a->B, a->H, a->W, a->D, grid_ctas, block_threads, smem_bytes);
  } else if (use_tma(a)) {
    magnet::tma_launch_info(a->B, a->H, a->W, a->D, grid_ctas, block_threads, smem_bytes);
  } else if (use_cells(a)) {
    magnet::cells_launch_info(a->B, a->H, a->W, a->D, grid_ctas, block_threads, smem_bytes);
  } else {
    *grid_ctas = ((a->H * a->W + 127) / 128) * a->D * a->B;
    *block_threads = 128;
    *smem_bytes = 0;
  }
  return MAGNET_OK;
}

int magnet_cost_volume_f32(const magnet_cost_args* a, void* stream) {
  const int st = validate_cost(a);
  if (st != MAGNET_OK) return st;
  magnet::CostParams p;
  p.B = a->B; p.V = a->V; p.D = a->D; p.H = a->H; p.W = a->W; p.HW = a->H * a->W;
  p.kappa = a->kappa;
  p.vf = (float)a->V;
  p.inv_v_exact = ((a->V & (a->V - 1)) == 0) ? 1.0f / (float)a->V : 0.0f;
  p.ref_feat = a->ref_feat; p.src_feat = a->src_feat; p.src_gmm = a->src_gmm; p.rays = a->rays;
  p.cams = a->cams; p.d_volume = a->d_volume; p.ref_gmm = a->ref_gmm; p.out = a->out;
  for (int j = 0; j < MAGNET_MAX_PLANES; ++j)
    p.k[j] = (a->depth_mode != MAGNET_DEPTH_VOLUME && j < a->D) ? a->k_host[j] : 0.0f;
  p.k_sorted = 1;
  for (int j = 1; j < a->D; ++j)
    if (!(p.k[j] >= p.k[j - 1])) p.k_sorted = 0;
  int launches = 0;
  cudaError_t e;
  if (use_mma(a) || use_tma(a) || use_cells(a)) {
    if (use_mma(a))
      e = magnet::launch_cost_mma(p, a->depth_mode, a->consistency != 0, (cudaStream_t)stream);
    else if (use_tma(a))
      e = magnet::launch_cost_tma(p, a->depth_mode, a->C, a->consistency != 0, (cudaStream_t)stream);
    else
      e = magnet::launch_cost_cells(p, a->depth_mode, a->C, a->consistency != 0,
                                    a->variant != MAGNET_VARIANT_CELLS_NOREUSE, (cudaStream_t)stream);
    launches = 1;
    if (e == cudaSuccess && a->softmax) {          // homography.py:46, in place on the 1/V-averaged scores
      e = magnet::launch_softmax_planes(a->out, a->B, a->D, a->H * a->W, (cudaStream_t)stream);
      launches = 2;
    }
  } else {
    e = magnet::launch_cost_direct(p, a->depth_mode, a->src_layout, a->C, a->consistency != 0, a->softmax != 0,
                                   (cudaStream_t)stream, &launches);
  }
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += launches;
  return MAGNET_OK;
}

int magnet_cost_volume_f_bwd_f32(const magnet_cost_f_bwd_args* b, void* stream) {
  if (!b || !b->fwd) return MAGNET_ERR_NULL;
  const magnet_cost_args* a = b->fwd;
  if (a->B <= 0 || a->V <= 0 || a->D <= 0 || a->C <= 0 || a->H <= 0 || a->W <= 0) return MAGNET_ERR_SHAPE;
  if (a->D > MAGNET_MAX_PLANES) return MAGNET_ERR_UNSUPPORTED;
  if (!a->ref_feat || !a->src_feat || !a->rays || !a->cams || !a->k_host) return MAGNET_ERR_NULL;
  if (!b->grad_out || !b->workspace || !b->grad_ref || !b->grad_src) return MAGNET_ERR_NULL;
  if (a->softmax && !b->prob) return MAGNET_ERR_NULL;
  if (a->consistency || a->depth_mode != MAGNET_DEPTH_PLANES || a->src_layout != MAGNET_SRC_NCHW)
    return MAGNET_ERR_UNSUPPORTED;
  if (a->C != 8 && a->C != 16 && a->C != 32 && a->C != 64) return MAGNET_ERR_UNSUPPORTED;
  magnet::BwdParams p;
  p.B = a->B; p.V = a->V; p.D = a->D; p.C = a->C; p.H = a->H; p.W = a->W; p.HW = a->H * a->W;
  p.softmax = a->softmax != 0;
  p.vf = (float)a->V;
  p.ref_feat = a->ref_feat; p.src_feat = a->src_feat; p.rays = a->rays; p.cams = a->cams;
  p.prob = b->prob; p.grad_out = b->grad_out; p.g_score = b->workspace; p.grad_ref = b->grad_ref; p.grad_src = b->grad_src;
  for (int j = 0; j < MAGNET_MAX_PLANES; ++j) p.k[j] = j < a->D ? a->k_host[j] : 0.0f;
  int launches = 0;
  cudaError_t e = magnet::launch_cost_f_bwd(p, (cudaStream_t)stream, &launches);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += launches;
  return MAGNET_OK;
}

int magnet_pack_cameras_f32(const float* intM, const float* R, int64_t r_sb, int64_t r_sv, int64_t r_si,
                            int64_t r_sj, const float* t, int64_t t_sb, int64_t t_sv, int64_t t_si,
                            const int32_t* is_valid, int32_t B, int32_t V, magnet_camera* cams_out,
                            void* stream) {
  if (!intM || !R || !t || !is_valid || !cams_out) return MAGNET_ERR_NULL;
  if (B <= 0 || V <= 0) return MAGNET_ERR_SHAPE;
  cudaError_t e = magnet::launch_pack_cameras(intM, R, r_sb, r_sv, r_si, r_sj, t, t_sb, t_sv, t_si, is_valid, B, V,
                                              cams_out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_repack_tiled32_f32(const float* src_nchw, float* dst, int32_t N, int32_t C, int32_t H, int32_t W,
                            void* stream) {
  if (!src_nchw || !dst) return MAGNET_ERR_NULL;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return MAGNET_ERR_SHAPE;
  if (C % 4 != 0 || C / 4 > 65535 || N > 65535) return MAGNET_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(dst) % 16 != 0) return MAGNET_ERR_ALIGN;
  cudaError_t e = magnet::launch_repack(src_nchw, dst, N, C, H, W, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_repack_pixc_f32(const float* src_nchw, const float* src_gmm, float* dst, int32_t N, int32_t C, int32_t H,
                           int32_t W, void* stream) {
  if (!src_nchw || !dst) return MAGNET_ERR_NULL;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || N > 65535) return MAGNET_ERR_SHAPE;
  if (C != 16 && C != 32 && C != 64) return MAGNET_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(dst) % 16 != 0) return MAGNET_ERR_ALIGN;
  cudaError_t e = magnet::launch_repack_pixc(src_nchw, src_gmm, dst, N, C, H, W, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

size_t magnet_split16_bytes(int32_t N, int32_t H, int32_t W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return magnet::split16_buffer_bytes(N, H, W);
}

int magnet_repack_split16_f32(const float* src_nchw, const float* src_gmm, void* dst, int32_t N, int32_t C, int32_t H,
                              int32_t W, void* stream) {
  if (!src_nchw || !dst) return MAGNET_ERR_NULL;
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || N > 65535) return MAGNET_ERR_SHAPE;
  if (C != 64) return MAGNET_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(dst) % 16 != 0 || reinterpret_cast<uintptr_t>(src_nchw) % 16 != 0) return MAGNET_ERR_ALIGN;
  int launches = 0;
  cudaError_t e = magnet::launch_repack_split16(src_nchw, src_gmm, dst, N, C, H, W, (cudaStream_t)stream, &launches);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += launches;
  return MAGNET_OK;
}

#ifdef MAGNET_MMA_DEBUG
void magnet_mma_debug_buffer(float* p) { magnet::mma_set_debug_buffer(p); }
#endif

int magnet_sample_depths_f32(const float* gmm, const float* k_host, int32_t B, int32_t D, int32_t HW,
                             float* d_volume, void* stream) {
  if (!gmm || !k_host || !d_volume) return MAGNET_ERR_NULL;
  if (B <= 0 || D <= 0 || HW <= 0) return MAGNET_ERR_SHAPE;
  if (D > MAGNET_MAX_PLANES || B > 65535) return MAGNET_ERR_UNSUPPORTED;
  cudaError_t e = magnet::launch_sample(gmm, k_host, B, D, HW, d_volume, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_gaussian_update_fwd_f32(const float* d_output, const float* ref_gmm, int32_t B, int32_t HW,
                                   float* out, void* stream) {
  if (!d_output || !ref_gmm || !out) return MAGNET_ERR_NULL;
  if (B <= 0 || HW <= 0 || B > 65535) return MAGNET_ERR_SHAPE;
  cudaError_t e = magnet::launch_update_fwd(d_output, ref_gmm, B, HW, out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_gaussian_update_bwd_f32(const float* grad_out, const float* d_output, const float* ref_gmm,
                                   int32_t B, int32_t HW, float* grad_d_output, void* stream) {
  if (!grad_out || !d_output || !ref_gmm || !grad_d_output) return MAGNET_ERR_NULL;
  if (B <= 0 || HW <= 0 || B > 65535) return MAGNET_ERR_SHAPE;
  cudaError_t e = magnet::launch_update_bwd(grad_out, d_output, ref_gmm, B, HW, grad_d_output, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_relative_poses_f32(const float* ext_ref, const float* ext_nghbr, int32_t B, int32_t V, float* poses_out,
                              int32_t* is_valid_out, void* stream) {
  if (!ext_ref || !ext_nghbr || !poses_out || !is_valid_out) return MAGNET_ERR_NULL;
  if (B <= 0 || V <= 0) return MAGNET_ERR_SHAPE;
  cudaError_t e = magnet::launch_relative_poses(ext_ref, ext_nghbr, B, V, poses_out, is_valid_out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_camera_rays_f32(const double* raw_intrinsics, int32_t B, int32_t H, int32_t W, float* intM_out,
                           float* rays_out, void* stream) {
  if (!raw_intrinsics || !intM_out || !rays_out) return MAGNET_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535) return MAGNET_ERR_SHAPE;
  cudaError_t e = magnet::launch_camera_rays(raw_intrinsics, B, H, W, intM_out, rays_out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_convex_upsample_fwd_f32(const float* depth, const float* up_mask, int32_t B, int32_t CH, int32_t H,
                                   int32_t W, int32_t k, float* out, void* stream) {
  if (!depth || !up_mask || !out) return MAGNET_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || k <= 0 || B > 65535 || H * k > 65535) return MAGNET_ERR_SHAPE;
  if (CH != 1 && CH != 2) return MAGNET_ERR_UNSUPPORTED;
  cudaError_t e = magnet::launch_upsample_fwd(depth, up_mask, B, CH, H, W, k, out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_convex_upsample_bwd_f32(const float* grad_out, const float* depth, const float* up_mask, int32_t B,
                                   int32_t CH, int32_t H, int32_t W, int32_t k, float* grad_depth, float* grad_mask,
                                   void* stream) {
  if (!grad_out || !depth || !up_mask || !grad_depth || !grad_mask) return MAGNET_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || k <= 0 || B > 65535 || H * k > 65535) return MAGNET_ERR_SHAPE;
  if (CH != 1 && CH != 2) return MAGNET_ERR_UNSUPPORTED;
  cudaError_t e = magnet::launch_upsample_bwd(grad_out, depth, up_mask, B, CH, H, W, k, grad_depth, grad_mask,
                                              (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_upsample_nll_partials(int32_t B, int32_t H, int32_t W, int32_t k) {
  if (B <= 0 || H <= 0 || W <= 0 || k <= 0) return MAGNET_ERR_SHAPE;
  return B * H * k * ((W * k + 127) / 128);
}

int magnet_upsample_nll_fwd_f32(const float* depth, const float* up_mask, const float* gt, const uint8_t* gt_mask,
                                int32_t B, int32_t H, int32_t W, int32_t k, float* partial, void* stream) {
  if (!depth || !up_mask || !gt || !gt_mask || !partial) return MAGNET_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || k <= 0 || B > 65535 || H * k > 65535) return MAGNET_ERR_SHAPE;
  cudaError_t e = magnet::launch_upsample_nll_fwd(depth, up_mask, gt, gt_mask, B, H, W, k, partial, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

int magnet_upsample_nll_bwd_f32(const float* depth, const float* up_mask, const float* gt, const uint8_t* gt_mask,
                                float scale, int32_t B, int32_t H, int32_t W, int32_t k, float* grad_depth,
                                float* grad_mask, void* stream) {
  if (!depth || !up_mask || !gt || !gt_mask || !grad_depth || !grad_mask) return MAGNET_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || k <= 0 || B > 65535 || H * k > 65535) return MAGNET_ERR_SHAPE;
  cudaError_t e = magnet::launch_upsample_nll_bwd(depth, up_mask, gt, gt_mask, scale, B, H, W, k, grad_depth, grad_mask,
                                                  (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e);
  g_launches += 1;
  return MAGNET_OK;
}

}  // extern "C"
