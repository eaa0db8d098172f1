/*
 * magnet_b200.h — C ABI of the B200-native multi-view matching hot path of MaGNet.
 *
 * The reference (baegwangbin/MaGNet) has no FFI layer: its boundary is two Python
 * module-level functions and two inlined blocks (SURVEY §8 b).  Every entry point
 * below names the reference interface it replaces.  A maintainer binds these with
 * ctypes / cffi / pybind; magnet_b200/_lib.py is such a binding.
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a DEVICE pointer unless the field says
 *     "host"; all tensors are fp32, dense, in the layout stated per field;
 *   - caller owns all memory; no entry point allocates, frees or synchronises; every
 *     launch goes to the cudaStream_t passed as `stream` (NULL = legacy default
 *     stream), so calls are CUDA-graph capturable;
 *   - re-entrant and stateless (one-time cudaFuncSetAttribute calls are idempotent);
 *   - return value: MAGNET_OK (0) or a negative magnet_status; never throws.
 */
#ifndef MAGNET_B200_H_
#define MAGNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAGNET_ABI_VERSION 3
#define MAGNET_MAX_PLANES 256   /* max depth hypotheses / planes per call (D) */

typedef enum magnet_status {
  MAGNET_OK = 0,
  MAGNET_ERR_NULL = -1,        /* a required pointer is NULL */
  MAGNET_ERR_SHAPE = -2,       /* a dimension is <= 0 or inconsistent */
  MAGNET_ERR_UNSUPPORTED = -3, /* C, D, layout or variant not supported by this build */
  MAGNET_ERR_CUDA = -4,        /* a CUDA runtime call failed (see magnet_last_cuda_error) */
  MAGNET_ERR_ALIGN = -5        /* a pointer violates the required 16-byte alignment */
} magnet_status;

/* Where the per-pixel depth hypotheses come from. */
typedef enum magnet_depth_mode {
  MAGNET_DEPTH_VOLUME = 0, /* read d_volume (B,D,H,W): drop-in for est_costvolume_CW         */
  MAGNET_DEPTH_GAUSS = 1,  /* d_j = mu + sigma*k_j from ref_gmm + k: sampler fused            */
  MAGNET_DEPTH_PLANES = 2  /* d_j = k_j for every pixel: fronto-parallel planes (est_costvolume_F) */
} magnet_depth_mode;

/* Memory layout of the source-view feature maps. */
typedef enum magnet_src_layout {
  MAGNET_SRC_NCHW = 0,   /* (V*B, C, H, W), the reference layout                               */
  MAGNET_SRC_TILED32 = 1, /* (V*B, H, ceil(W/32), C/4, 32, 4): per row, tiles of 32 pixels; inside a tile
                            the C/4 channel quads are 512 B apart and the 32 pixels of one quad are
                            contiguous (see magnet_repack_tiled32_f32).  Pixels x >= W are padding. */
  MAGNET_SRC_PIXC = 2,    /* (V*B, H, W, C+4): pixel-major, per pixel the C channels followed by the source
                            Gaussian (mu, sigma) and two zeros (see magnet_repack_pixc_f32): the layout the
                            TMA-staged CUDA-core kernel fetches its windows from.  With this layout
                            magnet_cost_args.src_gmm is ignored (the Gaussians travel inside src_feat). */
  MAGNET_SRC_SPLIT16 = 3  /* tensor-core layout (C == 64): a 256-byte header (power-of-two scale s), two fp16 planes
                            (V*B, 2, H, W, 64) with x*s = hi + lo, and a (V*B, H, W+1, 4) table whose entry x+1 holds (mu, sigma) of
                            pixel x and of pixel x+1 (zeros outside the row); see
                            magnet_repack_split16_f32 / magnet_split16_bytes.  With this layout ref_feat must ALSO
                            point to a split buffer (of the B reference feature maps, Gaussians NULL) and src_gmm is
                            ignored. */
} magnet_src_layout;

/* Kernel selection (for parity cross-checks and profiling). */
typedef enum magnet_variant {
  MAGNET_VARIANT_AUTO = 0,   /* production choice: MMA for MAGNET_SRC_SPLIT16, TMA for MAGNET_SRC_PIXC, CELLS for
                                MAGNET_SRC_TILED32, DIRECT otherwise                           */
  MAGNET_VARIANT_DIRECT = 1, /* one thread per output, 4 taps x C channels per hypothesis,
                                reference operation order, fp64 view accumulation             */
  MAGNET_VARIANT_CELLS = 2,  /* tap-sharing kernel: per-lane bilinear-cell records             */
  MAGNET_VARIANT_CELLS_NOREUSE = 3, /* diagnostic: as CELLS, but every cell gathers all 4 taps
                                       (MAGNET_DEPTH_GAUSS only)                               */
  MAGNET_VARIANT_TMA = 4,    /* CUDA-core tap-sharing kernel, 4 lanes per pixel, the CTA's source window
                                staged in shared memory by TMA (MAGNET_SRC_PIXC only)          */
  MAGNET_VARIANT_MMA = 5     /* tensor-core kernel: all (reference pixel, window cell) channel dot products of an
                                8x8 tile by tcgen05.mma into tensor memory (MAGNET_SRC_SPLIT16 only) */
} magnet_variant;

/* Per (batch element, view) camera constants, 16 floats, produced by magnet_pack_cameras_f32.
 * index = b*V + v.  Replaces the per-pair matmuls at homography.py:98-102. */
typedef struct magnet_camera {
  float valid;  /* 1.0f when is_valid[b,v] == 1, else 0.0f (homography.py:97)                  */
  float a[3];   /* K_b * t_bv          ('term1_pix', homography.py:101)                        */
  float A[9];   /* K_b * R_bv row-major; A*ray = 'term2_pix' (homography.py:102).  Row 2 equals
                   R_bv[2,:] and a[2] equals t_bv[2], i.e. the z rows of 'term1_cam/term2_cam' */
  float pad[3];
} magnet_camera;

/*
 * Arguments of the fused warp + bilinear sample + consistency weight + view fusion kernel.
 * Replaces models/submodules/homography.py:79-161 (est_costvolume_CW + _compute_cost_CW) and,
 * with consistency == 0, homography.py:10-75 (est_costvolume_F + _compute_cost_F);
 * with depth_mode == MAGNET_DEPTH_GAUSS it also absorbs the sampler of models/MAGNET.py:154-156.
 */
typedef struct magnet_cost_args {
  int32_t B, V, D, C, H, W;
  int32_t depth_mode;      /* magnet_depth_mode                                                 */
  int32_t src_layout;      /* magnet_src_layout                                                 */
  int32_t consistency;     /* 1: CW weighting |z - mu~| < kappa*sigma~ ; 0: plain plane sweep   */
  int32_t softmax;         /* 1: softmax over the D planes after the 1/V mean (est_costvolume_F)*/
  int32_t variant;         /* magnet_variant                                                    */
  float kappa;             /* 'thres' of est_costvolume_CW (float(int))                         */
  const float* ref_feat;   /* (B, C, H, W) NCHW; a split buffer with MAGNET_SRC_SPLIT16           */
  const float* src_feat;   /* (V*B, ...) view-major, layout = src_layout                        */
  const float* src_gmm;    /* (V*B, 2, H, W) [mu, sigma]; required when consistency == 1        */
  const float* rays;       /* (B, 3, H*W) 'unit_ray_array_2D'                                   */
  const magnet_camera* cams; /* (B*V)                                                           */
  const float* d_volume;   /* (B, D, H, W); MAGNET_DEPTH_VOLUME only                            */
  const float* ref_gmm;    /* (B, 2, H, W) [mu, sigma]; MAGNET_DEPTH_GAUSS only                 */
  const float* k_host;     /* HOST pointer, D floats: k_j (GAUSS) or plane depths (PLANES)      */
  float* out;              /* (B, D, H, W)                                                      */
} magnet_cost_args;

int magnet_abi_version(void);
const char* magnet_strerror(int status);
/* Text of the last CUDA error seen by this library on the calling thread ("" if none). */
const char* magnet_last_cuda_error(void);
/* Number of kernels this library has launched so far in this process (all entry points). */
uint64_t magnet_launch_count(void);

/* Bytes of dynamic shared memory and threads per CTA the selected variant will use (for reports). */
int magnet_cost_launch_info(const magnet_cost_args* args, int* grid_ctas, int* block_threads, int* smem_bytes);

/* Cost volume, forward.  Replaces homography.est_costvolume_CW / est_costvolume_F (see above). */
int magnet_cost_volume_f32(const magnet_cost_args* args, void* stream);

/*
 * Backward of the plane-sweep volume (est_costvolume_F) w.r.t. both feature maps — what autograd derives for
 * homography.py:10-75 during F-Net training (train_FNet.py:95-114): through the softmax, the 1/V mean, the channel
 * dot product and grid_sample's bilinear gather (scatter-add into the source features).
 * Geometry fields of `fwd` as in the forward call (consistency must be 0, depth_mode MAGNET_DEPTH_PLANES,
 * src_layout MAGNET_SRC_NCHW, C in {8,16,32,64}); `fwd->out` is ignored.
 */
typedef struct magnet_cost_f_bwd_args {
  const magnet_cost_args* fwd;
  const float* prob;      /* (B,D,H,W) forward output; used when fwd->softmax == 1                          */
  const float* grad_out;  /* (B,D,H,W) gradient w.r.t. the forward output                                   */
  float* workspace;       /* (B,D,H,W) scratch (gradient w.r.t. the pre-softmax scores)                     */
  float* grad_ref;        /* (B,C,H,W)   written                                                            */
  float* grad_src;        /* (V*B,C,H,W) NCHW, ACCUMULATED with atomics: the caller zeroes it               */
} magnet_cost_f_bwd_args;
int magnet_cost_volume_f_bwd_f32(const magnet_cost_f_bwd_args* args, void* stream);

/*
 * Camera constants.  Replaces homography.py:89,98-102 (IntM/R/t products, done there per pair per
 * iteration).  intM (B,3,3) dense; R and t are addressed with element strides so that the
 * non-contiguous views nghbr_poses[:,:,:3,:3] / [:,:,:3,3] of MAGNET.py:147-148 can be passed as is:
 *   R[b,v,i,j] = R[b*r_sb + v*r_sv + i*r_si + j*r_sj],  t[b,v,i] = t[b*t_sb + v*t_sv + i*t_si].
 * is_valid (B,V) int32 on the device.
 */
int magnet_pack_cameras_f32(const float* intM, const float* R, int64_t r_sb, int64_t r_sv, int64_t r_si,
                            int64_t r_sj, const float* t, int64_t t_sb, int64_t t_sv, int64_t t_si,
                            const int32_t* is_valid, int32_t B, int32_t V, magnet_camera* cams_out,
                            void* stream);

/* Source repack (N, C, H, W) features [+ (N, 2, H, W) Gaussians, may be NULL -> zeros] -> MAGNET_SRC_PIXC
 * (N, H, W, C+4); C in {16, 32, 64}, dst 16-byte aligned.  Once per forward (the features do not change across the
 * N_iter iterations, MAGNET.py:150-169). */
int magnet_repack_pixc_f32(const float* src_nchw, const float* src_gmm, float* dst, int32_t N, int32_t C, int32_t H,
                           int32_t W, void* stream);

/* Feature split (N, 64, H, W) features [+ (N, 2, H, W) Gaussians, may be NULL -> zeros] -> MAGNET_SRC_SPLIT16 buffer of
 * magnet_split16_bytes(N, H, W) bytes; src and dst 16-byte aligned.  Three stream operations (header memset, max |x|
 * reduction, split).  Once per forward for the source views and once for the reference features. */
size_t magnet_split16_bytes(int32_t N, int32_t H, int32_t W);
int magnet_repack_split16_f32(const float* src_nchw, const float* src_gmm, void* dst, int32_t N, int32_t C, int32_t H,
                              int32_t W, void* stream);

/* Source-feature repack (N, C, H, W) -> MAGNET_SRC_TILED32 (N, H, ceil(W/32), C/4, 32, 4);
 * C % 4 == 0, dst 16-byte aligned, padding pixels are written as zeros. */
int magnet_repack_tiled32_f32(const float* src_nchw, float* dst, int32_t N, int32_t C, int32_t H, int32_t W,
                            void* stream);

/*
 * Depth-candidate sampler alone.  Replaces models/MAGNET.py:154-156:
 *   d_volume[b,j,y,x] = mu[b,y,x] + sigma[b,y,x] * k_j   (separate multiply and add).
 * gmm (B,2,H,W); k_host: HOST pointer, D floats; d_volume (B,D,H,W).
 */
int magnet_sample_depths_f32(const float* gmm, const float* k_host, int32_t B, int32_t D, int32_t HW,
                             float* d_volume, void* stream);

/*
 * Gaussian update, forward.  Replaces models/MAGNET.py:60,65-69 (inside GNET.forward):
 *   mu' = mu0 + mu1*sigma0 ; sigma' = (elu(sigma1) + 1 + 1e-10) * sigma0.
 * d_output (B,2,H,W) = G-Net raw output (mu1, sigma1); ref_gmm (B,2,H,W); out (B,2,H,W).
 */
int magnet_gaussian_update_fwd_f32(const float* d_output, const float* ref_gmm, int32_t B, int32_t HW,
                                   float* out, void* stream);
/* Backward of the update w.r.t. d_output (ref_gmm is detached in the reference, MAGNET.py:168):
 *   g_mu1 = g_mu' * sigma0 ; g_sigma1 = g_sigma' * sigma0 * (sigma1 > 0 ? 1 : exp(sigma1)). */
int magnet_gaussian_update_bwd_f32(const float* grad_out, const float* d_output, const float* ref_gmm,
                                   int32_t B, int32_t HW, float* grad_d_output, void* stream);

/*
 * Caller-side camera preparation on the device.
 * magnet_relative_poses_f32 replaces utils/utils.py:72-98 (data_preprocess): poses_out[b,v] = ext_nghbr[v,b] *
 * inv(ext_ref[b]) (ref-camera -> source-camera), is_valid_out[b,v] = 0 when either extrinsic or the product holds a
 * NaN (the pose is then all zeros).  ext_ref (B,4,4); ext_nghbr (V,B,4,4) view-major; poses_out (B,V,4,4).
 */
int magnet_relative_poses_f32(const float* ext_ref, const float* ext_nghbr, int32_t B, int32_t V, float* poses_out,
                              int32_t* is_valid_out, void* stream);
/*
 * magnet_camera_rays_f32 replaces get_ray_array + get_cam_intrinsics (data/dataloader_scannet.py:113-153 and the
 * crop-margin variant data/dataloader_kitti.py:94-127): raw_intrinsics (B,8) float64 on the device =
 * fx, fy, cx, cy of the raw image; img_W, img_H = size of the (cropped) image the grid spans; left_margin, top_margin
 * = crop offsets in raw pixels (ScanNet: the raw size and 0, 0; KITTI: 1216, 352, (raw_W-1216)/2, raw_H-352).
 * intM_out (B,3,3) = intrinsics of the H x W grid; rays_out (B,3,H*W) = K_raw^-1 (pixel centre), z = 1.
 * Evaluated in fp64 and rounded once, bit-identical to the numpy originals.
 */
int magnet_camera_rays_f32(const double* raw_intrinsics, int32_t B, int32_t H, int32_t W, float* intM_out,
                           float* rays_out, void* stream);

/*
 * Learned convex upsampling.  Replaces upsample_depth_via_mask (models/MAGNET.py:15-27): softmax over the 9
 * neighbours of up_mask (B, 9*k*k, H, W) viewed (B,1,9,k,k,H,W), weighted sum of the zero-padded 3x3
 * neighbourhood of depth (B,CH,H,W), pixel shuffle -> out (B,CH,k*H,k*W).  CH in {1,2}.
 */
int magnet_convex_upsample_fwd_f32(const float* depth, const float* up_mask, int32_t B, int32_t CH, int32_t H,
                                   int32_t W, int32_t k, float* out, void* stream);
/* Backward: grad_mask (B,9*k*k,H,W) is written; grad_depth (B,CH,H,W) is ACCUMULATED (caller zeroes it). */
int magnet_convex_upsample_bwd_f32(const float* grad_out, const float* depth, const float* up_mask, int32_t B,
                                   int32_t CH, int32_t H, int32_t W, int32_t k, float* grad_depth, float* grad_mask,
                                   void* stream);

/*
 * Convex upsampling fused with the Gaussian negative log-likelihood — replaces, per prediction of pred_list,
 * upsample_depth_via_mask (models/MAGNET.py:15-27,172-173) followed by MagnetLoss's term (utils/losses.py:39-49):
 *   nll = (mu - gt)^2 / (2 var) + 0.5 log(var),  var = max(sigma^2, 1e-10),  over the pixels where gt_mask != 0.
 * depth (B,2,H,W) quarter-resolution [mu, sigma]; up_mask (B,9*k*k,H,W); gt (B,1,k*H,k*W); gt_mask (B,1,k*H,k*W)
 * uint8.  The (B,2,k*H,k*W) prediction is never materialised.
 * forward: partial[magnet_upsample_nll_partials(B,H,W,k)] receives one partial sum of nll per CTA (the caller adds
 *   them — deterministic — and divides by the number of supervised pixels).
 * backward: scale = upstream gradient * gamma^(n-i-1) / number of supervised pixels; grad_depth (B,2,H,W) is
 *   ACCUMULATED (the caller zeroes it), grad_mask (B,9*k*k,H,W) is written.
 */
int magnet_upsample_nll_partials(int32_t B, int32_t H, int32_t W, int32_t k);
int magnet_upsample_nll_fwd_f32(const float* depth, const float* up_mask, const float* gt, const uint8_t* gt_mask,
                                int32_t B, int32_t H, int32_t W, int32_t k, float* partial, void* stream);
int magnet_upsample_nll_bwd_f32(const float* depth, const float* up_mask, const float* gt, const uint8_t* gt_mask,
                                float scale, int32_t B, int32_t H, int32_t W, int32_t k, float* grad_depth,
                                float* grad_mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGNET_B200_H_ */
