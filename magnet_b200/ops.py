"""Tensor-level wrappers over the C ABI (one Python call = one kernel launch) and the
``torch.autograd.Function`` wrappers the reference-facing modules use.

Everything here requires CUDA tensors and the built library; nothing falls back to PyTorch.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import CostArgs, check, lib


def _stream(dev=None) -> int:
    """The current stream of the device the operands live on (not of whatever device happens to be current)."""
    return torch.cuda.current_stream(dev).cuda_stream


def _same_device(*named):
    """All operands of one launch must live on one CUDA device; returns it."""
    dev = None
    for name, x in named:
        if x is None:
            continue
        if dev is None:
            dev = x.device
        elif x.device != dev:
            raise _lib.MagnetError(f"{name} is on {x.device}, expected {dev} (all operands of a launch share one device)")
    return dev


def _expect(name: str, x: torch.Tensor, shape) -> None:
    if tuple(x.shape) != tuple(shape):
        raise _lib.MagnetError(f"{name} must have shape {tuple(shape)}, got {tuple(x.shape)}")


def _need_cuda_f32(name: str, x: torch.Tensor, contiguous: bool = True) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not x.is_cuda:
        raise _lib.MagnetError(f"{name} must be a CUDA tensor (magnet_b200 has no CPU path)")
    if x.dtype != torch.float32:
        raise _lib.MagnetError(f"{name} must be float32 (the reference path is fp32-only, homography.py:130)")
    if contiguous and not x.is_contiguous():
        x = x.contiguous()
    return x


def k_array(k: Sequence[float]):
    """Python / numpy / tensor sequence -> host float[D] (rounded to fp32 like torch does for
    tensor * python-scalar, MAGNET.py:155)."""
    if isinstance(k, torch.Tensor):
        k = k.detach().cpu().flatten().tolist()
    vals = [float(v) for v in k]
    if len(vals) > _lib.MAGNET_MAX_PLANES:
        raise _lib.MagnetError(f"at most {_lib.MAGNET_MAX_PLANES} hypotheses per call, got {len(vals)}")
    return (C.c_float * len(vals))(*vals)


def pack_cameras(intM: torch.Tensor, R: torch.Tensor, t: torch.Tensor, is_valid: torch.Tensor) -> torch.Tensor:
    """(B,3,3) intrinsics, (B,V,3,3) / (B,V,3) pose views (any strides), (B,V) int32 validity — all on the
    device — -> (B*V, 16) float32 camera-constant table (struct magnet_camera)."""
    intM = _need_cuda_f32("intM", intM)
    R = _need_cuda_f32("R", R, contiguous=False)
    t = _need_cuda_f32("t", t, contiguous=False)
    B, V = R.shape[0], R.shape[1]
    if is_valid.dtype != torch.int32 or not is_valid.is_cuda:
        is_valid = is_valid.to(device=intM.device, dtype=torch.int32)
    is_valid = is_valid.contiguous()
    cams = torch.empty(B * V, 16, device=intM.device, dtype=torch.float32)
    rs, ts = R.stride(), t.stride()
    dev = _same_device(("intM", intM), ("R", R), ("t", t), ("is_valid", is_valid))
    with torch.cuda.device(dev):
        check(lib().magnet_pack_cameras_f32(intM.data_ptr(), R.data_ptr(), rs[0], rs[1], rs[2], rs[3],
                                            t.data_ptr(), ts[0], ts[1], ts[2], is_valid.data_ptr(), B, V,
                                            cams.data_ptr(), _stream(dev)), "magnet_pack_cameras_f32")
    return cams


def repack_tiled32(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N,C,H,W) -> TILED32 (N, H, ceil(W/32), C/4, 32, 4): the source-feature layout the tap-sharing
    kernel gathers from (channel quads of a pixel 512 B apart, 32 neighbouring pixels contiguous)."""
    x = _need_cuda_f32("x", x)
    N, Cc, H, W = x.shape
    if out is None:
        out = torch.empty(N, H, (W + 31) // 32, Cc // 4, 32, 4, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib().magnet_repack_tiled32_f32(x.data_ptr(), out.data_ptr(), N, Cc, H, W, _stream(x.device)),
              "magnet_repack_tiled32_f32")
    return out


def repack_pixc(x: torch.Tensor, gmm: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N,C,H,W) features [+ (N,2,H,W) Gaussians] -> PIXC (N, H, W, C+4): pixel-major, per pixel the C channels then
    (mu, sigma, 0, 0) — the layout the TMA-staged CUDA-core kernel fetches its windows from.  C in {16, 32, 64}."""
    x = _need_cuda_f32("x", x)
    N, Cc, H, W = x.shape
    gptr = None
    if gmm is not None:
        gmm = _need_cuda_f32("gmm", gmm)
        if tuple(gmm.shape) != (N, 2, H, W):
            raise _lib.MagnetError(f"gmm must be (N,2,H,W) = {(N, 2, H, W)}, got {tuple(gmm.shape)}")
        gptr = gmm.data_ptr()
    if out is None:
        out = torch.empty(N, H, W, Cc + 4, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib().magnet_repack_pixc_f32(x.data_ptr(), gptr, out.data_ptr(), N, Cc, H, W, _stream(x.device)),
              "magnet_repack_pixc_f32")
    return out


def repack_split16(x: torch.Tensor, gmm: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N,64,H,W) features [+ (N,2,H,W) Gaussians] -> SPLIT16 buffer (uint8): header with the power-of-two scale s, fp16
    planes (N,2,H,W,64) with x*s = hi + lo, table (N,H,W,4) = (mu, sigma, 0, 0) — what the tensor-core kernel's TMA boxes
    fetch (reference features: gmm=None)."""
    x = _need_cuda_f32("x", x)
    N, Cc, H, W = x.shape
    gptr = None
    if gmm is not None:
        gmm = _need_cuda_f32("gmm", gmm)
        if tuple(gmm.shape) != (N, 2, H, W):
            raise _lib.MagnetError(f"gmm must be (N,2,H,W) = {(N, 2, H, W)}, got {tuple(gmm.shape)}")
        gptr = gmm.data_ptr()
    nbytes = int(lib().magnet_split16_bytes(N, H, W))
    if out is None:
        out = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    elif out.numel() * out.element_size() < nbytes or out.device != x.device:
        raise _lib.MagnetError(f"out must hold {nbytes} bytes on {x.device}")
    with torch.cuda.device(x.device):
        check(lib().magnet_repack_split16_f32(x.data_ptr(), gptr, out.data_ptr(), N, Cc, H, W, _stream(x.device)),
              "magnet_repack_split16_f32")
    return out


def sample_depths(gmm: torch.Tensor, k, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sampler alone (MAGNET.py:154-156): gmm (B,2,H,W) -> d_volume (B,D,H,W)."""
    gmm = _need_cuda_f32("gmm", gmm)
    karr = k if isinstance(k, C.Array) else k_array(k)
    B, _, H, W = gmm.shape
    D = len(karr)
    if out is None:
        out = torch.empty(B, D, H, W, device=gmm.device, dtype=torch.float32)
    with torch.cuda.device(gmm.device):
        check(lib().magnet_sample_depths_f32(gmm.data_ptr(), C.cast(karr, C.c_void_p), B, D, H * W, out.data_ptr(),
                                             _stream(gmm.device)), "magnet_sample_depths_f32")
    return out


def cost_volume(ref_feat: torch.Tensor, src_feat: torch.Tensor, rays: torch.Tensor, cams: torch.Tensor, *,
                V: int, src_layout: int, consistency: bool, src_gmm: Optional[torch.Tensor] = None,
                kappa: float = 5.0, d_volume: Optional[torch.Tensor] = None,
                ref_gmm: Optional[torch.Tensor] = None, k=None, planes: bool = False, softmax: bool = False,
                variant: int = _lib.VARIANT_AUTO, out: Optional[torch.Tensor] = None,
                ref_split: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One launch of magnet_cost_volume_f32.  Depth source: ``d_volume`` (drop-in), or ``ref_gmm`` + ``k``
    (fused sampler), or ``k`` with ``planes=True`` (fronto-parallel planes).  With ``src_layout=SRC_SPLIT16`` both
    ``src_feat`` and ``ref_split`` are ``repack_split16`` buffers (``ref_feat`` then only supplies the shape)."""
    ref_feat = _need_cuda_f32("ref_feat", ref_feat)
    if src_layout == _lib.SRC_SPLIT16:
        for nm, buf in (("src_feat", src_feat), ("ref_split", ref_split)):
            if not (isinstance(buf, torch.Tensor) and buf.is_cuda and buf.dtype == torch.uint8 and buf.is_contiguous()):
                raise _lib.MagnetError(f"{nm} must be a contiguous uint8 CUDA buffer from repack_split16")
    else:
        src_feat = _need_cuda_f32("src_feat", src_feat)
    rays = _need_cuda_f32("rays", rays)
    cams = _need_cuda_f32("cams", cams)
    if ref_feat.dim() != 4:
        raise _lib.MagnetError(f"ref_feat must be (B,C,H,W), got {tuple(ref_feat.shape)}")
    B, Cc, H, W = ref_feat.shape
    if V <= 0:
        raise _lib.MagnetError(f"V must be positive, got {V}")
    # every operand against (B, V, D, C, H, W): a mismatch would read out of bounds, the reference raises instead
    src_shape = {_lib.SRC_NCHW: (V * B, Cc, H, W), _lib.SRC_TILED32: (V * B, H, (W + 31) // 32, Cc // 4, 32, 4),
                 _lib.SRC_PIXC: (V * B, H, W, Cc + 4),
                 _lib.SRC_SPLIT16: (int(lib().magnet_split16_bytes(V * B, H, W)),)}.get(src_layout)
    if src_shape is None:
        raise _lib.MagnetError(f"unknown src_layout {src_layout}")
    _expect("src_feat", src_feat, src_shape)
    if src_layout == _lib.SRC_SPLIT16:
        _expect("ref_split", ref_split, (int(lib().magnet_split16_bytes(B, H, W)),))
    _expect("rays", rays, (B, 3, H * W))
    if cams.numel() != B * V * 16:
        raise _lib.MagnetError(f"cams must hold B*V = {B * V} camera records of 16 floats, got {tuple(cams.shape)}")
    dev = _same_device(("ref_feat", ref_feat), ("src_feat", src_feat), ("rays", rays), ("cams", cams),
                       ("src_gmm", src_gmm), ("d_volume", d_volume), ("ref_gmm", ref_gmm), ("out", out),
                       ("ref_split", ref_split))
    a = CostArgs()
    a.B, a.V, a.C, a.H, a.W = B, V, Cc, H, W
    a.src_layout = src_layout
    a.consistency = 1 if consistency else 0
    a.softmax = 1 if softmax else 0
    a.variant = variant
    a.kappa = float(kappa)
    a.ref_feat, a.src_feat, a.rays, a.cams = ref_feat.data_ptr(), src_feat.data_ptr(), rays.data_ptr(), cams.data_ptr()
    keep = [ref_feat, src_feat, rays, cams]
    if src_layout == _lib.SRC_SPLIT16:
        a.ref_feat = ref_split.data_ptr()
        keep.append(ref_split)
    if consistency and src_layout not in (_lib.SRC_PIXC, _lib.SRC_SPLIT16):   # those carry the source Gaussians inside src_feat
        src_gmm = _need_cuda_f32("src_gmm", src_gmm)
        _expect("src_gmm", src_gmm, (V * B, 2, H, W))
        a.src_gmm = src_gmm.data_ptr()
        keep.append(src_gmm)
    karr = None
    if d_volume is not None:
        d_volume = _need_cuda_f32("d_volume", d_volume)
        if d_volume.dim() != 4 or d_volume.shape[0] != B or tuple(d_volume.shape[2:]) != (H, W):
            raise _lib.MagnetError(f"d_volume must be (B,D,H,W) = ({B},D,{H},{W}), got {tuple(d_volume.shape)}")
        a.depth_mode, a.D, a.d_volume = _lib.DEPTH_VOLUME, d_volume.shape[1], d_volume.data_ptr()
        keep.append(d_volume)
    else:
        karr = k if isinstance(k, C.Array) else k_array(k)
        a.D = len(karr)
        a.k_host = C.cast(karr, C.c_void_p)
        if planes:
            a.depth_mode = _lib.DEPTH_PLANES
        else:
            ref_gmm = _need_cuda_f32("ref_gmm", ref_gmm)
            _expect("ref_gmm", ref_gmm, (B, 2, H, W))
            a.depth_mode, a.ref_gmm = _lib.DEPTH_GAUSS, ref_gmm.data_ptr()
            keep.append(ref_gmm)
    if out is None:
        out = torch.empty(B, a.D, H, W, device=ref_feat.device, dtype=torch.float32)
    else:
        out = _need_cuda_f32("out", out)
        _expect("out", out, (B, a.D, H, W))
    a.out = out.data_ptr()
    with torch.cuda.device(dev):
        check(lib().magnet_cost_volume_f32(C.byref(a), _stream(dev)), "magnet_cost_volume_f32")
    return out


def cost_volume_f_bwd(ref_feat, src_feat_nchw, rays, cams, planes, V, prob, grad_out, softmax=True):
    """Gradients of the plane-sweep volume w.r.t. (ref_feat, src_feat) — one magnet_cost_volume_f_bwd_f32 call.
    src_feat_nchw (V*B,C,H,W) view-major; prob = forward output; returns (grad_ref, grad_src) in NCHW."""
    ref_feat = _need_cuda_f32("ref_feat", ref_feat)
    src = _need_cuda_f32("src_feat", src_feat_nchw)
    rays = _need_cuda_f32("rays", rays)
    cams = _need_cuda_f32("cams", cams)
    prob = _need_cuda_f32("prob", prob)
    grad_out = _need_cuda_f32("grad_out", grad_out)
    B, Cc, H, W = ref_feat.shape
    karr = planes if isinstance(planes, C.Array) else k_array(planes)
    _expect("src_feat", src, (V * B, Cc, H, W))
    _expect("rays", rays, (B, 3, H * W))
    _expect("prob", prob, (B, len(karr), H, W))
    _expect("grad_out", grad_out, (B, len(karr), H, W))
    if cams.numel() != B * V * 16:
        raise _lib.MagnetError(f"cams must hold B*V = {B * V} camera records of 16 floats, got {tuple(cams.shape)}")
    dev = _same_device(("ref_feat", ref_feat), ("src_feat", src), ("rays", rays), ("cams", cams), ("prob", prob),
                       ("grad_out", grad_out))
    a = CostArgs()
    a.B, a.V, a.D, a.C, a.H, a.W = B, V, len(karr), Cc, H, W
    a.depth_mode, a.src_layout, a.consistency, a.softmax = _lib.DEPTH_PLANES, _lib.SRC_NCHW, 0, 1 if softmax else 0
    a.ref_feat, a.src_feat, a.rays, a.cams = ref_feat.data_ptr(), src.data_ptr(), rays.data_ptr(), cams.data_ptr()
    a.k_host = C.cast(karr, C.c_void_p)
    work = torch.empty_like(prob)
    g_ref = torch.empty_like(ref_feat)
    g_src = torch.zeros_like(src)
    bw = _lib.CostFBwdArgs()
    bw.fwd = C.pointer(a)
    bw.prob, bw.grad_out, bw.workspace = prob.data_ptr(), grad_out.data_ptr(), work.data_ptr()
    bw.grad_ref, bw.grad_src = g_ref.data_ptr(), g_src.data_ptr()
    with torch.cuda.device(dev):
        check(lib().magnet_cost_volume_f_bwd_f32(C.byref(bw), _stream(dev)), "magnet_cost_volume_f_bwd_f32")
    return g_ref, g_src


def cost_launch_info(B, V, D, Cc, H, W, variant=_lib.VARIANT_AUTO):
    """(grid CTAs, threads per CTA, dynamic smem bytes) the cost kernel would use for these sizes."""
    a = CostArgs()
    a.B, a.V, a.D, a.C, a.H, a.W = B, V, D, Cc, H, W
    layout = {_lib.VARIANT_TMA: _lib.SRC_PIXC, _lib.VARIANT_MMA: _lib.SRC_SPLIT16}.get(variant, _lib.SRC_TILED32)
    a.depth_mode, a.src_layout, a.consistency, a.variant = _lib.DEPTH_PLANES, layout, 0, variant
    one = C.c_void_p(0x1000)                       # never dereferenced: validated for non-NULL / alignment only
    a.ref_feat = a.src_feat = a.rays = a.cams = a.out = a.k_host = one
    g, b, s = C.c_int(), C.c_int(), C.c_int()
    check(lib().magnet_cost_launch_info(C.byref(a), C.byref(g), C.byref(b), C.byref(s)), "magnet_cost_launch_info")
    return g.value, b.value, s.value


class GaussianUpdate(torch.autograd.Function):
    """mu' = mu0 + mu1*sigma0 ; sigma' = (elu(sigma1) + 1 + 1e-10)*sigma0  (MAGNET.py:60,65-69).
    Differentiable w.r.t. the G-Net output only; ``ref_gmm`` is detached in the reference (MAGNET.py:168)."""

    @staticmethod
    def forward(ctx, d_output: torch.Tensor, ref_gmm: torch.Tensor) -> torch.Tensor:
        d_output = _need_cuda_f32("d_output", d_output)
        ref_gmm = _need_cuda_f32("ref_gmm", ref_gmm.detach())
        B, _, H, W = d_output.shape
        out = torch.empty_like(d_output)
        _expect("ref_gmm", ref_gmm, d_output.shape)
        dev = _same_device(("d_output", d_output), ("ref_gmm", ref_gmm))
        with torch.cuda.device(dev):
            check(lib().magnet_gaussian_update_fwd_f32(d_output.data_ptr(), ref_gmm.data_ptr(), B, H * W,
                                                       out.data_ptr(), _stream(dev)), "magnet_gaussian_update_fwd_f32")
        ctx.save_for_backward(d_output, ref_gmm)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        d_output, ref_gmm = ctx.saved_tensors
        grad_out = _need_cuda_f32("grad_out", grad_out)
        B, _, H, W = d_output.shape
        gin = torch.empty_like(d_output)
        with torch.cuda.device(d_output.device):
            check(lib().magnet_gaussian_update_bwd_f32(grad_out.data_ptr(), d_output.data_ptr(), ref_gmm.data_ptr(),
                                                       B, H * W, gin.data_ptr(), _stream(d_output.device)),
                  "magnet_gaussian_update_bwd_f32")
        return gin, None


def gaussian_update(d_output: torch.Tensor, ref_gmm: torch.Tensor) -> torch.Tensor:
    return GaussianUpdate.apply(d_output, ref_gmm)


class ConvexUpsample(torch.autograd.Function):
    """upsample_depth_via_mask (MAGNET.py:15-27) as one kernel each way; differentiable in depth and mask."""

    @staticmethod
    def forward(ctx, depth: torch.Tensor, up_mask: torch.Tensor, k: int) -> torch.Tensor:
        depth = _need_cuda_f32("depth", depth)
        up_mask = _need_cuda_f32("up_mask", up_mask)
        B, CH, H, W = depth.shape
        if up_mask.shape != (B, 9 * k * k, H, W):
            raise _lib.MagnetError(f"up_mask must be (B, 9*k*k, H, W) = {(B, 9 * k * k, H, W)}, got {tuple(up_mask.shape)}")
        out = torch.empty(B, CH, k * H, k * W, device=depth.device, dtype=torch.float32)
        dev = _same_device(("depth", depth), ("up_mask", up_mask))
        with torch.cuda.device(dev):
            check(lib().magnet_convex_upsample_fwd_f32(depth.data_ptr(), up_mask.data_ptr(), B, CH, H, W, k,
                                                       out.data_ptr(), _stream(dev)), "magnet_convex_upsample_fwd_f32")
        ctx.save_for_backward(depth, up_mask)
        ctx.k = k
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        depth, up_mask = ctx.saved_tensors
        grad_out = _need_cuda_f32("grad_out", grad_out)
        B, CH, H, W = depth.shape
        g_depth = torch.zeros_like(depth)
        g_mask = torch.empty_like(up_mask)
        with torch.cuda.device(depth.device):
            check(lib().magnet_convex_upsample_bwd_f32(grad_out.data_ptr(), depth.data_ptr(), up_mask.data_ptr(), B, CH, H,
                                                       W, ctx.k, g_depth.data_ptr(), g_mask.data_ptr(),
                                                       _stream(depth.device)), "magnet_convex_upsample_bwd_f32")
        return g_depth, g_mask, None


def convex_upsample(depth: torch.Tensor, up_mask: torch.Tensor, k: int) -> torch.Tensor:
    return ConvexUpsample.apply(depth, up_mask, k)


class UpsampleNLL(torch.autograd.Function):
    """mean over the supervised pixels of the Gaussian NLL of ONE convex-upsampled prediction — upsample_depth_via_mask
    (MAGNET.py:15-27) + the per-prediction term of MagnetLoss (utils/losses.py:39-49) in one kernel each way; the
    (B,2,kH,kW) prediction never reaches HBM.  Differentiable in depth (B,2,H,W) and up_mask (B,9k^2,H,W)."""

    @staticmethod
    def forward(ctx, depth, up_mask, gt, gt_mask_u8, k, count):
        depth = _need_cuda_f32("depth", depth)
        up_mask = _need_cuda_f32("up_mask", up_mask)
        gt = _need_cuda_f32("gt", gt)
        B, CH, H, W = depth.shape
        if CH != 2:
            raise _lib.MagnetError(f"depth must be (B,2,H,W) [mu, sigma], got {tuple(depth.shape)}")
        _expect("up_mask", up_mask, (B, 9 * k * k, H, W))
        _expect("gt", gt, (B, 1, k * H, k * W))
        if gt_mask_u8.dtype != torch.uint8 or not gt_mask_u8.is_cuda or tuple(gt_mask_u8.shape) != (B, 1, k * H, k * W):
            raise _lib.MagnetError("gt_mask must be a CUDA uint8 tensor of shape (B,1,k*H,k*W)")
        gt_mask_u8 = gt_mask_u8.contiguous()
        dev = _same_device(("depth", depth), ("up_mask", up_mask), ("gt", gt), ("gt_mask", gt_mask_u8))
        partial = torch.empty(lib().magnet_upsample_nll_partials(B, H, W, k), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(lib().magnet_upsample_nll_fwd_f32(depth.data_ptr(), up_mask.data_ptr(), gt.data_ptr(), gt_mask_u8.data_ptr(),
                                                    B, H, W, k, partial.data_ptr(), _stream(dev)), "magnet_upsample_nll_fwd_f32")
        ctx.save_for_backward(depth, up_mask, gt, gt_mask_u8)
        ctx.k, ctx.count = k, float(count)
        return partial.sum(dtype=torch.float64).to(torch.float32) / ctx.count

    @staticmethod
    def backward(ctx, grad_out):
        depth, up_mask, gt, gtm = ctx.saved_tensors
        B, _, H, W = depth.shape
        g_depth = torch.zeros_like(depth)
        g_mask = torch.empty_like(up_mask)
        scale = float(grad_out) / ctx.count
        with torch.cuda.device(depth.device):
            check(lib().magnet_upsample_nll_bwd_f32(depth.data_ptr(), up_mask.data_ptr(), gt.data_ptr(), gtm.data_ptr(), scale,
                                                    B, H, W, ctx.k, g_depth.data_ptr(), g_mask.data_ptr(),
                                                    _stream(depth.device)), "magnet_upsample_nll_bwd_f32")
        return g_depth, g_mask, None, None, None, None


def magnet_loss(pred_list, up_mask, gt, gt_mask, k: int, gamma: float = 0.8):
    """MagnetLoss 'gaussian' (utils/losses.py:34-50) on the QUARTER-RESOLUTION predictions of the matching loop and the
    shared upsampling mask: sum_i gamma^(n-i-1) * mean NLL(upsample(pred_i)), each term one fused kernel (f-2).
    pred_list: the (B,2,H,W) Gaussians pred_1..pred_n; gt (B,1,kH,kW); gt_mask bool / uint8 of the same shape."""
    gtm = gt_mask.to(torch.uint8)
    count = int(gtm.sum().item())          # one host read per step; the reference's boolean indexing syncs 3x per term
    if count == 0:
        raise _lib.MagnetError("gt_mask selects no pixel")
    n = len(pred_list)
    loss = 0.0
    for i, pred in enumerate(pred_list):
        loss = loss + gamma ** (n - i - 1) * UpsampleNLL.apply(pred, up_mask, gt, gtm, k, count)
    return loss


def relative_poses(ext_ref: torch.Tensor, ext_nghbr: torch.Tensor):
    """data_preprocess (utils/utils.py:72-98) on the device: ext_ref (B,4,4), ext_nghbr (V,B,4,4) ->
    (nghbr_poses (B,V,4,4), is_valid (B,V) int32)."""
    ext_ref = _need_cuda_f32("ext_ref", ext_ref)
    ext_nghbr = _need_cuda_f32("ext_nghbr", ext_nghbr)
    V, B = ext_nghbr.shape[0], ext_nghbr.shape[1]
    poses = torch.empty(B, V, 4, 4, device=ext_ref.device, dtype=torch.float32)
    valid = torch.empty(B, V, device=ext_ref.device, dtype=torch.int32)
    dev = _same_device(("ext_ref", ext_ref), ("ext_nghbr", ext_nghbr))
    with torch.cuda.device(dev):
        check(lib().magnet_relative_poses_f32(ext_ref.data_ptr(), ext_nghbr.data_ptr(), B, V, poses.data_ptr(),
                                              valid.data_ptr(), _stream(dev)), "magnet_relative_poses_f32")
    return poses, valid


def camera_rays(raw_intrinsics: torch.Tensor, H: int, W: int):
    """get_cam_intrinsics (data/dataloader_scannet.py:113-153, data/dataloader_kitti.py:94-127) on the device:
    raw_intrinsics (B,8) float64 [fx, fy, cx, cy, img_W, img_H, left_margin, top_margin] (a (B,6) tensor
    [fx, fy, cx, cy, raw_W, raw_H] is accepted as the ScanNet case: no crop) -> cam_intrins dict
    {'intM' (B,3,3), 'unit_ray_array_2D' (B,3,H*W)} (device)."""
    if not raw_intrinsics.is_cuda or raw_intrinsics.dtype != torch.float64 or raw_intrinsics.dim() != 2 \
            or raw_intrinsics.shape[1] not in (6, 8):
        raise _lib.MagnetError("raw_intrinsics must be a CUDA float64 tensor (B,8) (or (B,6) without crop margins)")
    if raw_intrinsics.shape[1] == 6:
        raw_intrinsics = torch.cat([raw_intrinsics, torch.zeros_like(raw_intrinsics[:, :2])], dim=1)
    raw_intrinsics = raw_intrinsics.contiguous()
    B = raw_intrinsics.shape[0]
    intM = torch.empty(B, 3, 3, device=raw_intrinsics.device, dtype=torch.float32)
    rays = torch.empty(B, 3, H * W, device=raw_intrinsics.device, dtype=torch.float32)
    with torch.cuda.device(raw_intrinsics.device):
        check(lib().magnet_camera_rays_f32(raw_intrinsics.data_ptr(), B, H, W, intM.data_ptr(), rays.data_ptr(),
                                           _stream(raw_intrinsics.device)), "magnet_camera_rays_f32")
    return {"intM": intM, "unit_ray_array_2D": rays}
