"""Drop-in replacements for ``models.submodules.homography`` of the reference.

Same function names, argument order, argument meaning and return contract as
``est_costvolume_CW`` (homography.py:79-121) and ``est_costvolume_F`` (homography.py:10-47), so
that ``MAGNET.forward`` (models/MAGNET.py:160-164) and ``MAGNET_F.forward`` (:197-200) run
unchanged after ``magnet_b200.install()``.

What the wrapper does around the single kernel launch (all of it hoisted out of the reference's
per-(batch, view) Python loop):
  * ``cam_intrins`` / ``is_valid`` arrive as CPU tensors (test_MaGNet.py:36-50): uploaded once and
    cached across the N_iter calls of one forward (keyed by object identity + version counter);
  * ``R`` / ``t`` arrive as non-contiguous views of ``nghbr_poses`` (MAGNET.py:147-148): passed to
    ``magnet_pack_cameras_f32`` with their strides, no copy;
  * ``nghbr_feat`` arrives NCHW: repacked once per forward into the (N, C/4, H, W, 4) layout the
    tap-sharing kernel gathers from (cached the same way).
The CW volume is not differentiable (its inputs never require grad in the reference);
the F volume is differentiable w.r.t. both feature maps (magnet_cost_volume_f_bwd_f32) for F-Net training.
"""
from __future__ import annotations

import weakref
from typing import Dict, Tuple

import torch

from . import _lib, ops


class _PrepCache:
    """Identity + version keyed cache of per-forward preparations (uploads, repacks, camera tables).

    An entry is valid only while the *same tensor object* is alive and unmodified: the key holds a
    weakref to every source tensor and its ``_version``; a freed-and-reallocated tensor at the same
    address can therefore never alias a stale entry."""

    def __init__(self, capacity: int = 8):
        self.capacity = capacity
        self._items: Dict[Tuple, Tuple[tuple, object]] = {}

    @staticmethod
    def _sig(tensors):
        return tuple((id(t), t._version, tuple(t.shape), str(t.device)) for t in tensors)

    def get(self, kind: str, tensors, extra=()):
        key = (kind,) + self._sig(tensors) + tuple(extra)
        hit = self._items.get(key)
        if hit is not None:
            refs, value = hit
            if all(r() is t for r, t in zip(refs, tensors)):
                return value
            del self._items[key]
        return None

    def put(self, kind: str, tensors, value, extra=()):
        key = (kind,) + self._sig(tensors) + tuple(extra)
        for dead in [k for k, (refs, _) in self._items.items() if any(r() is None for r in refs)]:
            del self._items[dead]                      # drop preparations whose source tensor is gone
        if len(self._items) >= self.capacity:
            self._items.pop(next(iter(self._items)))
        self._items[key] = (tuple(weakref.ref(t) for t in tensors), value)
        return value

    def clear(self):
        self._items.clear()


_cache = _PrepCache()


def clear_cache() -> None:
    _cache.clear()


def _device_intrinsics(cam_intrins, device):
    intM, rays = cam_intrins['intM'], cam_intrins['unit_ray_array_2D']
    hit = _cache.get("intr", (intM, rays), (str(device),))
    if hit is not None:
        return hit
    value = (intM.to(device=device, dtype=torch.float32).contiguous(),
             rays.to(device=device, dtype=torch.float32).contiguous())
    return _cache.put("intr", (intM, rays), value, (str(device),))


def _camera_table(cam_intrins, R, t, is_valid, device):
    intM_d, _ = _device_intrinsics(cam_intrins, device)
    # R and t are fresh views on every call (MAGNET.py:147-148 slices once per forward, but a caller
    # may re-slice); key on the storage they view + their layout instead of the view object.
    base = R._base if R._base is not None else R
    hit = _cache.get("cams", (base, is_valid, cam_intrins['intM']),
                     (R.data_ptr(), R.stride(), t.data_ptr(), t.stride()))
    if hit is not None:
        return hit
    valid_d = is_valid.to(device=device, dtype=torch.int32)
    cams = ops.pack_cameras(intM_d, R, t, valid_d)
    return _cache.put("cams", (base, is_valid, cam_intrins['intM']), cams,
                      (R.data_ptr(), R.stride(), t.data_ptr(), t.stride()))


def _packed_source(nghbr_feat):
    if nghbr_feat.shape[1] % 4 != 0:
        return nghbr_feat.contiguous(), _lib.SRC_NCHW
    hit = _cache.get("tiled32", (nghbr_feat,))
    if hit is None:
        hit = _cache.put("tiled32", (nghbr_feat,), ops.repack_tiled32(nghbr_feat))
    return hit, _lib.SRC_TILED32


def est_costvolume_CW(d_volume, ref_feat, nghbr_feat, ref_gmms, nghbr_gmms,
                      R, t, is_valid, cam_intrins, thres, variant=_lib.VARIANT_AUTO):
    """Consistency-weighted multi-view cost volume — drop-in for homography.est_costvolume_CW.

    d_volume (B,D,H,W); ref_feat (B,C,H,W); nghbr_feat (V*B,C,H,W) view-major; ref_gmms unused (as in
    the reference); nghbr_gmms (V*B,2,H,W) [mu, sigma]; R (B,V,3,3), t (B,V,3) device
    views; is_valid (B,V) int CPU or device; cam_intrins dict of 'intM' (B,3,3) and
    'unit_ray_array_2D' (B,3,H*W), CPU or device; thres int.  Returns (B,D,H,W) float32 on
    ref_feat.device, freshly allocated, detached."""
    device = ref_feat.device
    B = d_volume.shape[0]
    V = int(nghbr_feat.shape[0] / B)
    with torch.no_grad():
        _, rays_d = _device_intrinsics(cam_intrins, device)
        cams = _camera_table(cam_intrins, R, t, is_valid, device)
        src, layout = _packed_source(nghbr_feat.detach())
        return ops.cost_volume(ref_feat.detach(), src, rays_d, cams, V=V, src_layout=layout, consistency=True,
                               src_gmm=nghbr_gmms.detach(), kappa=float(thres), d_volume=d_volume.detach(),
                               variant=variant)


class _CostVolumeF(torch.autograd.Function):
    """Plane-sweep probability volume for F-Net training (homography.py:10-75), forward + backward kernels."""

    @staticmethod
    def forward(ctx, ref_feat, nghbr_feat, planes, rays_d, cams, V, variant):
        src, layout = _packed_source(nghbr_feat.detach())
        out = ops.cost_volume(ref_feat.detach(), src, rays_d, cams, V=V, src_layout=layout, consistency=False,
                              k=planes, planes=True, softmax=True, variant=variant)
        ctx.save_for_backward(ref_feat.detach(), nghbr_feat.detach(), out, rays_d, cams)
        ctx.planes, ctx.V = planes, V
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ref_feat, nghbr_feat, out, rays_d, cams = ctx.saved_tensors
        g_ref, g_src = ops.cost_volume_f_bwd(ref_feat, nghbr_feat, rays_d, cams, ctx.planes, ctx.V, out,
                                             grad_out.contiguous(), softmax=True)
        return g_ref, g_src, None, None, None, None, None


def est_costvolume_F(d_center, ref_feat, nghbr_feat, R, t, is_valid, cam_intrins, variant=_lib.VARIANT_AUTO):
    """Fronto-parallel plane-sweep volume with softmax over planes — drop-in (forward) for
    homography.est_costvolume_F.  d_center (1,D,1,1); the rest as in est_costvolume_CW."""
    device = ref_feat.device
    B = ref_feat.shape[0]
    V = int(nghbr_feat.shape[0] / B)
    planes = d_center.detach().reshape(-1).cpu().tolist()
    _, rays_d = _device_intrinsics(cam_intrins, device)
    cams = _camera_table(cam_intrins, R, t, is_valid, device)
    return _CostVolumeF.apply(ref_feat, nghbr_feat, planes, rays_d, cams, V, variant)
