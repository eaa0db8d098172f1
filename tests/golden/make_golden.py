"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

Inputs are not stored: they are rebuilt from the seed by magnet_b200.synthetic (numpy Generator
streams are version-stable); each file carries a sha256 of the inputs so a drifting generator is
detected instead of silently comparing against the wrong reference output.

Reference entry points exercised:
  models/submodules/homography.py  est_costvolume_CW (:79), est_costvolume_F (:10)
  models/MAGNET.py                 GNET.forward update equations (:58-70), upsample_depth_via_mask (:15-27),
                                   MAGNET.depth_sampling (:120-128), the sampler expression (:154-156)
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

CASES = {
    # name: (make_inputs kwargs)
    "cw_small_random": dict(B=2, V=3, D=8, H=24, W=32, C=16, seed=1, depth="random", invalid=[(1, 2)]),
    "cw_small_smooth": dict(B=2, V=3, D=8, H=24, W=32, C=16, seed=2, depth="smooth"),
    "cw_c64_d64": dict(B=1, V=2, D=64, H=16, W=64, C=64, seed=3, depth="smooth"),
    "cw_kitti": dict(B=1, V=2, D=12, H=22, W=76, C=32, seed=4, depth="smooth", family="kitti"),
    "cw_cfg1": dict(B=1, V=2, D=16, H=128, W=160, C=64, seed=0, depth="random"),
}
F_PLANES = 12


def input_digest(inp) -> str:
    h = hashlib.sha256()
    for tsr in (inp.ref_feat, inp.nghbr_feat, inp.ref_gmms, inp.nghbr_gmms, inp.nghbr_poses, inp.is_valid,
                inp.cam_intrins['intM'], inp.cam_intrins['unit_ray_array_2D'], inp.k):
        h.update(np.ascontiguousarray(tsr.numpy()).tobytes())
    return h.hexdigest()


def f_planes(n=F_PLANES, d_min=0.5, d_max=8.0):
    """SID plane centres as train_FNet.py:56-66 builds them (n planes instead of 80)."""
    idx = np.arange(n + 1)
    gamma = 1 - d_min
    bounds = np.exp(np.log(d_max + gamma) * idx / n) - gamma
    return ((bounds[:-1] + bounds[1:]) / 2).astype(np.float32)


def main():
    if not os.path.isdir(REF):
        raise SystemExit("/root/reference not present: golden vectors can only be generated in the build container")
    sys.path.insert(0, REF)
    # utils/utils.py:5-7 imports matplotlib, which is absent; the hot path never touches it.
    for name in ("matplotlib", "matplotlib.pyplot"):
        m = types.ModuleType(name)
        m.use = lambda *a, **k: None
        sys.modules.setdefault(name, m)
    import models.submodules.homography as refh
    from models.MAGNET import GNET, MAGNET, upsample_depth_via_mask
    from magnet_b200.synthetic import make_inputs

    torch.set_num_threads(4)
    for name, kw in CASES.items():
        inp = make_inputs(**kw)
        # the sampler exactly as MAGNET.py:154-156 writes it (k_list = python/numpy floats)
        mu, sigma = torch.split(inp.ref_gmms, 1, dim=1)
        holder = types.SimpleNamespace(sampling_range=3, n_samples=kw["D"])
        k_list = MAGNET.depth_sampling(holder)
        dvol = torch.cat([mu + sigma * k for k in k_list], dim=1)
        out = refh.est_costvolume_CW(dvol, inp.ref_feat, inp.nghbr_feat, inp.ref_gmms, inp.nghbr_gmms,
                                     inp.R, inp.t, inp.is_valid, inp.cam_intrins, inp.thres)
        save = dict(cost_cw=out.numpy(), k_list=np.asarray(k_list, dtype=np.float64),
                    digest=np.array(input_digest(inp)), kwargs=np.array(repr(kw)))
        if name != "cw_cfg1":
            save["d_volume"] = dvol.numpy()
            dc = torch.from_numpy(f_planes()).view(1, F_PLANES, 1, 1)
            save["planes"] = dc.numpy().reshape(-1)
            save["cost_f"] = refh.est_costvolume_F(dc, inp.ref_feat, inp.nghbr_feat, inp.R, inp.t,
                                                    inp.is_valid, inp.cam_intrins).numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        print(name, out.shape, "nonzero", float((out != 0).float().mean()))

    # Gaussian update: the reference's GNET.forward with the conv stack replaced by identity,
    # forward value and autograd gradient w.r.t. the (would-be) conv output.
    g = torch.Generator().manual_seed(7)
    d_output = (torch.randn(2, 2, 9, 11, generator=g) * 1.5).requires_grad_(True)
    ref_gmm = torch.stack([torch.rand(2, 9, 11, generator=g) * 4 + 0.5, torch.rand(2, 9, 11, generator=g) + 0.05], 1)
    gn = GNET(ch_in=2)
    gn.gnet = torch.nn.Identity()
    new = gn(d_output, ref_gmm)
    gout = torch.randn(new.shape, generator=g)
    (new * gout).sum().backward()
    # learned convex upsampling
    depth = torch.rand(2, 2, 6, 7, generator=g) * 3
    mask = torch.randn(2, 9 * 16, 6, 7, generator=g)
    up = upsample_depth_via_mask(depth, mask, 4)
    ks = {f"k_{b}_{n}": np.asarray(MAGNET.depth_sampling(types.SimpleNamespace(sampling_range=b, n_samples=n)))
          for (b, n) in ((3, 5), (3, 16), (3, 64), (2, 7))}
    np.savez_compressed(os.path.join(HERE, "update_upsample.npz"),
                        d_output=d_output.detach().numpy(), ref_gmm=ref_gmm.numpy(), new_gmm=new.detach().numpy(),
                        grad_out=gout.numpy(), grad_d_output=d_output.grad.numpy(),
                        depth=depth.numpy(), mask=mask.numpy(), up=up.numpy(), **ks)
    print("update / upsample / k_list written")


if __name__ == "__main__":
    main()
