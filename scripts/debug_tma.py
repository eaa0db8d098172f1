"""Diagnostics for the TMA kernel against the direct kernel: where (plane / tile position / view) do they differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import magnet_b200
from magnet_b200 import _lib, ops
from magnet_b200.synthetic import make_inputs

def run(tag, inp, open_mask=False, cw=True):
    if open_mask:
        inp.nghbr_gmms[:, 1] = 1e6
    g = inp.to("cuda")
    plan = magnet_b200.MatchingPlan(g.ref_feat, g.nghbr_feat, g.nghbr_gmms, g.nghbr_poses, inp.is_valid, inp.cam_intrins, thres=5)
    k = inp.k.tolist()
    a = plan.cost(g.ref_gmms, k, variant=_lib.VARIANT_TMA).cpu().numpy()
    d = plan.cost(g.ref_gmms, k, variant=_lib.VARIANT_DIRECT).cpu().numpy()
    sc = np.abs(d).max()
    bad = np.abs(a - d) > 1e-4 * sc
    B, D, H, W = a.shape
    print(f"[{tag}] shape {a.shape} bad frac {bad.mean():.4f} max rel {np.abs(a-d).max()/sc:.3g} nonfinite {int((~np.isfinite(a)).sum())}")
    if bad.any():
        print("   by plane j:", np.round(bad.mean(axis=(0, 2, 3)), 3).tolist())
        print("   by batch b:", np.round(bad.mean(axis=(1, 2, 3)), 3).tolist())
        by = np.zeros((4, 16)); cnt = np.zeros((4, 16))
        ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        pb = bad.mean(axis=(0, 1))
        for r in range(4):
            for c in range(16):
                m = (ys % 4 == r) & (xs % 16 == c)
                by[r, c] = pb[m].mean() if m.any() else 0
        print("   by tile position (row%4 x col%16):"); print(np.round(by, 2))
        idx = np.argwhere(bad)[:6]
        for (b, j, y, x) in idx:
            print(f"   sample b={b} j={j} y={y} x={x}: tma {a[b,j,y,x]:.5f} direct {d[b,j,y,x]:.5f}")
        ratio = a[bad] / np.where(d[bad] == 0, np.nan, d[bad])
        print("   ratio tma/direct quantiles:", np.nanpercentile(ratio, [5, 25, 50, 75, 95]).round(3).tolist(), " tma==0 frac", float((a[bad] == 0).mean()), " direct==0 frac", float((d[bad] == 0).mean()))

def ident(inp):
    inp.nghbr_poses.zero_()
    for i in range(4):
        inp.nghbr_poses[:, :, i, i] = 1.0
    return inp

run("identity V1 open", ident(make_inputs(B=1, V=1, D=16, H=16, W=32, C=64, seed=1, depth="smooth")), open_mask=True)
run("identity V1 mask", ident(make_inputs(B=1, V=1, D=16, H=16, W=32, C=64, seed=1, depth="smooth")))
run("V1 open", make_inputs(B=1, V=1, D=16, H=16, W=32, C=64, seed=1, depth="smooth"), open_mask=True)
run("V1 mask", make_inputs(B=1, V=1, D=16, H=16, W=32, C=64, seed=1, depth="smooth"))
run("V3 D16 smoke-like", make_inputs(B=2, V=3, D=16, H=32, W=48, C=64, seed=3, depth="smooth", invalid=[(1, 1)]))
run("V2 D64", make_inputs(B=1, V=2, D=64, H=24, W=32, C=64, seed=5, depth="smooth"))
run("V1 D5", make_inputs(B=1, V=1, D=5, H=16, W=32, C=64, seed=6, depth="smooth"))
run("C16", make_inputs(B=1, V=2, D=16, H=16, W=32, C=16, seed=7, depth="smooth"))
run("random", make_inputs(B=1, V=1, D=16, H=16, W=32, C=32, seed=8, depth="random"))
