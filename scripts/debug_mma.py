"""Diagnostics for the tensor-core kernel (MAGNET_VARIANT_MMA): the split16 buffer, the accumulator rows of CTA (0,0)
against an fp64 all-pairs product (debug build: MAGNET_B200_LIB=magnet_b200/libmagnet_b200_mmadbg.so), and the volume
against the direct kernel with the differences broken down by plane / tile position."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import magnet_b200
from magnet_b200 import _lib, ops
from magnet_b200.synthetic import make_inputs

L = _lib.lib()
DEBUG = hasattr(L, "magnet_mma_debug_buffer")
print("debug build:", DEBUG, flush=True)


def check_split(x, gmm=None):
    buf = ops.repack_split16(x, gmm)
    N, Cc, H, W = x.shape
    hdr = buf[:16].view(torch.float32).cpu()
    s, inv = float(hdr[0]), float(hdr[1])
    planes = buf[256:256 + N * 2 * H * W * 64 * 2].view(torch.float16).view(N, 2, H, W, 64).float()
    rec = (planes[:, 0] + planes[:, 1]) * inv
    want = x.permute(0, 2, 3, 1)
    err = (rec - want).abs().max() / want.abs().max()
    print(f"  split16: scale {s:g} inv {inv:g} s*inv {s*inv:g} absmax {float(x.abs().max()):.4f} scaled max {float(x.abs().max())*s:.1f} "
          f"reconstruction max err / max {float(err):.3g}")
    if gmm is not None:
        meta = buf[256 + N * H * W * 256:].view(torch.float32).view(N, H, W + 1, 4)
        pad = torch.nn.functional.pad(gmm, (1, 1))           # (N,2,H,W+2): zeros left and right
        ok = (torch.equal(meta[..., 0], pad[:, 0, :, :-1]) and torch.equal(meta[..., 1], pad[:, 1, :, :-1])
              and torch.equal(meta[..., 2], pad[:, 0, :, 1:]) and torch.equal(meta[..., 3], pad[:, 1, :, 1:]))
        print("  table ok:", bool(ok))
    return buf


def run(tag, inp, open_mask=False, gdump=False):
    if open_mask:
        inp.nghbr_gmms[:, 1] = 1e6
    g = inp.to("cuda")
    plan = magnet_b200.MatchingPlan(g.ref_feat, g.nghbr_feat, g.nghbr_gmms, g.nghbr_poses, inp.is_valid, inp.cam_intrins, thres=5)
    k = inp.k.tolist()
    dbg = None
    if DEBUG and gdump:
        dbg = torch.full((16 + 64 * 256,), float("nan"), device="cuda")
        L.magnet_mma_debug_buffer(C.c_void_p(dbg.data_ptr()))
    try:
        a_t = plan.cost(g.ref_gmms, k, variant=_lib.VARIANT_MMA)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"[{tag}] FAILED: {e}", flush=True)
        raise
    finally:
        if DEBUG:
            L.magnet_mma_debug_buffer(C.c_void_p(0))
    a = a_t.cpu().numpy()
    d = plan.cost(g.ref_gmms, k, variant=_lib.VARIANT_DIRECT).cpu().numpy()
    if dbg is not None:
        h = dbg[:16].cpu().numpy()
        wx0, wy0, nseg, rows, npad, gp, v, mask = [int(t) for t in h[:8]]
        sr, ss = float(h[8]), float(h[9])
        print(f"  G dump: window origin ({wx0},{wy0}) nseg {nseg} rows {rows} npad {npad} gp {gp} view {v} mask {mask} scales {sr:g} {ss:g}")
        G = dbg[16:].view(64, 256).cpu().numpy().astype(np.float64)[:, :nseg * rows * 8]
        B, Cc, H, W = inp.ref_feat.shape
        ref = inp.ref_feat[0].numpy().astype(np.float64)          # (C,H,W)
        src = inp.nghbr_feat[v * B + 0].numpy().astype(np.float64)
        pitch = nseg * 8
        exp = np.zeros_like(G)
        for p in range(64):
            ry, rx = p // 8, p % 8
            if ry >= H or rx >= W:
                continue
            for c in range(nseg * rows * 8):
                y, x = wy0 + c // pitch, wx0 + c % pitch
                if 0 <= y < H and 0 <= x < W:
                    exp[p, c] = ref[:, ry, rx] @ src[:, y, x]
        exp *= sr * ss
        sc = np.abs(exp).max()
        err = np.abs(G - exp)
        print(f"  G vs fp64 all-pairs: max |err| / max |G| = {err.max() / sc:.3g}; nan {int(np.isnan(G).sum())}; exp max {sc:.4g}")
        if err.max() > 1e-4 * sc:
            bad = err > 1e-4 * sc
            print("   bad fraction", bad.mean(), " by row (pixel) head:", np.round(bad.mean(axis=1)[:16], 2).tolist())
            print("   by column head:", np.round(bad.mean(axis=0)[:32], 2).tolist())
            print("   G[0,:8]", np.round(G[0, :8], 1).tolist(), "\n   E[0,:8]", np.round(exp[0, :8], 1).tolist())
            print("   G[1,:8]", np.round(G[1, :8], 1).tolist(), "\n   E[1,:8]", np.round(exp[1, :8], 1).tolist())
            print("   G[9,8:16]", np.round(G[9, 8:16], 1).tolist(), "\n   E[9,8:16]", np.round(exp[9, 8:16], 1).tolist())
    sc = np.abs(d).max()
    bad = np.abs(a - d) > 1e-4 * sc
    B, D, H, W = a.shape
    print(f"[{tag}] shape {a.shape} bad frac {bad.mean():.5f} max rel {np.abs(a-d).max()/sc:.3g} nonfinite {int((~np.isfinite(a)).sum())}", flush=True)
    if bad.any():
        print("   by plane j:", np.round(bad.mean(axis=(0, 2, 3)), 3).tolist())
        print("   by batch b:", np.round(bad.mean(axis=(1, 2, 3)), 3).tolist())
        by = np.zeros((8, 8))
        ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        pb = bad.mean(axis=(0, 1))
        for r in range(8):
            for c in range(8):
                m = (ys % 8 == r) & (xs % 8 == c)
                by[r, c] = pb[m].mean() if m.any() else 0
        print("   by tile position (row%8 x col%8):"); print(np.round(by, 3))
        for (b, j, y, x) in np.argwhere(bad)[:6]:
            print(f"   sample b={b} j={j} y={y} x={x}: mma {a[b,j,y,x]:.5f} direct {d[b,j,y,x]:.5f}")
    return bad.mean()


def ident(inp):
    inp.nghbr_poses.zero_()
    for i in range(4):
        inp.nghbr_poses[:, :, i, i] = 1.0
    return inp


torch.manual_seed(0)
x = torch.randn(3, 64, 16, 32, device="cuda") * 3.0
gm = torch.rand(3, 2, 16, 32, device="cuda")
check_split(x, gm)
check_split(x * 1e-3)
run("identity V1 open", ident(make_inputs(B=1, V=1, D=64, H=16, W=32, C=64, seed=1, depth="smooth")), open_mask=True, gdump=True)
run("identity V1 mask", ident(make_inputs(B=1, V=1, D=16, H=16, W=32, C=64, seed=1, depth="smooth")))
run("V1 open", make_inputs(B=1, V=1, D=64, H=16, W=32, C=64, seed=1, depth="smooth"), open_mask=True, gdump=True)
run("V1 mask", make_inputs(B=1, V=1, D=16, H=16, W=32, C=64, seed=1, depth="smooth"))
run("V3 D16 smoke-like", make_inputs(B=2, V=3, D=16, H=32, W=48, C=64, seed=3, depth="smooth", invalid=[(1, 1)]))
run("V2 D64", make_inputs(B=1, V=2, D=64, H=24, W=32, C=64, seed=5, depth="smooth"))
run("V1 D5", make_inputs(B=1, V=1, D=5, H=16, W=32, C=64, seed=6, depth="smooth"))
run("V2 D80 ragged", make_inputs(B=1, V=2, D=80, H=21, W=37, C=64, seed=7, depth="smooth"))
run("random depth (slow path)", make_inputs(B=1, V=2, D=16, H=16, W=32, C=64, seed=8, depth="random"))
run("cfg2-like", make_inputs(B=2, V=4, D=64, H=120, W=160, C=64, seed=1, depth="smooth"))
