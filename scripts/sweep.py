"""BASELINE.json configs[4]: stress sweep views x hypotheses at the 640x480 (120x160) grid, one GPU:
kernel time, algorithmic HBM GB/s and fraction of the measured roofline per point.
usage: python scripts/sweep.py [out.md]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import magnet_b200
from magnet_b200 import ops
from magnet_b200.synthetic import make_inputs
from bench import algorithmic_bytes, measured_peak

peak, src = measured_peak()
rows = []
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for V in (2, 4, 8):
    for D in (32, 64, 128, 256):
        inp = make_inputs(B=8, V=V, D=D, H=120, W=160, C=64, seed=1, depth="smooth")
        g = inp.to("cuda")
        plan = magnet_b200.MatchingPlan(g.ref_feat, g.nghbr_feat, g.nghbr_gmms, g.nghbr_poses, inp.is_valid,
                                        inp.cam_intrins, thres=5)
        k = ops.k_array(inp.k.tolist())
        out = torch.empty(8, D, 120, 160, device="cuda")
        for _ in range(3):
            plan.cost(g.ref_gmms, k, out=out)
        ts = []
        for _ in range(15):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.cost(g.ref_gmms, k, out=out); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        ab = algorithmic_bytes(8, V, D, 64, 120 * 160, fused=True)
        gbs = ab / (ms * 1e-3) / 1e9
        rows.append((V, D, ms, ab / 1e6, gbs, gbs / peak, 8 / (3 * ms * 1e-3)))
        print(rows[-1], flush=True)
        del plan, g, out
        torch.cuda.empty_cache()
md = ["# stress sweep (BASELINE.json configs[4]): fused cost kernel, B=8, 120x160 grid (640x480), C=64, 1x B200",
      f"peak = {peak:.0f} GB/s ({src}); frames/s = 8 frames / (3 iterations x kernel time), cost kernel only\n",
      "| views | hypotheses | kernel ms | algorithmic MB | GB/s | frac of HBM roofline | frames/s (kernel only) |", "|---:|---:|---:|---:|---:|---:|---:|"]
for r in rows:
    md.append("| %d | %d | %.3f | %.1f | %.0f | %.3f | %.0f |" % r)
open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep.md", "w").write("\n".join(md) + "\n")
