#!/usr/bin/env python
"""gpurun_out/<visit>/ -> committed evidence under profiles/ (round 2): launch list, one ncu summary per captured kernel
(metrics + hottest source lines), traffic.json.   usage: python scripts/summarize_r2.py gpurun_out/r2_06 r2"""
import collections
import csv
import json
import os
import subprocess
import sys

src, name = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
           'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
           'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
           'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
           'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
           'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
           'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg']

traffic = {"_source": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, ncu --set full (profiles/%s_ncu_*.md)" % name}
try:                                                       # keep the entries of earlier visits (other kernels)
    traffic.update({k: v for k, v in json.load(open(os.path.join(out, "traffic.json"))).items() if k != "_source"})
except Exception:
    pass
for rep, title, key in (("mma_cfg2.ncu-rep", "cost_mma_kernel<GAUSS,CW> at cfg2 (fused sampler, SPLIT16 planes, tcgen05.mma + TMA "
                         "windows) — the kernel bench.py's headline runs", "cfg2:mma"),
                        ("mma_cfg3.ncu-rep", "cost_mma_kernel<GAUSS,CW> at cfg3 (KITTI shape)", "cfg3:mma"),
                        ("mma_cfg2_volume.ncu-rep", "cost_mma_kernel<VOLUME,CW> at cfg2 (drop-in d_volume mode) — the kernel "
                         "est_costvolume_CW runs", "cfg2:mma:dropin"),
                        ("cells_cfg2.ncu-rep", "cost_cells_kernel<64,GAUSS,CW> at cfg2 (fused sampler, TILED32 global gather) — the "
                         "kernel bench.py's headline runs", "cfg2"),
                        ("cells_cfg3.ncu-rep", "cost_cells_kernel<64,GAUSS,CW> at cfg3 (KITTI shape)", "cfg3"),
                        ("tma_cfg2_volume.ncu-rep", "cost_tma_kernel<64,VOLUME,CW> at cfg2 (drop-in d_volume mode, PIXC + TMA window) — the "
                         "kernel est_costvolume_CW runs", "cfg2:dropin"),
                        ("cost_tma.ncu-rep", "cost_tma_kernel<64,GAUSS,CW> at cfg2 (fused sampler, PIXC + TMA window)", "cfg2:tma")):
    path = os.path.join(src, rep)
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u, d = rr[0], rr[1], rr[2]
    val = lambda k: d[h.index(k)] if k in h else "n/a"
    try:
        traffic[key] = (float(val('dram__bytes_read.sum')) + float(val('dram__bytes_write.sum'))) * 1e6
    except ValueError:
        pass
    srccsv = os.path.join("/tmp", rep + ".src.csv")
    with open(srccsv, "w") as f:
        f.write(subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                               capture_output=True, text=True).stdout)
    lines = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_lines.py"), srccsv, "28"], capture_output=True,
                           text=True).stdout
    with open(os.path.join(out, f"{name}_ncu_{rep.split('.')[0]}.md"), "w") as f:
        f.write(f"# {name}: `ncu --set full --clock-control none --import-source on` — {title}\n\n| metric | value | unit |\n|---|---:|---|\n")
        f.write(f"| Kernel Name | {val('Kernel Name')} |  |\n")
        for m in METRICS:
            if m in h:
                f.write(f"| {m} | {d[h.index(m)]} | {u[h.index(m)]} |\n")
        for i, m in enumerate(h):
            if 'issue_stalled' in m and 'per_issue_active' in m:
                try:
                    if float(d[i]) > 0.05:
                        f.write(f"| {m} | {d[i]} | {u[i]} |\n")
                except ValueError:
                    pass
        f.write("\n## hottest source lines (warp-stall samples)\n\n```\n" + lines + "```\n")
    print("wrote", rep)

lc = os.path.join(src, "launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(out, f"{name}_launches.md"), "w") as f:
        f.write(f"# {name}: launches of `bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gnet` under\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60` (cold-cache, serialised: compare SHARES)\n\n"
                "The window covers hot-path steps (repack, camera table, fused-sampler cost kernel, update kernel) — eager and "
                "graph-replayed launches look the same to ncu.\n\n| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:100]}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.2f} | {100 * a[1] / tot:.1f}% |\n")
json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1)
print(traffic)
