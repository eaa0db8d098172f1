#!/usr/bin/env python
"""Turn one GPU-box visit (gpurun_out/<tag>/: launches.csv, cost_cells.ncu-rep, bench*.json) into the committed
evidence under profiles/: <name>_launches.md, <name>_cost_kernel.md, traffic.json.
usage: python scripts/summarize_profiles.py gpurun_out/r1d r1"""
import collections, csv, json, os, subprocess, sys

src, name = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list -------------------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(src, "launches.csv"))) if len(r) > 5]
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
    f.write(f"# {name}: every launch of `bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gnet` under\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n\n"
            "The run contains the warm-up + timed hot-path steps (fused-sampler kernel `<64,1,...>`) and the e2e steps\n"
            "(drop-in kernel `<64,0,...>` that reads d_volume, plus sample_depths).\n\n"
            "| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k[:90]}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / a[0] / 1e3:.2f} | {100 * a[1] / tot:.1f}% |\n")

# ---- full capture of the cost kernel -------------------------------------------------------------------------
rep = os.path.join(src, "cost_cells.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, d = rr[0], rr[1], rr[2]
get = lambda k: (d[h.index(k)], u[h.index(k)]) if k in h else ("n/a", "")
keys = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "sm__cycles_elapsed.max",
]
with open(os.path.join(out, f"{name}_cost_kernel.md"), "w") as f:
    f.write(f"# {name}: `ncu --set full --clock-control none --import-source on -k regex:cost_cells` (one launch, cfg2)\n\n"
            "| metric | value | unit |\n|---|---:|---|\n")
    for k in keys:
        v, un = get(k)
        f.write(f"| {k} | {v} | {un} |\n")
    src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                             capture_output=True, text=True).stdout
    tmp = os.path.join(src, "src.csv")
    open(tmp, "w").write(src_csv)
    lines = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_lines.py"), tmp, "22"],
                           capture_output=True, text=True).stdout
    f.write("\n## hottest source lines (warp-stall samples)\n\n```\n" + "\n".join(l[:170] for l in lines.splitlines()) + "\n```\n")

def mb(k):
    v, un = get(k)
    v = float(v.replace(",", ""))
    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(un, 1)
tj = os.path.join(out, "traffic.json")
traffic = json.load(open(tj)) if os.path.exists(tj) else {}
traffic["cfg2"] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
traffic["_source"] = f"profiles/{name}_cost_kernel.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
json.dump(traffic, open(tj, "w"), indent=1)
for b in ("bench.json", "bench_cfg3.json", "bench_direct.json"):
    p = os.path.join(src, b)
    if os.path.exists(p) and os.path.getsize(p):
        open(os.path.join(out, f"{name}_{b}"), "w").write(open(p).read())
print("wrote", sorted(os.listdir(out)))
