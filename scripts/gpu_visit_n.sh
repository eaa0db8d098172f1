#!/bin/bash
# N-GPU visit (gpurun --gpus N): bench at N ranks with per-rank diagnostics, head training at N.  usage: gpu_visit_n.sh <tag> <N>
set -u
TAG=${1:-vn}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1; nproc >> "$OUT/gpu.txt"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
echo "bench n=$N rc=$?"
python - "$OUT/bench_n$N.json" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); r=d["roofline"]; p=d["repeats"]
    print("  N=%d value %.0f f/s ms/step %.4f (median of 20 regions %.4f, min %.4f max %.4f, eager %.4f) kern %.4f (max over ranks %.4f) e2e %.0f (%.3f ms) gnet %s aff %s" % (
        d["n_gpus"], d["value"], d["ms_per_step"], p["median_ms_per_step"], p["min_ms_per_step"], p["max_ms_per_step"], p["eager_ms_per_step"], r["kernel_ms"], r["kernel_ms_max_over_ranks"],
        d["e2e"]["value"], d["e2e"]["ms_per_step"], d["with_gnet"] and round(d["with_gnet"]["value"]), p["cpu_affinity"]))
    print("  per rank:", d.get("per_rank"))
except Exception as e: print("  no json", e)
PY
echo "== train head N=$N"; timeout 600 $TR --nproc-per-node $N --master-port $((29650+N)) examples/train_head.py --global-batch $((4*N)) --steps 20 2>&1 | tail -1 | tee "$OUT/train_n$N.txt"
