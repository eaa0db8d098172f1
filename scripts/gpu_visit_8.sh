#!/bin/bash
# 8-GPU visit: bench at N = 1, 2, 4, 8 (weak scaling), stress sweep at N = 2, 4, 8, head training at N = 8.
set -u
TAG=${1:-v8}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1; nproc >> "$OUT/gpu.txt"; nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"
  else timeout 900 $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"; fi
  echo "bench n=$n rc=$?"
  python - "$OUT/bench_n$n.json" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1]); r=d["roofline"]; p=d["repeats"]
    print("  N=%d value %.0f f/s ms/step %.4f (median of 20 regions %.4f, min %.4f max %.4f, eager %.4f) kern %.4f e2e %.0f (%.3f ms, %d regions %.3f..%.3f) gnet %s aff %s" % (
        d["n_gpus"], d["value"], d["ms_per_step"], p["median_ms_per_step"], p["min_ms_per_step"], p["max_ms_per_step"], p["eager_ms_per_step"], r["kernel_ms"],
        d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["regions"], d["e2e"]["min_ms_per_step"], d["e2e"]["max_ms_per_step"], d["with_gnet"] and round(d["with_gnet"]["value"]), p["cpu_affinity"]))
except Exception as e: print("  no json", e)
PY
done
echo "== train head N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29650 examples/train_head.py --global-batch 32 --steps 20 2>&1 | tail -1 | tee "$OUT/train_n8.txt"
timeout 600 $TR --nproc-per-node 8 --master-port 29651 examples/train_head.py --global-batch 32 --steps 20 --unfused-loss 2>&1 | tail -1 | tee -a "$OUT/train_n8.txt"
timeout 600 python examples/train_head.py --steps 20 2>&1 | tail -1 | tee "$OUT/train_n1.txt"
for n in 2 4 8; do
  echo "== sweep N=$n"; timeout 900 $TR --nproc-per-node $n --master-port $((29700+n)) scripts/sweep.py "$OUT/sweep_n$n.md" > "$OUT/sweep_n$n.log" 2>&1; echo "rc=$?"
done
cat "$OUT/sweep_n8.md"
