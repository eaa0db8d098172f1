#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench, ncu launch list + full capture of the cost kernel.
# Usage (from the build container):  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc >> "$OUT/gpu.txt"
echo "== smoke" | tee -a "$OUT/summary.txt"
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/smoke.log"
echo "== pytest gpu" | tee -a "$OUT/summary.txt"
timeout 1200 python -m pytest tests -m gpu -x -q -s > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -25 "$OUT/pytest_gpu.log"
echo "== bench" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"
timeout 600 python bench.py --config cfg3 --no-cpu-baseline --steps 100 > "$OUT/bench_cfg3.json" 2>> "$OUT/bench.err"
cat "$OUT/bench_cfg3.json"
timeout 600 python bench.py --variant direct --no-cpu-baseline --no-gnet --steps 10 --warmup 3 > "$OUT/bench_direct.json" 2>> "$OUT/bench.err"
cat "$OUT/bench_direct.json"
echo "== ncu launch list" | tee -a "$OUT/summary.txt"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gnet > "$OUT/ncu_launches.log" 2>&1; echo "ncu-list rc=$?" | tee -a "$OUT/summary.txt"
echo "== ncu full (cost kernel)" | tee -a "$OUT/summary.txt"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:cost_cells -s 9 -c 2 -f -o "$OUT/cost_cells" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gnet > "$OUT/ncu_full.log" 2>&1; echo "ncu-full rc=$?" | tee -a "$OUT/summary.txt"
ls -la "$OUT"
