#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench, tuning A/B, ncu launch list + full capture of the cost kernel.
# Usage (from the build container):  gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh <tag> [tuning libs...]'
set -u
TAG=${1:-r1}; shift || true
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
S="$OUT/summary.txt"
pick() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("  value=%.0f f/s ms/step=%.3f kern_ms=%.4f frac=%.3f e2e=%.0f gnet=%s launches=%s clocks=%s" % (
        d["value"], d["ms_per_step"], r.get("kernel_ms",0), r.get("frac",0), (d.get("e2e") or {}).get("value",0),
        ("%.0f"%d["with_gnet"]["value"]) if d.get("with_gnet") else None, d.get("gpu_launches"), (d.get("clocks") or {}).get("sm_mhz")))
except Exception as e:
    print("  (no json)", e)
PY
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > "$OUT/gpu.txt" 2>&1; nproc >> "$OUT/gpu.txt"
echo "== smoke" | tee -a "$S"
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$S"
tail -4 "$OUT/smoke.log" | tee -a "$S"
echo "== pytest gpu" | tee -a "$S"
timeout 1500 python -m pytest tests -m gpu -q -s > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$S"
tail -12 "$OUT/pytest_gpu.log" | cut -c1-300 | tee -a "$S"
echo "== bench default" | tee -a "$S"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$S"; pick "$OUT/bench.json" | tee -a "$S"
timeout 600 python bench.py --config cfg3 --no-cpu-baseline --steps 100 > "$OUT/bench_cfg3.json" 2>> "$OUT/bench.err"; echo "cfg3" | tee -a "$S"; pick "$OUT/bench_cfg3.json" | tee -a "$S"
for v in cells_noreuse window direct; do
  timeout 600 python bench.py --variant $v --no-cpu-baseline --no-gnet --steps 20 --warmup 3 > "$OUT/bench_$v.json" 2>> "$OUT/bench.err"; echo "variant $v" | tee -a "$S"; pick "$OUT/bench_$v.json" | tee -a "$S"
done
timeout 600 python scripts/sweep.py "$OUT/sweep.md" > "$OUT/sweep.log" 2>&1; echo "sweep rc=$?" | tee -a "$S"
for lib in "$@"; do
  MAGNET_B200_LIB=$PWD/magnet_b200/libmagnet_b200_$lib.so timeout 600 python bench.py --no-cpu-baseline --no-gnet --steps 50 --warmup 5 > "$OUT/bench_$lib.json" 2>> "$OUT/bench.err"
  echo "tuning $lib" | tee -a "$S"; pick "$OUT/bench_$lib.json" | tee -a "$S"
done
echo "== ncu launch list" | tee -a "$S"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gnet > "$OUT/ncu_launches.log" 2>&1; echo "ncu-list rc=$?" | tee -a "$S"
echo "== ncu full (cost kernel)" | tee -a "$S"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:cost_cells -s 9 -c 1 -f -o "$OUT/cost_cells" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gnet > "$OUT/ncu_full.log" 2>&1; echo "ncu-full rc=$?" | tee -a "$S"
tail -3 "$OUT/bench.err"
