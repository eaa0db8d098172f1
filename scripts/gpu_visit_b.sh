#!/bin/bash
# 1-GPU visit: all GPU tests (verbose for the full-size parity numbers), sanitizers on the smoke path, default bench.
set -u
TAG=${1:-vb}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -s > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?"; tail -4 "$OUT/pytest.log" | cut -c1-250
grep -E "^cfg[23] (fused|drop-in)|install on the real|reference non-finite" "$OUT/pytest.log" | cut -c1-300
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" $OUT/$tool.log | tail -3
done
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "golden or seeded or behind or known_answers or graph or non_finite or fused_upsample or camera_prep or mma" > $OUT/memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/memcheck_tests.log | tail -3
echo "== bench"; timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; tail -2 "$OUT/bench.err"
python - "$OUT/bench.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("value %.0f f/s ms/step %.3f kern_ms %.4f frac %.3f traffic %s e2e %.0f (%.2f ms) gnet %s launches %s clocks %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["traffic"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["with_gnet"] and round(d["with_gnet"]["value"]), d["gpu_launches"], d["clocks"]))
    print("cpu", d.get("cpu_baseline")); print("refcuda kind", d["reference_cuda"]["kind"], d["reference_cuda"]["ms_per_cost_volume"])
except Exception as e: print("no json", e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
