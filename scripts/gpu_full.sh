#!/bin/bash
# Full GPU visit: smoke, ALL gpu tests, kernel timings (both depth modes), bench default + cfg3 + reference arm.
set -u
TAG=${1:-full}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1; nproc >> "$OUT/gpu.txt"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -6 "$OUT/smoke.log"
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?"; tail -25 "$OUT/pytest.log" | cut -c1-250
echo "== kbench"
for a in ${KB:-"cfg2 mma 20 gauss" "cfg2 cells 20 gauss" "cfg2 tma 20 gauss" "cfg2 mma 20 volume" "cfg2 tma 20 volume" "cfg3 mma 20 gauss" "cfg3 cells 20 gauss" "cfg3 mma 20 volume"}; do timeout 300 python scripts/kbench.py $a 2>&1 | tail -1; done | tee "$OUT/kbench.txt"
echo "== bench"; timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; tail -3 "$OUT/bench.err"
python - "$OUT/bench.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("value %.0f f/s ms/step %.3f kern_ms %.4f frac %.3f e2e %.0f (%.2f ms) gnet %s launches %s clocks %s repeats %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["with_gnet"] and round(d["with_gnet"]["value"]), d["gpu_launches"], d["clocks"]["sm_mhz"], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["repeats"].items()}))
    print("refcuda", d.get("reference_cuda")); print("cpu", d.get("cpu_baseline"))
except Exception as e: print("no json", e)
PY
