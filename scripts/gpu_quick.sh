#!/bin/bash
# Quick GPU visit: smoke, the cost-volume parity tests, kernel-only timings.  usage: bash scripts/gpu_quick.sh <tag> [pytest -k expr]
set -u
TAG=${1:-q}; KEXPR=${2:-"golden or known_answers or seeded or behind or fused_sampler or full_size or f_identity"}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -6 "$OUT/smoke.log"
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x -k "$KEXPR" > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?"; tail -15 "$OUT/pytest.log" | cut -c1-250
echo "== kbench"
for a in "cfg2 tma" "cfg2 cells" "cfg3 tma" "cfg3 cells"; do timeout 300 python scripts/kbench.py $a 20 2>&1 | tail -1; done | tee "$OUT/kbench.txt"
