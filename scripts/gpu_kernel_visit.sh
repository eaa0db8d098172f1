#!/bin/bash
# debug diagnostics + parity subset + kernel timings + ncu capture of the TMA kernel.  usage: gpu_kernel_visit.sh <tag>
set -u
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 300 python scripts/debug_tma.py > $OUT/staged.txt 2>&1; echo "debug rc=$?"; grep "^\[" $OUT/staged.txt
MAGNET_B200_LIB=$PWD/magnet_b200/libmagnet_b200_noglob.so timeout 300 python scripts/debug_tma.py > $OUT/global.txt 2>&1; echo "debug-global rc=$?"; grep "^\[" $OUT/global.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -6 "$OUT/smoke.log"
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x -k "${2:-golden or known_answers or seeded or behind or fused_sampler or full_size or f_identity or non_finite}" > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?"; tail -15 "$OUT/pytest.log" | cut -c1-300
echo "== kbench"
for a in "cfg2 tma" "cfg2 cells" "cfg3 tma"; do timeout 300 python scripts/kbench.py $a 20 2>&1 | tail -1; done | tee "$OUT/kbench.txt"
MAGNET_B200_LIB=$PWD/magnet_b200/libmagnet_b200_noglob.so timeout 300 python scripts/kbench.py cfg2 tma 20 2>&1 | tail -1 | tee -a "$OUT/kbench.txt"
echo "== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cost_tma -s 3 -c 1 -f -o $OUT/cost_tma python scripts/kbench.py cfg2 tma 2 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log
