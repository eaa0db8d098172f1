#!/bin/bash
# 1-GPU evidence visit for the tensor-core kernel: cfg3 bench, sweep, shipped operating point, training step, ncu
# captures (fused cfg2 / cfg3, drop-in cfg2), launch list of a bench step.  usage: gpu_visit_c.sh <tag>
set -u
TAG=${1:-vc}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
echo "== bench cfg3"; timeout 600 python bench.py --config cfg3 --no-cpu-baseline --steps 100 > "$OUT/bench_cfg3.json" 2> "$OUT/bench_cfg3.err"; echo "rc=$?"; python -c "
import json;d=json.loads(open('$OUT/bench_cfg3.json').read().strip().splitlines()[-1]);r=d['roofline'];print('cfg3 value %.0f ms/step %.3f kern %.4f frac %.3f e2e %.0f'%(d['value'],d['ms_per_step'],r['kernel_ms'],r['frac'],d['e2e']['value']))"
echo "== ship"; timeout 600 python scripts/ship_point.py "$OUT/ship.md" > "$OUT/ship.log" 2>&1; echo "rc=$?"; cat "$OUT/ship.md"
echo "== train head"; for f in "" "--unfused-loss"; do timeout 600 python examples/train_head.py --steps 10 $f 2>&1 | tail -1; done | tee "$OUT/train.txt"
echo "== sweep"; timeout 900 python scripts/sweep.py "$OUT/sweep_n1.md" > "$OUT/sweep.log" 2>&1; echo "rc=$?"; cat "$OUT/sweep_n1.md"
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_mma -s 3 -c 1 -f -o $OUT/mma_cfg2 python scripts/kbench.py cfg2 mma 2 > $OUT/ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_mma -s 3 -c 1 -f -o $OUT/mma_cfg3 python scripts/kbench.py cfg3 mma 2 > $OUT/ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_mma -s 3 -c 1 -f -o $OUT/mma_cfg2_volume python scripts/kbench.py cfg2 mma 2 volume > $OUT/ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file "$OUT/launches.csv" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gnet > "$OUT/ncu_launches.log" 2>&1; echo "ncu-list rc=$?"
