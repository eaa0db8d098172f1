#!/bin/bash
# A/B matrix on the GPU box: tuning libs x resident CTAs per SM, kernel-only timing.
OUT=gpurun_out/${1:-tune}; mkdir -p $OUT; shift
for v in window cells direct; do
  timeout 300 python scripts/kbench.py cfg2 $v 20 2>&1 | tail -1 | tee -a $OUT/tune.txt
done
timeout 300 python scripts/kbench.py cfg3 window 20 2>&1 | tail -1 | tee -a $OUT/tune.txt
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log | cut -c1-250
