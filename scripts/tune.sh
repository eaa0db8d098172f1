#!/bin/bash
# A/B matrix on the GPU box: tuning libs x resident CTAs per SM, kernel-only timing.
OUT=gpurun_out/${1:-tune}; mkdir -p $OUT; shift
for lib in default "$@"; do
  for ctas in 0 2 3 4; do
    if [ "$lib" = default ]; then unset MAGNET_B200_LIB; else export MAGNET_B200_LIB=$PWD/magnet_b200/libmagnet_b200_$lib.so; fi
    MAGNET_CTAS_PER_SM=$ctas timeout 300 python scripts/kbench.py cfg2 auto 20 2>&1 | tail -1 | tee -a $OUT/tune.txt
  done
done
unset MAGNET_B200_LIB
MAGNET_CTAS_PER_SM=2 timeout 300 python scripts/kbench.py cfg3 auto 20 2>&1 | tail -1 | tee -a $OUT/tune.txt
MAGNET_CTAS_PER_SM=0 timeout 300 python scripts/kbench.py cfg3 auto 20 2>&1 | tail -1 | tee -a $OUT/tune.txt
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log | cut -c1-250
