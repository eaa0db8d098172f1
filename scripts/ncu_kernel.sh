#!/bin/bash
# ncu full capture of one launch of the cost kernel via the micro-bench.  usage: bash scripts/ncu_kernel.sh <tag>
OUT=gpurun_out/${1:-ncu}; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cost_cells -s 3 -c 1 -f -o $OUT/cost_cells \
    python scripts/kbench.py cfg2 cells 2 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log
