#!/bin/bash
# ncu capture of the tensor-core kernel at cfg2 (fused sampler).  usage: gpu_mma_ncu.sh <tag>
O=gpurun_out/$1; mkdir -p $O
timeout 200 python scripts/kbench.py cfg2 mma 20 2>&1 | tail -1 | tee $O/kbench.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cost_mma -s 3 -c 1 -f -o $O/cost_mma python scripts/kbench.py cfg2 mma 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu.log
