#!/bin/bash
# GPU visit of the tensor-core kernel: diagnostics with the debug build, timing with the production build, ncu capture
O=gpurun_out/$1; mkdir -p $O
timeout 400 env MAGNET_B200_LIB=$PWD/magnet_b200/libmagnet_b200_mmadbg.so python scripts/debug_mma.py > $O/debug.txt 2>&1
echo "debug exit $?" >> $O/debug.txt
grep "^\[\|G vs\|exit" $O/debug.txt
for v in mma cells; do timeout 200 python scripts/kbench.py cfg2 $v 20 >> $O/kbench.txt 2>&1; done
timeout 200 python scripts/kbench.py cfg2 mma 20 volume >> $O/kbench.txt 2>&1
timeout 200 python scripts/kbench.py cfg3 mma 20 >> $O/kbench.txt 2>&1
cat $O/kbench.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cost_mma -s 3 -c 1 -f -o $O/cost_mma python scripts/kbench.py cfg2 mma 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -1 $O/ncu.log
