#!/bin/bash
# development visit of the tensor-core kernel: diagnostics with the debug build, parity subset, timing, one ncu capture
O=gpurun_out/$1; mkdir -p $O
timeout 400 env MAGNET_B200_LIB=$PWD/magnet_b200/libmagnet_b200_mmadbg.so python scripts/debug_mma.py > $O/debug.txt 2>&1
echo "debug exit $?" >> $O/debug.txt
grep "^\[\|G vs\|exit\|table\|split16" $O/debug.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "mma or seeded or behind or non_finite" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log | cut -c1-200
for a in "cfg2 mma 20" "cfg2 mma 20 volume" "cfg3 mma 20"; do timeout 200 python scripts/kbench.py $a 2>&1 | tail -1; done | tee $O/kbench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_mma -s 3 -c 1 -f -o $O/cost_mma python scripts/kbench.py cfg2 mma 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -1 $O/ncu.log
