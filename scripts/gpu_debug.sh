#!/bin/bash
OUT=gpurun_out/${1:-dbg}; mkdir -p $OUT
timeout 300 python scripts/debug_tma.py > $OUT/staged.txt 2>&1; echo "rc=$?"; cat $OUT/staged.txt | head -150
echo "=========== FORCE GLOBAL"
MAGNET_B200_LIB=$PWD/magnet_b200/libmagnet_b200_noglob.so timeout 300 python scripts/debug_tma.py > $OUT/global.txt 2>&1; echo "rc=$?"; grep "^\[" $OUT/global.txt
