#!/bin/bash
# A/B on the GPU box, kernel-only timing.  usage: bash scripts/tune.sh <tag> "<env> <variant>" ...
OUT=gpurun_out/${1:-tune}; mkdir -p $OUT; shift
for spec in "$@"; do
  env $spec timeout 300 python scripts/kbench.py cfg2 $(echo $spec | grep -o 'VAR=[a-z]*' | cut -d= -f2) 20 2>&1 | tail -1 | sed "s|$| [$spec]|" | tee -a $OUT/tune.txt
done
