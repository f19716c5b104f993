#!/bin/bash
# GPU box: parity + chaos (random delays before every wait) for the experimental encoder kernels, then A/B timing
set -u
mkdir -p gpurun_out
for v in 3 4; do
  export AFT_ENCODER=$v
  unset AFT_B200_LIB
  timeout 300 python tools/tc_check.py fwd 2>&1 | tail -3 | cut -c1-140 | sed "s/^/v$v fwd /"
  for kind in forti ada; do
    AFT_B200_LIB=$PWD/adafortitran_b200/lib/libaft_b200_chaos.so timeout 300 python tools/chaos_check.py 296 $kind 2>&1 | tail -1 | cut -c1-200 | sed "s/^/v$v chaos /"
  done
done | tee gpurun_out/v34_check.txt
unset AFT_B200_LIB
bash tools/v3_ab.sh v2 main 2>&1 | grep "est/s"
bash tools/ab_ada.sh 2:main 3:main 2>&1 | grep "est/s"
