#!/bin/bash
# GPU box: bf16 parity of encoder v3, then one ncu --set full capture of the kernel (source-level samples)
set -u
mkdir -p gpurun_out
export AFT_ENCODER=3
timeout 300 python tools/tc_check.py fwd > gpurun_out/v3_fwd.txt 2>&1
tail -3 gpurun_out/v3_fwd.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:encoder3_kernel -s 2 -c 1 -f -o gpurun_out/v3_enc python bench.py --workload forti --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/v3_ncu.log 2>&1
tail -2 gpurun_out/v3_ncu.log | cut -c1-200
