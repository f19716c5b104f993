#!/bin/bash
# GPU box: encoder v4 bring-up -- wait-timeout diagnostics on a tiny batch, bf16 parity, short A/B bench.
set -u
mkdir -p gpurun_out
export AFT_ENCODER=4
AFT_DIAG_BATCH=2 timeout 120 python tools/tc_check.py diag > gpurun_out/v4_diag.txt 2>&1
tail -30 gpurun_out/v4_diag.txt
timeout 300 python tools/tc_check.py fwd > gpurun_out/v4_fwd.txt 2>&1
tail -5 gpurun_out/v4_fwd.txt
if grep -q rel_db gpurun_out/v4_fwd.txt; then
  for v in 4 3; do
    AFT_ENCODER=$v timeout 300 python bench.py --workload forti --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/v4_bench_$v.json 2> gpurun_out/v4_bench_$v.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v4_bench_$v.json").read().strip().splitlines()[-1])
    print("enc v$v", "est/s %.0f" % d["value"], "stages", {k: round(x,2) for k,x in d["stages_ms_per_step"].items()}, "frac %.4f" % d["roofline"]["frac"])
except Exception as e:
    print("bench v$v failed", e); print(open("gpurun_out/v4_bench_$v.err").read()[-600:])
PY
  done
fi
