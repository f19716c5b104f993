#!/bin/bash
# GPU box: timing only (results may be garbage: diagnostic builds).  usage: tools/v3_time.sh name...
set -u
export AFT_ENCODER=3
for name in "$@"; do
  unset AFT_B200_LIB
  if [ "$name" != "main" ]; then export AFT_B200_LIB=$PWD/adafortitran_b200/lib/libaft_b200_$name.so; fi
  timeout 300 python bench.py --workload forti --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/t_$name.json 2> gpurun_out/t_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t_$name.json").read().strip().splitlines()[-1])
    print("$name", "encoder ms %.2f" % d["stages_ms_per_step"]["encoder"])
except Exception as e:
    print("bench $name failed", e); print(open("gpurun_out/t_$name.err").read()[-400:])
PY
done
