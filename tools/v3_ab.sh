#!/bin/bash
# GPU box: A/B of encoder v3 builds.  usage: tools/v3_ab.sh name1 name2 ...   ("main" = product library, "v2" = product library with AFT_ENCODER=2)
set -u
mkdir -p gpurun_out
for name in "$@"; do
  export AFT_ENCODER=3
  unset AFT_B200_LIB
  if [ "$name" = "v2" ]; then export AFT_ENCODER=2; elif [ "$name" != "main" ]; then export AFT_B200_LIB=$PWD/adafortitran_b200/lib/libaft_b200_$name.so; fi
  timeout 300 python tools/tc_check.py fwd > gpurun_out/ab_$name.fwd 2>&1
  echo "== $name"; tail -3 gpurun_out/ab_$name.fwd | cut -c1-150
  timeout 300 python bench.py --workload forti --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("$name", "est/s %.0f" % d["value"], "stages", {k: round(x,2) for k,x in d["stages_ms_per_step"].items()}, "frac %.4f" % d["roofline"]["frac"])
except Exception as e:
    print("bench $name failed", e); print(open("gpurun_out/ab_$name.err").read()[-600:])
PY
done
