#!/bin/bash
# GPU box: driver-workload (AdaFortiTran, B=65536) A/B.  usage: tools/ab_ada.sh "<enc>:<libname|main>" ...
set -u
mkdir -p gpurun_out
for spec in "$@"; do
  v=${spec%%:*}; name=${spec##*:}
  export AFT_ENCODER=$v
  unset AFT_B200_LIB
  if [ "$name" != "main" ]; then export AFT_B200_LIB=$PWD/adafortitran_b200/lib/libaft_b200_$name.so; fi
  timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ada_${v}_$name.json 2> gpurun_out/ada_${v}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ada_${v}_$name.json").read().strip().splitlines()[-1])
    print("enc $v lib $name", "est/s %.0f" % d["value"], "stages", {k: round(x,2) for k,x in d["stages_ms_per_step"].items()}, "frac %.4f" % d["roofline"]["frac"])
except Exception as e:
    print("bench $spec failed", e); print(open("gpurun_out/ada_${v}_$name.err").read()[-600:])
PY
done
