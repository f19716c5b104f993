#!/bin/bash
# Runs on the GPU box: for every experiment library lib/libaft_b200_<name>.so given on the command line, a short
# bf16 parity check and a short bench line.  Results land in gpurun_out/var_<name>.{fwd,json}.
# usage: tools/variants_bench.sh [--ada] name1 name2 ...   ("main" = the product library)
set -u
WL=forti
if [ "${1:-}" = "--ada" ]; then WL=ada; shift; fi
mkdir -p gpurun_out
for name in "$@"; do
  if [ "$name" = "main" ]; then unset AFT_B200_LIB; else export AFT_B200_LIB=$PWD/adafortitran_b200/lib/libaft_b200_$name.so; fi
  timeout 300 python tools/tc_check.py fwd > gpurun_out/var_$name.fwd 2>&1
  timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  echo "== $name"; cat gpurun_out/var_$name.fwd | tail -3
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var_$name.json").read().strip().splitlines()[-1])
    print("$name", "est/s %.0f" % d["value"], "stages", {k: round(v,2) for k,v in d["stages_ms_per_step"].items()}, "frac %.4f" % d["roofline"]["frac"])
except Exception as e:
    print("$name bench failed", e); print(open("gpurun_out/var_$name.err").read()[-800:])
PY
done
