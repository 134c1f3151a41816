#!/bin/bash
set -u
mkdir -p gpurun_out
CONFIGS="5 3" STEPS=5 tools/gpu_r2.sh r2p ab "pf:X=1"
for v in "c12:--chunks 12" "c16:--chunks 16" "c24:--chunks 24"; do
  name=${v%%:*}; a=${v#*:}
  for c in 5 4; do
  timeout 600 python bench.py --config $c --only --steps 5 --warmup 3 --no-cpu-baseline $a > gpurun_out/r2p_${name}_c$c.json 2> gpurun_out/r2p_${name}_c$c.err
  python - gpurun_out/r2p_${name}_c$c.json ${name}_c$c <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "e2e_ms %.2f" % d["e2e"]["ms_per_step"], "e2e %.3g" % d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
  done
done
