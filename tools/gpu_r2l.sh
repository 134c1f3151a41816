#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "lag2:--drain-lag 2" "lag1:--drain-lag 1" "lag3:--drain-lag 3" "c16:--chunks 16 --drain-lag 3"; do
  name=${v%%:*}; a=${v#*:}
  BENCH_TRACE=1 timeout 600 python bench.py --config ${CFG:-5} --only --steps 5 --warmup 3 --no-cpu-baseline $a > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err
  python - gpurun_out/r2l_$name.json $name <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "e2e_ms %.2f" % d["e2e"]["ms_per_step"], "e2e %.3g" % d["e2e"]["value"])
PY
  grep "e2e trace" gpurun_out/r2l_$name.err | tail -12
done
