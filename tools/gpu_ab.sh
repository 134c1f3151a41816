#!/bin/bash
# A/B of the decode kernels on one B200 (run under gpurun): value/roofline lines per variant into gpurun_out/.
# usage: tools/gpu_ab.sh "name:ENV=V ENV2=V" ...
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
[ $# -eq 0 ] && set -- "warp4:MMC_WARP_OCC=4" "warp3:MMC_WARP_OCC=3" "warp2:MMC_WARP_OCC=2" "general:MMC_DECODE_PATH=general"
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("$name", "kernel_ms %.3f" % d["device_timed"]["kernel_ms_mean"], "frac %.4f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"], "e2e_ms %.2f" % d["e2e"]["ms_per_step"], "value %.0f" % d["value"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/ab_$name.err").read()[-1500:])
PY
done
