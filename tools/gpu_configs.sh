#!/bin/bash
# bench lines of the other synthetic configs (run under gpurun)
mkdir -p gpurun_out
for c in 2 3 4; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/cfg$c.json 2> gpurun_out/cfg$c.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/cfg$c.json").read().strip().splitlines()[-1])
    print("config $c kernel_ms %.3f reads/s %.0f calls/s %.3g frac %.4f e2e %.0f deferred %s reads %d" % (d["device_timed"]["kernel_ms_mean"], d["device_timed"]["reads_per_s"], d["device_timed"]["calls_per_s"], d["roofline"]["frac"], d["e2e"]["value"], d["roofline"].get("reads_deferred_to_fallback_kernels"), d["config"]["reads_per_gpu"]))
except Exception as e:
    print("config $c failed", e); print(open("gpurun_out/cfg$c.err").read()[-1500:])
PY
done
