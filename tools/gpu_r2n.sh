#!/bin/bash
# round 2: full GPU suite with the CIGAR byte form as default, e2e A/B of the two CIGAR transports, ncu captures (C4: L2 atomics, C3)
set -u
mkdir -p gpurun_out
tools/gpu_r2.sh r2n pytest
for v in "cig8:--cigar-packing 8" "cig32:--cigar-packing 32"; do
  name=${v%%:*}; a=${v#*:}
  for c in 5 3; do
    timeout 600 python bench.py --config $c --only --steps 5 --warmup 3 --no-cpu-baseline $a > gpurun_out/r2n_${name}_c$c.json 2> gpurun_out/r2n_${name}_c$c.err
    python - gpurun_out/r2n_${name}_c$c.json ${name}_c$c <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "e2e_ms %.2f" % d["e2e"]["ms_per_step"], "e2e %.3g" % d["e2e"]["value"], "h2d %.3g" % d["e2e"]["h2d_bytes_per_step"], "kernel_ms %.3f" % d["device_timed"]["kernel_ms_mean"])
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
  done
done
NCU_COUNT=2 tools/gpu_r2.sh r2n ncu 4 "k_decode_stream|k_flat_setup"
NCU_COUNT=2 tools/gpu_r2.sh r2n ncu 3 "k_decode_stream|k_flat_setup"
