#!/bin/bash
# Round-2 GPU session (run under gpurun): GPU parity suite, bench lines per config and decode path, ncu captures.
# usage: tools/gpu_r2.sh TAG [pytest] [ab "name:ENV=V ..." ...] ; everything lands in gpurun_out/TAG_*
set -u
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
line() { python - "$1" "$2" <<'PY'
import json, sys
name, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(name, "kernel_ms %.3f" % d["device_timed"]["kernel_ms_mean"], "frac %.4f" % d["roofline"]["frac"], "reads/s %.3g" % d["device_timed"]["reads_per_s"],
          "e2e %.0f" % d["e2e"]["value"], "e2e_ms %.2f" % d["e2e"]["ms_per_step"], "fin_ms %.2f" % d["device_timed"]["finalize_ms"],
          "incl_fin %.3g" % d["value_incl_finalize"],
          "deferred", d["roofline"].get("reads_deferred_to_fallback_kernels"))
except Exception as e:
    print(name, "failed", e); print(open(path.replace(".json", ".err")).read()[-1500:])
PY
}
while [ $# -gt 0 ]; do
  case "$1" in
    pytest) shift; timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log ;;
    ab) shift
      while [ $# -gt 0 ] && [[ "$1" == *:* ]]; do
        v=$1; shift; name=${v%%:*}; envs=${v#*:}
        for c in ${CONFIGS:-2 3 4}; do
          env $envs timeout 900 python bench.py --config $c --only --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${name}_c$c.json 2> gpurun_out/${TAG}_${name}_c$c.err
          line "${name}_c$c" gpurun_out/${TAG}_${name}_c$c.json
        done
      done ;;
    ncu) shift; c=$1; pat=$2; shift 2
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$pat -c ${NCU_COUNT:-3} -o gpurun_out/${TAG}_prof_c$c -f \
        python bench.py --config $c --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c$c.log 2>&1; echo "ncu c$c rc=$?" ;;
    launches) shift; c=$1; shift
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c$c.csv \
        python bench.py --config $c --only --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_c$c.log 2>&1; echo "launches c$c rc=$?" ;;
    full) shift; timeout 1500 python bench.py > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err; echo "full bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_full.err; python - gpurun_out/${TAG}_bench_full.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def show(name, r):
    print(name, "kernel_ms %.3f" % r["device_timed"]["kernel_ms_mean"], "frac %.4f" % r["roofline"]["frac"], "value %.3g" % r["value"], "incl_fin %.3g" % r["value_incl_finalize"],
          "e2e %.3g" % r["e2e"]["value"], "fin_ms %.2f" % r["device_timed"]["finalize_ms"], "cpu", (r.get("cpu_baseline") or {}).get("value"))
show("c5", d)
for k, r in d.get("configs", {}).items(): show(k, r)
PY
      ;;
    *) echo "unknown $1"; shift ;;
  esac
done
