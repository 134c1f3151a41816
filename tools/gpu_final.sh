#!/bin/bash
# End-of-round GPU session (run under gpurun).  usage: tools/gpu_final.sh TAG [tests] [bench] [profiles] [cli]
# Everything lands in gpurun_out/TAG_*; copy what is to be judged into profiles/.
set -u
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
for what in "$@"; do
  case $what in
    tests)
      MINIMOD_FULLSIZE=all timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log ;;
    bench)
      timeout 1500 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench rc=$?"
      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "reference arm rc=$?"
      python - gpurun_out/${TAG}_bench_default.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def show(name, r):
    print(name, "kernel_ms %.3f" % r["device_timed"]["kernel_ms_mean"], "frac %.4f" % r["roofline"]["frac"], "value %.3g" % r["value"], "incl_fin %.3g" % r["value_incl_finalize"],
          "e2e %.3g (%.2f ms)" % (r["e2e"]["value"], r["e2e"]["ms_per_step"]), "fin_ms %.2f" % r["device_timed"]["finalize_ms"], "cpu", (r.get("cpu_baseline") or {}).get("value"))
show("c5", d)
for k, r in d.get("configs", {}).items(): show(k, r)
print("clocks", d.get("clocks"))
PY
      ;;
    profiles)
      # (the reports are ~20 MB each and gpurun brings back 64 MB: they are summarised here and only config 5's is kept)
      declare -A READS=([2]=99616 [3]=149425 [4]=29885 [5]=409850)
      for c in 5 2 3 4; do
        rep=gpurun_out/${TAG}_prof_c$c
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_decode_stream|k_flat_setup|k_decode_warp" -c 2 -o $rep -f \
          python bench.py --config $c --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c$c.log 2>&1; echo "ncu c$c rc=$?"
        python tools/ncu_kernel_table.py $rep.ncu-rep > gpurun_out/${TAG}_ncu_kernels_config$c.txt 2>&1
        python tools/ncu_l2_atomics.py $rep.ncu-rep > gpurun_out/${TAG}_ncu_l2_config$c.txt 2>&1
        python tools/ncu_lines_by_file.py $rep.ncu-rep 1.0 "k_decode_stream|k_decode_warp" > gpurun_out/${TAG}_ncu_hot_lines_config$c.txt 2>&1
        TRAFFIC_JSON=gpurun_out/${TAG}_traffic.json python tools/ncu_traffic.py $c:$rep.ncu-rep:${READS[$c]} > /dev/null 2>&1
        [ $c != 5 ] && rm -f $rep.ncu-rep
      done
      for c in 5 3; do
        timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c$c.csv \
          python bench.py --config $c --only --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_c$c.log 2>&1; echo "launches c$c rc=$?"
      done ;;
    cli)
      MINIMOD_TRACE=1 MMC_TRACE_CREATE=1 tools/cli_e2e.sh 200000 16 2 > gpurun_out/${TAG}_cli_c2.log 2>&1; tail -25 gpurun_out/${TAG}_cli_c2.log
      MINIMOD_TRACE=1 tools/cli_e2e.sh 200000 16 3 > gpurun_out/${TAG}_cli_c3.log 2>&1; grep -E "wall|identical|trace" gpurun_out/${TAG}_cli_c3.log
      MMC_TRACE_CREATE=1 tools/cli_startup.sh > gpurun_out/${TAG}_startup.log 2>&1 ;;
  esac
done
