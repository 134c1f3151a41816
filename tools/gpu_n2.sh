#!/bin/bash
# two GPUs: smoke, goldens, the 2-rank bench (config 5 contig-sharded + region-sharded leg with the NCCL halo reduce)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/n2_smoke.log
timeout 600 python -m pytest tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench_2gpu.json 2> gpurun_out/n2_bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/n2_bench_2gpu.json").read().strip().splitlines()[-1])
print("N=2 value %.3g incl_fin %.3g e2e %.3g ms %.2f e2e_ms %.2f" % (d["value"], d["value_incl_finalize"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d["config"]["rows"], d["config"]["rows_checksum"])
r = d.get("region_shard", {})
print("region:", r.get("halo"), "%.3g %.3g" % (r.get("value", 0), r.get("value_incl_halo_reduce", 0)))
PY
tail -3 gpurun_out/n2_bench_2gpu.err
