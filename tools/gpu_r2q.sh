#!/bin/bash
# two GPUs: the one-process-several-devices CLI over NCCL, and the 2-rank bench (config 5 contig-sharded + region-sharded leg)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2q_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_device.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2q_bench_2gpu.json 2> gpurun_out/r2q_bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2q_bench_2gpu.json").read().strip().splitlines()[-1])
print("N=2 value %.3g e2e %.3g ms %.2f e2e_ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d["config"]["rows"], d["config"]["rows_checksum"])
r = d.get("region_shard", {})
print("region:", r.get("halo"), r.get("value"), r.get("value_incl_halo_reduce"))
PY
tail -5 gpurun_out/r2q_bench_2gpu.err
