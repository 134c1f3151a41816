#!/bin/bash
# round 2, drain A/B (run under gpurun): GPU tests of the streaming read-back, e2e with and without drains per config,
# whole-tool wall clock with and without early rows
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_drain.py tests/test_cli_edges.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
CONFIGS="${CONFIGS:-5 2 3 4}" STEPS=5 tools/gpu_r2.sh r2k ab "drain:X=1" "nodrain:BENCH_NO_DRAIN=1"
tools/cli_e2e.sh 200000 16 2 > gpurun_out/r2k_cli_c2.log 2>&1; tail -12 gpurun_out/r2k_cli_c2.log
MINIMOD_NO_DRAIN=1 tools/cli_e2e.sh 200000 16 2 > gpurun_out/r2k_cli_c2_nodrain.log 2>&1; grep "wall" gpurun_out/r2k_cli_c2_nodrain.log
