#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_drain.py tests/test_cli_edges.py tests/test_gpu_golden.py tests/test_integration_glue.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
CONFIGS="3 5 2 4" STEPS=5 tools/gpu_r2.sh r2v ab "sd:X=1"
MINIMOD_TRACE=1 MMC_TRACE_CREATE=1 tools/cli_e2e.sh 200000 16 3 > gpurun_out/r2v_cli_c3.log 2>&1; grep -E "wall|identical|trace|mmc_destroy" gpurun_out/r2v_cli_c3.log
tools/cli_e2e.sh 200000 16 2 > gpurun_out/r2v_cli_c2.log 2>&1; grep -E "wall|identical|Real" gpurun_out/r2v_cli_c2.log
