#!/bin/bash
set -u
mkdir -p gpurun_out
CONFIGS="5 2 3 4" STEPS=5 tools/gpu_r2.sh r2m ab "lag1:X=1"
MMC_TRACE_CREATE=1 tools/cli_startup.sh > gpurun_out/r2m_startup.log 2>&1; grep -E "mmc_create|Real time|Entries" gpurun_out/r2m_startup.log | head -40
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2m_fullsize.log 2>&1; echo "fullsize rc=$?"; tail -5 gpurun_out/r2m_fullsize.log
