#!/bin/bash
set -u
mkdir -p gpurun_out
CONFIGS="4" STEPS=5 tools/gpu_r2.sh r2u ab "split:X=1" "nosplit:MMC_STREAM_SPLIT=0"
CONFIGS="3 5" STEPS=5 tools/gpu_r2.sh r2u ab "split:MMC_STREAM_SPLIT=1"
timeout 900 python -m pytest tests/test_gpu_synth.py -m gpu -x -q -k "split_blocks or paths" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log
NCU_COUNT=2 tools/gpu_r2.sh r2u ncu 4 "k_decode_stream|k_flat_setup"
