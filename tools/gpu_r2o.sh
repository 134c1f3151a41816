#!/bin/bash
set -u
mkdir -p gpurun_out
CONFIGS="4" STEPS=5 tools/gpu_r2.sh r2o ab "haps3:X=1" "minb4:MMC_STREAM_MINB=4" "minb8:MMC_STREAM_MINB=8"
CONFIGS="5 3" STEPS=5 tools/gpu_r2.sh r2o ab "minb4:MMC_STREAM_MINB=4"
NCU_COUNT=2 tools/gpu_r2.sh r2o ncu 4 "k_decode_stream|k_flat_setup"
timeout 600 python -m pytest tests/test_gpu_synth.py tests/test_gpu_golden.py -m gpu -x -q -k "4 or hap or 5c or 2c" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
