#!/bin/bash
set -u
mkdir -p gpurun_out
CONFIGS="5 2 3 4" STEPS=5 tools/gpu_r2.sh r2t ab "fin:X=1"
tools/gpu_r2.sh r2t pytest
