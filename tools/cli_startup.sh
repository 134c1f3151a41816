#!/bin/bash
# Where the start-up time of the tool goes: a small config-2 job, full stderr with its wall-clock stamps (run under gpurun).
D=$(mktemp -d /tmp/mm_start.XXXX)
python - <<PY
import sys
sys.path.insert(0, ".")
from minimod_b200.synth import Synth
s = Synth(2)
s.write_fasta("$D/ref.fa"); s.write_bam("$D/reads.bam", 40000, 3000, threads=16)
PY
for i in 1 2; do
  minimod_b200/bin/minimod freq -c "m[CG]" -m 0.8 -b -t 16 -K 4092 -B 100M -o $D/out.bed $D/ref.fa $D/reads.bam 2> $D/err.txt
  grep -v "processed" $D/err.txt | head -40
done
MMC_DECODE_PATH=split CUDA_MODULE_LOADING=LAZY minimod_b200/bin/minimod freq -c "m[CG]" -m 0.8 -b -t 16 -K 4092 -B 100M -o $D/out.bed $D/ref.fa $D/reads.bam 2>&1 | grep -E "Real time"
CUDA_MODULE_LOADING=EAGER minimod_b200/bin/minimod freq -c "m[CG]" -m 0.8 -b -t 16 -K 4092 -B 100M -o $D/out.bed $D/ref.fa $D/reads.bam 2>&1 | grep -E "Real time"
rm -rf $D
