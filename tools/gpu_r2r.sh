#!/bin/bash
set -u
mkdir -p gpurun_out
D=$(mktemp -d /tmp/mm_md.XXXX)
python - <<PY
import sys
sys.path.insert(0, ".")
from minimod_b200.synth import Synth
s = Synth(5, contigs=(("t2", 300000), ("t10", 200000), ("t1_x", 150000), ("t7", 90000)), coverage=2.0)
s.write_fasta("$D/ref.fa"); s.write_bam("$D/reads.bam"); s.close()
PY
env | grep -i nccl > gpurun_out/r2r_env.txt
minimod_b200/bin/minimod freq -c "m[CG]" -m 0.8 -K 37 $D/ref.fa $D/reads.bam > $D/one.tsv 2> $D/one.err; echo "one rc=$?"
minimod_b200/bin/minimod freq -c "m[CG]" -m 0.8 -K 37 --devices 0,1 $D/ref.fa $D/reads.bam > $D/two.tsv 2> $D/two.err; echo "two rc=$?"
minimod_b200/bin/minimod freq -c "m[CG]" -m 0.8 -K 37 --devices 0,1 --shard-regions $D/ref.fa $D/reads.bam > $D/reg.tsv 2> $D/reg.err; echo "reg rc=$?"
wc -l $D/one.tsv $D/two.tsv $D/reg.tsv
cmp $D/one.tsv $D/two.tsv && echo "contig-sharded identical"
cmp $D/one.tsv $D/reg.tsv && echo "region-sharded identical"
head -c 300 $D/reg.tsv | cat -A | head -8
diff <(cat $D/one.tsv) <(cat $D/reg.tsv) | head -20
tail -5 $D/reg.err
cp $D/reg.err gpurun_out/r2r_reg.err
rm -rf $D
