#!/bin/bash
# Whole-tool wall clock on one B200 box: `minimod freq` (this repo) vs the unmodified reference binary on the same
# synthetic BGZF BAM + FASTA (config 2 shape, N reads).  Everything is inside the time: BAM inflate, packing, H2D,
# kernels, compaction, D2H, text output.  usage: tools/cli_e2e.sh [n_reads] [threads]
set -u
N=${1:-24000}; T=${2:-$(nproc)}
D=$(mktemp -d /tmp/mm_cli_e2e.XXXX)
python - <<PY
import sys
sys.path.insert(0, ".")
from minimod_b200.synth import Synth
s = Synth(2)
s.write_fasta("$D/ref.fa")
first = (s.n_reads - $N) // 2
st = s.write_bam("$D/reads.bam", first, $N, threads=$T)
print("wrote", st["n_reads"], "reads,", st["bases"] // 1000000, "Mbase")
PY
ls -la $D | tail -2
wall() { local t0=$(date +%s.%N); "$@"; local t1=$(date +%s.%N); echo "$(echo "$t1 - $t0" | python -c "print('%.2f' % eval(input()))") s wall"; }
for i in 1 2; do
  echo -n "minimod-b200 freq -b (run $i): "; wall minimod_b200/bin/minimod freq -c "m[CG]" -m 0.8 -b -t $T -K 4092 -B 100M -o $D/mine.bed $D/ref.fa $D/reads.bam 2> $D/mine.err
done
echo -n "minimod_ref   freq -b: "; wall oracle/_ref/minimod_ref freq -c "m[CG]" -m 0.8 -b -t $T -K 4092 -B 100M -o $D/ref.bed $D/ref.fa $D/reads.bam 2> $D/ref.err
grep -E "Data loading time|Data processing time|Data merging time|Data output time|Sorting" $D/ref.err | sed 's/^/  ref: /'
grep -E "time|GPU" $D/mine.err | tail -8 | sed 's/^/  mine: /'
cmp $D/mine.bed $D/ref.bed && echo "outputs byte-identical ($(wc -l < $D/mine.bed) rows)"
rm -rf $D
