#!/bin/bash
# Whole-tool wall clock on one B200 box: `minimod freq` (this repo) vs the unmodified reference binary on the same
# synthetic BGZF BAM + FASTA (config 2 shape, N reads).  Everything is inside the time: BAM inflate, packing, H2D,
# kernels, compaction, D2H, text output.  usage: tools/cli_e2e.sh [n_reads] [threads] [config]
set -u
N=${1:-24000}; T=${2:-$(nproc)}; CFG=${3:-2}
D=$(mktemp -d /tmp/mm_cli_e2e.XXXX)
python - <<PY
import sys
sys.path.insert(0, ".")
from minimod_b200.synth import Synth, cli_args
s = Synth($CFG)
s.write_fasta("$D/ref.fa")
n = min($N, s.n_reads)
first = (s.n_reads - n) // 2
st = s.write_bam("$D/reads.bam", first, n, threads=$T)
print("config $CFG: wrote", st["n_reads"], "reads,", st["bases"] // 1000000, "Mbase")
open("$D/args", "w").write(" ".join(cli_args($CFG)))
PY
ARGS=$(cat $D/args)
ls -la $D | tail -2
wall() { local t0=$(date +%s.%N); "$@"; local t1=$(date +%s.%N); echo "$(echo "$t1 - $t0" | python -c "print('%.2f' % eval(input()))") s wall"; }
for i in 1 2; do
  echo -n "minimod-b200 freq (run $i): "; wall minimod_b200/bin/minimod freq $ARGS -t $T -K 4092 -B 100M -o $D/mine.bed $D/ref.fa $D/reads.bam 2> $D/mine.err
done
echo -n "minimod_ref   freq: "; wall oracle/_ref/minimod_ref freq $ARGS -t $T -K 4092 -B 100M -o $D/ref.bed $D/ref.fa $D/reads.bam 2> $D/ref.err
grep -E "Data loading time|Data processing time|Data merging time|Data output time|Sorting" $D/ref.err | sed 's/^/  ref: /'
grep -E "time|GPU|trace|mmc_create|Rows" $D/mine.err | tail -24 | sed 's/^/  mine: /'
if cmp -s $D/mine.bed $D/ref.bed; then echo "outputs byte-identical ($(wc -l < $D/mine.bed) rows)"; else LC_ALL=C sort $D/mine.bed > $D/a; LC_ALL=C sort $D/ref.bed > $D/b; cmp $D/a $D/b && echo "outputs identical after LC_ALL=C sort ($(wc -l < $D/mine.bed) rows; rows sharing (contig,pos) have no defined order in the reference)"; fi
rm -rf $D
