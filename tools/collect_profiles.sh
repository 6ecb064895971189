#!/bin/bash
# Copy the evidence of one measurement pass (tools/final_measure.sh <tag>, run under gpurun) from gpurun_out/ (scratch)
# into profiles/ (tracked): bench JSON lines, the ncu launch list, and text summaries of the .ncu-rep captures
# (the reports themselves are too large to track).  Usage, here in the build container: bash tools/collect_profiles.sh <tag>
set -u
T=${1:?tag}
O=gpurun_out
P=profiles
for f in $O/${T}_bench_*.json $O/${T}_c5_parts.txt $O/${T}_launches_bench_c4.csv $O/${T}_score_shapes.txt $O/${T}_score_c1_stats.txt; do
  [ -s "$f" ] && cp "$f" $P/
done
for k in score score_c2 score_c1 flash_fwd flash_bwd; do
  r=$O/prof_${k}_${T}.ncu-rep
  [ -s "$r" ] || continue
  { echo "# ncu --set full --clock-control none --import-source on, one launch; python tools/ncu_keys.py $r"; python tools/ncu_keys.py "$r"; } > $P/${T}_${k}_ncu.txt 2>&1
done
ls -la $P | grep "${T}_"
