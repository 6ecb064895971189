#!/bin/bash
# One-GPU measurement pass of a round: bench lines, then (only after the plain runs exited 0) the ncu launch list and
# the full captures of the dominant kernels.  Usage (on the GPU box): bash tools/final_measure.sh <tag>
set -u
T=${1:-rX}
O=gpurun_out
python bench.py > $O/${T}_bench_c2_n1.json 2> $O/${T}_bench_c2_n1.err || { echo "bench failed"; tail -5 $O/${T}_bench_c2_n1.err; exit 1; }
python bench.py --workload c1 --steps 50 > $O/${T}_bench_c1_n1.json 2> $O/${T}_bench_c1_n1.err
python bench.py --workload c3 --steps 5 --warmup 3 > $O/${T}_bench_c3_n1.json 2> $O/${T}_bench_c3_n1.err
python bench.py --workload c5 --steps 20 > $O/${T}_bench_c5_n1.json 2> $O/${T}_bench_c5_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python tools/bench_train_hyp.py > $O/${T}_train_hyp_c5.json 2> $O/${T}_train_hyp_c5.err
python tools/c5_parts.py > $O/${T}_c5_parts.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_topk -s 3 -c 1 -o $O/prof_score_${T} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_score.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_dist -s 2 -c 1 -o $O/prof_gram_${T} \
    python tools/bench_train_hyp.py > $O/${T}_ncu_gram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pairdist_bwd_fused -s 2 -c 1 -o $O/prof_bwd_${T} \
    python tools/bench_train_hyp.py > $O/${T}_ncu_bwd.log 2>&1
for f in c2_n1 c1_n1 c3_n1 c5_n1 reference; do echo "== $f"; cut -c1-600 $O/${T}_bench_$f.json; done
cat $O/${T}_train_hyp_c5.json | head -3
