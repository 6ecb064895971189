#!/bin/bash
# One-GPU measurement pass of a round: bench lines first, then (only after the plain runs exited 0) the ncu launch list
# of the default bench command and the full captures of the dominant kernels.
# Usage (on the GPU box): bash tools/final_measure.sh <tag>       -> gpurun_out/<tag>_*
set -u
T=${1:-rX}
O=gpurun_out
python bench.py > $O/${T}_bench_c4_n1.json 2> $O/${T}_bench_c4_n1.err || { echo "bench failed"; tail -5 $O/${T}_bench_c4_n1.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py --workload c2 > $O/${T}_bench_c2_n1.json 2> $O/${T}_bench_c2_n1.err
python bench.py --workload c1 --steps 50 > $O/${T}_bench_c1_n1.json 2> $O/${T}_bench_c1_n1.err
python bench.py --workload c5 --steps 20 > $O/${T}_bench_c5_n1.json 2> $O/${T}_bench_c5_n1.err
python bench.py --workload c3 --steps 5 --warmup 3 > $O/${T}_bench_c3_n1.json 2> $O/${T}_bench_c3_n1.err
python tools/c5_parts.py > $O/${T}_c5_parts.txt 2>&1
# ncu: the launch list of the DEFAULT bench command (its own run above exited 0), then one full capture per kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches_bench_c4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_topk -s 3 -c 1 -f -o $O/prof_score_${T} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_score.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_topk -s 3 -c 1 -f -o $O/prof_score_c2_${T} \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_score_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_topk -s 3 -c 1 -f -o $O/prof_score_c1_${T} \
    python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_score_c1.log 2>&1
python tools/bench_score_shapes.py > $O/${T}_score_shapes.txt 2>&1
HYPRET_STATS=1 python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep "hypret stats" | tail -2 > $O/${T}_score_c1_stats.txt
ncu --set full --clock-control none --import-source on -k flash_tile_kernel -s 4 -c 1 -f -o $O/prof_flash_fwd_${T} \
    python tools/c5_parts.py > $O/${T}_ncu_flash_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k flash_tile_kernel -s 30 -c 1 -f -o $O/prof_flash_bwd_${T} \
    python tools/c5_parts.py > $O/${T}_ncu_flash_bwd.log 2>&1
for f in c4_n1 reference c2_n1 c1_n1 c5_n1 c3_n1; do echo "== $f"; cut -c1-400 $O/${T}_bench_$f.json; done
