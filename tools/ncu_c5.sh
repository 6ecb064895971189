#!/bin/bash
# ncu capture of the flash train_hyp kernels (one GPU). usage: tools/ncu_c5.sh <tag>
# tools/c5_parts.py launches flash_tile_kernel 23 times forward (lse), then 46 times backward (dA, dP)
set -e
tag=${1:-r2}
python tools/c5_parts.py > gpurun_out/${tag}_c5_parts.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k flash_tile_kernel -s 4 -c 1 -o gpurun_out/prof_flash_fwd_${tag} -f python tools/c5_parts.py > gpurun_out/${tag}_ncu_flash_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k flash_tile_kernel -s 30 -c 1 -o gpurun_out/prof_flash_bwd_${tag} -f python tools/c5_parts.py > gpurun_out/${tag}_ncu_flash_bwd.log 2>&1
