#!/bin/bash
# Run on the GPU box (under gpurun): produces gpurun_out/{launches_rN.csv, kernels_rN.ncu-rep}.
# Usage: bash scripts/profile.sh r1 [samples]
set -u
TAG=${1:-r1}
S=${2:-1024}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --samples $S --no-cpu-baseline --no-e2e --quick"
# (1) every launch of the timed steps with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/launches_$TAG.log 2>&1
# (2) one full capture per kernel family (skip the warm-up step launches)
ncu --set full --clock-control none --import-source on -k regex:'pcl_fwd|pcl_bwd_mid|pcl_bwd_img|mano_skin_fwd|mano_skin_bwd|mano_pose|mano_blend_tc|mano_gfeat_tc' -s 36 -c 16 -o gpurun_out/kernels_$TAG $CMD > gpurun_out/kernels_$TAG.log 2>&1
tail -2 gpurun_out/kernels_$TAG.log | cut -c1-200
