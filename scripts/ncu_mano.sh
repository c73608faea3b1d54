#!/bin/bash
# Usage (under gpurun): bash scripts/ncu_mano.sh <tag> [B]  -> gpurun_out/<tag>.ncu-rep: one full capture of each mano_* kernel at batch B
B=${2:-8192}
ncu --set full --clock-control none --import-source on -k regex:'mano_' -s 60 -c 12 -o gpurun_out/$1 python scripts/mano_stage_times.py $B > gpurun_out/$1.log 2>&1
tail -2 gpurun_out/$1.log | cut -c1-200
