#!/bin/bash
# Usage (under gpurun): bash scripts/ncu_one.sh <kernel-regex> <tag>   -> gpurun_out/<tag>.ncu-rep (one full capture)
ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -o gpurun_out/$2 python scripts/pcl_stage_times.py > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log | cut -c1-200
