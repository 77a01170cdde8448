#!/bin/bash
# ncu --set full of ONE kernel (regex $2) from a python command ($3...), exported to raw/source CSV. Usage:
#   bash tools/gpu_profile_kernel.sh <tag> <kernel-regex> <skip> <count> python tools/microbench.py --only roi
set -u
TAG=$1; KRE=$2; SKIP=$3; CNT=$4; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $CNT -o gpurun_out/prof_$TAG -f "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_$TAG.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_$TAG.source.csv 2>/dev/null
gzip -f gpurun_out/prof_$TAG.source.csv
rm -f gpurun_out/prof_$TAG.ncu-rep
