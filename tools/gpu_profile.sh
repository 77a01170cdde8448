#!/bin/bash
# ncu --set full captures of the hot-path kernels inside one bench step (1 GPU only), exported to CSV on the box so that
# gpurun_out stays small.  Usage: bash tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
cap() {  # name, kernel regex, launch count
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$2" -c $3 \
      -o gpurun_out/prof_${TAG}_$1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sfod-step 0 --profiler-range > gpurun_out/ncu_full_${TAG}_$1.log 2>&1
  echo "ncu $1 rc=$?"
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_$1.source.csv 2>/dev/null
  gzip -f gpurun_out/prof_${TAG}_$1.source.csv
  sz=$(stat -c %s gpurun_out/prof_${TAG}_$1.ncu-rep)
  if [ "$sz" -gt 6000000 ]; then rm -f gpurun_out/prof_${TAG}_$1.ncu-rep; fi
}
cap det 'roi_align|roi_sep|ema_multi|nms_|bitonic|rpn_|frcnn_|transpose' 41
cap bn 'bn_apply|bn_stats|bn_finalize|bn_fused|bn_frozen' 39
du -sh gpurun_out
