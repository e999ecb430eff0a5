#!/bin/sh
# ncu captures of one predict step (run under gpurun). Keeps gpurun_out under the 64 MiB merge limit:
# the all-kernel capture is exported to CSV on the box and the big report is deleted.
K="regex:conv_tc|conv_row|maxpool|prepass|head_quant"
# usage: sh scripts/gpu_profile.sh [tag] [single-kernel indices...]; RSB_PRECISION selects strict (default) / fast
TAG=${1:-r2}; shift 2>/dev/null
IDXS=${*:-"43 54 56 57"}
SKIP=180   # 3 warm-up steps x 60 launches
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s $SKIP -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1
timeout 900 ncu --set full --clock-control none -k "$K" -s $SKIP -c 60 -o /tmp/full_step python scripts/profile_step.py >> gpurun_out/profile_step.log 2>&1
ncu -i /tmp/full_step.ncu-rep --page raw --csv > gpurun_out/full_step_raw_${TAG}.csv 2>> gpurun_out/profile_step.log
for IDX in $IDXS; do   # default: layer4.0.conv2 (pair, N=256), dec1 (pair, N=256), dec3 (pair, N=128), dec4
  timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -s $((SKIP+IDX)) -c 1 -o gpurun_out/prof_${TAG}_k$IDX python scripts/profile_step.py >> gpurun_out/profile_step.log 2>&1
done
grep LAUNCH_ORDER gpurun_out/profile_step.log | head -1 > gpurun_out/launch_order_${TAG}.txt
ls -la gpurun_out
