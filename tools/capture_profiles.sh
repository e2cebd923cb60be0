#!/bin/bash
# Run ON THE GPU BOX (gpurun): ncu evidence of one bench step, written to gpurun_out/ (summarised into profiles/ back in the build container
# with tools/summarize_profiles.py).  Usage: bash tools/capture_profiles.sh <tag>
#   1. launch list of one step with duration / DRAM bytes / tensor-pipe % per launch (cold-cache, serialised: compare shares)
#   2. SpeedOfLight / workload / launch sections for EVERY conv_tc launch of that step (csv)
#   3. `--set full --import-source on` of representative launches: L0..L3 (index 0-3), P3 1x1 128->64 (8), P5 1x1 512->256 (30), FFM 3x3 (63)
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
SKIP=$(python tools/one_step.py --skip-count)
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $OUT/launches_raw_$TAG.csv python tools/one_step.py > $OUT/one_step_$TAG.log 2>&1
timeout 900 ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
    --clock-control none -k regex:conv_tc_kernel -s $SKIP -c 65 --csv --page raw --log-file $OUT/conv_sections_$TAG.csv python tools/one_step.py >> $OUT/one_step_$TAG.log 2>&1
for W in "0 4" "8 1" "30 1" "63 1"; do
  set -- $W
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $((SKIP + $1)) -c $2 -f -o $OUT/conv_full_${TAG}_$1 \
      python tools/one_step.py >> $OUT/one_step_$TAG.log 2>&1
done
ls -la $OUT | tail -12
# back in the build container:  python tools/summarize_profiles.py last-step $TAG gpurun_out/launches_raw_$TAG.csv <kernels per step, last line of one_step log>
