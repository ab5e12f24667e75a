#!/bin/bash
# Reproduces the profiles/ evidence on a B200 box:  gpurun --timeout 1500 -- 'bash profiles/run_profiles.sh r01'
# (never a bench value: numbers printed under ncu are not reported anywhere)
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 7 --warmup 3 --no-e2e --no-cpu --no-ref-config"
# 1. launch list of the bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_${TAG}.csv $BENCH > $OUT/launches_${TAG}.log 2>&1
# 2. full capture of the dominant kernel (two steps per launch) and of the one-step kernel
ncu --set full --clock-control none --import-source on -k regex:k_step2x -s 1 -c 1 -o $OUT/prof_2x_${TAG} -f $BENCH > $OUT/prof_2x_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_pair -s 1 -c 1 -o $OUT/prof_pair_${TAG} -f $BENCH --single-step > $OUT/prof_pair_${TAG}.log 2>&1
# 2b. BC-bearing workload with two steps per pass: launch list (k_step2x on the clean rows + mask launches on the strips)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_step -s 2 -c 16 --csv \
    --log-file $OUT/karman_fused_${TAG}.csv $BENCH --workload karman > $OUT/karman_fused_${TAG}.log 2>&1
[ "${QUICK:-0}" = 1 ] && { ls -la $OUT; exit 0; }
# 3. BC-bearing workload, one step per pass: flag mask folded into the kernel vs mask-free kernel + edge kernel
BENCH="$BENCH --single-step"
for mode in mask edge; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
      --clock-control none -k regex:k_step -s 3 -c 6 --csv --log-file $OUT/karman_${mode}_${TAG}.csv \
      $BENCH --workload karman --bc-mode $mode > $OUT/karman_${mode}_${TAG}.log 2>&1
done
ls -la $OUT
