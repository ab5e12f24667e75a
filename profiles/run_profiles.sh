#!/bin/bash
# Reproduces the profiles/ evidence on a B200 box:  gpurun --timeout 1500 -- 'bash profiles/run_profiles.sh r02'
# (never a bench value: numbers printed under ncu are not reported anywhere)
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 9 --warmup 3 --no-e2e --no-cpu --no-ref-config --no-sub --no-parity"
# 1. launch list of the bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
# 2. full captures: the three-step kernel (dominant), the two-step kernel (remainder passes, lattices with boundary cells)
#    and the one-step kernel
ncu --set full --clock-control none --import-source on -k regex:k_stepNx -s 1 -c 1 -o $OUT/${TAG}_k_stepNx3 -f $BENCH > $OUT/${TAG}_k_stepNx3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step2x -s 1 -c 1 -o $OUT/${TAG}_k_step2x -f $BENCH --depth 2 > $OUT/${TAG}_k_step2x.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_pair -s 1 -c 1 -o $OUT/${TAG}_k_step_pair -f $BENCH --single-step > $OUT/${TAG}_k_step_pair.log 2>&1
for k in k_stepNx3 k_step2x k_step_pair; do
  ncu -i $OUT/${TAG}_$k.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_full.csv 2>/dev/null
done
python tools/make_traffic_json.py $OUT/${TAG}_k_stepNx3_full.csv $OUT/${TAG}_k_step2x_full.csv $OUT/${TAG}_k_step_pair_full.csv > $OUT/traffic.json
# 3. BC-bearing workload with two steps per pass: launch list (k_step2x on the clean rows + mask launches on the strips)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_step -s 2 -c 16 --csv \
    --log-file $OUT/${TAG}_karman_fused.csv $BENCH --workload karman > $OUT/${TAG}_karman_fused.log 2>&1
# 4. the cluster kernel on configs 1-3 (100 x 50 periodic, 100 x 100 Couette, 100 x 50 Poiseuille; 2000 steps in one launch)
for c in periodic couette poiseuille; do
  ncu --set full --clock-control none -k regex:k_cluster -c 1 -o $OUT/${TAG}_k_cluster_$c -f python tools/profile_cluster.py $c > $OUT/${TAG}_k_cluster_$c.log 2>&1
  ncu -i $OUT/${TAG}_k_cluster_$c.ncu-rep --page raw --csv > $OUT/${TAG}_k_cluster_${c}_full.csv 2>/dev/null
done
# gpurun brings back at most 64 MiB: keep the report of the headline kernel only (the CSV pages of the others are above)
rm -f $OUT/${TAG}_k_step2x.ncu-rep $OUT/${TAG}_k_step_pair.ncu-rep $OUT/${TAG}_k_cluster_*.ncu-rep
ls -la $OUT
