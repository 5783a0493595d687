#!/bin/bash
# Round-end ncu evidence, run on the GPU box through gpurun (one GPU).  The .ncu-rep files stay in
# /tmp on the box (they exceed gpurun's 64 MiB return limit); CSV exports come back in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r1d'
set -u
TAG=${1:-r1x}
OUT=gpurun_out
B="python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu --no-secondary"
$B > $OUT/plain_${TAG}.log 2>&1 || { echo "plain bench failed"; tail -5 $OUT/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv $B > $OUT/ncu1_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hk_small_kernel --launch-skip 16 --launch-count 22 -o /tmp/prof_${TAG} -f $B > $OUT/ncu2_${TAG}.log 2>&1
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > $OUT/full_${TAG}.csv 2>/dev/null
./tools/tune_small > $OUT/plain_${TAG}_obs.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hk_small_kernel --launch-skip 22 --launch-count 20 -o /tmp/prof_${TAG}_obs -f ./tools/tune_small > $OUT/ncu_${TAG}_obs.log 2>&1
ncu -i /tmp/prof_${TAG}_obs.ncu-rep --page raw --csv > $OUT/full_${TAG}_obs.csv 2>/dev/null
python tools/profile_c5.py > $OUT/plain_${TAG}_c5.log 2>&1 && \
ncu --set full --clock-control none -k regex:hk_generic_kernel --launch-count 21 -o /tmp/prof_${TAG}_c5 -f python tools/profile_c5.py > $OUT/ncu_${TAG}_c5.log 2>&1
ncu -i /tmp/prof_${TAG}_c5.ncu-rep --page raw --csv > $OUT/full_${TAG}_c5.csv 2>/dev/null
tail -2 $OUT/plain_${TAG}.log | cut -c1-200; ls -la $OUT
