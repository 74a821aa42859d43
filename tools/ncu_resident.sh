#!/bin/bash
# ncu captures of k_resident: fixed steps (tools/profile_target.py: 3rd k_resident launch) and the
# stop modes (tools/stopmode.py: minimise with a criterion that never fires)
set -e
TAG=${1:-r2c}
ncu --set full --clock-control none --import-source on -k regex:k_resident -s 2 -c 1 \
    -o gpurun_out/${TAG}_ncu_resident_fixed -f python tools/profile_target.py 592 > gpurun_out/${TAG}_ncu_fixed.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:k_resident -s 3 -c 1 \
    -o gpurun_out/${TAG}_ncu_resident_stop -f python tools/stopmode.py 592 > gpurun_out/${TAG}_ncu_stop.log 2>&1 || true
for k in fixed stop; do
  ncu -i gpurun_out/${TAG}_ncu_resident_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_resident_$k.csv 2>/dev/null || true
done
tail -n 2 gpurun_out/${TAG}_ncu_fixed.log; tail -n 2 gpurun_out/${TAG}_ncu_stop.log
