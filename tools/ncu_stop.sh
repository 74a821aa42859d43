#!/bin/bash
# ncu capture of k_resident in the stop modes (minimise with a criterion that never fires):
# tools/stopmode.py 592 launches minimise(), kick steps, timeSteps x2, then minimise-mode x2
set -e
OUT=gpurun_out/r2_ncu_resident_stop
ncu --set full --clock-control none --import-source on -k regex:k_resident -s 3 -c 1 -o $OUT -f \
    python tools/stopmode.py 592 > gpurun_out/r2_ncu_stop.log 2>&1 || true
ncu -i $OUT.ncu-rep --page raw --csv > $OUT.csv 2>/dev/null || true
tail -3 gpurun_out/r2_ncu_stop.log
