mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resident_thermal -o gpurun_out/r1h_thermal python tools/profile_thermal.py > gpurun_out/ncu_thermal.log 2>&1
tail -5 gpurun_out/ncu_thermal.log
