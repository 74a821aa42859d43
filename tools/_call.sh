mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log)
timeout 200 python tools/blocked_bench.py > gpurun_out/blocked.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_blocked -c 10 -o gpurun_out/r1g_blocked python tools/profile_blocked.py > gpurun_out/ncu_blocked.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/blocked.log
