mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_thermal.py -x -q > gpurun_out/pytest_thermal.log 2>&1; echo rc=$? >> gpurun_out/pytest_thermal.log)
tail -12 gpurun_out/pytest_thermal.log
timeout 200 python tools/thermal_bench.py 2>&1 | tail -2
