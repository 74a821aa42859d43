mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_thermal.py tests/test_cpp_host.py -x -q > gpurun_out/pytest_thermal.log 2>&1; echo rc=$? >> gpurun_out/pytest_thermal.log)
timeout 200 python tools/thermal_bench.py > gpurun_out/thermal_bench.log 2>&1
tail -30 gpurun_out/pytest_thermal.log; cat gpurun_out/thermal_bench.log
