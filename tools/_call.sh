mkdir -p gpurun_out
timeout 200 python tools/thermal_bench.py > gpurun_out/thermal_bench.log 2>&1
cat gpurun_out/thermal_bench.log
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log)
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
