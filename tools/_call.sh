mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log)
tail -4 gpurun_out/pytest.log
timeout 200 python tools/thermal_bench.py 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stream_2d_bulk --launch-skip 6 -c 2 -o gpurun_out/r1k_stream2d_bulk python tools/line2d.py 4096 4096 10 > gpurun_out/ncu_2d.log 2>&1
tail -2 gpurun_out/ncu_2d.log
