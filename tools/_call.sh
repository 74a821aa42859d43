mkdir -p gpurun_out
timeout 100 python tools/line2d.py 2>&1 | grep -E "stream_2d|minimise|rror" > gpurun_out/v2d_halo.log; cat gpurun_out/v2d_halo.log
(timeout 600 python -m pytest tests -m gpu -x -q -k "2d or Line2d or fullsize or slab or golden" > gpurun_out/pytest2d.log 2>&1; echo rc=$? >> gpurun_out/pytest2d.log)
tail -3 gpurun_out/pytest2d.log
