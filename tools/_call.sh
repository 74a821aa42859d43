mkdir -p gpurun_out
timeout 100 python tools/line2d.py 2>&1 | grep -E "stream_2d|nopassing|rror" > gpurun_out/v2d_w7.log; cat gpurun_out/v2d_w7.log
(timeout 600 python -m pytest tests -m gpu -x -q -k "2d or Line2d or nopassing or Nopassing or slab" > gpurun_out/pytest2d.log 2>&1; echo rc=$? >> gpurun_out/pytest2d.log)
tail -3 gpurun_out/pytest2d.log
