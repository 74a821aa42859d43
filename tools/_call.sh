mkdir -p gpurun_out
timeout 100 python tools/line2d.py 2>&1 | grep -E "nopassing|rror" > gpurun_out/np_ilp.log; cat gpurun_out/np_ilp.log
(timeout 600 python -m pytest tests -m gpu -x -q -k "nopassing or Nopassing or fullsize or slab or golden" > gpurun_out/pytest2d.log 2>&1; echo rc=$? >> gpurun_out/pytest2d.log)
tail -3 gpurun_out/pytest2d.log
