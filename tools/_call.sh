mkdir -p gpurun_out
for b in 1 0; do echo "BULK=$b"; FQSB_S2_BULK=$b timeout 100 python tools/line2d.py 2>&1 | grep -E "stream_2d|minimise|rror"; done > gpurun_out/v2d_bulk.log 2>&1; cat gpurun_out/v2d_bulk.log
(timeout 600 python -m pytest tests -m gpu -x -q -k "2d or Line2d or fullsize or slab or golden" > gpurun_out/pytest2d.log 2>&1; echo rc=$? >> gpurun_out/pytest2d.log)
tail -3 gpurun_out/pytest2d.log
