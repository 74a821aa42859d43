mkdir -p gpurun_out
for v in 0 1 2 3; do echo "VARIANT=$v"; FQSB_S2_NP_BULK_VARIANT=$v timeout 120 python tools/line2d.py 2>&1 | grep "fixed point"; done > gpurun_out/np_bulk_var.log 2>&1
cat gpurun_out/np_bulk_var.log
