mkdir -p gpurun_out
for ty in 8 14 28 56 113; do echo "NP TY=$ty"; FQSB_S2_TY_NP=$ty timeout 120 python tools/line2d.py 2>&1 | grep "fixed point"; done > gpurun_out/np_ty.log 2>&1
cat gpurun_out/np_ty.log
