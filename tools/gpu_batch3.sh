#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
BODYFIT_BWD_FILL=0 timeout 300 python tools/dense_breakdown.py 2>&1 | head -n 1 > gpurun_out/b3_fill0_$i.log
BODYFIT_BWD_FILL=1 timeout 300 python tools/dense_breakdown.py 2>&1 | head -n 1 > gpurun_out/b3_fill1_$i.log
done
cat gpurun_out/b3_fill0_1.log gpurun_out/b3_fill1_1.log gpurun_out/b3_fill0_2.log gpurun_out/b3_fill1_2.log
