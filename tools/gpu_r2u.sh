#!/bin/bash
# session u (1 GPU): cheap env sweeps with the final code: re-sort schedule, pose kernel frames per CTA
mkdir -p gpurun_out
: > gpurun_out/r2u.log
run() { echo "## $*" >> gpurun_out/r2u.log; env "$@" SWEEP_PARTS=2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 10000 2>&1 | grep "^{" >> gpurun_out/r2u.log; }
run BODYFIT_RESORT=6,16,36
run BODYFIT_RESORT=4,10,20,40
run BODYFIT_RESORT=5,12,24,48
run BODYFIT_RESORT=8,24,50
run BODYFIT_RESORT=3,8,16,32,64
run BODYFIT_POSE_WPB=1
run BODYFIT_POSE_WPB=3
run BODYFIT_POSE_WPB=4
run BODYFIT_RESORT=6,16,36
cat gpurun_out/r2u.log
