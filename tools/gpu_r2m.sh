#!/bin/bash
# session m (1 GPU): persistent per-frame kernel with next-frame prefetch + length-sorted joint slots: parity suite, A/B kernel times, sweep, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2m_kernels.log 2>&1
BODYFIT_FRAME_PERSIST=0 timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2m_kernels_nopersist.log 2>&1
SWEEP_PARTS=1,2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 10000 > gpurun_out/r2m_sweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame_loss_bwd -s 4 -c 1 -f -o gpurun_out/r2m_k_frame_loss_bwd python tools/profile_step.py --iters 4 > gpurun_out/r2m_ncu.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2m_tests.log | tail -5; grep "^{" gpurun_out/r2m_kernels.log gpurun_out/r2m_kernels_nopersist.log gpurun_out/r2m_sweep.log
