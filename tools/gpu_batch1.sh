#!/bin/bash
# GPU batch: parity of the TMA-staged per-frame kernel, A/B kernel times, end-to-end part layouts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "loss_and_gradients or fit_trajectory or golden or edge_cases or concurrent" > gpurun_out/b1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/b1_tests.log
BODYFIT_FRAME_TMA=0 timeout 300 python tools/time_kernels.py > gpurun_out/b1_time_tma0.log 2>&1
BODYFIT_FRAME_TMA=1 timeout 300 python tools/time_kernels.py > gpurun_out/b1_time_tma1.log 2>&1
BODYFIT_FRAME_TMA=0 timeout 300 python tools/time_kernels.py > gpurun_out/b1_time_tma0b.log 2>&1
BODYFIT_FRAME_TMA=1 timeout 300 python tools/time_kernels.py > gpurun_out/b1_time_tma1b.log 2>&1
timeout 900 python tools/e2e_parts.py > gpurun_out/b1_e2e.log 2>&1
tail -3 gpurun_out/b1_tests.log; tail -1 gpurun_out/b1_time_tma*.log; tail -12 gpurun_out/b1_e2e.log
