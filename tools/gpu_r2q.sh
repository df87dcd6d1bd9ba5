#!/bin/bash
# session q (1 GPU): per-frame kernel launch shapes: 256x5 (baseline), 224x5 (56 regs), 224x6 (40 regs, compact live-vertex buffer)
mkdir -p gpurun_out
for v in 2 1; do
  BODYFIT_FRAME_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "fit_trajectory or golden or full_size or row_sorted or c_abi_host" > gpurun_out/r2q_tests_v$v.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2q_tests_v$v.log
done
for v in 0 1 2; do
  BODYFIT_FRAME_VARIANT=$v timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2q_kernels_v$v.log 2>&1
done
for v in 0 1 2; do
  BODYFIT_FRAME_VARIANT=$v SWEEP_PARTS=2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 10000 > gpurun_out/r2q_sweep_v$v.log 2>&1
done
grep -E "passed|failed|FAILED|Error" gpurun_out/r2q_tests_v*.log | tail -6; grep "^{" gpurun_out/r2q_kernels_v*.log gpurun_out/r2q_sweep_v*.log | cut -c1-330
