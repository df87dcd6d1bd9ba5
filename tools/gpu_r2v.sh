#!/bin/bash
# session v (1 GPU): per-frame kernel with 160 threads / 8 frames per SM (keypoints from global memory) vs the default
mkdir -p gpurun_out
BODYFIT_FRAME_THREADS=160 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "fit_trajectory or golden or full_size or row_sorted" > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2v_tests.log
: > gpurun_out/r2v.log
for r in 1 2; do for t in 224 160; do
  echo "## threads $t round $r" >> gpurun_out/r2v.log
  BODYFIT_FRAME_THREADS=$t timeout 300 python tools/kernels_at.py 1250 10000 2>&1 | grep "^{" | cut -c1-220 >> gpurun_out/r2v.log
  BODYFIT_FRAME_THREADS=$t SWEEP_PARTS=2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 10000 2>&1 | grep "^{" >> gpurun_out/r2v.log
done; done
grep -E "passed|failed|FAILED|Error" gpurun_out/r2v_tests.log | tail -4; cat gpurun_out/r2v.log
