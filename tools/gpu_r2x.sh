#!/bin/bash
# session x (1 GPU): L2 prefetch of the keypoint rows A/B; ncu capture + launch list of the final per-frame kernel
mkdir -p gpurun_out
: > gpurun_out/r2x.log
for r in 1 2; do for pfx in 1 0; do
  echo "## prefetch $pfx round $r" >> gpurun_out/r2x.log
  BODYFIT_FRAME_PREFETCH=$pfx timeout 300 python tools/kernels_at.py 1250 10000 2>&1 | grep "^{" | cut -c1-220 >> gpurun_out/r2x.log
  BODYFIT_FRAME_PREFETCH=$pfx SWEEP_PARTS=2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 10000 2>&1 | grep "^{" >> gpurun_out/r2x.log
done; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "fit_trajectory or golden or full_size" > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2x_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches_fit_smplx_10k.csv python tools/profile_step.py --iters 8 > gpurun_out/r2x_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frame_loss_bwd -s 4 -c 1 -f -o gpurun_out/r2f_k_frame_loss_bwd python tools/profile_step.py --iters 8 > gpurun_out/r2x_ncu.log 2>&1
cat gpurun_out/r2x.log | cut -c1-200; grep -E "passed|failed" gpurun_out/r2x_tests.log | tail -2
