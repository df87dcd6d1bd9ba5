#!/bin/bash
# evidence session (1 GPU): ncu launch list of a fit, full captures of the top kernels, dense-operator bench, final bench line
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_fit_smplx_10k.csv python tools/profile_step.py --iters 8 > gpurun_out/r2g_launch.log 2>&1
for k in k_frame_loss_bwd k_pose_bwd k_blend_fwd_tc_blk k_blend_bwd_tc; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/r2_$k python tools/profile_step.py --iters 8 > gpurun_out/r2g_ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_skin_frame -s 2 -c 2 -f -o gpurun_out/r2_k_skin_frame python tools/profile_step.py --iters 2 --frames 1024 --dense > gpurun_out/r2g_ncu_skin.log 2>&1
timeout 300 python tools/dense_breakdown.py > gpurun_out/r2g_dense_breakdown.log 2>&1
BODYFIT_SKIN=rows timeout 300 python tools/dense_breakdown.py > gpurun_out/r2g_dense_breakdown_rows.log 2>&1
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2g_dense_breakdown.log; tail -3 gpurun_out/r2g_dense_breakdown_rows.log; tail -3 gpurun_out/r2g_bench.err; head -c 600 gpurun_out/r2g_bench.json
timeout 600 python tools/e2e_sweep.py > gpurun_out/r2g_e2e_sweep.log 2>&1; grep "^{" gpurun_out/r2g_e2e_sweep.log
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log; grep -E "passed|failed|FAILED" gpurun_out/r2g_tests.log | tail -5
