#!/bin/bash
# final evidence session of round 2 (1 GPU): ncu launch list of a fit + full captures of the four kernels of an iteration
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches_fit_smplx_10k.csv python tools/profile_step.py --iters 8 > gpurun_out/r2t_launch.log 2>&1
for k in k_frame_loss_bwd k_pose_bwd k_blend_fwd_tc_blk k_blend_bwd_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r2f_$k python tools/profile_step.py --iters 8 > gpurun_out/r2t_ncu_$k.log 2>&1
done
ls -la gpurun_out/r2f_*; tail -2 gpurun_out/r2t_launch.log
