#!/bin/bash
# end-of-session evidence: full GPU test suite, bench line, ncu launch list, ncu full captures of the two per-frame kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python tools/profile_step.py --iters 4 > gpurun_out/f_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame_loss_bwd -s 4 -c 1 -f -o gpurun_out/r1c_frame python tools/profile_step.py --iters 4 > gpurun_out/f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pose_bwd -s 4 -c 1 -f -o gpurun_out/r1c_pose_fused python tools/profile_step.py --iters 4 > gpurun_out/f_ncu2.log 2>&1
tail -n 3 gpurun_out/f_tests.log; cut -c1-600 gpurun_out/f_bench.json; tail -n 2 gpurun_out/f_bench.err
