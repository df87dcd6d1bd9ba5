#!/bin/bash
# 2-GPU session f: dist invariance again (masks / re-sort / halo), strong-scaling bench at N=2 with the weak leg and progress log
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2f_dist.log 2>&1; echo "dist rc=$?" >> gpurun_out/r2f_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --weak > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2f_bench_n2.err
tail -8 gpurun_out/r2f_dist.log; grep "bench rank" gpurun_out/r2f_bench_n2.err | tail -30; tail -3 gpurun_out/r2f_bench_n2.err; head -c 1200 gpurun_out/r2f_bench_n2.json
