#!/bin/bash
# 2-GPU session: sharded == single-GPU (plain, NVLink halo in a graph, host NCCL halo), then the strong-scaling bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c_topo.log 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2c_dist.log 2>&1; echo "dist rc=$?" >> gpurun_out/r2c_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2c_bench_n2.err
tail -15 gpurun_out/r2c_dist.log; tail -5 gpurun_out/r2c_bench_n2.err; head -c 1500 gpurun_out/r2c_bench_n2.json
