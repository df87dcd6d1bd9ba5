#!/bin/bash
# N-GPU bench session: usage  gpu_r2h.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err; echo "bench rc=$?" >> gpurun_out/r2h_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2h_ref_n$N.json 2> gpurun_out/r2h_ref_n$N.err
grep "bench rank 0" gpurun_out/r2h_bench_n$N.err | tail -12; grep -i "failed\|error" gpurun_out/r2h_bench_n$N.err | head -5; tail -2 gpurun_out/r2h_bench_n$N.err; head -c 300 gpurun_out/r2h_bench_n$N.json; echo; cat gpurun_out/r2h_ref_n$N.json | head -c 400
