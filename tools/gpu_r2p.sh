#!/bin/bash
# N-GPU bench (own arm only): usage  gpu_r2p.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r2p_bench_n$N.json 2> gpurun_out/r2p_bench_n$N.err; echo "bench rc=$?" >> gpurun_out/r2p_bench_n$N.err
grep "bench rank 0" gpurun_out/r2p_bench_n$N.err | tail -6; grep -i "failed\|error" gpurun_out/r2p_bench_n$N.err | head -5; tail -2 gpurun_out/r2p_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2p_bench_n$N.json').read().strip().splitlines()[-1])
e=d['e2e']
print('N=$N value %.0f ms %.2f e2e %.0f e2e_ms %.2f' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step']), e.get('host_link'))
print({k: d[k] for k in ('config4','weak') if k in d})
PY
