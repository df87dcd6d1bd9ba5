#!/bin/bash
# N=8 bench, main leg only (budget): value + e2e
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 --no-extras > gpurun_out/r2p_bench_n8_noextras.json 2> gpurun_out/r2p_bench_n8_noextras.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_n8_noextras.json').read().strip().splitlines()[-1])
e=d['e2e']
print('N=8 value %.0f ms %.2f e2e %.0f e2e_ms %.2f' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step']), e.get('host_link'), d['config']['streams'][:50])
PY
