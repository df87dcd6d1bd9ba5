#!/bin/bash
# session j (1 GPU): run-merged TMA boxes: bit-identity tests, per-kernel times, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider -k "sorted or full_size or concurrent or graph or trajectory or golden" > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2j_kernels.log 2>&1
SWEEP_PARTS=1,2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 10000 > gpurun_out/r2j_sweep.log 2>&1
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
grep -E "passed|failed|FAILED" gpurun_out/r2j_tests.log | tail -3; grep "^{" gpurun_out/r2j_kernels.log gpurun_out/r2j_sweep.log; tail -2 gpurun_out/r2j_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), json.dumps(d.get('lbs_dense'))[:1200])
PY
