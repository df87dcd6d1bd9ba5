#!/bin/bash
# session l (1 GPU): fused or two-CTA half-reductions in the masked backward GEMM: full parity suite, per-kernel times, sweep, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
timeout 300 python tools/kernels_at.py 1250 2500 5000 10000 > gpurun_out/r2l_kernels.log 2>&1
SWEEP_PARTS=1,2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 5000 10000 > gpurun_out/r2l_sweep.log 2>&1
timeout 900 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
grep -E "passed|failed|FAILED" gpurun_out/r2l_tests.log | tail -5; grep "^{" gpurun_out/r2l_kernels.log gpurun_out/r2l_sweep.log; tail -2 gpurun_out/r2l_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))
PY
