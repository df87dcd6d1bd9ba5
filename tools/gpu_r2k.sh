#!/bin/bash
# session k (1 GPU): two half-reductions in the masked backward GEMM: full parity suite, per-kernel times, sweep, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log
timeout 300 python tools/kernels_at.py 416 1250 2500 10000 > gpurun_out/r2k_kernels.log 2>&1
BODYFIT_BWD_BN=128 timeout 300 python tools/kernels_at.py 1250 > gpurun_out/r2k_kernels_bn128.log 2>&1
BODYFIT_BWD_BN=256 timeout 300 python tools/kernels_at.py 1250 > gpurun_out/r2k_kernels_bn256.log 2>&1
SWEEP_PARTS=1,2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 10000 > gpurun_out/r2k_sweep.log 2>&1
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
grep -E "passed|failed|FAILED" gpurun_out/r2k_tests.log | tail -5; grep "^{" gpurun_out/r2k_kernels.log gpurun_out/r2k_kernels_bn128.log gpurun_out/r2k_kernels_bn256.log gpurun_out/r2k_sweep.log; tail -2 gpurun_out/r2k_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))
PY
