#!/bin/bash
# session r (1 GPU): 224-thread per-frame kernel as default: parity suite, 192-thread A/B, sweep, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x -p no:cacheprovider > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_tests.log
timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2r_kernels.log 2>&1
BODYFIT_FRAME_THREADS=192 timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2r_kernels_192.log 2>&1
BODYFIT_FRAME_THREADS=192 SWEEP_PARTS=2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 10000 > gpurun_out/r2r_sweep_192.log 2>&1
SWEEP_PARTS=1,2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 10000 > gpurun_out/r2r_sweep.log 2>&1
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r2r_tests.log | tail -5; grep "^{" gpurun_out/r2r_kernels.log gpurun_out/r2r_kernels_192.log gpurun_out/r2r_sweep_192.log gpurun_out/r2r_sweep.log | cut -c1-330; tail -2 gpurun_out/r2r_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f roofline %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['e2e'].get('host_link'))
PY
