#!/bin/bash
# session y (1 GPU): per-frame kernel shape chosen by batch size: full parity suite, sweep over shard sizes, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x -p no:cacheprovider > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2y_tests.log
SWEEP_PARTS=1,2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 2500 5000 10000 > gpurun_out/r2y_sweep.log 2>&1
timeout 900 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r2y_tests.log | tail -4; grep "^{" gpurun_out/r2y_sweep.log; tail -2 gpurun_out/r2y_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f roofline %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['e2e'].get('host_link'))
PY
