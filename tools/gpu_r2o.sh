#!/bin/bash
# session o (1 GPU): bf_model_create path + host-result parts for small batches: parity suite, e2e sweep at shard sizes, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x -p no:cacheprovider > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2o_tests.log
SWEEP_PARTS=1,2,4 SWEEP_E2E=1 timeout 300 python tools/sweep_parts.py 1250 2500 > gpurun_out/r2o_sweep.log 2>&1
timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
grep -E "passed|failed|FAILED|Error|bf_model_create fit" gpurun_out/r2o_tests.log | tail -6; grep "^{" gpurun_out/r2o_sweep.log; tail -2 gpurun_out/r2o_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))
PY
