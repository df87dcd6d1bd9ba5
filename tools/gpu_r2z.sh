#!/bin/bash
# session z (1 GPU): final validation of round 2: full parity suite + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x -p no:cacheprovider > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_tests.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r2z_tests.log | tail -4; tail -2 gpurun_out/r2z_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f roofline %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['e2e'].get('host_link'), d['config']['streams'][:60])
PY
