#!/bin/bash
# session i (1 GPU): full suite after the 4-frame skin kernel + kid model, dense breakdown, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
timeout 300 python tools/dense_breakdown.py > gpurun_out/r2i_dense_breakdown.log 2>&1
timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r2i_tests.log | tail -5; grep "kid" gpurun_out/r2i_tests.log | head -8; tail -3 gpurun_out/r2i_dense_breakdown.log; tail -2 gpurun_out/r2i_bench.err; tail -2 gpurun_out/r2i_smoke.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), json.dumps(d['e2e'].get('host_link')), json.dumps(d.get('lbs_dense'))[:1500])
PY
