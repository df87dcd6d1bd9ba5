#!/bin/bash
# session w (1 GPU): 160-thread / 8-frames-per-SM per-frame kernel as default: full parity suite, A/B vs 192x6 and 224x5, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x -p no:cacheprovider > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2w_tests.log
: > gpurun_out/r2w.log
for r in 1 2; do for t in 160 192 224; do
  echo "## threads $t round $r" >> gpurun_out/r2w.log
  BODYFIT_FRAME_THREADS=$t timeout 300 python tools/kernels_at.py 2500 10000 2>&1 | grep "^{" | cut -c1-220 >> gpurun_out/r2w.log
  BODYFIT_FRAME_THREADS=$t SWEEP_PARTS=2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 1250 2500 5000 10000 2>&1 | grep "^{" >> gpurun_out/r2w.log
done; done
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r2w_tests.log | tail -4; cat gpurun_out/r2w.log | cut -c1-200; tail -2 gpurun_out/r2w_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f roofline %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['e2e'].get('host_link'))
PY
