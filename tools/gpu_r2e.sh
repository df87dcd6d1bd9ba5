#!/bin/bash
# round-2 session e (1 GPU): full parity suite (masks, re-sort, priorities), A/B of the masks, BK=32 library, timeline, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2e_kernels_mask.log 2>&1
BODYFIT_BLKMASK=0 timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2e_kernels_nomask.log 2>&1
BODYFIT_LIB=$PWD/bodyfitting_b200/libbodyfit_b200_bk32.so timeout 300 python tools/kernels_at.py 1250 10000 > gpurun_out/r2e_kernels_bk32.log 2>&1
SWEEP_PARTS=1,2,4 timeout 600 python tools/sweep_parts.py 1250 10000 > gpurun_out/r2e_sweep.log 2>&1
BODYFIT_SORT=0 SWEEP_PARTS=1,2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 10000 > gpurun_out/r2e_sweep_nosort.log 2>&1
BODYFIT_PARTS=4 timeout 300 python tools/e2e_timeline.py 10000 > gpurun_out/r2e_timeline4.log 2>&1
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
grep -E "passed|failed" gpurun_out/r2e_tests.log | tail -3; grep FAILED gpurun_out/r2e_tests.log; cat gpurun_out/r2e_kernels_mask.log gpurun_out/r2e_kernels_nomask.log gpurun_out/r2e_kernels_bk32.log | grep "^{"; cat gpurun_out/r2e_sweep.log gpurun_out/r2e_sweep_nosort.log | grep "^{"; tail -2 gpurun_out/r2e_timeline4.log; tail -3 gpurun_out/r2e_bench.err; head -c 700 gpurun_out/r2e_bench.json
