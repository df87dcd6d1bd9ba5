#!/bin/bash
# session s (1 GPU): interleaved A/B of the per-frame kernel's launch shape (256 / 224 / 224 + full carve-out), three rounds
mkdir -p gpurun_out
: > gpurun_out/r2s_ab.log
for r in 1 2 3; do
  for cfg in "256 0" "224 0" "224 1"; do
    set -- $cfg
    echo "threads $1 carveout $2 round $r" >> gpurun_out/r2s_ab.log
    BODYFIT_FRAME_THREADS=$1 BODYFIT_FRAME_CARVEOUT=$2 SWEEP_PARTS=2 SWEEP_E2E=0 timeout 300 python tools/sweep_parts.py 10000 2>&1 | grep "^{" >> gpurun_out/r2s_ab.log
    BODYFIT_FRAME_THREADS=$1 BODYFIT_FRAME_CARVEOUT=$2 timeout 300 python tools/kernels_at.py 10000 2>&1 | grep "^{" | cut -c1-200 >> gpurun_out/r2s_ab.log
  done
done
cat gpurun_out/r2s_ab.log | cut -c1-260
