#!/bin/bash
# LayerNorm launch-geometry sweep (warps per row x rows per CTA) measured in situ with the launch timeline.
mkdir -p gpurun_out
: > gpurun_out/r02_ln_sweep.txt
for g in 1x8 1x2 2x4 2x2 4x2 4x1; do
  echo "== AVSR_B200_LN=$g" >> gpurun_out/r02_ln_sweep.txt
  AVSR_B200_LN=$g timeout 120 python scripts/timeline_probe.py S2 3 - brief 2>&1 | grep -E "layernorm|sum of exposed|ms/forward" >> gpurun_out/r02_ln_sweep.txt
done
cat gpurun_out/r02_ln_sweep.txt
