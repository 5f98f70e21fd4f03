#!/bin/bash
# Sweep of the deferred split-K plan of the FFN w_2 GEMMs (AVSR_B200_W2SPLIT = "0" | "<pair tile>:<k slices>").
mkdir -p gpurun_out
: > gpurun_out/w2split.log
for cfg in 0 256:2 256:3 384:4 0 256:3; do
  echo "== W2SPLIT=$cfg" >> gpurun_out/w2split.log
  AVSR_B200_W2SPLIT=$cfg timeout 300 python bench.py --no-cpu --steps 40 --warmup 5 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'])" >> gpurun_out/w2split.log 2>&1
done
cat gpurun_out/w2split.log
