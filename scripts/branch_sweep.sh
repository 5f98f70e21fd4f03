#!/bin/bash
# Sweep of the number of concurrent batch-slice branches in the plan graph (AVSR_B200_BRANCHES).
mkdir -p gpurun_out
: > gpurun_out/branches.log
for nb in 1 2 4 2 1; do
  echo "== BRANCHES=$nb" >> gpurun_out/branches.log
  AVSR_B200_BRANCHES=$nb timeout 300 python bench.py --no-cpu --steps 40 --warmup 5 2>gpurun_out/branches.err | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])" >> gpurun_out/branches.log 2>&1
  tail -3 gpurun_out/branches.err >> gpurun_out/branches.log
done
cat gpurun_out/branches.log
