#!/bin/bash
# One GPU call that settles everything written blind at the end of round 1 (run under gpurun; ~4 min):
#   1. first B200 run of the row-8f#1 tests and of the experimental variants (with --runxfail so failures print)
#   2. A/B of the variants on the S2 bench (device-resident frames/s)
#   3. phase trace of the default path (build the trace library BEFORE the call: python scripts/build_trace.py)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_head.py tests/test_zz_gpu_experimental.py -q -m gpu --runxfail -x 2>&1 \
  | tail -40 > gpurun_out/r02_unverified_tests.txt
: > gpurun_out/r02_variants.txt
for v in "X=0" "AVSR_B200_PREB=1" "AVSR_B200_ATTN=x4" "AVSR_B200_PREB=1 AVSR_B200_ATTN=x4"; do
  echo "== $v" >> gpurun_out/r02_variants.txt
  env $v timeout 200 python bench.py --no-cpu --steps 40 --warmup 5 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['us_per_launch'])" \
    >> gpurun_out/r02_variants.txt 2>&1
done
if [ -f auto_avsr_b200/csrc/libavsr_b200_trace.so ]; then
  timeout 200 python scripts/phase_probe.py 2 graph > gpurun_out/r02_phase_probe.txt 2>&1
fi
cat gpurun_out/r02_unverified_tests.txt gpurun_out/r02_variants.txt; head -60 gpurun_out/r02_phase_probe.txt 2>/dev/null
