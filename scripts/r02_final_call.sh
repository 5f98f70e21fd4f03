#!/bin/bash
# Final validation of a round (one GPU): the whole GPU suite, the bench line with extras, the in-situ launch timeline and
# the ncu launch list.  Outputs under gpurun_out/ (copied to profiles/ in the build container).
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py 2> gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_final.json
python -c "import json; d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['us_per_launch'], d['roofline']['frac'], d['clocks'])"
timeout 120 python scripts/timeline_probe.py S2 3 gpurun_out/timeline_final.txt > /dev/null 2>&1
grep "sum of exposed" gpurun_out/timeline_final.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 405 -c 180 --csv --log-file gpurun_out/r02_launches_f16.csv \
    python scripts/profile_launches.py 2 f16 > gpurun_out/r02_launches.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_f16.csv > gpurun_out/r02_launches_f16.summary.txt 2>&1
head -12 gpurun_out/r02_launches_f16.summary.txt
