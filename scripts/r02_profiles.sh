#!/bin/bash
# ncu evidence of round 2 (run under gpurun, one GPU): the launch list of one forward (direct launches, f16 path) and a
# --set full capture of every kernel kind of a layer.  Summaries are made here, copied to profiles/ in the build container.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 405 -c 180 --csv --log-file gpurun_out/r02_launches_f16.csv \
    python scripts/profile_launches.py 2 f16 > gpurun_out/r02_launches.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_f16.csv > gpurun_out/r02_launches_f16.summary.txt 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_tc2_kernel|ln_kernel|dwconv_bn_silu_kernel|attention_f16_kernel" -s 20 -c 20 \
    -o gpurun_out/r02_layer_kernels python scripts/profile_launches.py 2 f16 > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_ncu_full.log
head -16 gpurun_out/r02_launches_f16.summary.txt
