#!/bin/bash
# ncu launch list + full capture of the lstsq config (C3) on the final tree
mkdir -p gpurun_out
python scripts/bench_configs.py c3 2048 > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_c3_launches.csv \
    python scripts/bench_configs.py c3 2048 > gpurun_out/ncu_c3a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_conv_fwd|k_raytrace_comps|k_gram|k_pinv|k_lstsq|k_raytrace_bwd' -s 24 -c 8 -f -o gpurun_out/r02_c3_full \
    python scripts/bench_configs.py c3 512 > gpurun_out/ncu_c3b.log 2>&1
echo "ncu c3 rc=$?"
