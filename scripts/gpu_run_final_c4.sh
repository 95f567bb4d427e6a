#!/bin/bash
# ncu launch list + full capture of the cluster config (one .ncu-rep per call: gpurun_out is limited to 64 MiB)
mkdir -p gpurun_out
python scripts/bench_configs.py c4 256 > gpurun_out/plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_c4_launches.csv \
    python scripts/bench_configs.py c4 256 > gpurun_out/ncu_c4a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_raytrace' -s 12 -c 2 -f -o gpurun_out/r02_c4_full \
    python scripts/bench_configs.py c4 256 > gpurun_out/ncu_c4b.log 2>&1
echo "ncu c4 rc=$?"
