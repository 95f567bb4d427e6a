#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_configs.py c3 2048 > /dev/null 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02n_c3_launches.csv \
    python scripts/bench_configs.py c3 2048 > gpurun_out/ncu_c3n.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gram_tc -c 1 -f -o gpurun_out/r02n_gram \
    python scripts/bench_configs.py c3 512 > gpurun_out/ncu_c3n2.log 2>&1
echo "ncu full rc=$?"
