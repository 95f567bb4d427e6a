#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest.log 2>&1
tail -4 gpurun_out/r02d_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
echo "bench rc=$?"
python scripts/bench_configs.py c4 256 > gpurun_out/plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02d_c4_launches.csv \
    python scripts/bench_configs.py c4 256 > gpurun_out/ncu_c4a.log 2>&1
echo "ncu c4 launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_raytrace' -s 12 -c 2 -f -o gpurun_out/r02d_c4_full \
    python scripts/bench_configs.py c4 256 > gpurun_out/ncu_c4b.log 2>&1
echo "ncu c4 full rc=$?"
