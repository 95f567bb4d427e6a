#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_configs.py c3 2048 > gpurun_out/r02g_c3.log 2>&1; tail -1 gpurun_out/r02g_c3.log
python -m pytest tests/test_gpu_parity.py -q -k "lstsq or shapelets" > gpurun_out/r02g_pytest.log 2>&1
tail -4 gpurun_out/r02g_pytest.log
python scripts/bench_configs.py c3 2048 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02g_c3_launches.csv \
    python scripts/bench_configs.py c3 2048 > gpurun_out/ncu_c3g.log 2>&1
echo "ncu rc=$?"
