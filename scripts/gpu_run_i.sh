#!/bin/bash
mkdir -p gpurun_out
tests/cuda/_build/gram_tc_check | tail -8
python scripts/ab_c3.py 2048 > gpurun_out/r02i_ab_c3.jsonl 2> gpurun_out/r02i_ab_c3.err; cat gpurun_out/r02i_ab_c3.jsonl; tail -3 gpurun_out/r02i_ab_c3.err
python -m pytest tests/test_gpu_parity.py -q -k "lstsq or shapelets or gram" > gpurun_out/r02i_pytest.log 2>&1
tail -4 gpurun_out/r02i_pytest.log
python scripts/bench_configs.py c3 2048 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02i_c3_launches.csv \
    python scripts/bench_configs.py c3 2048 > gpurun_out/ncu_c3i.log 2>&1
echo "ncu rc=$?"
