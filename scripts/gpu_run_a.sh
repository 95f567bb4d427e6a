#!/bin/bash
# first GPU pass of the round: parity suite, bench line, C2 launch list + full capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"
python bench.py --steps 2 --warmup 1 --headline-only > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 40 --csv --log-file gpurun_out/r02a_c2_launches.csv \
    python bench.py --steps 2 --warmup 1 --headline-only > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_raytrace|k_conv' -s 12 -c 4 -f -o gpurun_out/r02a_c2_full \
    python bench.py --steps 2 --warmup 1 --headline-only > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
tail -5 gpurun_out/r02a_pytest.log
