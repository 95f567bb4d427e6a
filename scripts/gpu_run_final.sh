#!/bin/bash
# round-2 evidence run: parity suite, bench line, ncu launch lists + full captures for C2 / C4 / C3
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1
tail -3 gpurun_out/r02_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_ref.json 2> gpurun_out/r02_ref.err
echo "ref rc=$?"
python bench.py --steps 2 --warmup 1 --headline-only > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 40 --csv --log-file gpurun_out/r02_c2_launches.csv \
    python bench.py --steps 2 --warmup 1 --headline-only > gpurun_out/ncu1.log 2>&1
echo "ncu c2 launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_raytrace|k_conv' -s 12 -c 4 -f -o gpurun_out/r02_c2_full \
    python bench.py --steps 2 --warmup 1 --headline-only > gpurun_out/ncu2.log 2>&1
echo "ncu c2 full rc=$?"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/conv_pair_probe scripts/probes/conv_pair_probe.cu 2>/dev/null && /tmp/conv_pair_probe > gpurun_out/r02_conv_pair_probe.txt 2>&1; cat gpurun_out/r02_conv_pair_probe.txt
