#!/bin/bash
# evidence run of the tree: GPU parity suite (+ parity report), bench lines at N = 1 (both arms), C3 launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02p_pytest.log 2>&1
tail -3 gpurun_out/r02p_pytest.log
cp gpurun_out/parity_gpu.json gpurun_out/r02p_parity.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02p_ref.json 2> gpurun_out/r02p_ref.err
echo "ref rc=$?"
python scripts/bench_configs.py c3 2048 > gpurun_out/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02p_c3_launches.csv \
    python scripts/bench_configs.py c3 2048 > gpurun_out/ncu_c3a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_conv_fwd|k_raytrace_comps|k_gram|k_pinv|k_lstsq|k_raytrace_bwd' -s 24 -c 8 -f -o gpurun_out/r02p_c3_full \
    python scripts/bench_configs.py c3 512 > gpurun_out/ncu_c3b.log 2>&1
echo "ncu c3 rc=$?"
