#!/bin/bash
mkdir -p gpurun_out
python scripts/ab_options.py c2 4096 conv_const_taps=1 conv_const_taps=0 > gpurun_out/r02b_ab_conv.jsonl 2> gpurun_out/r02b_ab_conv.err
cat gpurun_out/r02b_ab_conv.jsonl
python -m pytest tests/test_gpu_parity.py -q -k "c4_cluster_positions or uniform_datapath or conv_ or lstsq" > gpurun_out/r02b_pytest.log 2>&1
tail -4 gpurun_out/r02b_pytest.log
python bench.py --steps 20 --warmup 5 --c5-svi-steps 10 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$?"
