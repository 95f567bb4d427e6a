#!/bin/bash
# full GPU parity suite + the bench line at N = 1 (both arms) of the current tree
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1
tail -5 gpurun_out/r02k_pytest.log
cp gpurun_out/parity_gpu.json gpurun_out/r02k_parity.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02k_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02k_ref.json 2> gpurun_out/r02k_ref.err
echo "ref rc=$?"
