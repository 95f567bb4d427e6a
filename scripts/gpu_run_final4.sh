#!/bin/bash
# last evidence run of the round: smoke, parity suite, bench line (both arms) of the final tree
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02g_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1
tail -2 gpurun_out/r02g_pytest.log
cp gpurun_out/parity_gpu.json gpurun_out/r02g_parity.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02g_ref.json 2> gpurun_out/r02g_ref.err
echo "ref rc=$?"
