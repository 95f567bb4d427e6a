#!/bin/bash
mkdir -p gpurun_out
python scripts/ab_options.py c4 1024 tape=1 > gpurun_out/r02e_ab.jsonl 2> gpurun_out/r02e_ab.err
cat gpurun_out/r02e_ab.jsonl; tail -3 gpurun_out/r02e_ab.err
python -m pytest tests/test_gpu_parity.py -q -k "c4_ or cluster or tape or nfw or mass_profiles or hessian" > gpurun_out/r02e_pytest.log 2>&1
tail -6 gpurun_out/r02e_pytest.log
