#!/bin/bash
mkdir -p gpurun_out
python scripts/ab_options.py c4 1024 tape=1 tape=0 > gpurun_out/r02c_ab_tape.jsonl 2> gpurun_out/r02c_ab_tape.err
cat gpurun_out/r02c_ab_tape.jsonl; tail -3 gpurun_out/r02c_ab_tape.err
python -m pytest tests/test_gpu_parity.py -q -k "c4_ or cluster or tape" > gpurun_out/r02c_pytest.log 2>&1
tail -4 gpurun_out/r02c_pytest.log
