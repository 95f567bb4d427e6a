#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -s -k "full_batch_per_sample" > gpurun_out/r02l_pytest.log 2>&1
grep -E "C2 bs 4096|passed|failed|Error" gpurun_out/r02l_pytest.log | cut -c1-900
