#!/bin/bash
mkdir -p gpurun_out
timeout 120 tests/cuda/_build/gram_tc_check | tail -12
timeout 300 python scripts/ab_c3.py 2048 > gpurun_out/r02m_ab_c3.jsonl 2> gpurun_out/r02m_ab_c3.err; cat gpurun_out/r02m_ab_c3.jsonl; tail -3 gpurun_out/r02m_ab_c3.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "lstsq or shapelets or gram" > gpurun_out/r02m_pytest.log 2>&1
tail -5 gpurun_out/r02m_pytest.log
