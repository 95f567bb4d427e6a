#!/bin/bash
mkdir -p gpurun_out
python scripts/ab_c3.py 2048 > gpurun_out/r02j_ab_c3.jsonl 2> gpurun_out/r02j_ab_c3.err; cat gpurun_out/r02j_ab_c3.jsonl; tail -3 gpurun_out/r02j_ab_c3.err
python -m pytest tests/test_gpu_parity.py -q -x -k "lstsq or shapelets or gram" > gpurun_out/r02j_pytest.log 2>&1
tail -15 gpurun_out/r02j_pytest.log
