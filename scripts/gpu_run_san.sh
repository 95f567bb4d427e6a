#!/bin/bash
# usage: gpu_run_san.sh memcheck|racecheck|synccheck
tool=$1
mkdir -p gpurun_out
python scripts/sanitize_step.py > gpurun_out/san_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/san_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_step.py > gpurun_out/r02_sanitizer_${tool}.txt 2>&1
echo "sanitizer step rc=$?"
timeout 300 compute-sanitizer --tool $tool --print-limit 20 tests/cuda/_build/gram_tc_check > gpurun_out/r02_sanitizer_${tool}_gram_probe.txt 2>&1
echo "sanitizer probe rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done|PASS" gpurun_out/r02_sanitizer_${tool}.txt gpurun_out/r02_sanitizer_${tool}_gram_probe.txt
