#!/bin/bash
# round-2 third-session evidence run of the final tree: smoke, parity suite, bench line (both arms), ncu launch list + full capture (C2)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02f_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1
tail -2 gpurun_out/r02f_pytest.log
cp gpurun_out/parity_gpu.json gpurun_out/r02f_parity.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02f_ref.json 2> gpurun_out/r02f_ref.err
echo "ref rc=$?"
python bench.py --steps 2 --warmup 1 --headline-only > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 36 -c 40 --csv --log-file gpurun_out/r02f_c2_launches.csv \
    python bench.py --steps 2 --warmup 1 --headline-only > gpurun_out/ncu1.log 2>&1
echo "ncu c2 launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_raytrace|k_conv' -s 12 -c 4 -f -o gpurun_out/r02f_c2_full \
    python bench.py --steps 2 --warmup 1 --headline-only > gpurun_out/ncu2.log 2>&1
echo "ncu c2 full rc=$?"
