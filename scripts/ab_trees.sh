# development aid: alternate two source trees (this one and a copy of another commit with its built library under _ab_prev/) on the
# same GPU box -- how the CoreSersic / ABI-2 change was checked for speed against its parent commit (identical but for 6 us in the per-sample kernels)
for i in 1 2 3; do
  for d in _ab_prev .; do
    tag=$(basename $(realpath $d))
    (cd $d && python scripts/ab_bench.py 4096 row_flush=1 2>&1 | tail -1; python scripts/bench_configs.py all 2>/dev/null | grep -o '"ms_fwd_bwd": [0-9.]*\|"evals_per_s": [0-9.]*' | tr '\n' ' ') >> gpurun_out/ab_$tag.txt 2>&1
    echo >> gpurun_out/ab_$tag.txt
  done
done
