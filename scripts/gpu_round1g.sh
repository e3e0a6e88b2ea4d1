set -x
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"count_ngrams|ngram_filter" -s 6 -c 6 -f -o gpurun_out/prof_count_r01b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"constrained_match|constrained_unigram|constrained_stats" -s 6 -c 6 -f -o gpurun_out/prof_constrained_r01b python scripts/ncu_constrained.py > gpurun_out/ncu_constrained.log 2>&1
tail -2 gpurun_out/ncu_constrained.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
