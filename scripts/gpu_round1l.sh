set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"count_ngrams|ngram_filter" -s 6 -c 6 -f -o gpurun_out/prof_count_r01c python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench2.log 2>&1
timeout 400 python scripts/probe_big2.py > gpurun_out/probe_big2.log 2>&1; cat gpurun_out/probe_big2.log | cut -c1-600
