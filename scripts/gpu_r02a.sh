# Round 2, first GPU pass: parity suite with the forced code paths, the bench line (list mode on / off), launch list + full ncu of the count family.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r02a.sh'
set -x
mkdir -p gpurun_out
T=${TAG:-r02a}
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$T.log
tail -8 gpurun_out/pytest_gpu_$T.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; cut -c1-3000 gpurun_out/bench_$T.json
COLIBRI_B200_SPARSE_DIV=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_${T}_nolist.json 2> gpurun_out/bench_${T}_nolist.err; cut -c1-1500 gpurun_out/bench_${T}_nolist.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$T.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"count_ngrams|ngram_filter|relabel" -s 11 -c 11 -f -o gpurun_out/prof_count_$T python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$T.log 2>&1
tail -3 gpurun_out/ncu_full_$T.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
