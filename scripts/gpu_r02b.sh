# Round 2, second GPU pass: parity suite, bench line with extras, A/B of the new knobs, launch list + full ncu of the count family.
set -x
mkdir -p gpurun_out
T=${TAG:-r02b}
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$T.log
tail -8 gpurun_out/pytest_gpu_$T.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; cut -c1-6000 gpurun_out/bench_$T.json; tail -5 gpurun_out/bench_$T.err
Q="--no-cpu-baseline --no-e2e --no-extra --no-digest"
for V in "COLIBRI_B200_NO_L2_PIN=1" "COLIBRI_B200_FILTER_MIN=33554432" "COLIBRI_B200_DENSE=4096" "COLIBRI_B200_DENSE=0" "COLIBRI_B200_SPARSE_DIV=0" "COLIBRI_B200_FILTER_LOG2=26" "COLIBRI_B200_HOT=0"; do
  echo "== $V"; env $V timeout 200 python bench.py $Q 2> gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_ms_per_step'], {k:(v['count_ms'],v['items']) for k,v in d['levels_last_step'].items()}, d['parity'].get('checksum_ok'))"
done 2>&1 | tee gpurun_out/ab_$T.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 $Q > gpurun_out/ncu_launch_$T.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"count_ngrams|ngram_filter|relabel|prune_dense" -s 13 -c 13 -f -o gpurun_out/prof_count_$T python bench.py --steps 2 --warmup 1 $Q > gpurun_out/ncu_full_$T.log 2>&1
tail -3 gpurun_out/ncu_full_$T.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
