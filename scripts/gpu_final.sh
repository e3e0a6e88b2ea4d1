# Round-2 closing pass on one B200: the whole GPU suite, the default bench line, the ncu launch list of the same command and a full capture of the
# count family (DRAM traffic per step), racecheck of the hot-key cache on a small corpus.  Results land in gpurun_out/ (copied into profiles/ by hand).
mkdir -p gpurun_out
T=${TAG:-r02f}
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu_$T.log 2>&1; tail -3 gpurun_out/pytest_gpu_$T.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; cut -c1-1500 gpurun_out/bench_$T.json; tail -3 gpurun_out/bench_$T.err
Q="--no-cpu-baseline --no-e2e --no-extra --no-digest"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 $Q > gpurun_out/ncu_launch_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"make_id1_hist|part_hist|part_split|part_count|part_gather|count_ngrams|ngram_filter|filter_to_bitmap" -s 14 -c 28 -f -o gpurun_out/prof_count_$T python bench.py --steps 2 --warmup 1 $Q > gpurun_out/ncu_full_$T.log 2>&1
tail -2 gpurun_out/ncu_full_$T.log | cut -c1-200
COLIBRI_B200_HOT=2 COLIBRI_B200_FILTER_MIN=0 COLIBRI_B200_FILTER_LOG2=16 COLIBRI_B200_DENSE_MIN=0 COLIBRI_B200_DENSE=64 COLIBRI_B200_PART_MIN=0 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -c "
import colibri_core_b200 as cb
c = cb.Corpus.synthetic(200000, vocab=3000, seed=5)
m = cb.train(c, MINTOKENS=2, MAXLENGTH=4, QUIET=1)
print('patterns', len(m))
import os
os.environ['COLIBRI_B200_PART_MIN'] = '999999999999'
m = cb.train(c, MINTOKENS=2, MAXLENGTH=4, QUIET=1)
print('patterns', len(m))
" > gpurun_out/racecheck_$T.log 2>&1; tail -6 gpurun_out/racecheck_$T.log | cut -c1-300
