# ncu passes of the bench step (development aid): TAG=r02d bash scripts/gpu_prof.sh  -> gpurun_out/${TAG}_launches.csv, prof_count_${TAG}.ncu-rep
mkdir -p gpurun_out
T=${TAG:-r02x}
Q="--no-cpu-baseline --no-e2e --no-extra --no-digest ${BENCH_ARGS:-}"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 $Q > gpurun_out/ncu_launch_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-count_ngrams|ngram_filter|relabel|prune}" -s ${SKIP:-20} -c ${COUNT:-20} -f -o gpurun_out/prof_count_$T python bench.py --steps 2 --warmup 1 $Q > gpurun_out/ncu_full_$T.log 2>&1
tail -3 gpurun_out/ncu_full_$T.log
