set -x
mkdir -p gpurun_out
run() { env "$@" timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1]); print('$*', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phase_ms_per_step'].items() if k in ('count','prune','export')})"; }
run A=1
run COLIBRI_B200_MLP_FILTER=2
run COLIBRI_B200_MLP_FILTER=4
run COLIBRI_B200_MLP_COUNT=2
run COLIBRI_B200_NO_FILTER=1
run COLIBRI_B200_FILTER_LOG2=27
run COLIBRI_B200_FILTER_LOG2=26
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
