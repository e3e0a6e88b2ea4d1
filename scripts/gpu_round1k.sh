set -x
mkdir -p gpurun_out
run() { env "$@" timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1]); print('$*', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phase_ms_per_step'].items() if k in ('count','prune','export')}, d['config']['patterns'])"; }
run COLIBRI_B200_DENSE=0
run COLIBRI_B200_DENSE=2048
run COLIBRI_B200_DENSE=4096
run COLIBRI_B200_DENSE=1024
run COLIBRI_B200_DENSE=0
run COLIBRI_B200_DENSE=2048
