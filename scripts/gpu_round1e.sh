set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_constrained.py tests/test_host_cli.py -m gpu -q --maxfail=25 --timeout 150 -p no:cacheprovider > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_e.log
tail -8 gpurun_out/pytest_gpu_e.log
timeout 240 python scripts/probe_constrained.py 1e8 > gpurun_out/probe_constrained3.log 2>&1; grep -v lookup gpurun_out/probe_constrained3.log | cut -c1-420 | tail -7
COLIBRI_B200_NO_CHAIN=1 timeout 240 python scripts/probe_constrained.py 1e8 > gpurun_out/probe_constrained3_nochain.log 2>&1; grep -v lookup gpurun_out/probe_constrained3_nochain.log | cut -c1-420 | sed -n 3,4p
run() { env "$@" timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1]); print('$*', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phase_ms_per_step'].items()})"; }
run A=1
run COLIBRI_B200_COUNT_BPS=6
run COLIBRI_B200_COUNT_BPS=4
run COLIBRI_B200_COUNT_BPS=3
run COLIBRI_B200_COUNT_BPS=2
run COLIBRI_B200_FILTER_BPS=4
run COLIBRI_B200_FILTER_BPS=2
run A=2
