set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_constrained.py tests/test_host_cli.py -m gpu -q --maxfail=25 --timeout 150 -p no:cacheprovider > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
tail -8 gpurun_out/pytest_gpu_c.log
timeout 240 python scripts/probe_constrained.py 1e8 > gpurun_out/probe_constrained2.log 2>&1; grep -v lookup gpurun_out/probe_constrained2.log | cut -c1-420 | tail -7
for i in 1 2; do
(cd scripts/ab_old && timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 10 > ../../gpurun_out/bench_old_$i.json 2> ../../gpurun_out/bench_old_$i.err); python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_old_$i.json').read().strip().splitlines()[-1]); print('OLD', d['ms_per_step'], d['phase_ms_per_step'])"
timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_new_$i.json 2> gpurun_out/bench_new_$i.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_new_$i.json').read().strip().splitlines()[-1]); print('NEW', d['ms_per_step'], d['phase_ms_per_step'])"
done
