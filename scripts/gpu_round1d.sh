set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --maxfail=25 --timeout 150 -p no:cacheprovider -x > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
tail -8 gpurun_out/pytest_gpu_d.log
for u in 1 2 4 1 2 4; do
COLIBRI_B200_MLP=$u timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_mlp$u.json 2> gpurun_out/bench_mlp$u.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_mlp$u.json').read().strip().splitlines()[-1]); print('MLP$u', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['phase_ms_per_step'].items()})"
done
