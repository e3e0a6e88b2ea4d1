# One validation pass on a B200 box (development aid): `gpurun --timeout 900 -- 'bash scripts/gpu_validate.sh'`.
# Full GPU parity suite, the default bench line, one constrained (two-stage) run, smoke().  Results land in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --maxfail=25 --timeout 200 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-2500 gpurun_out/bench_n1.json
timeout 100 python scripts/ncu_constrained.py 1e8 2>&1 | tail -1 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
