set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --maxfail=25 --timeout 200 -p no:cacheprovider > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_m.log
tail -6 gpurun_out/pytest_gpu_m.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json | cut -c1-2500
timeout 100 python scripts/ncu_constrained.py 1e8 2>&1 | tail -1 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
