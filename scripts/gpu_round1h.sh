set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_constrained.py tests/test_host_cli.py -m gpu -q --maxfail=25 --timeout 150 -p no:cacheprovider > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log
tail -12 gpurun_out/pytest_gpu_h.log
timeout 240 python scripts/probe_constrained.py 1e8 > gpurun_out/probe_constrained4.log 2>&1; grep -v lookup gpurun_out/probe_constrained4.log | cut -c1-420 | tail -7; grep lookup gpurun_out/probe_constrained4.log | tail -2
