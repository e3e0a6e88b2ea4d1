set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_constrained.py -m gpu -q --maxfail=10 --timeout 150 -p no:cacheprovider -k flex > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_n.log
tail -40 gpurun_out/pytest_gpu_n.log | cut -c1-400
