# quick check of a kernel change on one B200: a slice of the parity suite, then the A/B line of the bench (ms per step, phases, checksum)
mkdir -p gpurun_out
T=${TAG:-quick}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -p no:cacheprovider -k "${KEXPR:-tokenis or golden or overflow or bench_corpus or streamed or compact}" > gpurun_out/pytest_$T.log 2>&1; tail -3 gpurun_out/pytest_$T.log | cut -c1-300
TAG=$T bash scripts/gpu_ab.sh
