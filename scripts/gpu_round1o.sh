set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 500 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --maxfail=10 --timeout 200 -p no:cacheprovider > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log
tail -15 gpurun_out/pytest_gpu_o.log | cut -c1-400
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -1 gpurun_out/bench_n2.json | cut -c1-1200
