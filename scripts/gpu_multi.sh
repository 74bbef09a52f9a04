mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_host_api.py tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_dist.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_2gpu.log | cut -c1-1500
