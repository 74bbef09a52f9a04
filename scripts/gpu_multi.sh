mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout -s KILL 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_dist.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -4 | tee gpurun_out/bench_2gpu.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-400
