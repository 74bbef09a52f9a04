# round-2 call 15: tests touching the rewritten finish kernels (MSE / moments / channel-last backward) + host-time profile
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_observers.py tests/test_gpu_bwd.py tests/test_gpu_next_rows.py tests/test_gpu_calibration.py tests/test_gpu_reference_ext.py -m gpu -q --tb=short -x > gpurun_out/pytest_finish.log 2>&1
tail -n 6 gpurun_out/pytest_finish.log | cut -c1-300
timeout 120 python scripts/exp/prof_calibration.py > gpurun_out/prof_calibration.txt 2>&1
head -n 50 gpurun_out/prof_calibration.txt | cut -c1-200
