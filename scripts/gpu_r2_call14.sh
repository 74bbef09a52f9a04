# round-2 call 14: decode kernel launch modes (programmatic launch strict / static weights, L2 prefetch, CTAs per SM)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gptq.py -m gpu -q --tb=short -x -k "decode or batch_launch or linear_f16 or quant_linear" > gpurun_out/pytest_decode.log 2>&1
tail -n 8 gpurun_out/pytest_decode.log | cut -c1-300
timeout 300 python scripts/exp/ab_r02.py decode > gpurun_out/ab_decode2.jsonl 2> gpurun_out/ab_decode2.err
tail -c 300 gpurun_out/ab_decode2.err; grep decode_layer gpurun_out/ab_decode2.jsonl | cut -c1-200
