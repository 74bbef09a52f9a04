# usage: bash scripts/gpu_tests.sh [pytest args]   -- GPU parity suite + smoke, logs into gpurun_out/
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --maxfail=60 "$@" 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
