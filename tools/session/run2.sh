set -x
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -6 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench2.log 2>&1; tail -3 gpurun_out/bench2.log
