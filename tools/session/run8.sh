set -x
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench8.log 2>&1; tail -1 gpurun_out/bench8.log | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench8_ref.log 2>&1; tail -1 gpurun_out/bench8_ref.log | cut -c1-400
