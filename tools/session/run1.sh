set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv
./tools/probe_device > gpurun_out/probe_device.txt 2>&1; tail -80 gpurun_out/probe_device.txt | head -5
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -8 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.log 2>&1; tail -3 gpurun_out/bench1.log
