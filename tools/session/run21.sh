for e in 1 ""; do
KD_BENCH_NO_CLOCKS=$e timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT noclocks=$e ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],'clocks',d['clocks'])"
done
python tools/time_pieces.py | head -2
