for f in 1 2; do for sp in 1; do
KD_FUSED=$f KD_SPARE_SMS=$sp timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/b27.log 2>&1; tail -1 gpurun_out/b27.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('RESULT fused=$f spare=$sp ms/step',d['ms_per_step'],'emit',d['roofline']['kernel_ms'],d['roofline']['frac'])"
done; done
KD_FUSED=1 python tools/time_pieces.py 2>&1 | tail -6
KD_FUSED=2 python tools/time_pieces.py 2>&1 | tail -6
