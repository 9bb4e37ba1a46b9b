for i in 1 2 3; do
for cfg in "" "KD_CLUSTER_SIZE=8"; do
  env $cfg python bench.py --steps 50 --warmup 5 --mode per_row --no-secondary --no-e2e --no-cpu-baseline --no-parity 2>/dev/null | grep -E "^\{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[per_row $cfg] ms/step %.3f emit-alone %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
done
done
