for cfg in "" "KD_SPARE_SMS=1" "KD_SPARE_SMS=1 KD_HELPER_RESERVE=4" "KD_SPARE_SMS=1 KD_HELPER_RESERVE=20" ""; do
  env $cfg python bench.py --steps 50 --warmup 5 --no-secondary --no-e2e --no-cpu-baseline --no-parity 2>/dev/null | grep -E "^\{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$cfg] ms/step %.3f emit %.3f ms frac %.3f loss %.9f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['loss']))"
done
