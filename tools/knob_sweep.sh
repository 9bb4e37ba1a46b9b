# runs tools/knob_sweep.py for the default library and every variant under tools/variants/ (rg* = regrad, ms* = mse, pl* = pooling)
python tools/knob_sweep.py 2>&1 | grep "^\["
for d in tools/variants/*/; do
  n=$(basename $d)
  case $n in rg*) w=regrad;; ms*) w=mse;; pl*) w=pool;; *) continue;; esac
  KD_B200_LIB=$d/libkd_b200.so python tools/knob_sweep.py $w 2>&1 | grep "^\["
done
python tools/knob_sweep.py 2>&1 | grep "^\["
