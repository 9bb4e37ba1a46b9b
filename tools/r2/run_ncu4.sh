#!/bin/bash
# ncu captures, last batch of round 2: the re-emission and MSE kernels after the batch-size sweep (tools/knob_sweep.py).
# Each op first runs plain, then under ncu; the reports are summarised ON the box into gpurun_out/round2_*.md.
cd $GRAFT_REPO_ROOT
export KD_COMMIT=${KD_COMMIT:-unknown}
NCU="ncu --set full --clock-control none --import-source on"
V=151936; N=16384
cap() {  # name op kernel-regex skip count "what" alg kind [extra]
  python tools/profile_ops.py $2 > gpurun_out/plain_$1.log 2>&1 && $NCU -k regex:$3 -s $4 -c $5 -o /tmp/$1 python tools/profile_ops.py $2 > gpurun_out/ncu_$1.log 2>&1
  python tools/summarize_ncu.py gpurun_out/round2_$1.md /tmp/$1.ncu-rep "$6" $7 $8 $9 ${10} > /dev/null 2>&1 || echo "summary failed: $1"
  rm -f /tmp/$1.ncu-rep
}
cap mse mse "mse_kernel" 2 1 "feature MSE 65536 x 3584 bf16, forward + gradient in one pass (eight vector pairs in flight per thread)" $((3*65536*3584*2)) bytes
cap regrad regrad "regrad_kernel" 2 1 "backward re-emission (grad_output != 1), 16384 x 151936, as-written LoCa (two vector pairs + table slice per thread)" $((6*V*N)) bytes
ls -la gpurun_out | grep "round2_mse\|round2_regrad"
