#!/bin/bash
# ncu captures of round 2: each op first runs plain (exit code checked), then under ncu; the reports are summarised ON the box
# (gpurun brings back at most 64 MiB) into gpurun_out/*.md
cd $GRAFT_REPO_ROOT
export KD_COMMIT=b50a703
NCU="ncu --set full --clock-control none --import-source on"
V=151936; N=16384
cap() {  # name op kernel-regex skip count "what" alg kind [extra]
  python tools/profile_ops.py $2 > gpurun_out/plain_$1.log 2>&1 && $NCU -k regex:$3 -s $4 -c $5 -o /tmp/$1 python tools/profile_ops.py $2 > gpurun_out/ncu_$1.log 2>&1
  python tools/summarize_ncu.py gpurun_out/round2_$1.md /tmp/$1.ncu-rep "$6" $7 $8 $9 ${10} > /dev/null 2>&1 || echo "summary failed: $1"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/round2_$1_raw.csv 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
cap emit_reference loca_reference row_emit_fused 2 1 "as-written LoCa emit kernel, 16384 x 151936 bf16, T=2 (tools/profile_ops.py loca_reference)" $((6*V*N)) bytes --emit-traffic $N
cap emit_per_row loca_per_row row_emit_fused 2 1 "per-row LoCa, single sweep (no phase A), 16384 x 151936 bf16, T=2" $((6*V*N)) bytes
cap emit_reference_T08 loca_reference_T08 row_emit_fused 2 1 "as-written LoCa emit kernel, peaked logits, T=0.8: every row takes pass B'" $((6*V*N)) bytes
cap emit_log_target kl_log_target row_emit_cluster 2 1 "log_target KL (FB:206-219) on the three-sweep cluster kernel, T=0.8" $((6*V*N)) bytes
cap proj_mse proj_mse kd_gemm_kernel 3 3 "projected feature MSE 65536 x 896 -> 3584: forward (MSE epilogue), d input, d proj" $((2*65536*896*3584)) flops
cap gemm_dh gemm_dh kd_gemm_kernel 2 1 "dH = g . W (A K-major, B MN-major, fp32 out), 65536 x 18992 x 896" $((2*65536*18992*896)) flops
cap gemm_dw gemm_dw kd_gemm_kernel 2 1 "dW = g^T . hidden (both MN-major), 18992 x 65536 x 896" $((2*65536*18992*896)) flops
cap mse mse mse_kernel 2 1 "feature MSE on given tensors, 65536 x 3584 bf16" $((6*65536*3584)) bytes
cap regrad regrad regrad_kernel 2 1 "backward re-emission (grad_output != 1), 16384 x 151936" $((6*V*N)) bytes
ls -la gpurun_out | head -40
timeout 600 python -m pytest tests/test_gpu_module.py -x -q -m gpu 2>&1 | tail -8
