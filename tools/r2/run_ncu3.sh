#!/bin/bash
# ncu captures, third batch of round 2 (kernels changed in the last session: balanced pooling, projected InfoNCE, batched MSE /
# re-emission / vocabulary-sharded sweeps) + a fresh capture of the headline emit kernel for profiles/emit_traffic.json.
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
cap emit_reference loca_reference row_emit_fused 2 1 "as-written LoCa emit kernel, 16384 x 151936 bf16, T=2 (tools/profile_ops.py loca_reference)" $((6*V*N)) bytes --emit-traffic $N
cap info_nce_chain info_nce "pool_|similarity_row|nce_finalize|broadcast_rows" 5 5 "pooled InfoNCE 64 x 729 x 1152: the five launches of one forward+backward (balanced pooling)" $((3*64*729*1152*2)) bytes
cap info_nce_proj info_nce_proj "pool_|small_gemm|similarity_row|nce_finalize|broadcast_rows" 9 8 "pooled InfoNCE through a 1152 -> 896 projection (kd_info_nce_proj_fwd_bwd): the eight launches of one forward+backward" $(((2*64*729*1152+64*729*896)*2)) bytes
cap mse mse "mse_kernel" 2 1 "feature MSE 65536 x 3584 bf16, forward + gradient in one pass (four vector pairs in flight per thread)" $((3*65536*3584*2)) bytes
cap regrad regrad "regrad_kernel" 2 1 "backward re-emission (grad_output != 1), 16384 x 151936, as-written LoCa" $((6*V*N)) bytes
cap vshard_phases vshard_phases "vshard_" 4 4 "vocabulary-sharded KD phases on one shard (65536 x 18992 of 151936), both tensors streamed: stats, combine, accum, emit" $((2*65536*18992*2)) bytes
ls -la gpurun_out | grep round2_ | head -40
