#!/bin/bash
# ncu captures, second batch of round 2 (kernels added after run_ncu1.sh): each op first runs plain, then under ncu; the reports
# are summarised ON the box into gpurun_out/round2_*.md (copied to profiles/ afterwards)
cd $GRAFT_REPO_ROOT
export KD_COMMIT=${KD_COMMIT:-unknown}
NCU="ncu --set full --clock-control none --import-source on"
V=151936; N=16384; M=46656
cap() {  # name op kernel-regex skip count "what" alg kind [extra]
  python tools/profile_ops.py $2 > gpurun_out/plain_$1.log 2>&1 && $NCU -k regex:$3 -s $4 -c $5 -o /tmp/$1 python tools/profile_ops.py $2 > gpurun_out/ncu_$1.log 2>&1
  python tools/summarize_ncu.py gpurun_out/round2_$1.md /tmp/$1.ncu-rep "$6" $7 $8 $9 ${10} > /dev/null 2>&1 || echo "summary failed: $1"
  rm -f /tmp/$1.ncu-rep
}
cap emit_reference loca_reference row_emit_fused 2 1 "as-written LoCa emit kernel, 16384 x 151936 bf16, T=2 (tools/profile_ops.py loca_reference)" $((6*V*N)) bytes --emit-traffic $N
cap emit_reference_ce loca_reference_ce row_emit_fused 2 1 "as-written LoCa emit + fused student cross-entropy (half of the rows carry a target), 16384 x 151936 bf16, T=2" $((6*V*N)) bytes
cap emit_log_target kl_log_target row_emit_fused 2 1 "log_target KL (FB:206-219) on the fused single-reduction kernel (power sums), T=0.8" $((6*V*N)) bytes
cap emit_per_row loca_per_row row_emit_fused 2 1 "per-row LoCa, single sweep (no phase A), 16384 x 151936 bf16, T=2" $((6*V*N)) bytes
cap nce_tc nce_tc kd_gemm_kernel 3 3 "token-level InfoNCE 46656 x 1152: similarity statistics (no store), softmax - I from a recomputation, G . u_t" $((2*M*M*1152)) flops
cap info_nce_chain info_nce "pool_|similarity_row|nce_finalize|broadcast_rows" 5 5 "pooled InfoNCE 64 x 729 x 1152: the five launches of one forward+backward" $((3*64*729*1152*2)) bytes
cap vshard vshard "vshard_" 0 4 "vocabulary-sharded KD phases on one shard (65536 x 18992 of 151936): stats, combine, accum, emit" $((2*65536*18992*2)) bytes
ls -la gpurun_out | grep round2_ | head -40
