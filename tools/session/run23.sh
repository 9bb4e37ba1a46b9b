set -x
python tools/profile_emit.py 4096 reference 2 > gpurun_out/plain.log 2>&1 && \
KD_SPARE_SMS=0 ncu --set full --clock-control none --import-source on -k regex:row_emit_fused -s 1 -c 1 -o gpurun_out/fused_r1 python tools/profile_emit.py 4096 reference 2 > gpurun_out/ncu_fused.log 2>&1
tail -3 gpurun_out/ncu_fused.log
