set -x
python tools/profile_emit.py 16384 reference 2 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:teacher_stats -s 1 -c 1 -o gpurun_out/stats_r1 python tools/profile_emit.py 16384 reference 2 > gpurun_out/ncu_stats.log 2>&1
tail -2 gpurun_out/ncu_stats.log
