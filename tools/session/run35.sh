KD_STATS_TMA=0 python tools/time_stats.py 2>&1 | tail -1
python tools/time_stats.py 2>&1 | tail -1
for v in 384_4_2 512_3_2 768_4_1 512_6_1; do echo $v; KD_B200_LIB=$PWD/tools/variants/libkd_tma_$v.so python tools/time_stats.py 2>&1 | tail -1; done
