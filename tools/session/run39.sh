python tools/time_stats.py 2>&1 | tail -1
for v in 4_8 4_4 5_4 3_6 6_4; do echo $v; KD_B200_LIB=$PWD/tools/variants/libkd_st_$v.so python tools/time_stats.py 2>&1 | tail -1; done
