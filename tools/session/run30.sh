KD_SPARE_SMS=0 KD_B200_LIB=$PWD/tools/variants/libkd_trace.so KD_TRACE_DUMP=gpurun_out/trace_fused.bin python tools/profile_emit.py 16384 reference 4 > gpurun_out/trace.log 2>&1; tail -1 gpurun_out/trace.log
python tools/trace_fused.py gpurun_out/trace_fused.bin
