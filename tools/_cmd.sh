cd /root/repo; export PYTHONUNBUFFERED=1
AKI_MMA_LIB=build/libaki_trace.so timeout 120 python tools/bwd_trace.py 5 > gpurun_out/bwd_trace_q3.txt 2>&1
grep -E "(cmp_h0|mma_A|mma_B|drain) it=(4|5|6|7|8|9):" gpurun_out/bwd_trace_q3.txt
