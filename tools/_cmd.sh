cd /root/repo; export PYTHONUNBUFFERED=1
for i in 1 2; do
for v in r1bwd bg2 bg4 bg16 bg64; do
  AKI_MMA_LIB=$PWD/build/libaki_$v.so timeout 120 python tools/bwd_time.py 2>&1 | tail -1
done
timeout 120 python tools/bwd_time.py 2>&1 | tail -1
done
AKI_MMA_LIB=$PWD/build/libaki_trace.so timeout 120 python tools/bwd_trace.py 0 > gpurun_out/bwd_trace_r2.txt 2>&1
