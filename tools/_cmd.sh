cd /root/repo; export PYTHONUNBUFFERED=1
AKI_MMA_LIB=$PWD/build/libaki_trap.so timeout 400 python tools/bwd_check.py > gpurun_out/bwd_check.log 2>&1
rc=$?; echo "bwd_check rc=$rc"; tail -14 gpurun_out/bwd_check.log
if [ $rc -ne 0 ]; then exit $rc; fi
for i in 1 2; do
  timeout 120 python tools/bwd_time.py 2>&1 | tail -1
  AKI_MMA_BWD_SCHED=static timeout 120 python tools/bwd_time.py 2>&1 | tail -1 | sed 's/^/static /'
done
