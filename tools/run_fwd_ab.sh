#!/bin/bash
# One gpurun call: forward bring-up check on the trap build, then same-box timing of the variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
AKI_MMA_LIB=$PWD/build/libaki_trap.so timeout 400 python tools/fwd_check.py > gpurun_out/fwd_check.log 2>&1
rc=$?
echo "fwd_check rc=$rc"; tail -25 gpurun_out/fwd_check.log
if [ $rc -ne 0 ]; then exit $rc; fi
for i in 1 2; do
  AKI_MMA_LIB_COMPAT=1 AKI_MMA_LIB=$PWD/build/libaki_r1base.so timeout 120 python tools/fwd_time.py 2>&1 | tail -1
  timeout 120 python tools/fwd_time.py 2>&1 | tail -1
  AKI_MMA_FWD_SCHED=static timeout 120 python tools/fwd_time.py 2>&1 | tail -1 | sed 's/^/static-sched /'
  AKI_MMA_PLAN_FLAGS=0 timeout 120 python tools/fwd_time.py 2>&1 | tail -1 | sed 's/^/plan-flags-0 /'
  AKI_MMA_PLAN_FLAGS=1 timeout 120 python tools/fwd_time.py 2>&1 | tail -1 | sed 's/^/plan-flags-1 /'
done
