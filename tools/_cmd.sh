cd /root/repo; export PYTHONUNBUFFERED=1
timeout 200 python tools/stress.py 60 2 2>&1 | grep -v rope_param | tail -4
AKI_MMA_LIB=$PWD/build/libaki_trap.so timeout 200 python tools/stress.py 45 3 2>&1 | grep -v rope_param | tail -4
for i in 1 2; do timeout 120 python tools/step_repro.py 8 2>&1 | tail -2; done
