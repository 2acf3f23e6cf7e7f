cd /root/repo; export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_meta.py -q -m gpu --timeout 100 -k plan 2>&1 | tail -4
AKI_MMA_LIB=$PWD/build/libaki_trap.so timeout 300 python tools/fwd_check.py 2>&1 | tail -3
timeout 100 python tools/fwd_time.py 2>&1 | tail -1
timeout 300 python tools/sweep_cfg3.py 2>&1 | grep "^| 8192\|^| 4096\|^| 16384" | cut -c1-60
