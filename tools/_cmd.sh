cd /root/repo; export PYTHONUNBUFFERED=1
timeout 200 python tools/stress_decode.py 50 1 2>&1 | grep -v rope_param | tail -6
