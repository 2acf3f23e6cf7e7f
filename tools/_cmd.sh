cd /root/repo; export PYTHONUNBUFFERED=1
timeout 300 python tools/decode_time.py 2>&1 | grep -v rope_param | tail -12
