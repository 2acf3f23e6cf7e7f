cd /root/repo; export PYTHONUNBUFFERED=1
timeout 120 python tools/fwd_time.py 2>&1 | tail -1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 120 python tools/fwd_time.py 2>&1 | tail -1
