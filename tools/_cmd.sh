cd /root/repo; export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests/test_gpu_attention.py -q -m gpu --timeout 200 -k "at_scale" 2>&1 | tail -15
