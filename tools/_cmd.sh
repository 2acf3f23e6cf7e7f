cd /root/repo; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 150 2>&1 | tail -6
timeout 300 python tools/decode_time.py 2>&1 | grep -v rope_param | tee gpurun_out/r02_decode_time.txt | tail -12
