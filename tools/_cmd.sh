cd /root/repo; export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_module.py tests/test_gpu_soak.py -m gpu -x -q --timeout 150 2>&1 | tail -3
timeout 200 python tools/stress_decode.py 30 7 2>&1 | grep -v rope_param | tail -3
timeout 300 python tools/decode_time.py 2>&1 | grep -v rope_param | grep decode
AKI_MMA_PDL=0 timeout 300 python tools/decode_time.py 2>&1 | grep -v rope_param | grep "fused" | sed 's/^/PDL=0 /'
