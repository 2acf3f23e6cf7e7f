cd /root/repo; export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_module.py -m gpu -q --timeout 150 -k "fused_elementwise or hf_generate" 2>&1 | tail -8
