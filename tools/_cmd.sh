cd /root/repo; export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_module.py -m gpu -q --timeout 150 -k "add_rmsnorm or swiglu or fused_prefill or fused_decode" 2>&1 | tail -12
timeout 200 python tools/prefill_profile.py 8 655 2>&1 | grep -v rope_param | grep -A8 "ms of kernels"
timeout 200 python tools/prefill_profile.py 2 8192 2>&1 | grep -v rope_param | grep -A10 "ms of kernels"
