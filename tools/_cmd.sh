cd /root/repo; export PYTHONUNBUFFERED=1
timeout 300 python tools/sft_profile.py 32 2>&1 | grep -v rope_param | grep -A32 "ms of kernels"
