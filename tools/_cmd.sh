cd /root/repo; export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_module.py -m gpu -q --timeout 150 -k "cross_entropy or sft_loss" 2>&1 | tail -8
timeout 300 python bench.py --workload sft --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('value'), d.get('ms_per_step'), d.get('sft',{}).get('step_ms'))"
