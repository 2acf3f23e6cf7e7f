#!/bin/bash
# perf sensitivity experiments for the forward kernel (results are wrong by design when a debug bit is set)
for d in 0 16 32 48 63; do
  echo "== AKI_MMA_FWD_DEBUG=$d"
  AKI_MMA_FWD_DEBUG=$d timeout 120 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fwd_ms', d['kernels']['fwd_ms'], 'fwd_tflops', d['kernels']['fwd_tflops'], 'bwd_ms', d['kernels']['bwd_ms'])"
done
