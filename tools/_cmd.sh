cd /root/repo; export PYTHONUNBUFFERED=1
export AKI_MMA_LIB=$PWD/build/libaki_trap.so
timeout 200 python tools/bwd_check.py 2>&1 | tail -4
timeout 200 python tools/fwd_check.py quick 2>&1 | grep -c "^ok"
F="--no-cpu --no-e2e --no-prefill --no-longctx --no-sft"
for s in 5 6 7 5 6 30; do
  timeout 100 python bench.py --steps $s --warmup 3 $F > gpurun_out/s.out 2> gpurun_out/s.err; echo "steps $s rc=$? $(python -c "import json;d=json.load(open('gpurun_out/s.out'));print(round(d['value'],1), round(d['kernels']['attn_fwd_sm100_kernel_ms'],3), round(d['kernels']['attn_bwd_sm100_kernel_ms'],3))" 2>/dev/null)"
done
