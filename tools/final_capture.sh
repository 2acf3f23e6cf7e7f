#!/bin/bash
# One gpurun call that re-captures the evidence under profiles/ from the current build (outputs land in gpurun_out/final/;
# copy what is to be judged into profiles/).  usage: gpurun --timeout 1500 -- 'bash tools/final_capture.sh [tag]'
cd "$(dirname "$0")/.."
tag=${1:-r02}
out=gpurun_out/final; mkdir -p $out
export PYTHONUNBUFFERED=1
S=$SECONDS
timeout 600 python -m pytest tests -m gpu -x -q --timeout 150 > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$? t=$((SECONDS-S))s"; tail -2 $out/${tag}_pytest_gpu.log
timeout 400 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; echo "bench rc=$? t=$((SECONDS-S))s"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference arm rc=$? t=$((SECONDS-S))s"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_step.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-longctx --no-sft > $out/ncu_launches.log 2>&1; echo "launch list rc=$? t=$((SECONDS-S))s"
timeout 500 ncu --set full --clock-control none --import-source on \
  -k regex:'attn_fwd_sm100|attn_bwd_sm100|bwd_preprocess|dq_finalize|decode_partial|decode_combine|rope_kv_write|fwd_plan|skinny_linear|add_rmsnorm|swiglu' \
  -c 15 -f -o $out/${tag}_prof python tools/profile_case.py > $out/ncu_full.log 2>&1; echo "ncu full rc=$? t=$((SECONDS-S))s"
python tools/ncu_summary.py $out/${tag}_prof.ncu-rep $out/${tag}_ncu_full_summary.csv $out/ncu_traffic.json > /dev/null 2>&1; echo "summary rc=$?"
timeout 200 python tools/prefill_profile.py 8 655 > $out/${tag}_prefill_profile.txt 2>&1; timeout 200 python tools/prefill_profile.py 2 8192 >> $out/${tag}_prefill_profile.txt 2>&1; echo "prefill profile t=$((SECONDS-S))s"
timeout 300 python tools/decode_time.py > $out/${tag}_decode_time.txt 2>&1; echo "decode time t=$((SECONDS-S))s"
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool python tools/sanitizer_case.py > $out/${tag}_sanitizer_$tool.log 2>&1; echo "$tool rc=$? t=$((SECONDS-S))s"; tail -1 $out/${tag}_sanitizer_$tool.log
done
cat $out/${tag}_bench_1gpu.json | head -c 600; echo
cat $out/${tag}_bench_reference.json | head -c 600; echo
