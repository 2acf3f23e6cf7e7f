cd /root/repo; export PYTHONUNBUFFERED=1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload sft --steps 10 --warmup 3 2>gpurun_out/sft.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())['sft']
print('$1', 'step', round(d['step_ms'],1), 'compute-only', round(d['step_ms_without_allreduce'],1), 'exposed', round(d['exposed_allreduce_ms'],1), 'busbw', round(d['nccl_allreduce_busbw_gbs_1gib'],0))"; }
run highprio
AKI_DDP_BUCKET_MB=100 run highprio_b100
AKI_DDP_BUCKET_MB=25 run highprio_b25
grep -i "error\|Traceback" gpurun_out/sft.err | head -3
