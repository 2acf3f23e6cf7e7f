cd /root/repo; export PYTHONUNBUFFERED=1
run() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload sft --steps 6 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d.get('sft',d)
print({k:round(s[k],2) for k in ('step_ms','step_ms_without_allreduce','exposed_allreduce_ms','nccl_allreduce_busbw_gbs_1gib') if k in s})"; }
run A=1
run NCCL_MAX_NCHANNELS=4
run NCCL_MAX_NCHANNELS=8
run NCCL_MAX_NCHANNELS=16 AKI_DDP_BUCKET_MB=64
