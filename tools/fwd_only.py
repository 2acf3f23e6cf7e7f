"""One forward of the bench workload (for ncu --metrics dram__bytes_read.sum)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, aki_b200
from aki_b200 import ops
dev = torch.device("cuda", 0); H, D, T, B = 32, 96, 8192, 2
rope = aki_b200.LongRope(device=dev)
lang, am = bench.make_prompt(B, T, 4)
segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), 128, bench.MEDIA_ID, t_cap=T, exact_shape=False)
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B, T, 3 * H * D, generator=g, device=dev).to(torch.bfloat16)
cos, sin = rope.tables(torch.arange(T, device=dev)[None], max_position=T - 1)
q4 = qkv[..., :H * D].unflatten(-1, (H, D)); v4 = qkv[..., 2 * H * D:].unflatten(-1, (H, D))
k_rot = torch.empty(B, H, T, D, dtype=torch.bfloat16, device=dev)
ops.rope_kv_write(qkv, cos, sin, k_rot, None, 0, H)
for _ in range(3):
    ops.attn_fwd_raw(q4, k_rot.transpose(1, 2), v4, cos, sin, ops.meta_tuple(segs), D ** -0.5)
torch.cuda.synchronize()
