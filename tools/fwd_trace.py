"""clock64 stamps of one forward CTA's first item (trace build: tools/build_variant.sh trace -DAKI_FWD_TRACE, run with
AKI_MMA_LIB=build/libaki_trace.so).  usage: python tools/fwd_trace.py [cta] [n_img]   (headline shape T=8192 B=2 H=32)
columns  sm_t*:  0 pass start | 1 S_FULL seen | 2 S in registers, S_FREE | 3 max / rescale decided | 4 exp2(keys 0-63) done |
                 5 P buffer free | 6 P(a) published | 7 exp2(keys 64-127) done | 8 PV(a) landed | 9 P(b) published
         mma_t*: 0 QK^T wait S_FREE | 1 issue QK^T | 2 wait P(a) | 3 issue PV(a) | 4 wait P(b) | 5 issue PV(b)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import aki_b200
from aki_b200 import ops
dev = torch.device("cuda", 0)
H, D, T, B = 32, 96, 8192, 2
cta = sys.argv[1] if len(sys.argv) > 1 else "0"
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rope = aki_b200.LongRope(device=dev)
lang, am = bench.make_prompt(B, T, n_img) if n_img else (np.random.default_rng(0).integers(3, 31000, size=(B, T)).astype(np.int64), np.ones((B, T), dtype=np.int64))
segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), 128, bench.MEDIA_ID, t_cap=T, exact_shape=False)
meta = ops.meta_tuple(segs)
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B, T, 3 * H * D, generator=g, device=dev).to(torch.bfloat16)
cos, sin = rope.tables(torch.arange(T, device=dev)[None], max_position=T - 1)
q4 = qkv[..., :H * D].unflatten(-1, (H, D)); v4 = qkv[..., 2 * H * D:].unflatten(-1, (H, D))
k_rot = torch.empty(B, H, T, D, dtype=torch.bfloat16, device=dev)
ops.rope_kv_write(qkv, cos, sin, k_rot, None, 0, H)
for _ in range(3):
    ops.attn_fwd_raw(q4, k_rot.transpose(1, 2), v4, cos, sin, meta, D ** -0.5)
torch.cuda.synchronize()
os.environ["AKI_MMA_FWD_TRACE"] = cta
ops.attn_fwd_raw(q4, k_rot.transpose(1, 2), v4, cos, sin, meta, D ** -0.5)
torch.cuda.synchronize()
