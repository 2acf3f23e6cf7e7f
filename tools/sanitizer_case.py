import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as Hp
from aki_b200 import ops
dev = "cuda"
H, D = 32, 96
lang, am = Hp.make_prompt(2, 300, 144, 1, pad_right=37)
segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), 144, Hp.MEDIA_ID)
T = segs.T
q, k, v = (x.to(dev) for x in Hp.qkv_inputs(2, T, H, D, seed=1))
d_o = torch.randn(2, T, H, D, device=dev).to(torch.bfloat16)
meta = ops.meta_tuple(segs)
o, lse = ops.attn_fwd_raw(q, k, v, None, None, meta, D ** -0.5)
dq = torch.empty_like(q); dk = torch.empty_like(q); dv = torch.empty_like(q)
ops.attn_bwd_raw(d_o, q, k, v, o, lse, None, None, meta, D ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
print("ran", T, float(o.float().abs().max()), float(dq.float().abs().max()))
