"""One forward + backward on a minimal grid for compute-sanitizer (memcheck / synccheck / racecheck):
    compute-sanitizer --tool synccheck python tools/sanitizer_case.py [B H T n_img N]
Default: B=1, H=1, T~256 with one 64-token image span (a couple of items per kernel, every warp role exercised)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as Hp
from aki_b200 import ops
dev = "cuda"
B, H, L, n_img, N = (int(x) for x in (sys.argv[1:6] + ["1", "1", "193", "1", "64"][len(sys.argv) - 1:]))
D = 96
lang, am = Hp.make_prompt(B, L, N, n_img)
segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
T = segs.T
q, k, v = (x.to(dev) for x in Hp.qkv_inputs(B, T, H, D, seed=1))
d_o = torch.randn(B, T, H, D, device=dev).to(torch.bfloat16)
meta = ops.meta_tuple(segs)
o, lse = ops.attn_fwd_raw(q, k, v, None, None, meta, D ** -0.5)
dq = torch.empty_like(q); dk = torch.empty_like(q); dv = torch.empty_like(q)
ops.attn_bwd_raw(d_o, q, k, v, o, lse, None, None, meta, D ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
os_, lse_s = ops.attn_fwd_raw(q, k, v, None, None, meta, D ** -0.5, simt=True)
print("ran", B, H, T, "max|o - o_simt|", float((o.float() - os_.float()).abs().max()), "max|dq|", float(dq.float().abs().max()))
