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
# the HBM-bound kernels of the decode / prefill paths on small shapes (odd row counts: every bounds check is exercised)
import aki_b200
rope = aki_b200.LongRope(device=dev)
cos, sin = rope.tables(torch.arange(T, device=dev)[None], max_position=T - 1)
o2, _ = ops.attn_fwd_raw(q, k, v, cos, sin, meta, D ** -0.5)                       # RoPE instantiation of the forward
x = torch.randn(3, 1024, device=dev).to(torch.bfloat16); w = torch.randn(48, 1024, device=dev).to(torch.bfloat16)
gam = torch.ones(1024, device=dev, dtype=torch.bfloat16); res = torch.randn(3, 48, device=dev).to(torch.bfloat16)
ops.skinny_linear(x, w, gam, 1e-5); ops.skinny_linear(x, w, residual=res)
ops.skinny_linear(x, torch.randn(96, 1024, device=dev).to(torch.bfloat16), swiglu=True)
xm = torch.randn(13, 1024, device=dev).to(torch.bfloat16)
ops.add_rmsnorm(xm, gam, 1e-5, residual=torch.randn(13, 1024, device=dev).to(torch.bfloat16))
ops.swiglu(torch.randn(13, 2 * 72, device=dev).to(torch.bfloat16))
lg = torch.randn(2, 7, 40, device=dev).to(torch.bfloat16).requires_grad_(True)
lab = torch.randint(0, 40, (2, 7), device=dev); lab[:, :2] = -100
ops.cross_entropy_shifted(lg, lab).backward()
kc = torch.randn(2, H, 40, D, device=dev).to(torch.bfloat16); vc = torch.randn_like(kc)
ops.decode_op(torch.randn(2, H, D, device=dev).to(torch.bfloat16), kc, vc, torch.tensor([33, 40], dtype=torch.int32, device=dev), 40,
              D ** -0.5, torch.tensor([5, 0], dtype=torch.int32, device=dev))
hh = torch.randn(13, 1024, device=dev).requires_grad_(True); aa = torch.randn(13, 1024, device=dev).to(torch.bfloat16).requires_grad_(True)
ww = torch.ones(1024, device=dev).requires_grad_(True)
hn, xx = ops.add_rmsnorm_amp(hh, aa, ww, 1e-5)
(xx.float().sum() + hn.sum()).backward()
gg = torch.randn(13, 2 * 72, device=dev).to(torch.bfloat16).requires_grad_(True)
ops.swiglu_train(gg).float().sum().backward()
torch.cuda.synchronize()
print("helper kernels ran")
