"""One forward + backward of the bench workload (T=8192, B=2, 4 image spans) for ncu:
  ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_sm100|attn_bwd_sm100|bwd_preprocess|dq_finalize|decode_partial|decode_combine|rope_kv_write|fwd_plan|skinny_linear|add_rmsnorm|swiglu' \
      -c 15 -o gpurun_out/prof python tools/profile_case.py
then tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>.csv"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import aki_b200
from aki_b200 import ops
dev = torch.device("cuda", 0)
H, D, T, B = 32, 96, 8192, 2
rope = aki_b200.LongRope(device=dev)
lang, am = bench.make_prompt(B, T, 4)
segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), 128, bench.MEDIA_ID, t_cap=T, exact_shape=False)
meta = ops.meta_tuple(segs)
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(B, T, 3 * H * D, generator=g, device=dev).to(torch.bfloat16)
d_o = torch.randn(B, T, H, D, generator=g, device=dev).to(torch.bfloat16)
cos, sin = rope.tables(torch.arange(T, device=dev)[None], max_position=T - 1)
q4 = qkv[..., :H * D].unflatten(-1, (H, D)); v4 = qkv[..., 2 * H * D:].unflatten(-1, (H, D))
k_rot = torch.empty(B, H, T, D, dtype=torch.bfloat16, device=dev)
dq = torch.empty_like(q4.contiguous()); dk = torch.empty_like(dq); dv = torch.empty_like(dq)
ops.rope_kv_write(qkv, cos, sin, k_rot, None, 0, H)
o, lse = ops.attn_fwd_raw(q4, k_rot.transpose(1, 2), v4, cos, sin, meta, D ** -0.5)
ops.attn_bwd_raw(d_o, q4, k_rot.transpose(1, 2), v4, o, lse, cos, sin, meta, D ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
# decode attention against an 8K cache (HBM-bound): B=8 sequences x 32 heads
Bd, Td = 8, 8192
kc = torch.randn(Bd, H, Td, D, device=dev).to(torch.bfloat16); vc = torch.randn(Bd, H, Td, D, device=dev).to(torch.bfloat16)
qd = torch.randn(Bd, H, D, device=dev).to(torch.bfloat16)
ops.decode_op(qd, kc, vc, torch.full((Bd,), Td, dtype=torch.int32, device=dev), Td, D ** -0.5)
torch.cuda.synchronize()
# the decode step's linear layers at B=8 tokens (HBM-bound weight streams): qkv (RMSNorm prologue), o_proj (+residual),
# gate_up (RMSNorm, SwiGLU), down (+residual)
x = torch.randn(8, 3072, device=dev).to(torch.bfloat16); xi = torch.randn(8, 8192, device=dev).to(torch.bfloat16)
gamma = torch.ones(3072, device=dev, dtype=torch.bfloat16); res = torch.randn(8, 3072, device=dev).to(torch.bfloat16)
w_qkv = torch.randn(9216, 3072, device=dev).to(torch.bfloat16); w_o = torch.randn(3072, 3072, device=dev).to(torch.bfloat16)
w_gu = torch.randn(16384, 3072, device=dev).to(torch.bfloat16); w_d = torch.randn(3072, 8192, device=dev).to(torch.bfloat16)
ops.skinny_linear(x, w_qkv, gamma, 1e-5)
ops.skinny_linear(x, w_o, residual=res)
ops.skinny_linear(x, w_gu, gamma, 1e-5, swiglu=True)
ops.skinny_linear(xi, w_d, residual=res)
torch.cuda.synchronize()
# the prefill-size element-wise kernels of the decoder layer (B=8, T=655: 5240 tokens)
xm = torch.randn(5240, 3072, device=dev).to(torch.bfloat16); rm = torch.randn(5240, 3072, device=dev).to(torch.bfloat16)
gu = torch.randn(5240, 16384, device=dev).to(torch.bfloat16)
ops.add_rmsnorm(xm, gamma, 1e-5, residual=rm)
ops.swiglu(gu)
torch.cuda.synchronize()
