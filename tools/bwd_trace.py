import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aki_b200 import ops
B, T, H, D = 1, 2048, 32, 96
q = torch.randn(B, T, H, D, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q); do = torch.randn_like(q)
o, lse = ops.attn_fwd_raw(q, k, v, None, None, None, D ** -0.5)
dq = torch.empty_like(q); dk = torch.empty_like(q); dv = torch.empty_like(q)
for _ in range(2):
    ops.attn_bwd_raw(do, q, k, v, o, lse, None, None, None, D ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
os.environ["AKI_MMA_BWD_TRACE"] = sys.argv[1] if len(sys.argv) > 1 else "0"
ops.attn_bwd_raw(do, q, k, v, o, lse, None, None, None, D ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
