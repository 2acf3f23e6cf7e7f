import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aki_b200 import ops
B, T, H, D = 1, 2048, 32, 96
q = torch.randn(B, T, H, D, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
for _ in range(3):
    ops.attn_fwd_raw(q, k, v, None, None, None, D ** -0.5)
torch.cuda.synchronize()
os.environ["AKI_MMA_FWD_TRACE"] = sys.argv[1] if len(sys.argv) > 1 else "0"
ops.attn_fwd_raw(q, k, v, None, None, None, D ** -0.5)
torch.cuda.synchronize()
