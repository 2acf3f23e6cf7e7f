import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aki_b200 import ops
T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
q = torch.randn(1, T, 32, 96, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
o, lse = ops.attn_fwd_raw(q, k, v, None, None, None, 96 ** -0.5)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
