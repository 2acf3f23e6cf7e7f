"""clock64 stamps of one backward CTA's first item (trace build: tools/build_variant.sh trace -DAKI_FWD_TRACE, run with
AKI_MMA_LIB=build/libaki_trace.so).  usage: python tools/bwd_trace.py [cta]   (headline shape T=8192 B=2 H=32, causal)
columns  cmp_h*: 0 phase a start | 1 S_FULL seen | 2 exponentials done | 3 P_READY | 4 phase b start | 5 DP_FULL seen | 6 DS_READY
         mma_A:  0 wait dO / P | 1 issue dV | 2 dV issued | 3 Q(next) landed | 4 S^T(next) issued
         mma_B:  0 wait DS_READY | 1 issue dQ, dK | 2 issued | 3 dO / Q (next) landed | 4 dQ drained | 5 dP^T(next) issued
         drain:  0 wait DQ_FULL | 1 seen | 2 dQ in registers, DQ_DRAINED"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aki_b200
from aki_b200 import ops
dev = torch.device("cuda", 0)
H, D, T, B = 32, 96, 8192, 2
g = torch.Generator(device=dev).manual_seed(0)
q, k, v, d_o = (torch.randn(B, T, H, D, generator=g, device=dev).to(torch.bfloat16) for _ in range(4))
o, lse = ops.attn_fwd_raw(q, k, v, None, None, None, D ** -0.5)
dq = torch.empty_like(q); dk = torch.empty_like(q); dv = torch.empty_like(q)
for _ in range(2):
    ops.attn_bwd_raw(d_o, q, k, v, o, lse, None, None, None, D ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
os.environ["AKI_MMA_BWD_TRACE"] = sys.argv[1] if len(sys.argv) > 1 else "0"
ops.attn_bwd_raw(d_o, q, k, v, o, lse, None, None, None, D ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
