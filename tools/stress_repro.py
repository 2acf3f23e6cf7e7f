"""Focused companion of tools/stress.py for a handful of short-sequence geometries (3-4 samples x 32 heads, images, RoPE
on): forward + backward back to back WITHOUT host synchronisation, copies of the forward outputs taken in stream right
after each forward; reports whether an input was modified, whether an output changed after its forward, and whether the
forward outputs differ run to run (this is how the skipped-phase wait of round 2 was narrowed down).
usage: python tools/stress_repro.py [trials]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as Hp
import aki_b200
from aki_b200 import ops
dev = torch.device("cuda", 0)
D = 96
rope = aki_b200.LongRope(device=dev)
n_bad = 0
shapes = [(3, 461, 1, 53), (3, 957, 4, 0), (4, 399, 1, 0), (3, 248, 1, 0), (4, 972, 1, 0), (4, 605, 3, 46), (4, 472, 1, 0)]
for trial in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
    B, L, n_img, pad = shapes[trial % len(shapes)]
    H = 32
    lang, am = Hp.make_prompt(B, L, 128, n_img, pad_right=pad)
    segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), 128, Hp.MEDIA_ID)
    T = segs.T
    meta = ops.meta_tuple(segs)
    g = torch.Generator(device=dev).manual_seed(trial)
    q, k, v, d_o = (torch.randn(B, T, H, D, generator=g, device=dev).to(torch.bfloat16) for _ in range(4))
    cos, sin = rope.tables(torch.arange(T, device=dev)[None], max_position=T - 1)
    named = dict(q=q, k=k, v=v, d_o=d_o, cos=cos, sin=sin, row_lo=segs.row_lo, row_hi=segs.row_hi, plan=segs.fwd_plan,
                 seq_len=segs.seq_len)
    saved = {n: t.clone() for n, t in named.items()}
    scale = D ** -0.5
    msgs = []
    recs = []
    for rep in range(4):                                   # no host synchronisation inside the loop
        o, lse = ops.attn_fwd_raw(q, k, v, cos, sin, meta, scale)
        o_c, lse_c = o.clone(), lse.clone()
        dq, dk, dv = (torch.empty_like(q) for _ in range(3))
        ops.attn_bwd_raw(d_o, q, k, v, o, lse, cos, sin, meta, scale, dq, dk, dv)
        recs.append((o, lse, o_c, lse_c))
    torch.cuda.synchronize()
    for n, t in named.items():
        if not torch.equal(t, saved[n]):
            idx = torch.nonzero(t != saved[n])
            msgs.append(f"INPUT {n} modified n={idx.shape[0]} first={idx[0].tolist()} last={idx[-1].tolist()}")
    for rep, (o, lse, o_c, lse_c) in enumerate(recs):
        if not torch.equal(o, o_c):
            idx = torch.nonzero(o != o_c)
            msgs.append(f"rep{rep}: o changed AFTER the forward n={idx.shape[0]} first={idx[0].tolist()} last={idx[-1].tolist()}")
        if not torch.equal(lse, lse_c):
            msgs.append(f"rep{rep}: lse changed AFTER the forward")
        if not torch.equal(recs[0][2], o_c):
            idx = torch.nonzero(recs[0][2] != o_c)
            msgs.append(f"rep{rep}: forward output (copy taken right after it) differs from rep0 n={idx.shape[0]} first={idx[0].tolist()} last={idx[-1].tolist()}")
    if msgs:
        n_bad += 1
        print(f"trial {trial} B={B} T={T} img={n_img} pad={pad}: " + "; ".join(msgs[:5]), flush=True)
print(f"{n_bad} bad trials")
