"""Backward bring-up check: tcgen05 backward vs the SIMT verification kernels on geometries that exercise the item ring
(many items per CTA, empty key tiles), unaligned image visits and ragged lengths.  usage: python tools/bwd_check.py [quick]"""
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
bad = 0
def case(name, B, L, N, n_img, H, pad_right=0, use_rope=True, plain=False):
    global bad
    if plain:
        segs = None; T = L
    else:
        lang, am = Hp.make_prompt(B, L, N, n_img, pad_right=pad_right)
        segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
        T = segs.T
    g = torch.Generator(device=dev).manual_seed(1)
    q, k, v, d_o = (torch.randn(B, T, H, D, generator=g, device=dev).to(torch.bfloat16) for _ in range(4))
    cos = sin = None
    if use_rope:
        cos, sin = rope.tables(torch.arange(T, device=dev)[None], max_position=T - 1)
    meta = ops.meta_tuple(segs)
    o, lse = ops.attn_fwd_raw(q, k, v, cos, sin, meta, D ** -0.5, simt=True)
    outs = []
    for simt in (False, True):
        dq, dk, dv = (torch.full_like(q, float("nan")) for _ in range(3))
        ops.attn_bwd_raw(d_o, q, k, v, o, lse, cos, sin, meta, D ** -0.5, dq, dk, dv, simt=simt)
        torch.cuda.synchronize()
        outs.append((dq, dk, dv))
    errs = []
    ok = True
    for a, r_, nm in zip(outs[0], outs[1], ("dq", "dk", "dv")):
        fin = torch.isfinite(a).all().item()
        e = (a.float() - r_.float()).abs().max().item()
        scale = r_.float().abs().max().item()
        errs.append(f"{nm} {e:.4f}/{scale:.2f}")
        ok = ok and fin and e <= 2e-2 * max(scale, 1.0)
    bad += (not ok)
    print(f"{'ok ' if ok else 'BAD'} {name:30s} T={T:5d} B={B} H={H}  " + "  ".join(errs), flush=True)

case("plain causal T=128", 1, 128, 0, 0, 1, plain=True, use_rope=False)
case("plain causal T=300 rope", 1, 300, 0, 0, 2, plain=True)
case("plain causal T=1024 rope", 2, 1024, 0, 0, 4, plain=True)
case("cfg1 1 image", 1, 257, 128, 1, 32)
case("sft pad 1 image N=144", 4, 513, 144, 1, 8, pad_right=70)
case("2 images N=128", 2, 600, 128, 2, 4)
case("3 images N=144 ragged", 3, 1000, 144, 3, 2, pad_right=133)
case("heavy pad (empty key tiles)", 4, 1200, 128, 1, 4, pad_right=700)
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sys.exit(1 if bad else 0)
case("4 images T~2.5K", 2, 2048, 128, 4, 4)
case("many items B=4 H=32 T=2K", 4, 2048 - 3 * 127, 128, 3, 32)
print("FAILED" if bad else "ALL OK", flush=True)
sys.exit(1 if bad else 0)
