"""Forward bring-up check: tcgen05 forward vs the SIMT verification kernel (same predicate, no tensor cores) on a set
of geometries that exercise the plan (span-aligned tiles, pairing), the item ring (many items per CTA) and ragged
lengths; then kernel-only timing of the headline shape.  usage: python tools/fwd_check.py [quick]"""
import os, sys, statistics
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as Hp
import aki_b200
from aki_b200 import ops
from aki_b200._lib import lib
dev = torch.device("cuda", 0)
D = 96
rope = aki_b200.LongRope(device=dev)
bad = 0
def case(name, B, L, N, n_img, H, pad_right=0, use_rope=True, first_img=8, q_frac=0.85, plain=False):
    global bad
    if plain:
        segs = None; T = L
    else:
        lang, am = Hp.make_prompt(B, L, N, n_img, q_frac=q_frac, pad_right=pad_right, first_img=first_img)
        segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
        T = segs.T
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(B, T, H, D, generator=g, device=dev).to(torch.bfloat16)
    k = torch.randn(B, T, H, D, generator=g, device=dev).to(torch.bfloat16)
    v = torch.randn(B, T, H, D, generator=g, device=dev).to(torch.bfloat16)
    cos = sin = None
    if use_rope:
        cos, sin = rope.tables(torch.arange(T, device=dev)[None], max_position=T - 1)
    meta = ops.meta_tuple(segs)
    o, lse = ops.attn_fwd_raw(q, k, v, cos, sin, meta, D ** -0.5)
    o2, lse2 = ops.attn_fwd_raw(q, k, v, cos, sin, meta, D ** -0.5, simt=True)
    torch.cuda.synchronize()
    e = (o.float() - o2.float()).abs().max().item()
    fin = torch.isfinite(lse2)
    el = (lse[fin] - lse2[fin]).abs().max().item() if fin.any() else 0.0
    same_inf = bool((torch.isfinite(lse) == fin).all())
    ok = e < 2e-2 and el < 2e-2 and same_inf
    bad += (not ok)
    print(f"{'ok ' if ok else 'BAD'} {name:34s} T={T:5d} B={B} H={H} max|do|={e:.4f} max|dlse|={el:.4f} inf-pattern={'same' if same_inf else 'DIFF'}", flush=True)

case("plain causal T=128", 1, 128, 0, 0, 1, plain=True, use_rope=False)
case("plain causal T=256", 1, 256, 0, 0, 2, plain=True, use_rope=False)
case("plain causal T=300 rope", 1, 300, 0, 0, 2, plain=True)
case("plain causal T=1024 rope", 2, 1024, 0, 0, 4, plain=True)
case("cfg1 1 image", 1, 257, 128, 1, 32)
case("sft pad 1 image N=144", 4, 513, 144, 1, 8, pad_right=70)
case("2 images N=128", 2, 600, 128, 2, 4)
case("4 images T~2.5K", 2, 2048, 128, 4, 4)
case("3 images N=144 ragged", 3, 1000, 144, 3, 2, pad_right=133)
case("no rope 2 images", 2, 700, 128, 2, 2, use_rope=False)
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sys.exit(1 if bad else 0)
case("4 images T~8.5K", 1, 8192, 128, 4, 4)
case("many items B=4 H=32 T=2K", 4, 2048 - 3 * 127, 128, 3, 32)
print("FAILED" if bad else "ALL OK", flush=True)
sys.exit(1 if bad else 0)
