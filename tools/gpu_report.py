"""Bring-up report (run on the B200 box): prints errors of every kernel against the oracle without stopping at the
first failure.  The pytest suite (-m gpu) is the gate; this is the wide-angle view used while developing.
    python tools/gpu_report.py [fwd] [bwd] [meta] [decode] [big]
"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as Hp                     # noqa: E402
from oracle import mma_oracle as O       # noqa: E402
import aki_b200                          # noqa: E402
from aki_b200 import ops                 # noqa: E402

dev = "cuda"
what = set(sys.argv[1:]) or {"meta", "fwd", "decode", "bwd"}


def section(name):
    print(f"\n=== {name}", flush=True)


def guarded(fn):
    def run(*a, **k):
        try:
            return fn(*a, **k)
        except Exception:
            traceback.print_exc()
            print("   -> EXCEPTION", flush=True)
    return run


def segs_for(lang, am, N, T=None):
    S = O.segments_ref(lang, am, N, Hp.MEDIA_ID)
    segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
    return S, segs


@guarded
def report_meta():
    section("segments vs reference golden masks")
    g = np.load(os.path.join(ROOT, "tests/golden/mask_reference.npz"))
    bad = 0
    for n in range(int(g["n_cases"])):
        lang, am, N = g[f"c{n}_lang"], g[f"c{n}_am"], int(g[f"c{n}_N"])
        shape = g[f"c{n}_mask_shape"]
        ref = np.unpackbits(g[f"c{n}_mask_bits"], axis=-1)[..., :shape[-1]].reshape(shape).astype(np.int64)
        segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
        got = segs.expand_to_4d().cpu().numpy()
        if got.shape != ref.shape or not np.array_equal(got, ref):
            bad += 1
            print("   mismatch case", n, got.shape, ref.shape)
    print(f"   {int(g['n_cases'])} cases, mismatches: {bad}")


@guarded
def fwd_case(name, B, L, N, n_img, rope, simt_only=False, pad_right=0, std=1.0, causal_only=False, check_simt=True):
    lang, am = Hp.make_prompt(B, L, N, n_img, pad_right=pad_right)
    if causal_only:
        S, segs = None, None
        T = L
    else:
        S, segs = segs_for(lang, am, N)
        T = segs.T
    H, D = 32, 96
    q, k, v = Hp.qkv_inputs(B, T, H, D, seed=1, std=std)
    cos = sin = None
    if rope:
        inv = O.longrope_inv_freq(96, 10000.0, np.ones(48, dtype=np.float32))
        cos, sin = O.rope_cos_sin(torch.arange(T)[None], inv, 1.19)
        cos, sin = cos[..., :48].contiguous(), sin[..., :48].contiguous()
    scaling = D ** -0.5
    ref32 = Hp.oracle_attention(q, k, v, S, scaling, cos, sin, torch.float32)
    ref16 = Hp.oracle_attention(q, k, v, S, scaling, cos, sin, torch.bfloat16) if T <= 1024 else None
    qd, kd, vd = q.to(dev), k.to(dev), v.to(dev)
    cd = None if cos is None else cos.to(dev); sd = None if sin is None else sin.to(dev)
    if rope:   # the kernel takes post-RoPE K
        kr = torch.empty(B, H, T, D, dtype=torch.bfloat16, device=dev)
        packed = torch.cat([qd.reshape(B, T, -1), kd.reshape(B, T, -1), vd.reshape(B, T, -1)], -1).contiguous()
        ops.rope_kv_write(packed, cd, sd, kr, None, 0, H)
        kd_in = kr.transpose(1, 2)
    else:
        kd_in = kd
    rows = Hp.live_rows(S, B, T)
    meta = ops.meta_tuple(segs)
    for impl in (["simt"] if simt_only else (["simt", "sm100"] if check_simt else ["sm100"])):
        torch.cuda.synchronize(); t0 = time.time()
        o, lse = ops.attn_fwd_raw(qd, kd_in, vd, cd, sd, meta, scaling, simt=(impl == "simt"))
        torch.cuda.synchronize(); dt = time.time() - t0
        ek, rk, rms = Hp.err_stats(o, ref32, rows)
        eb = Hp.err_stats(ref16, ref32, rows)[0] if ref16 is not None else float("nan")
        dead = float(o.float().cpu()[~rows].abs().max()) if (~rows).any() else 0.0
        nan = bool(torch.isnan(o.float()).any())
        print(f"   {name:34s} {impl:6s} T={T:5d} max_err={ek:.4e} rms_err={rk:.3e} (bf16-eager max_err={eb:.3e}) "
              f"ref_rms={rms:.3f} dead_rows_max={dead:.1e} nan={nan} {dt*1e3:.1f} ms", flush=True)


def report_fwd(big):
    section("forward vs fp32 oracle")
    fwd_case("causal T=128", 1, 128, 4, 0, False, causal_only=True)
    fwd_case("causal T=256", 1, 256, 4, 0, False, causal_only=True)
    fwd_case("causal T=384 B=2", 2, 384, 4, 0, False, causal_only=True)
    fwd_case("causal T=200 (ragged)", 1, 200, 4, 0, False, causal_only=True)
    fwd_case("cfg1 1img N=128 L=257", 1, 257, 128, 1, False)
    fwd_case("cfg1 + rope", 1, 257, 128, 1, True)
    fwd_case("sft-like B=2 pad N=144", 2, 300, 144, 1, True, pad_right=37)
    fwd_case("2 images N=128 L=600", 1, 600, 128, 2, True)
    fwd_case("large scores std=3", 1, 300, 64, 1, False, std=3.0)
    if big:
        fwd_case("4 images T~2k", 1, 1540, 128, 4, True)
        fwd_case("causal T=4096", 1, 4096, 4, 0, False, causal_only=True)


@guarded
def bwd_case(name, B, L, N, n_img, rope, pad_right=0, causal_only=False, impls=("simt", "sm100"), H=32):
    D = 96
    lang, am = Hp.make_prompt(B, L, N, n_img, pad_right=pad_right)
    if causal_only:
        S, segs, T = None, None, L
    else:
        S, segs = segs_for(lang, am, N)
        T = segs.T
    q, k, v = Hp.qkv_inputs(B, T, H, D, seed=2)
    g = torch.Generator().manual_seed(5)
    d_o = torch.randn(B, T, H, D, generator=g).to(torch.bfloat16)
    cos = sin = None
    if rope:
        inv = O.longrope_inv_freq(96, 10000.0, np.ones(48, dtype=np.float32))
        cos, sin = O.rope_cos_sin(torch.arange(T)[None], inv, 1.19)
        cos, sin = cos[..., :48].contiguous(), sin[..., :48].contiguous()
    scaling = D ** -0.5
    # fp32 oracle with autograd (gradients w.r.t. the PRE-RoPE q, k)
    qf = q.float().requires_grad_(True); kf = k.float().requires_grad_(True); vf = v.float().requires_grad_(True)
    qh, kh, vh = qf.transpose(1, 2), kf.transpose(1, 2), vf.transpose(1, 2)
    if rope:
        c = torch.cat([cos, cos], -1).expand(B, -1, -1); s_ = torch.cat([sin, sin], -1).expand(B, -1, -1)
        qh = O.apply_rope(qh, c, s_); kh = O.apply_rope(kh, c, s_)
    m4 = torch.from_numpy(O.expand_segments_to_4d(S, t_out=T)) if S is not None else torch.tril(torch.ones(T, T, dtype=torch.int64))[None, None]
    rows = Hp.live_rows(S, B, T)
    out = O.eager_attention(qh, kh, vh, O.invert_4d_mask(m4, torch.float32), scaling)
    d_o_eff = d_o.float() * rows[:, :, None, None]          # fully masked rows are don't-care: no gradient flows
    out.backward(d_o_eff)
    qd, kd, vd, dod = q.to(dev), k.to(dev), v.to(dev), d_o.to(dev)
    cd = None if cos is None else cos.to(dev); sd = None if sin is None else sin.to(dev)
    if rope:
        kr = torch.empty(B, H, T, D, dtype=torch.bfloat16, device=dev)
        packed = torch.cat([qd.reshape(B, T, -1), kd.reshape(B, T, -1), vd.reshape(B, T, -1)], -1).contiguous()
        ops.rope_kv_write(packed, cd, sd, kr, None, 0, H)
        kd_in = kr.transpose(1, 2)
    else:
        kd_in = kd
    meta = ops.meta_tuple(segs)
    for impl in impls:
        simt = impl == "simt"
        o, lse = ops.attn_fwd_raw(qd, kd_in, vd, cd, sd, meta, scaling, simt=simt)
        dq = torch.full((B, T, H, D), float("nan"), dtype=torch.bfloat16, device=dev); dk = dq.clone(); dv = dq.clone()
        torch.cuda.synchronize(); t0 = time.time()
        ops.attn_bwd_raw(dod, qd, kd_in, vd, o, lse, cd, sd, meta, scaling, dq, dk, dv, simt=simt)
        torch.cuda.synchronize(); dt = time.time() - t0
        msg = f"   {name:30s} {impl:6s} T={T:5d}"
        for nm, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
            e, r, rms = Hp.err_stats(got, ref)
            msg += f" | {nm} max={e:.3e} rms={r:.2e} ref_rms={rms:.3f} nan={bool(torch.isnan(got.float()).any())}"
        print(msg + f" | {dt*1e3:.1f} ms", flush=True)


def report_bwd(big):
    section("backward vs fp32 autograd oracle")
    bwd_case("causal T=128", 1, 128, 4, 0, False, causal_only=True)
    bwd_case("causal T=384 B=2", 2, 384, 4, 0, False, causal_only=True)
    bwd_case("causal T=200 (ragged)", 1, 200, 4, 0, False, causal_only=True)
    bwd_case("cfg1 1img N=128 L=257", 1, 257, 128, 1, False)
    bwd_case("cfg1 + rope", 1, 257, 128, 1, True)
    bwd_case("sft-like B=2 pad N=144", 2, 300, 144, 1, True, pad_right=37)
    bwd_case("2 images N=128 L=600", 1, 600, 128, 2, True)
    if big:
        bwd_case("4 images T~2k H=8", 1, 1540, 128, 4, True, H=8)


@guarded
def report_decode():
    section("decode vs oracle")
    for B, H, tcap, lens in ((1, 32, 1024, [777]), (3, 32, 4096, [4096, 1, 513]), (2, 32, 600, [600, 599])):
        D = 96
        g = torch.Generator().manual_seed(3)
        q = torch.randn(B, H, D, generator=g).to(torch.bfloat16)
        kc = torch.randn(B, H, tcap, D, generator=g).to(torch.bfloat16)
        vc = torch.randn(B, H, tcap, D, generator=g).to(torch.bfloat16)
        ref = O.decode_attention(q.float()[:, :, None], kc.float(), vc.float(), lens, D ** -0.5)[:, 0]
        kv_len = torch.tensor(lens, dtype=torch.int32, device=dev)
        out = ops.decode_op(q.to(dev), kc.to(dev), vc.to(dev), kv_len, max(lens), D ** -0.5)
        e, r, rms = Hp.err_stats(out, ref)
        print(f"   B={B} lens={lens}: max_err={e:.3e} rms_err={r:.2e} ref_rms={rms:.3f}", flush=True)


@guarded
def report_rope():
    section("rope table / kv write vs oracle")
    inv = O.longrope_inv_freq(96, 10000.0, 1.0 + np.arange(48, dtype=np.float32) / 10)
    pos = torch.stack([torch.arange(300), torch.arange(5000, 5300)])
    cos_ref, sin_ref = O.rope_cos_sin(pos, inv, 1.19)
    cos, sin = ops.rope_table(pos.to(dev), inv.to(dev), 1.19)
    print(f"   table: cos err {float((cos.cpu() - cos_ref[..., :48]).abs().max()):.2e} sin err {float((sin.cpu() - sin_ref[..., :48]).abs().max()):.2e}")
    B, T, H, D = 2, 300, 32, 96
    qkv = torch.randn(B, T, 3 * H * D).to(torch.bfloat16)
    kc = torch.zeros(B, H, 512, D, dtype=torch.bfloat16, device=dev); vc = torch.zeros_like(kc)
    qr = torch.empty(B, H, T, D, dtype=torch.bfloat16, device=dev)
    ops.rope_kv_write(qkv.to(dev), cos, sin, kc, vc, 100, H, q_rot=qr)
    kk = qkv[..., H * D:2 * H * D].view(B, T, H, D).transpose(1, 2).float()
    qq = qkv[..., :H * D].view(B, T, H, D).transpose(1, 2).float()
    k_ref = O.apply_rope(kk, cos_ref, sin_ref); q_ref = O.apply_rope(qq, cos_ref, sin_ref)
    v_ref = qkv[..., 2 * H * D:].view(B, T, H, D).transpose(1, 2)
    print(f"   kv_write: k err {float((kc[:, :, 100:400].float().cpu() - k_ref).abs().max()):.3e} "
          f"q err {float((qr.float().cpu() - q_ref).abs().max()):.3e} v exact {bool(torch.equal(vc[:, :, 100:400].cpu(), v_ref))} "
          f"untouched {float(kc[:, :, :100].abs().max()) == 0.0}")


if "meta" in what:
    report_meta()
if "rope" in what:
    report_rope()
if "decode" in what:
    report_decode()
if "bwd" in what:
    report_bwd("big" in what)
if "fwd" in what:
    report_fwd("big" in what)
print("\ndone", flush=True)
