"""BASELINE config 3: MMA attention fwd+bwd sweep over seq 1K-16K with 1-4 interleaved image spans (plus a pure-causal
control), bf16, one B200.  B = 16384 / T; kernel-only times through aki_mma_set_timing_events; TFLOP/s from the exact
number of visible pairs.  Prints one markdown table row per point."""
import os, sys, statistics, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import aki_b200
from aki_b200 import ops
from aki_b200._lib import lib
from oracle import mma_oracle as O

dev = torch.device("cuda", 0)
H, D = 32, 96
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0
rope = aki_b200.LongRope(device=dev)
print(f"| T | B | images | nnz (M) | fwd ms | fwd TFLOP/s | bwd ms | bwd TFLOP/s | fwd+bwd TFLOP/s | pct of {peak:.0f} |")
print("|---|---|---|---|---|---|---|---|---|---|")
for T in (1024, 2048, 4096, 8192, 16384):
    B = max(1, 16384 // T)
    for n_img, use_rope in ((0, True), (1, True), (2, True), (3, True), (4, True), (4, False)):
        if n_img * 127 + 200 > T:
            continue
        lang, am = bench.make_prompt(B, T, n_img) if n_img else (np.random.default_rng(0).integers(3, 31000, size=(B, T)).astype(np.int64), np.ones((B, T), dtype=np.int64))
        segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), 128, bench.MEDIA_ID, t_cap=T, exact_shape=False)
        nnz = O.count_allowed(O.segments_ref(lang, am, 128, bench.MEDIA_ID))
        meta = ops.meta_tuple(segs)
        g = torch.Generator(device=dev).manual_seed(0)
        qkv = torch.randn(B, T, 3 * H * D, generator=g, device=dev).to(torch.bfloat16)
        d_o = torch.randn(B, T, H, D, generator=g, device=dev).to(torch.bfloat16)
        cos, sin = rope.tables(torch.arange(T, device=dev)[None], max_position=T - 1)
        cos_a, sin_a = (cos, sin) if use_rope else (None, None)      # RoPE-off variant: Q used as is (K is given rotated or not alike)
        q4 = qkv[..., :H * D].unflatten(-1, (H, D)); v4 = qkv[..., 2 * H * D:].unflatten(-1, (H, D))
        k_rot = torch.empty(B, H, T, D, dtype=torch.bfloat16, device=dev)
        dq = torch.empty_like(q4.contiguous()); dk = torch.empty_like(dq); dv = torch.empty_like(dq)
        ops.rope_kv_write(qkv, cos, sin, k_rot, None, 0, H)
        tf, tb = [], []
        for it in range(8):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            for e in ev: e.record()
            lib.aki_mma_set_timing_events(ev[0].cuda_event, ev[1].cuda_event)
            o, lse = ops.attn_fwd_raw(q4, k_rot.transpose(1, 2), v4, cos_a, sin_a, meta, D ** -0.5)
            lib.aki_mma_set_timing_events(ev[2].cuda_event, ev[3].cuda_event)
            ops.attn_bwd_raw(d_o, q4, k_rot.transpose(1, 2), v4, o, lse, cos_a, sin_a, meta, D ** -0.5, dq, dk, dv)
            torch.cuda.synchronize()
            if it >= 3:
                tf.append(ev[0].elapsed_time(ev[1])); tb.append(ev[2].elapsed_time(ev[3]))
        f, b_ = statistics.median(tf), statistics.median(tb)
        tot = 43008.0 * nnz / ((f + b_) * 1e-3) / 1e12
        print(f"| {T} | {B} | {n_img}{'' if use_rope else ' (RoPE off)'} | {nnz / 1e6:.1f} | {f:.3f} | {12288.0 * nnz / (f * 1e-3) / 1e12:.0f} | {b_:.3f} | "
              f"{30720.0 * nnz / (b_ * 1e-3) / 1e12:.0f} | {tot:.0f} | {100 * tot / peak:.1f} |")
        assert not torch.isnan(dq.float()).any() and not torch.isnan(o.float()).any()
