"""Decode step of the AKI-4B language model (32 layers, B=8, T~655 then B=2, T=8192): graph-replayed step through HF's
decoder layers vs through the fused layers (ops.skinny_linear); plus the weight-streaming kernel alone on the four
projection shapes (GB/s against the measured copy bandwidth).  usage: python tools/decode_time.py"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aki_b200
from aki_b200 import ops
from aki_b200.model import AkiPhi3Runner, phi35_mini_config
dev = torch.device("cuda", 0)
ev = lambda: torch.cuda.Event(enable_timing=True)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for (B, K, N, kw) in [(8, 3072, 9216, {}), (8, 3072, 3072, {"res": 1}), (8, 3072, 8192, {"swiglu": 1}), (8, 8192, 3072, {"res": 1}), (8, 3072, 32064, {})]:
    x = torch.randn(B, K, device=dev).to(torch.bfloat16)
    ws = [torch.randn((2 * N if kw.get("swiglu") else N), K, device=dev).to(torch.bfloat16) for _ in range(8)]   # > L2 in total
    res = torch.randn(B, N, device=dev).to(torch.bfloat16) if kw.get("res") else None
    gamma = torch.ones(K, device=dev, dtype=torch.bfloat16) if not kw else None
    for w in ws: ops.skinny_linear(x, w, gamma, 1e-5, res, bool(kw.get("swiglu")))
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()                      # replayed: the launch path of 40 ctypes calls is not the subject
    with torch.cuda.graph(graph):
        for rep in range(5):
            for w in ws: ops.skinny_linear(x, w, gamma, 1e-5, res, bool(kw.get("swiglu")))
    graph.replay(); torch.cuda.synchronize()
    a, b_ = ev(), ev(); a.record()
    graph.replay()
    b_.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b_) / 40
    gb = ws[0].numel() * 2 / 1e9
    print(f"skinny_linear B={B} K={K} N={N} {kw}: {ms * 1e3:.1f} us, {gb / (ms * 1e-3):.0f} GB/s = {gb / (ms * 1e-3) / peak:.2f} of {peak:.0f}", flush=True)
runner = AkiPhi3Runner(phi35_mini_config(), device=dev, seed=0)
for (B, T, n_dec) in [(8, 655, 32), (2, 8192, 32)]:
    emb = (torch.randn(B, T, 3072, device=dev) * 0.05).to(torch.bfloat16)
    for fused in (False, True):
        cache = runner.new_cache(B, T + n_dec + 4)
        tok = runner.prefill(emb, None, cache)[:, -1].argmax(-1, keepdim=True)
        tok = runner.decode_step_graphed(tok, cache, fused=fused)
        torch.cuda.synchronize()
        a, b_ = ev(), ev(); a.record()
        for _ in range(n_dec - 1):
            tok = runner.decode_step_graphed(tok, cache, fused=fused)
        b_.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b_) / (n_dec - 1)
        floor = (7.64e9 + 2 * B * 32 * (T + n_dec / 2) * 96 * 2 * 32) / (peak * 1e9) * 1e3
        print(f"decode B={B} T={T} {'fused layers' if fused else 'HF layers  '}: {ms:.2f} ms per step (HBM floor {floor:.2f} ms: 7.6 GB weights + K/V)", flush=True)
        del cache
