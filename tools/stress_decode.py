"""Soak test of the decode path: prefill + graph-replayed fused decode steps (skinny_linear chained by programmatic dependent
launch, rope_kv_write, decode attention) on random batch sizes / prompt lengths, each scenario run TWICE from scratch with
no host synchronisation inside; the path has no atomics, so the greedy tokens and the K/V caches must be bit-identical.
Every 4th scenario is also checked against the step through HF's decoder layers.
usage: python tools/stress_decode.py [seconds] [seed]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aki_b200
from aki_b200.model import AkiPhi3Runner, phi35_mini_config
dev = torch.device("cuda", 0)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
runner = AkiPhi3Runner(phi35_mini_config(num_layers=3), device=dev, seed=0)
t_end = time.time() + budget
n_case = bad = 0
while time.time() < t_end:
    B = int(rng.integers(1, 9)); T = int(rng.integers(5, 700)); n_new = int(rng.integers(2, 24))
    emb = (torch.randn(B, T, 3072, generator=torch.Generator().manual_seed(n_case)) * 0.05).to(torch.bfloat16).to(dev)
    runs = []
    for rep in range(2):
        cache = runner.new_cache(B, T + n_new + 2)
        tok = runner.prefill(emb, None, cache)[:, -1].argmax(-1, keepdim=True)
        toks = [tok]
        for _ in range(n_new):
            tok = runner.decode_step_graphed(tok, cache, fused=True)
            toks.append(tok)
        runs.append((torch.cat(toks, 1), cache))
    torch.cuda.synchronize()
    why = []
    if not torch.equal(runs[0][0], runs[1][0]):
        why.append(f"tokens differ run to run: {runs[0][0].tolist()} vs {runs[1][0].tolist()}")
    n = T + n_new
    for l in range(3):
        if not (torch.equal(runs[0][1].k[l][:, :, :n], runs[1][1].k[l][:, :, :n]) and
                torch.equal(runs[0][1].v[l][:, :, :n], runs[1][1].v[l][:, :, :n])):
            why.append(f"layer {l} K/V cache differs run to run")
    if n_case % 4 == 0:
        cache = runner.new_cache(B, T + n_new + 2)
        tok = runner.prefill(emb, None, cache)[:, -1].argmax(-1, keepdim=True)
        cache_f = runner.new_cache(B, T + n_new + 2)
        runner.prefill(emb, None, cache_f)
        for step in range(n_new):
            ref = runner.decode_step(tok, cache).float()
            got = runner.decode_step_fused(tok, cache_f).float()
            e = float((got - ref).abs().max())
            if e > 3e-2 * max(1.0, float(ref.abs().max())):
                why.append(f"step {step}: fused vs HF layers err {e:.3g}")
                break
            tok = ref[:, -1].argmax(-1, keepdim=True)
    if why:
        bad += 1
        print(f"BAD case {n_case}: B={B} T={T} n_new={n_new}: " + "; ".join(why[:3]), flush=True)
    n_case += 1
print(f"{n_case} scenarios, {bad} bad", flush=True)
sys.exit(1 if bad else 0)
