"""Where the AKI-4B prefill (B=8, T=655; and B=2, T=8192) spends its time: torch profiler kernel table of one prefill
through AkiPhi3Runner (HF decoder layers + this library's attention).  usage: python tools/prefill_profile.py [B T]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aki_b200
from aki_b200.model import AkiPhi3Runner, phi35_mini_config
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 655
runner = AkiPhi3Runner(phi35_mini_config(), device=dev, seed=0)
emb = (torch.randn(B, T, 3072, device=dev) * 0.05).to(torch.bfloat16)
for _ in range(2):
    cache = runner.new_cache(B, T + 8)
    runner.prefill(emb, None, cache)
torch.cuda.synchronize()
cache = runner.new_cache(B, T + 8)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    runner.prefill(emb, None, cache)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"B={B} T={T}: {tot / 1e3:.2f} ms of kernels")
for k, t, n in rows[:18]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  {k[:110]}")
