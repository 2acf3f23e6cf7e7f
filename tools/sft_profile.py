"""Where the SFT step (BASELINE config 4: B=4, T=656, amp_bf16, AdamW, clip) spends its time: torch profiler kernel table of
one step of bench.py's harness on one GPU.  usage: python tools/sft_profile.py [layers]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import aki_b200
from aki_b200.model import AkiPhi3SFT, phi35_mini_config
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Bp, L, N = 4, 513, 144
model = AkiPhi3SFT(phi35_mini_config(num_layers=layers), device=dev, seed=0)
opt = torch.optim.AdamW(model.parameters(), lr=2e-5, weight_decay=1e-4, fused=True)
g = np.random.default_rng(0)
me = type("M", (), {})()
me.lang_model = model.lm; me.media_token_id = bench.MEDIA_ID; me.num_tokens_per_vis = N; me.pad_token_id = 32000
lang = g.integers(3, 31000, size=(Bp, L)).astype(np.int64)
lang[:, 10] = bench.MEDIA_ID; lang[:, 120] = bench.ASST_ID
labels = lang.copy(); labels[:, :121] = -100
ids, am, lab = torch.from_numpy(lang).to(dev), torch.ones(Bp, L, dtype=torch.int64, device=dev), torch.from_numpy(labels).to(dev)
vis = (torch.randn(Bp, 1, N, 3072) * 0.02).to(torch.bfloat16).to(dev)
def step():
    pr = aki_b200.prepare_inputs_for_forward(me, vis, ids, am, labels=lab, padding_side="right")
    loss = model(pr["inputs_embeds"].float(), pr["mma_segments"], pr["labels"])
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"SFT step, {layers} layers: {tot / 1e3:.2f} ms of kernels")
for k, t, n in rows[:30]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  n={n:4d}  {k[:120]}")
