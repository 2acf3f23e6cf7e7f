"""Run-to-run determinism of the tcgen05 backward: dK and dV accumulate in TMEM in a fixed order, so two runs on the
same inputs must agree bit for bit; any difference is a synchronisation bug.  Prints where differences occur."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench, helpers as Hp
from aki_b200 import ops
dev = torch.device("cuda", 0)
B, T, H, D = 1, 8192, 32, 96
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lang, am = bench.make_prompt(B, T, 4)
segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), 128, Hp.MEDIA_ID, t_cap=T, exact_shape=False)
meta = ops.meta_tuple(segs)
g = torch.Generator(device=dev).manual_seed(0)
q, k, v, d_o = (torch.randn(B, T, H, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(4))
o, lse = ops.attn_fwd_raw(q, k, v, None, None, meta, D ** -0.5)
ref = None
noise = torch.empty(64 << 20, device=dev)
for it in range(n_rep):
    dq = torch.empty_like(q); dk = torch.empty_like(q); dv = torch.empty_like(q)
    if it % 3 == 1:
        noise.normal_()                      # perturb timing / L2 state between runs
    ops.attn_bwd_raw(d_o, q, k, v, o, lse, None, None, meta, D ** -0.5, dq, dk, dv)
    torch.cuda.synchronize()
    if ref is None:
        ref = (dk.clone(), dv.clone(), dq.clone())
        continue
    for name, a, b_ in (("dk", dk, ref[0]), ("dv", dv, ref[1])):
        if not torch.equal(a, b_):
            diff = (a.float() - b_.float()).abs()
            idx = (diff > 0).nonzero()
            ts = idx[:, 1]; hs = idx[:, 2]
            print(f"run {it}: {name} differs in {idx.shape[0]} elements, max {float(diff.max()):.4f}; rows {int(ts.min())}..{int(ts.max())} "
                  f"(key tiles {sorted(set((ts // 128).tolist()))[:8]}), heads {sorted(set(hs.tolist()))[:8]}, d range {int(idx[:,3].min())}..{int(idx[:,3].max())}")
    dqe = float((dq.float() - ref[2].float()).abs().max())
    if dqe > 0.05:
        print(f"run {it}: dq max diff {dqe:.4f}")
print("done", n_rep)
# --- the SIMT verification kernels on the same problem (tests use them as the on-device reference)
os_, lse_s = ops.attn_fwd_raw(q, k, v, None, None, meta, D ** -0.5, simt=True)
sref = None
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    dq = torch.empty_like(q); dk = torch.empty_like(q); dv = torch.empty_like(q)
    ops.attn_bwd_raw(d_o, q, k, v, os_, lse_s, None, None, meta, D ** -0.5, dq, dk, dv, simt=True)
    torch.cuda.synchronize()
    if sref is None:
        sref = (dk.clone(), dv.clone())
        print("simt vs tcgen05: dk", float((dk.float() - ref[0].float()).abs().max()), "dv", float((dv.float() - ref[1].float()).abs().max()))
        continue
    for name, a, b_ in (("dk", dk, sref[0]), ("dv", dv, sref[1])):
        if not torch.equal(a, b_):
            diff = (a.float() - b_.float()).abs(); idx = (diff > 0).nonzero()
            print(f"SIMT run {it}: {name} differs in {idx.shape[0]} elements, max {float(diff.max()):.4f}, rows {int(idx[:,1].min())}..{int(idx[:,1].max())} heads {sorted(set(idx[:,2].tolist()))[:6]}")
print("simt done")
