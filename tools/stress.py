"""Soak test of the two tcgen05 kernels' barrier protocols: random geometries (batch, heads, length, image count, ragged
padding, RoPE on / off) run forward + backward back to back without host synchronisation in between, the way a training
step does; every 8th case is checked against the SIMT verification kernels, every case for finite outputs and for
run-to-run identical forward outputs.  Meant for the trap build (tools/build_variant.sh trap -DAKI_MBAR_TRAP): a wait that
never completes becomes a launch failure after 4 s instead of a hang.
usage: AKI_MMA_LIB=build/libaki_trap.so python tools/stress.py [seconds] [seed]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as Hp
import aki_b200
from aki_b200 import ops
dev = torch.device("cuda", 0)
D = 96
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
rope = aki_b200.LongRope(device=dev)
t_end = time.time() + budget
n_case = n_checked = bad = 0
while time.time() < t_end:
    B = int(rng.integers(1, 5)); H = int(rng.choice([1, 2, 4, 8, 32]))
    many = rng.random() < 0.25                       # many short spans (more than the forward plan cuts at) vs 0-4 long ones
    n_img = int(rng.integers(5, 13)) if many else int(rng.integers(0, 5))
    N = int(rng.choice([16, 40])) if many else int(rng.choice([128, 144]))
    L = int(rng.integers(40 + 12 * n_img, 3000 if H <= 8 else 1200))
    pad = int(rng.integers(0, max(1, L // 2))) if rng.random() < 0.5 else 0
    use_rope = bool(rng.random() < 0.7)
    text_only = bool(rng.random() < 0.2)
    if n_img == 0 and rng.random() < 0.5:
        segs, T = None, L
    else:
        lang, am = Hp.make_prompt(B, L, N, n_img, pad_right=pad)
        if rng.random() < 0.3:                       # left padding on the even samples (AKI.generate pads on the left)
            lp = int(rng.integers(1, 8))
            lang[0::2, :lp] = Hp.PAD_ID; am[0::2, :lp] = 0
        segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID,
                                  text_only=text_only)
        T = segs.T
    g = torch.Generator(device=dev).manual_seed(n_case)
    q, k, v, d_o = (torch.randn(B, T, H, D, generator=g, device=dev).to(torch.bfloat16) for _ in range(4))
    cos = sin = None
    if use_rope:
        cos, sin = rope.tables(torch.arange(T, device=dev)[None], max_position=T - 1)
    meta = ops.meta_tuple(segs)
    scale = D ** -0.5
    reps = int(rng.integers(2, 6))
    outs = []
    try:
        for _ in range(reps):            # no synchronisation between the launches
            o, lse = ops.attn_fwd_raw(q, k, v, cos, sin, meta, scale)
            dq, dk, dv = (torch.empty_like(q) for _ in range(3))
            ops.attn_bwd_raw(d_o, q, k, v, o, lse, cos, sin, meta, scale, dq, dk, dv)
            outs.append((o, lse, dq, dk, dv))
        torch.cuda.synchronize()
    except Exception as e:
        print(f"FAILED case {n_case}: B={B} H={H} T={T} img={n_img} pad={pad} rope={use_rope}: {e}", flush=True)
        sys.exit(1)
    why = []
    for nm, t in zip(("o", "lse", "dq", "dk", "dv"), outs[0]):
        # lse is +inf by design on rows without a visible key (pad rows): only NaN is an error there
        bad_el = torch.isnan(t.float()) if nm == "lse" else ~torch.isfinite(t.float())
        if bad_el.any().item():
            idx = torch.nonzero(bad_el)
            why.append(f"{nm} non-finite n={idx.shape[0]} first={idx[0].tolist()} last={idx[-1].tolist()}")
    for r_i, o_ in enumerate(outs[1:]):
        for i, nm in ((0, "o"), (1, "lse")):
            if not torch.equal(outs[0][i], o_[i]):
                dd = (outs[0][i].float() - o_[i].float()).abs(); idx = torch.nonzero(dd > 0)
                why.append(f"{nm} rep{r_i + 1} differs n={idx.shape[0]} max={dd.max().item():.3g} first={idx[0].tolist()} last={idx[-1].tolist()}")
    ok = not why
    if n_case % 8 == 0:
        o_s, lse_s = ops.attn_fwd_raw(q, k, v, cos, sin, meta, scale, simt=True)
        dq_s, dk_s, dv_s = (torch.empty_like(q) for _ in range(3))
        ops.attn_bwd_raw(d_o, q, k, v, o_s, lse_s, cos, sin, meta, scale, dq_s, dk_s, dv_s, simt=True)
        for a, r_ in zip((outs[-1][0], outs[-1][2], outs[-1][3], outs[-1][4]), (o_s, dq_s, dk_s, dv_s)):
            e = (a.float() - r_.float()).abs().max().item()
            if e > 2e-2 * max(r_.float().abs().max().item(), 1.0):
                ok = False; why.append(f"vs SIMT err {e:.3g}")
        n_checked += 1
    if not ok:
        bad += 1
        # diagnostics: which repetition is wrong against the SIMT kernel, and does a forward-only rerun agree with it
        o_s, lse_s = ops.attn_fwd_raw(q, k, v, cos, sin, meta, scale, simt=True)
        errs = [f"{(o_[0].float() - o_s.float()).abs().max().item():.3g}" for o_ in outs]
        again = [ops.attn_fwd_raw(q, k, v, cos, sin, meta, scale)[0] for _ in range(3)]
        errs2 = [f"{(a_.float() - o_s.float()).abs().max().item():.3g}" for a_ in again]
        why.append(f"max|o - simt| per rep {errs}, forward-only rerun {errs2}, ptrs o={[hex(o_[0].data_ptr()) for o_ in outs]} q={hex(q.data_ptr())}")
        print(f"BAD case {n_case}: B={B} H={H} T={T} img={n_img} pad={pad} rope={use_rope} segs={segs is not None} text_only={text_only} N={N} reps={reps}: " + "; ".join(why[:3] + why[-1:]), flush=True)
    n_case += 1
print(f"{n_case} cases ({n_checked} checked against the SIMT kernels), {bad} bad", flush=True)
sys.exit(1 if bad else 0)
