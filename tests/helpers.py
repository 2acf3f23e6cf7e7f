"""Shared builders for the parity tests: seeded synthetic inputs in the reference's geometry, the oracle's
answer for them, and the tolerance rule of SURVEY.md section 8(d)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import mma_oracle as O

MEDIA_ID, ASST_ID, PAD_ID = 32012, 32001, 32000


def make_prompt(B, L, N, n_img, q_frac=0.85, seed=0, pad_right=0, first_img=8):
    """lang_x / attention_mask with n_img <image> placeholders at evenly spaced offsets (first at `first_img`),
    one <|assistant|> token at q_frac of the text, optional right padding (SFT collate)."""
    g = np.random.default_rng(seed)
    lang = g.integers(3, 31000, size=(B, L)).astype(np.int64)
    am = np.ones((B, L), dtype=np.int64)
    for b in range(B):
        n_valid = L - (pad_right if b % 2 == 1 else 0)
        if n_img:
            step = max(1, (int(n_valid * q_frac) - first_img) // n_img)
            for k in range(n_img):
                lang[b, first_img + k * step] = MEDIA_ID
        qpos = int(n_valid * q_frac)
        while lang[b, qpos] == MEDIA_ID:
            qpos += 1
        lang[b, qpos] = ASST_ID
        lang[b, n_valid:] = PAD_ID
        am[b, n_valid:] = 0
    return lang, am


def qkv_inputs(B, T, H=32, D=96, seed=0, device="cpu", std=1.0):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(B, T, H, D, generator=g) * std).to(torch.bfloat16)
    k = (torch.randn(B, T, H, D, generator=g) * std).to(torch.bfloat16)
    v = (torch.randn(B, T, H, D, generator=g) * std).to(torch.bfloat16)
    return q.to(device), k.to(device), v.to(device)


def oracle_attention(q, k, v, S, scaling, cos=None, sin=None, dtype=torch.float32, row_block=None):
    """q,k,v (B,T,H,D) bf16 (pre-RoPE if cos given: both q and k are rotated, mirroring Phi3Attention).
    Returns (B,T,H,D) in `dtype` arithmetic."""
    qf, kf, vf = (x.cpu().to(dtype).transpose(1, 2) for x in (q, k, v))
    if cos is not None:
        c = torch.cat([cos, cos], -1).cpu().to(dtype); s = torch.cat([sin, sin], -1).cpu().to(dtype)
        if c.shape[0] == 1:
            c = c.expand(qf.shape[0], -1, -1); s = s.expand(qf.shape[0], -1, -1)
        qf = O.apply_rope(qf, c, s); kf = O.apply_rope(kf, c, s)
    T = qf.shape[2]
    if S is not None:
        m4 = torch.from_numpy(O.expand_segments_to_4d(S, t_out=T))
    else:
        m4 = torch.tril(torch.ones(T, T, dtype=torch.int64))[None, None]
    add = O.invert_4d_mask(m4, dtype).to(dtype) if dtype != torch.float32 else O.invert_4d_mask(m4, dtype)
    return O.eager_attention(qf, kf, vf, add, scaling, row_block=row_block)


def live_rows(S, B, T):
    """(B,T) bool: rows with at least one visible key (fully masked rows are don't-care)."""
    if S is None:
        return torch.ones(B, T, dtype=torch.bool)
    return torch.from_numpy(~O.fully_masked_rows(S))[:, :T]


def err_stats(x, ref, rows=None):
    x = x.detach().float().cpu(); ref = ref.detach().float().cpu()
    if rows is not None:
        x = x[rows]; ref = ref[rows]
    d = (x - ref).abs()
    return float(d.max()), float(d.pow(2).mean().sqrt()), float(ref.pow(2).mean().sqrt())


def within_tolerance(kernel, bf16_ref, fp32_ref, rows=None, floor=1e-3):
    """kernel error vs fp32 oracle <= 2 x error of the reference-style bf16 eager path + floor * RMS(ref)."""
    ek, _, rms = err_stats(kernel, fp32_ref, rows)
    eb, _, _ = err_stats(bf16_ref, fp32_ref, rows)
    return ek <= 2.0 * eb + floor * max(rms, 1.0), ek, eb, rms
