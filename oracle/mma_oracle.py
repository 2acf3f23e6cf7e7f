"""CPU oracle for AKI's modality-mutual attention (MMA) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``aki_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the CPU arm -- never as the thing shipped.

It restates, in numpy / CPU torch, the algorithm of the reference path:

  a1  VLMWithLanguageStream._make_modality_mutual_mask   codes/open_flamingo/src/vlm.py:410-443
  a2  VLMWithLanguageStream._prepare_inputs_for_forward   codes/open_flamingo/src/vlm.py:445-603
  a3  stack_with_padding / stack_with_padding_2D_attention codes/open_flamingo/src/utils.py:62-108
  a6  _aki_update_model_kwargs_for_generation (decode)     codes/open_flamingo/src/aki_generation.py:36-86
  a7  _prepare_4d_causal_attention_mask, 4-D branch        transformers==4.41.2 (pinned codes/setup.py:12;
                                                           third party, same code still shipped in the
                                                           installed 5.5.0 modeling_attn_mask_utils.py:356-367)
  a8  Phi3Attention.forward (eager)                        microsoft/Phi-3.5-mini-instruct remote code,
                                                           hub revision unpinned (third party, absent from
                                                           /root/reference; arithmetic-equivalent to the
                                                           installed models/phi3/modeling_phi3.py:153-175,226-271)
  a9  Phi-3 longrope rotary                                same; modeling_rope_utils.py:462-547,
                                                           models/phi3/modeling_phi3.py:67-131

Pinning status (see DESIGN.md "Oracle"):
  * mask half (a1-a3): PINNED -- checked bit-for-bit against the reference's own code executed in the
    build container (oracle/gen_golden.py imports /root/reference with two stubs) and against the
    committed fixtures tests/golden/mask_*.npz generated that way.
  * attention half (a7-a9): the reference holds no test, golden vector or fixture for it and the
    arithmetic lives in un-vendored third-party code => "parity unpinned" at that boundary.  The
    restatement is validated against the installed transformers' Phi3Attention (eager) +
    _prepare_4d_causal_attention_mask, whose outputs are committed as tests/golden/attn_cfg1_small.npz.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

ASSISTANT_TOKEN_ID = 32001  # hard-coded in the reference, vlm.py:492
IGNORE_INDEX = -100


# --------------------------------------------------------------------------------------------------
# a1: the mask for one sample
# --------------------------------------------------------------------------------------------------
def make_modality_mutual_mask(attention_mask_2d: np.ndarray, image_start_idx: int, text_start_idx: int,
                              text_end_idx: int) -> np.ndarray:
    """vlm.py:410-443.  Returns (1, T, T) int64 with values {0, 1}."""
    am = np.asarray(attention_mask_2d)
    T = am.shape[0]
    idx = np.arange(T)
    mask = (idx[None, :] < (idx + 1)[:, None]).astype(np.int64)          # :424-426  j <= i
    # python slice semantics (negative / out-of-range ends clamp) are part of the behaviour    :429
    mask[slice(image_start_idx, text_start_idx), slice(text_start_idx, text_end_idx)] = 1
    # :434-436  inverted = 1 - m ; inverted.bool() is True wherever m != 1.  For the {0,1} masks the
    # reference feeds this equals m == 0; the literal expression is kept for any integer m.
    drop = (1.0 - am.astype(np.float32)) != 0.0
    mask[:, drop] = 0                                                     # :438
    return mask[None]


# --------------------------------------------------------------------------------------------------
# a2 + a3: splice + per-sample masks + stacking
# --------------------------------------------------------------------------------------------------
@dataclass
class PreparedInputs:
    """What _prepare_inputs_for_forward returns (vlm.py:599-603) plus the index map used to build it."""
    attention_mask_4d: np.ndarray        # (B, 1, Tmax, Tmax) int64 {0,1}
    spliced_mask_2d: list                # per sample (T_b,) int64: the 2-D mask after splicing (ones on vision)
    labels: np.ndarray | None            # (B, Tmax) int64, padded with -100 on padding_side
    src_index: np.ndarray                # (B, Tmax) int64: >=0 text token index l; -1-(k*N+v) vision v of image k;
                                         #   np.iinfo(int64).min for batch padding  (on padding_side)
    lengths: np.ndarray                  # (B,) spliced lengths T_b


def prepare_inputs_for_forward(lang_x: np.ndarray, attention_mask: np.ndarray, num_tokens_per_vis: int,
                               media_token_id: int, labels: np.ndarray | None = None,
                               padding_side: str = "left", assistant_token_id: int = ASSISTANT_TOKEN_ID,
                               multi_image: str = "reference") -> PreparedInputs:
    """vlm.py:445-603 (mask / label / index bookkeeping; embeddings are represented by src_index).

    multi_image:
      "reference"  -- literal reference behaviour; raises for a sample with >= 2 <image> tokens, as the
                      reference does (vlm.py:547-554 re-feeds a 3-D mask as a 1-D vector).
      "contiguous" -- SURVEY section 8a-note generalisation: every image span sees all later tokens of a
                      different segment up to and including <|assistant|>.
      "text_only"  -- stricter variant: image spans see only later *text* tokens up to <|assistant|>.
    """
    lang_x = np.asarray(lang_x)
    attention_mask = np.asarray(attention_mask)
    B, L = lang_x.shape
    N = int(num_tokens_per_vis)
    masks, m2d, labs, srcs = [], [], [], []
    for i in range(B):
        image_token_idxs = np.where(lang_x[i] == media_token_id)[0]             # :488
        q = np.where(lang_x[i] == assistant_token_id)[0]                         # :492-496
        q = int(q[0]) if len(q) else 0
        if len(image_token_idxs) == 0:                                            # :498-512
            masks.append(make_modality_mutual_mask(attention_mask[i], 0, 0, q))
            m2d.append(attention_mask[i].astype(np.int64).copy())
            srcs.append(np.arange(L, dtype=np.int64))
            if labels is not None:
                labs.append(np.asarray(labels[i]).astype(np.int64).copy())
            continue
        if len(image_token_idxs) > 1 and multi_image == "reference":
            raise RuntimeError("Tensors must have same number of dimensions: got 3 and 1 "
                               "(reference crashes on the 2nd image of a sample, vlm.py:547-554)")
        if multi_image == "reference":
            p = int(image_token_idxs[0])
            new_am = np.concatenate([attention_mask[i][:p].astype(np.int64), np.ones(N, dtype=np.int64),
                                     attention_mask[i][p + 1:].astype(np.int64)])                  # :547-554
            new_src = np.concatenate([np.arange(p, dtype=np.int64), -1 - np.arange(N, dtype=np.int64),
                                      np.arange(p + 1, L, dtype=np.int64)])                        # :539-546
            new_lab = None
            if labels is not None:                                                                  # :566-577
                lab = np.asarray(labels[i]).astype(np.int64)
                new_lab = np.concatenate([lab[:p], np.full(N, IGNORE_INDEX, dtype=np.int64), lab[p + 1:]])
            mask = make_modality_mutual_mask(new_am, p, p + N, q + N)                               # :556-564
        else:
            # The reference offsets later image idxs by N although one placeholder is consumed (:536-537)
            # and never reaches a second image; the generalisation uses true post-splice indices.
            new_am, new_src, _, new_lab = _true_splice(lang_x[i], attention_mask[i], N, media_token_id,
                                                       None if labels is None else labels[i])
            mask = _generalised_mask(lang_x[i], attention_mask[i], N, media_token_id, assistant_token_id,
                                     multi_image)
        masks.append(mask)
        m2d.append(new_am)
        srcs.append(new_src)
        if labels is not None:
            labs.append(new_lab)
    lengths = np.array([m.shape[-1] for m in masks], dtype=np.int64)
    Tmax = int(lengths.max())
    # a3: masks are always padded bottom/right with zeros (utils.py:99-108) whatever padding_side is
    m4 = np.zeros((B, 1, Tmax, Tmax), dtype=np.int64)
    for i, m in enumerate(masks):
        t = m.shape[-1]
        m4[i, :, :t, :t] = m
    # a3: 1-D tensors padded on padding_side (utils.py:62-96)
    pad_src = np.iinfo(np.int64).min

    def _stack(rows, value):
        out = np.full((B, Tmax), value, dtype=np.int64)
        for i, r in enumerate(rows):
            if padding_side == "right":
                out[i, :len(r)] = r
            else:
                out[i, Tmax - len(r):] = r
        return out

    return PreparedInputs(attention_mask_4d=m4, spliced_mask_2d=m2d,
                          labels=_stack(labs, IGNORE_INDEX) if labels is not None else None,
                          src_index=_stack(srcs, pad_src), lengths=lengths)


def _true_splice(lang_row, am_row, N, media_token_id, lab_row):
    am, src, seg, lab = [], [], [], []
    k = 0
    for l, tok in enumerate(np.asarray(lang_row)):
        if tok == media_token_id:
            am += [1] * N
            src += list(-1 - (k * N + np.arange(N)))
            seg += [k + 1] * N
            lab += [IGNORE_INDEX] * N
            k += 1
        else:
            am.append(int(am_row[l])); src.append(l); seg.append(0)
            lab.append(int(lab_row[l]) if lab_row is not None else 0)
    return (np.array(am, dtype=np.int64), np.array(src, dtype=np.int64), np.array(seg, dtype=np.int64),
            np.array(lab, dtype=np.int64) if lab_row is not None else None)


def _generalised_mask(lang_row, am_row, N, media_token_id, assistant_token_id, variant):
    am, src, seg, _ = _true_splice(lang_row, am_row, N, media_token_id, None)
    T = len(am)
    asst = np.where((src >= 0) & (np.asarray(lang_row)[np.clip(src, 0, None)] == assistant_token_id))[0]
    q_end = int(asst[0]) + 1 if len(asst) else 0
    i = np.arange(T)[:, None]
    j = np.arange(T)[None, :]
    causal = j <= i
    mutual = (seg[:, None] > 0) & (seg[None, :] != seg[:, None]) & (j < q_end) & (j > i)
    if variant == "text_only":
        mutual &= (seg[None, :] == 0)
    mask = (causal | mutual) & (am[None, :] != 0)
    return mask.astype(np.int64)[None]


# --------------------------------------------------------------------------------------------------
# The compact description the CUDA path uses -- restated on CPU so tests can check the
# device segments kernel field by field, and expanded back to the reference's 4-D tensor.
# --------------------------------------------------------------------------------------------------
@dataclass
class Segments:
    seq_len: np.ndarray      # (B,)   int32  spliced length T_b (mask coordinates: sample occupies [0, T_b))
    q_end: np.ndarray        # (B,)   int32  post-splice index of the first <|assistant|> + 1, 0 if absent
    seg: np.ndarray          # (B,T)  int32  0 text, k>=1 vision tokens of image k, -1 batch padding
    valid: np.ndarray        # (B,T)  uint8  spliced 2-D mask (key validity); 0 on batch padding
    row_lo: np.ndarray       # (B,T)  int32  mutual interval [row_lo, row_hi) of extra visible keys (0,0 = none)
    row_hi: np.ndarray       # (B,T)  int32
    src: np.ndarray          # (B,T)  int32  >=0 text token l; -1-(k*N+v) vision; INT32_MIN padding.  Laid out
                             #               in MASK coordinates (top-left aligned) -- see DESIGN.md


INT32_MIN = np.iinfo(np.int32).min


def segments_ref(lang_x, attention_mask, num_tokens_per_vis, media_token_id, t_cap=None,
                 assistant_token_id=ASSISTANT_TOKEN_ID, text_only=False) -> Segments:
    lang_x = np.asarray(lang_x); attention_mask = np.asarray(attention_mask)
    B, L = lang_x.shape
    N = int(num_tokens_per_vis)
    rows = []
    for b in range(B):
        am, src, seg, _ = _true_splice(lang_x[b], attention_mask[b], N, media_token_id, None)
        T = len(am)
        asst = np.where(lang_x[b] == assistant_token_id)[0]
        if len(asst):
            l0 = int(asst[0])
            q_end = int(np.where(src == l0)[0][0]) + 1
        else:
            q_end = 0
        lo = np.zeros(T, dtype=np.int64); hi = np.zeros(T, dtype=np.int64)
        k = 1
        while (seg == k).any():
            span = np.where(seg == k)[0]
            e = int(span[-1]) + 1
            if q_end > e:
                lo[span] = e; hi[span] = q_end
            k += 1
        rows.append((T, q_end, seg, am, lo, hi, src))
    Tmax = max(r[0] for r in rows) if t_cap is None else int(t_cap)
    S = Segments(seq_len=np.array([r[0] for r in rows], dtype=np.int32),
                 q_end=np.array([r[1] for r in rows], dtype=np.int32),
                 seg=np.full((B, Tmax), -1, dtype=np.int32), valid=np.zeros((B, Tmax), dtype=np.uint8),
                 row_lo=np.zeros((B, Tmax), dtype=np.int32), row_hi=np.zeros((B, Tmax), dtype=np.int32),
                 src=np.full((B, Tmax), INT32_MIN, dtype=np.int32))
    for b, (T, _, seg, am, lo, hi, src) in enumerate(rows):
        S.seg[b, :T] = seg; S.valid[b, :T] = (am != 0); S.row_lo[b, :T] = lo; S.row_hi[b, :T] = hi
        S.src[b, :T] = src
    return S


def expand_segments_to_4d(S: Segments, t_out=None, text_only=False) -> np.ndarray:
    """allowed(i,j) = i<len & j<len & valid[j] & (j<=i | row_lo[i]<=j<row_hi[i] [& seg[j]==0 if text_only])."""
    B, T = S.seg.shape
    t_out = T if t_out is None else t_out
    i = np.arange(t_out)[:, None]; j = np.arange(t_out)[None, :]
    out = np.zeros((B, 1, t_out, t_out), dtype=np.int64)
    for b in range(B):
        n = int(S.seq_len[b])
        lo = np.zeros(t_out, dtype=np.int64); hi = np.zeros(t_out, dtype=np.int64); va = np.zeros(t_out, dtype=bool)
        sg = np.full(t_out, -1, dtype=np.int64)
        m = min(T, t_out)
        lo[:m] = S.row_lo[b, :m]; hi[:m] = S.row_hi[b, :m]; va[:m] = S.valid[b, :m] != 0; sg[:m] = S.seg[b, :m]
        mutual = (j >= lo[:, None]) & (j < hi[:, None])
        if text_only:
            mutual &= (sg[None, :] == 0)
        ok = ((j <= i) | mutual) & va[None, :] & (i < n) & (j < n)
        out[b, 0] = ok
    return out


def count_allowed(S: Segments) -> int:
    """nnz = number of allowed (query, key) pairs, exact from the predicate."""
    total = 0
    for b in range(S.seg.shape[0]):
        n = int(S.seq_len[b])
        v = (S.valid[b, :n] != 0).astype(np.int64)
        c = np.concatenate([[0], np.cumsum(v)])
        i = np.arange(n)
        total += int(c[i + 1].sum())
        lo = np.clip(S.row_lo[b, :n].astype(np.int64), 0, n); hi = np.clip(S.row_hi[b, :n].astype(np.int64), 0, n)
        lo = np.maximum(lo, i + 1)
        ext = np.where(hi > lo, c[np.maximum(hi, lo)] - c[lo], 0)
        total += int(ext.sum())
    return total


# --------------------------------------------------------------------------------------------------
# a7: 4-D inversion
# --------------------------------------------------------------------------------------------------
def invert_4d_mask(mask4d: torch.Tensor, embeds_dtype: torch.dtype) -> torch.Tensor:
    """transformers 4.41.2 _prepare_4d_causal_attention_mask, 4-D branch: fp32 tensor of {0, finfo(dtype).min}."""
    inverted = 1.0 - mask4d
    return inverted.masked_fill(inverted.to(torch.bool), torch.finfo(embeds_dtype).min)


# --------------------------------------------------------------------------------------------------
# a9: longrope
# --------------------------------------------------------------------------------------------------
def longrope_inv_freq(head_dim: int, rope_theta: float, ext_factors) -> torch.Tensor:
    """inv_freq[k] = 1 / (ext[k] * theta^(2k/d)); modeling_rope_utils.py:541-545."""
    ext = torch.as_tensor(ext_factors, dtype=torch.float32)
    shape = torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim
    return 1.0 / (ext * rope_theta ** shape)


def longrope_attention_factor(max_position_embeddings: int, original_max_position_embeddings: int) -> float:
    """modeling_rope_utils.py:527-536; = 1.19024 for Phi-3.5-mini (131072 / 4096)."""
    factor = max_position_embeddings / original_max_position_embeddings
    return 1.0 if factor <= 1.0 else math.sqrt(1 + math.log(factor) / math.log(original_max_position_embeddings))


def select_ext_factors(position_ids: torch.Tensor, short_factor, long_factor, original_max_position_embeddings: int):
    """long factors iff max(position_ids)+1 > original_max (modeling_rope_utils.py:47-80, :538-541)."""
    seq_len = int(position_ids.max()) + 1
    return long_factor if seq_len > original_max_position_embeddings else short_factor


def rope_cos_sin(position_ids: torch.Tensor, inv_freq: torch.Tensor, attention_factor: float):
    """models/phi3/modeling_phi3.py:118-131, kept in fp32.  Returns cos, sin of shape (B, T, head_dim)."""
    freqs = position_ids[:, :, None].float() * inv_freq[None, None, :].float()
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos() * attention_factor, emb.sin() * attention_factor


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]; x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x (B,H,T,D); cos/sin (B,T,D).  models/phi3/modeling_phi3.py:178-205."""
    return x * cos[:, None] + rotate_half(x) * sin[:, None]


# --------------------------------------------------------------------------------------------------
# a8: eager attention core and module
# --------------------------------------------------------------------------------------------------
def eager_attention(q, k, v, additive_mask, scaling, row_block: int | None = None):
    """softmax_fp32(q k^T * scaling + mask) v;  q (B,H,Tq,D), k/v (B,H,Tk,D), mask (B,1,Tq,Tk) or None.
    Returns (B,Tq,H,D) like eager_attention_forward (modeling_phi3.py:153-175).  row_block evaluates the same
    formula in query-row blocks so 8K-16K contexts fit host RAM."""
    B, H, Tq, D = q.shape
    if row_block is None:
        w = torch.matmul(q, k.transpose(2, 3)) * scaling
        if additive_mask is not None:
            w = w + additive_mask
        w = torch.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        return torch.matmul(w, v).transpose(1, 2).contiguous()
    out = torch.empty(B, Tq, H, D, dtype=q.dtype)
    for s in range(0, Tq, row_block):
        e = min(Tq, s + row_block)
        w = torch.matmul(q[:, :, s:e], k.transpose(2, 3)) * scaling
        if additive_mask is not None:
            w = w + additive_mask[:, :, s:e]
        w = torch.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        out[:, s:e] = torch.matmul(w, v).transpose(1, 2)
    return out


def additive_mask_from_segments(S: Segments, dtype=torch.float32, rows=None) -> torch.Tensor:
    m4 = torch.from_numpy(expand_segments_to_4d(S))
    if rows is not None:
        m4 = m4[:, :, rows[0]:rows[1]]
    return invert_4d_mask(m4, dtype)


def attention_module_forward(hidden, w_qkv, w_o, cos, sin, additive_mask, num_heads=32, head_dim=96,
                             past_kv=None):
    """Phi3Attention.forward (modeling_phi3.py:226-271) in the dtype of its inputs.
    hidden (B,T,hidden); w_qkv (3*H*D, hidden); w_o (hidden, H*D); cos/sin (B,T,D).
    Returns (out (B,T,hidden), (k_cache, v_cache) each (B,H,T_kv,D) with K post-RoPE)."""
    B, T, _ = hidden.shape
    qkv = hidden @ w_qkv.t()
    hd = num_heads * head_dim
    q = qkv[..., :hd].view(B, T, num_heads, head_dim).transpose(1, 2)
    k = qkv[..., hd:2 * hd].view(B, T, num_heads, head_dim).transpose(1, 2)
    v = qkv[..., 2 * hd:].view(B, T, num_heads, head_dim).transpose(1, 2)
    cos = cos.to(hidden.dtype); sin = sin.to(hidden.dtype)
    q = apply_rope(q, cos, sin); k = apply_rope(k, cos, sin)
    if past_kv is not None:                                   # DynamicCache.update: append along dim 2
        k = torch.cat([past_kv[0], k], dim=2); v = torch.cat([past_kv[1], v], dim=2)
    o = eager_attention(q, k, v, additive_mask, head_dim ** -0.5)
    out = o.reshape(B, T, hd) @ w_o.t()
    return out, (k, v)


def decode_attention(q, k_cache, v_cache, kv_len, scaling):
    """a6: after prefill the mask is 2-D all-ones => one query sees every cached key [0, kv_len).
    q (B,H,1,D); caches (B,H,T_cap,D).  Returns (B,1,H,D)."""
    B = q.shape[0]
    outs = []
    for b in range(B):
        n = int(kv_len[b]) if hasattr(kv_len, "__len__") else int(kv_len)
        outs.append(eager_attention(q[b:b + 1], k_cache[b:b + 1, :, :n], v_cache[b:b + 1, :, :n], None, scaling))
    return torch.cat(outs, dim=0)


def attention_fwd_bwd_fp32(q, k, v, d_out, S: Segments | None, scaling, row_block=None):
    """fp32 forward + autograd backward of the core (used as the gradient oracle).  q,k,v (B,H,T,D),
    d_out (B,T,H,D).  Returns out, dq, dk, dv (fp32)."""
    q = q.detach().float().requires_grad_(True); k = k.detach().float().requires_grad_(True)
    v = v.detach().float().requires_grad_(True)
    mask = additive_mask_from_segments(S, torch.float32) if S is not None else _causal_additive(q.shape[2])
    out = eager_attention(q, k, v, mask, scaling, row_block=row_block)
    out.backward(d_out.float())
    return out.detach(), q.grad, k.grad, v.grad


def _causal_additive(T):
    m = torch.tril(torch.ones(T, T, dtype=torch.int64))[None, None]
    return invert_4d_mask(m, torch.float32)


def fully_masked_rows(S: Segments) -> np.ndarray:
    """(B,T) bool: rows with no visible key.  The reference gives them a uniform average over all Tmax keys
    (every score is finfo.min); the CUDA path writes zeros.  They are excluded from output parity."""
    m4 = expand_segments_to_4d(S)
    return m4[:, 0].sum(-1) == 0
