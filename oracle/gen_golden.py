"""Generate tests/golden/*.npz by EXECUTING the reference (mask half) and the installed transformers
(attention half).  Runs only in the build container (needs /root/reference); the fixtures it writes are
committed so the GPU box -- which has no /root/reference -- can check against them.

    python oracle/gen_golden.py            # rewrites tests/golden/

Mask half   : imports /root/reference/codes/open_flamingo/src/vlm.py with two in-process stubs for the
              missing `einops_exts` / `open_clip` packages and calls the reference's own
              VLMWithLanguageStream._prepare_inputs_for_forward / _make_modality_mutual_mask.
Attention   : transformers' _prepare_4d_causal_attention_mask (4.41.2 semantics, still shipped) +
              Phi3Attention(eager) + Phi3RotaryEmbedding(longrope), fp32, CPU.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
MEDIA_ID = 32012          # <image> is the first added token after Phi-3's 32011 ids (factory.py:140-150)
PAD_ID = 32000
ASST = 32001


def import_reference():
    m = types.ModuleType("einops_exts"); m.rearrange_many = lambda *a, **k: None
    sys.modules.setdefault("einops_exts", m)
    sys.modules.setdefault("open_clip", types.ModuleType("open_clip"))
    sys.path.insert(0, "/root/reference/codes")
    from open_flamingo.src.vlm import VLMWithLanguageStream
    return VLMWithLanguageStream


def run_reference_prepare(V, lang_x, attention_mask, labels, N, padding_side, hidden=4):
    emb = torch.nn.Embedding(32064, hidden)
    me = SimpleNamespace(lang_model=SimpleNamespace(get_input_embeddings=lambda: emb), media_token_id=MEDIA_ID,
                         num_tokens_per_vis=N, pad_token_id=PAD_ID,
                         _make_modality_mutual_mask=V._make_modality_mutual_mask)
    B = lang_x.shape[0]
    n_img = int((lang_x == MEDIA_ID).sum(1).max())
    vt = torch.randn(B, max(n_img, 1), N, hidden)
    with torch.no_grad():
        out = V._prepare_inputs_for_forward(me, vt, torch.from_numpy(lang_x), torch.from_numpy(attention_mask),
                                            labels=None if labels is None else torch.from_numpy(labels),
                                            padding_side=padding_side)
    return out


def random_case(rng, B, L, N, p_img=0.85, p_asst=0.8, pad="right"):
    lang = rng.integers(3, 31000, size=(B, L)).astype(np.int64)
    am = np.ones((B, L), dtype=np.int64)
    for b in range(B):
        n_valid = int(rng.integers(min(L, max(2, L // 2)), L + 1))
        if pad == "right":
            am[b, n_valid:] = 0; lang[b, n_valid:] = PAD_ID; lo, hi = 0, n_valid
        elif pad == "left":
            am[b, :L - n_valid] = 0; lang[b, :L - n_valid] = PAD_ID; lo, hi = L - n_valid, L
        else:
            lo, hi = 0, L
        if rng.random() < p_img:
            lang[b, int(rng.integers(lo, hi))] = MEDIA_ID
        if rng.random() < p_asst:
            pos = int(rng.integers(lo, hi))
            if lang[b, pos] != MEDIA_ID:
                lang[b, pos] = ASST
                if rng.random() < 0.3:                       # a second <|assistant|>: only the first counts
                    pos2 = int(rng.integers(lo, hi))
                    if lang[b, pos2] != MEDIA_ID:
                        lang[b, pos2] = ASST
    labels = lang.copy()
    return lang, am, labels


def pack_mask(m4):
    return np.packbits(m4.astype(np.uint8), axis=-1), np.array(m4.shape, dtype=np.int64)


def gen_mask_fixtures():
    V = import_reference()
    rng = np.random.default_rng(20261017)
    cases = []
    # hand-written corner cases ------------------------------------------------------------------
    def fixed(lang, am, N, side):
        lang = np.array(lang, dtype=np.int64); am = np.array(am, dtype=np.int64)
        return lang, am, lang.copy(), N, side
    I, A, P = MEDIA_ID, ASST, PAD_ID
    cases += [
        fixed([[1, 5, I, 7, 8, A, 9, 10, P, P]], [[1, 1, 1, 1, 1, 1, 1, 1, 0, 0]], 4, "right"),      # SFT-like
        fixed([[P, P, 1, I, 7, A, 9, 10]], [[0, 0, 1, 1, 1, 1, 1, 1]], 3, "left"),                    # generate-like
        fixed([[1, A, 6, I, 7, 8]], [[1, 1, 1, 1, 1, 1]], 4, "right"),                                # <|assistant|> before image
        fixed([[1, 2, I, 7, 8, 9]], [[1, 1, 1, 1, 1, 1]], 4, "right"),                                # no <|assistant|> (pre-training)
        fixed([[1, 2, 3, 7, A, 9]], [[1, 1, 1, 1, 1, 1]], 4, "right"),                                # no image
        fixed([[I, A]], [[1, 1]], 5, "right"),                                                        # image first, block of 1
        fixed([[1, I]], [[1, 1]], 2, "left"),                                                         # image last
        fixed([[1, 5, I, 7, A, 9], [1, 2, 3, 4, A, 6]], [[1] * 6, [1, 1, 1, 1, 1, 0]], 3, "right"),   # mixed batch -> padding rows
        fixed([[1, 5, I, 7, A, 9], [1, 2, 3, 4, A, 6]], [[1] * 6, [0, 1, 1, 1, 1, 1]], 3, "left"),
        fixed([[1, I, 0, 7, A, 9]], [[1, 1, 0, 1, 1, 1]], 2, "right"),                                # masked key inside the block
    ]
    # BASELINE config 1 geometry: N=128, L=257, <image> at 8, <|assistant|> at 224 -----------------
    lang = rng.integers(3, 31000, size=(1, 257)).astype(np.int64); lang[0, 8] = I; lang[0, 224] = A
    cases.append((lang, np.ones_like(lang), lang.copy(), 128, "right"))
    # SFT geometry (sft.yaml: L=513, N=144), reduced batch ---------------------------------------
    lang, am, lab = random_case(rng, 2, 513, 144, p_img=1.0, p_asst=1.0, pad="right")
    cases.append((lang, am, lab, 144, "right"))
    # randomised ------------------------------------------------------------------------------------
    for _ in range(60):
        B = int(rng.integers(1, 5)); L = int(rng.integers(2, 40)); N = int(rng.integers(1, 9))
        side = ["right", "left"][int(rng.integers(0, 2))]
        lang, am, lab = random_case(rng, B, L, N, pad=[side, "none"][int(rng.integers(0, 2))])
        cases.append((lang, am, lab, N, side))
    out = {}
    for n, (lang, am, lab, N, side) in enumerate(cases):
        ref = run_reference_prepare(V, lang, am, lab, N, side)
        packed, shape = pack_mask(ref["attention_mask"].numpy())
        assert ref["attention_mask"].dtype == torch.int64
        out[f"c{n}_lang"] = lang; out[f"c{n}_am"] = am; out[f"c{n}_N"] = np.int64(N)
        out[f"c{n}_side"] = np.array(side); out[f"c{n}_mask_bits"] = packed; out[f"c{n}_mask_shape"] = shape
        out[f"c{n}_labels"] = ref["labels"].numpy()
        # embeddings are pure gathers: record which rows are pad (= pad_token_id scalar fill, vlm.py:584-588)
        out[f"c{n}_embed_is_pad"] = (ref["inputs_embeds"] == float(PAD_ID)).all(-1).numpy()
    out["n_cases"] = np.int64(len(cases)); out["media_token_id"] = np.int64(MEDIA_ID)
    np.savez_compressed(os.path.join(GOLD, "mask_reference.npz"), **out)
    # direct calls of a1 with arbitrary integer arguments (slice clamping etc.) ---------------------
    d = {}
    k = 0
    for _ in range(80):
        T = int(rng.integers(1, 24))
        am = (rng.random(T) < 0.85).astype(np.int64)
        a, b, c = (int(x) for x in rng.integers(-3, T + 4, size=3))
        m = V._make_modality_mutual_mask(torch.from_numpy(am), a, b, c, am.shape, torch.int64, torch.device("cpu"))
        d[f"d{k}_am"] = am; d[f"d{k}_args"] = np.array([a, b, c], dtype=np.int64); d[f"d{k}_mask"] = m.numpy().astype(np.uint8)
        k += 1
    d["n_cases"] = np.int64(k)
    np.savez_compressed(os.path.join(GOLD, "mask_a1_direct.npz"), **d)
    print(f"mask fixtures: {len(cases)} prepare cases, {k} direct a1 cases")


def phi3_config(short, long):
    from transformers import Phi3Config
    return Phi3Config(hidden_size=3072, num_attention_heads=32, num_key_value_heads=32, intermediate_size=8192,
                      vocab_size=32064, max_position_embeddings=131072, original_max_position_embeddings=4096,
                      rms_norm_eps=1e-5, attention_dropout=0.0,
                      rope_parameters={"rope_type": "longrope", "rope_theta": 10000.0, "short_factor": short,
                                       "long_factor": long, "original_max_position_embeddings": 4096},
                      attn_implementation="eager")


def fixed_factors():
    g = np.random.default_rng(7)
    short = (1.0 + 0.5 * g.random(48)).round(4).tolist()
    long = (1.0 + 60.0 * np.sort(g.random(48))).round(4).tolist()
    return short, long


def gen_attention_fixture():
    """cfg1-like geometry at reduced T (N=32 image tokens, L=65 => T=96) so the fixture stays small; plus a
    long-factor case (positions beyond 4096).  Everything fp32 on CPU."""
    from transformers.modeling_attn_mask_utils import _prepare_4d_causal_attention_mask
    from transformers.models.phi3.modeling_phi3 import Phi3Attention, Phi3RotaryEmbedding
    V = import_reference()
    short, long = fixed_factors()
    cfg = phi3_config(short, long)
    torch.manual_seed(0)
    attn = Phi3Attention(cfg, layer_idx=0).float().eval()
    for p in attn.parameters():
        torch.nn.init.normal_(p, std=0.02)
    rot = Phi3RotaryEmbedding(cfg)
    out = {"short_factor": np.array(short, dtype=np.float32), "long_factor": np.array(long, dtype=np.float32),
           "w_qkv_sample": attn.qkv_proj.weight.detach()[::257, ::31].numpy(),
           "w_o_sample": attn.o_proj.weight.detach()[::129, ::29].numpy()}
    N, L = 32, 65
    lang = np.random.default_rng(3).integers(3, 31000, size=(1, L)).astype(np.int64)
    lang[0, 8] = MEDIA_ID; lang[0, 48] = ASST
    am = np.ones_like(lang)
    ref = run_reference_prepare(V, lang, am, None, N, "right")
    m4 = ref["attention_mask"]
    T = m4.shape[-1]
    for tag, pos0 in (("short", 0), ("long", 5000)):
        torch.manual_seed(1)
        hidden = torch.randn(1, T, 3072)
        add = _prepare_4d_causal_attention_mask(m4, (1, T), hidden, 0)
        pos = torch.arange(pos0, pos0 + T)[None]
        cos, sin = rot(hidden, pos)
        with torch.no_grad():
            y, w = attn(hidden, (cos, sin), add)
        out[f"{tag}_out"] = y.numpy()[:, :, ::16].copy()
        out[f"{tag}_out_norm"] = np.float64(y.double().norm())
        out[f"{tag}_weights_pos"] = np.packbits((w > 0).numpy(), axis=-1)
        out[f"{tag}_cos"] = cos.numpy()[:, ::7].copy(); out[f"{tag}_sin"] = sin.numpy()[:, ::7].copy()
        out[f"{tag}_inv_freq"] = rot.inv_freq.numpy().copy()
        out[f"{tag}_pos0"] = np.int64(pos0)
    out["attention_scaling"] = np.float64(rot.attention_scaling)
    out["lang"] = lang; out["N"] = np.int64(N); out["T"] = np.int64(T)
    out["mask_bits"], out["mask_shape"] = pack_mask(m4.numpy())
    np.savez_compressed(os.path.join(GOLD, "attn_cfg1_small.npz"), **out)
    print("attention fixture: T =", T, "attention_scaling =", rot.attention_scaling)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    gen_mask_fixtures()
    gen_attention_fixture()
