"""-m gpu: the drop-in boundary -- AkiMMAAttention (Phi3Attention signature / state-dict keys), the KV-cache
contract, the AttentionInterface plugin and a Phi-3 model with its attention swapped, against the fixture produced
by the installed transformers' eager Phi3Attention and against the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
import helpers as Hp
from oracle import mma_oracle as O

pytestmark = pytest.mark.gpu
dev = "cuda"


def _config(short=None, long=None, layers=2, intermediate=1024, vocab=1000):
    from transformers import Phi3Config
    rp = {"rope_type": "longrope", "rope_theta": 10000.0, "short_factor": [float(x) for x in short], "long_factor": [float(x) for x in long],
          "original_max_position_embeddings": 4096} if short is not None else None
    kw = dict(hidden_size=3072, num_attention_heads=32, num_key_value_heads=32, intermediate_size=intermediate,
              vocab_size=vocab, num_hidden_layers=layers, max_position_embeddings=131072,
              original_max_position_embeddings=4096, rms_norm_eps=1e-5, attention_dropout=0.0, pad_token_id=0,
              attn_implementation="eager")
    if rp is not None:
        kw["rope_parameters"] = rp
    return Phi3Config(**kw)


def _module_like_fixture(g):
    """Same construction + init order as oracle/gen_golden.py (o_proj is registered before qkv_proj)."""
    import aki_b200
    cfg = _config(g["short_factor"], g["long_factor"])
    torch.manual_seed(0)
    mod = aki_b200.AkiMMAAttention(cfg, layer_idx=0).float().eval()
    for p in mod.parameters():
        torch.nn.init.normal_(p, std=0.02)
    assert list(mod.state_dict().keys()) == ["o_proj.weight", "qkv_proj.weight"]
    assert mod.qkv_proj.weight.shape == (9216, 3072) and mod.o_proj.weight.shape == (3072, 3072)
    assert np.allclose(mod.qkv_proj.weight.detach()[::257, ::31].numpy(), g["w_qkv_sample"])
    assert np.allclose(mod.o_proj.weight.detach()[::129, ::29].numpy(), g["w_o_sample"])
    return cfg, mod


def test_module_matches_installed_phi3_eager_fixture():
    import aki_b200
    from aki_b200 import ops
    g = np.load(os.path.join(GOLDEN, "attn_cfg1_small.npz"))
    cfg, mod = _module_like_fixture(g)
    w_qkv, w_o = mod.qkv_proj.weight.detach().clone(), mod.o_proj.weight.detach().clone()
    mod = mod.to(dev).to(torch.bfloat16)
    T, N = int(g["T"]), int(g["N"])
    lang = torch.from_numpy(g["lang"]).to(dev)
    segs = ops.build_segments(lang, torch.ones_like(lang), N, Hp.MEDIA_ID)
    m4 = np.unpackbits(g["mask_bits"], axis=-1)[..., :T].reshape(g["mask_shape"]).astype(np.int64)
    assert np.array_equal(segs.expand_to_4d().cpu().numpy(), m4)
    rope = aki_b200.LongRope(short_factor=g["short_factor"], long_factor=g["long_factor"], device=dev)
    add = O.invert_4d_mask(torch.from_numpy(m4), torch.float32)
    for tag in ("short", "long"):
        pos0 = int(g[f"{tag}_pos0"])
        pos = torch.arange(pos0, pos0 + T, device=dev)[None]
        cos, sin = rope.tables(pos)
        torch.manual_seed(1)
        hidden = torch.randn(1, T, 3072)
        with torch.no_grad():
            out, w = mod(hidden.to(dev).to(torch.bfloat16), None, None, mma_segments=segs, mma_rope=(cos, sin))
        assert w is None and out.shape == (1, T, 3072)
        ref = torch.from_numpy(g[f"{tag}_out"])                                # fp32 HF eager, every 16th column
        # reference-style bf16 eager error of the same module (CPU oracle in bf16)
        c96 = torch.cat([cos, cos], -1).cpu(); s96 = torch.cat([sin, sin], -1).cpu()
        o16, _ = O.attention_module_forward(hidden.bfloat16(), w_qkv.bfloat16(), w_o.bfloat16(), c96, s96, add.bfloat16())
        ok, ek, eb, rms = Hp.within_tolerance(out[:, :, ::16], o16[:, :, ::16], ref, None, floor=2e-3)
        assert ok, f"{tag}: module err {ek:.3e} vs bf16-eager err {eb:.3e} (rms {rms:.3f})"
        # the HF-style call with position_embeddings=(cos, sin) of shape (B,T,96) gives the same result
        with torch.no_grad():
            out2, _ = mod(hidden.to(dev).to(torch.bfloat16), (c96.to(dev).bfloat16().float(), s96.to(dev).bfloat16().float()), None,
                          mma_segments=segs)
        assert float((out2.float() - out.float()).abs().max()) < 2e-2 * max(1.0, float(out.float().abs().max()))


def test_module_rejects_4d_mask_without_segments_and_wrong_head_dim():
    import aki_b200
    g = np.load(os.path.join(GOLDEN, "attn_cfg1_small.npz"))
    cfg = _config(g["short_factor"], g["long_factor"])
    mod = aki_b200.AkiMMAAttention(cfg, 0).to(dev).to(torch.bfloat16)
    x = torch.zeros(1, 8, 3072, device=dev, dtype=torch.bfloat16)
    cs = torch.ones(1, 8, 48, device=dev)
    with pytest.raises(ValueError):
        mod(x, None, torch.ones(1, 1, 8, 8, device=dev), mma_rope=(cs, cs))
    from transformers import Phi3Config
    with pytest.raises(ValueError):
        aki_b200.AkiMMAAttention(Phi3Config(hidden_size=4096, num_attention_heads=32, num_key_value_heads=32), 0)


def test_kv_cache_prefill_then_decode_equals_full_recompute():
    """Generate contract (aki_generation.py:36-86): prefill with the MMA mask writes post-RoPE K and V into the
    cache; each decode step sees every cached key (2-D all-ones mask), position id = past length."""
    import aki_b200
    from aki_b200 import ops
    g = np.load(os.path.join(GOLDEN, "attn_cfg1_small.npz"))
    cfg, mod = _module_like_fixture(g)
    mod = mod.to(dev).to(torch.bfloat16)
    T, N = int(g["T"]), int(g["N"])
    lang = torch.from_numpy(g["lang"]).to(dev)
    segs = ops.build_segments(lang, torch.ones_like(lang), N, Hp.MEDIA_ID)
    rope = aki_b200.LongRope(short_factor=g["short_factor"], long_factor=g["long_factor"], device=dev)
    n_new = 3
    torch.manual_seed(2)
    hidden = torch.randn(1, T + n_new, 3072, device=dev).to(torch.bfloat16)
    cos_all, sin_all = rope.tables(torch.arange(T + n_new, device=dev)[None])
    cache = aki_b200.AkiKVCache(1, 1, 32, 96, t_cap=T + 8, device=dev)
    with torch.no_grad():
        out_p, _ = mod(hidden[:, :T], None, None, past_key_values=cache, mma_segments=segs,
                       mma_rope=(cos_all[:, :T].contiguous(), sin_all[:, :T].contiguous()))
        ref_p, _ = mod(hidden[:, :T], None, None, mma_segments=segs,
                       mma_rope=(cos_all[:, :T].contiguous(), sin_all[:, :T].contiguous()))
    assert torch.equal(out_p, ref_p)                              # cache write does not change the prefill result
    assert cache[0][0].shape == (1, 32, T, 96) and cache.get_seq_length() == T
    for step in range(n_new):
        t = T + step
        with torch.no_grad():
            out_d, _ = mod(hidden[:, t:t + 1], None, None, past_key_values=cache,
                           mma_rope=(cos_all[:, t:t + 1].contiguous(), sin_all[:, t:t + 1].contiguous()))
        assert cache[0][0].shape[2] == t + 1
        # oracle for the step: row t of full attention where rows >= T are plain causal over all keys
        S = O.segments_ref(g["lang"], np.ones_like(g["lang"]), N, Hp.MEDIA_ID)
        m = np.zeros((1, 1, t + 1, t + 1), dtype=np.int64)
        m[:, :, :T, :T] = O.expand_segments_to_4d(S)
        for rr in range(T, t + 1):
            m[0, 0, rr, :rr + 1] = 1
        c96 = torch.cat([cos_all, cos_all], -1)[:, :t + 1].cpu(); s96 = torch.cat([sin_all, sin_all], -1)[:, :t + 1].cpu()
        ref, _ = O.attention_module_forward(hidden[:, :t + 1].float().cpu(), mod.qkv_proj.weight.float().cpu(),
                                            mod.o_proj.weight.float().cpu(), c96, s96,
                                            O.invert_4d_mask(torch.from_numpy(m), torch.float32))
        err = float((out_d[0, 0].float().cpu() - ref[0, t]).abs().max())
        assert err < 3e-2 * max(1.0, float(ref.abs().max())), (step, err)


def test_dynamic_cache_and_plugin_paths():
    """Foreign HF DynamicCache honoured through update(); AttentionInterface function on rotated q/k/v."""
    import aki_b200
    from aki_b200 import ops
    from transformers import DynamicCache
    g = np.load(os.path.join(GOLDEN, "attn_cfg1_small.npz"))
    cfg, mod = _module_like_fixture(g)
    mod = mod.to(dev).to(torch.bfloat16)
    T, N = int(g["T"]), int(g["N"])
    lang = torch.from_numpy(g["lang"]).to(dev)
    segs = ops.build_segments(lang, torch.ones_like(lang), N, Hp.MEDIA_ID)
    rope = aki_b200.LongRope(short_factor=g["short_factor"], long_factor=g["long_factor"], device=dev)
    cos, sin = rope.tables(torch.arange(T + 1, device=dev)[None])
    torch.manual_seed(4)
    hidden = torch.randn(1, T + 1, 3072, device=dev).to(torch.bfloat16)
    rp = lambda a, b_: (cos[:, a:b_].contiguous(), sin[:, a:b_].contiguous())
    with torch.no_grad():
        ref_p, _ = mod(hidden[:, :T], None, None, mma_segments=segs, mma_rope=rp(0, T))
        try:
            dc = DynamicCache(config=cfg)
        except TypeError:
            dc = DynamicCache()
        out_p, _ = mod(hidden[:, :T], None, None, past_key_values=dc, mma_segments=segs, mma_rope=rp(0, T))
        assert float((out_p.float() - ref_p.float()).abs().max()) < 2e-2 * max(1.0, float(ref_p.float().abs().max()))
        assert dc.get_seq_length() == T
        ak = aki_b200.AkiKVCache(1, 1, 32, 96, t_cap=T + 4, device=dev)
        mod(hidden[:, :T], None, None, past_key_values=ak, mma_segments=segs, mma_rope=rp(0, T))
        d1, _ = mod(hidden[:, T:], None, None, past_key_values=dc, mma_rope=rp(T, T + 1))
        d2, _ = mod(hidden[:, T:], None, None, past_key_values=ak, mma_rope=rp(T, T + 1))
        assert float((d1.float() - d2.float()).abs().max()) < 2e-2 * max(1.0, float(d2.float().abs().max()))
    # plugin function: same contract as eager_attention_forward
    q, k, v = Hp.qkv_inputs(1, T, 32, 96, seed=6, device=dev)
    o, w = aki_b200.aki_mma_attention(None, q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), None,
                                      scaling=96 ** -0.5, mma_segments=segs)
    assert w is None and o.shape == (1, T, 32, 96) and o.is_contiguous()
    S = O.segments_ref(g["lang"], np.ones_like(g["lang"]), N, Hp.MEDIA_ID)
    ref32 = Hp.oracle_attention(q.cpu(), k.cpu(), v.cpu(), S, 96 ** -0.5)
    ref16 = Hp.oracle_attention(q.cpu(), k.cpu(), v.cpu(), S, 96 ** -0.5, dtype=torch.bfloat16)
    ok, ek, eb, _ = Hp.within_tolerance(o, ref16, ref32)
    assert ok, (ek, eb)


def test_phi3_model_with_swapped_attention_matches_hf_eager_on_reference_mask():
    """2-layer random-init Phi-3 (hidden 3072, 32x96): HF eager fed the reference's inverted 4-D mask vs the same
    weights with every self_attn replaced by AkiMMAAttention fed mma_segments.  Logits within bf16 tolerance."""
    import aki_b200
    from aki_b200 import ops
    from transformers import Phi3ForCausalLM
    g = np.load(os.path.join(GOLDEN, "attn_cfg1_small.npz"))
    cfg = _config(g["short_factor"], g["long_factor"])
    torch.manual_seed(0)
    model = Phi3ForCausalLM(cfg).to(dev).to(torch.bfloat16).eval()
    B, L, N = 2, 80, 32
    lang, am = Hp.make_prompt(B, L, N, 1, pad_right=9)
    segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
    T = segs.T
    S = O.segments_ref(lang, am, N, Hp.MEDIA_ID)
    m4 = torch.from_numpy(O.expand_segments_to_4d(S)).to(dev)
    torch.manual_seed(5)
    embeds = (torch.randn(B, T, 3072, device=dev) * 0.5).to(torch.bfloat16)
    pos = torch.arange(T, device=dev)[None].expand(B, -1)
    add = O.invert_4d_mask(m4, torch.bfloat16).to(torch.bfloat16)          # a7: {0, finfo(bf16).min}
    with torch.no_grad():
        ref = model(inputs_embeds=embeds, attention_mask=add, position_ids=pos, use_cache=False).logits
        n = aki_b200.replace_phi3_attention(model)
        assert n == 2
        got = model(inputs_embeds=embeds, attention_mask=segs.spliced_mask_2d(), position_ids=pos, use_cache=False,
                    mma_segments=segs).logits
    rows = Hp.live_rows(S, B, T).to(dev)
    d = (got.float() - ref.float())[rows]
    scale = float(ref.float()[rows].abs().max())
    assert float(d.abs().max()) < 0.06 * max(scale, 1.0), (float(d.abs().max()), scale)
    assert float(d.pow(2).mean().sqrt()) < 0.01 * max(scale, 1.0)


def _small_lm(layers=2, vocab=512, seed=0):
    from transformers import Phi3ForCausalLM
    g = np.load(os.path.join(GOLDEN, "attn_cfg1_small.npz"))
    cfg = _config(g["short_factor"], g["long_factor"], layers=layers, vocab=vocab)
    torch.manual_seed(seed)
    return cfg, Phi3ForCausalLM(cfg).to(dev).to(torch.bfloat16).eval()


def _teacher_forced_reference_logits(ref_model, embeds, new_ids, S, T):
    """HF eager model (untouched attention) on prompt + generated tokens in ONE pass with the reference's mask: the MMA
    (B,1,T,T) block for the prompt, plain causal rows for the generated tokens (the generate loop's all-ones 2-D mask,
    aki_generation.py:56-62).  Returns the logits that predict each generated token and the one after the last."""
    n_new = new_ids.shape[1]
    full = torch.cat([embeds, ref_model.model.embed_tokens(new_ids)], 1)
    Tt = T + n_new
    m4 = torch.zeros(1, 1, Tt, Tt, dtype=torch.int64)
    m4[:, :, :T, :T] = torch.from_numpy(O.expand_segments_to_4d(S))
    for i in range(T, Tt):
        m4[0, 0, i, :i + 1] = 1
    add = O.invert_4d_mask(m4.to(dev), torch.bfloat16).to(torch.bfloat16)
    pos = torch.arange(Tt, device=dev)[None]
    with torch.no_grad():
        logits = ref_model(inputs_embeds=full, attention_mask=add, position_ids=pos, use_cache=False).logits
    return logits[:, T - 1:]


@pytest.mark.parametrize("mode", ["module", "plugin"])
def test_hf_generate_through_the_dropin(mode):
    """The reference's generate contract (codes/open_flamingo/src/aki.py:136-209, aki_generation.py:36-86) end to end
    through transformers' own generate(): inputs_embeds prefill with the MMA description, then greedy decode steps on
    the KV cache.  `module`: every self_attn replaced by AkiMMAAttention with an AkiKVCache (a transformers Cache) as
    past_key_values; `plugin`: the stock Phi3Attention with config._attn_implementation = "aki_mma" and HF's own
    DynamicCache.  Per-step logits against the untouched eager model fed the reference's mask, teacher-forced."""
    import copy
    import aki_b200
    from aki_b200 import ops
    cfg, model = _small_lm()
    ref_model = copy.deepcopy(model)
    B, L, N, n_new = 1, 70, 32, 5
    lang, am = Hp.make_prompt(B, L, N, 1)
    segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
    T = segs.T
    S = O.segments_ref(lang, am, N, Hp.MEDIA_ID)
    torch.manual_seed(9)
    embeds = (torch.randn(B, T, 3072, device=dev) * 0.5).to(torch.bfloat16)
    kw = dict(inputs_embeds=embeds, attention_mask=segs.spliced_mask_2d(), max_new_tokens=n_new, do_sample=False,
              use_cache=True, return_dict_in_generate=True, output_logits=True, pad_token_id=0)
    if mode == "module":
        assert aki_b200.replace_phi3_attention(model) == 2
        cache = aki_b200.AkiKVCache(2, B, 32, 96, t_cap=T + n_new + 1, device=dev)
        kw["past_key_values"] = cache
    else:
        aki_b200.register_attention_interface("aki_mma")
        model.config._attn_implementation = "aki_mma"
    with aki_b200.mma_context(segs):          # generate() rejects unknown model kwargs, so the description rides the context
        out = model.generate(**kw)
    new_ids = out.sequences[:, -n_new:]
    assert new_ids.shape == (B, n_new)
    if mode == "module":
        assert cache.get_seq_length() == T + n_new - 1     # the last generated token is never fed back
    got = torch.stack(out.logits, 1).float()               # (B, n_new, vocab)
    ref = _teacher_forced_reference_logits(ref_model, embeds, new_ids, S, T).float()[:, :n_new]
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) < 0.06 * max(scale, 1.0), (float((got - ref).abs().max()), scale)
    assert float((got - ref).pow(2).mean().sqrt()) < 0.012 * max(scale, 1.0)


def test_plugin_selected_through_config_matches_module_swap():
    """config._attn_implementation = "aki_mma" (AttentionInterface registry, modeling_utils.py:4832-4870) on the stock
    Phi3Attention: HF passes attention_mask=None and forwards mma_segments to every layer; logits equal the module swap."""
    import copy
    import aki_b200
    from aki_b200 import ops
    cfg, model = _small_lm(seed=2)
    swapped = copy.deepcopy(model)
    aki_b200.replace_phi3_attention(swapped)
    aki_b200.register_attention_interface("aki_mma")
    model.config._attn_implementation = "aki_mma"
    B, L, N = 2, 60, 16
    lang, am = Hp.make_prompt(B, L, N, 2, pad_right=7)
    segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
    T = segs.T
    torch.manual_seed(3)
    embeds = (torch.randn(B, T, 3072, device=dev) * 0.5).to(torch.bfloat16)
    pos = torch.arange(T, device=dev)[None].expand(B, -1)
    with torch.no_grad():
        a_ = model(inputs_embeds=embeds, attention_mask=segs.spliced_mask_2d(), position_ids=pos, use_cache=False,
                   mma_segments=segs).logits.float()
        b_ = swapped(inputs_embeds=embeds, attention_mask=segs.spliced_mask_2d(), position_ids=pos, use_cache=False,
                     mma_segments=segs).logits.float()
    rows = Hp.live_rows(O.segments_ref(lang, am, N, Hp.MEDIA_ID), B, T).to(dev)
    d = (a_ - b_)[rows]
    assert float(d.abs().max()) < 0.03 * max(1.0, float(b_[rows].abs().max()))


def test_transformers_441_call_signature():
    """The call the reference's pinned transformers 4.41.2 decoder layer makes (keywords attention_mask, position_ids,
    past_key_value, output_attentions, use_cache; three return values; no kwargs pass-through): rope tables from
    position_ids, MMA description from mma_context()."""
    import aki_b200
    from aki_b200 import ops
    g = np.load(os.path.join(GOLDEN, "attn_cfg1_small.npz"))
    cfg, mod = _module_like_fixture(g)
    mod = mod.to(dev).to(torch.bfloat16)
    T, N = int(g["T"]), int(g["N"])
    lang = torch.from_numpy(g["lang"]).to(dev)
    segs = ops.build_segments(lang, torch.ones_like(lang), N, Hp.MEDIA_ID)
    rope = aki_b200.LongRope(short_factor=g["short_factor"], long_factor=g["long_factor"], device=dev)
    pos = torch.arange(T, device=dev)[None]
    cos, sin = rope.tables(pos)
    torch.manual_seed(4)
    hidden = torch.randn(1, T, 3072, device=dev).to(torch.bfloat16)
    with torch.no_grad():
        ref, _ = mod(hidden, None, None, mma_segments=segs, mma_rope=(cos, sin))
        with aki_b200.mma_context(segs):
            ret = mod(hidden_states=hidden, attention_mask=None, position_ids=pos, past_key_value=None,
                      output_attentions=False, use_cache=False)
    assert isinstance(ret, tuple) and len(ret) == 3 and ret[1] is None and ret[2] is None
    assert torch.equal(ret[0], ref)


def test_sft_loss_and_parameter_gradients_match_hf_eager_on_reference_mask():
    """Model-level parity of the training step (codes/open_flamingo/src/aki.py:125-130, train/train_utils.py:143-158):
    a 2-layer random-init Phi-3 in the reference's amp_bf16 mode (fp32 master weights, bf16 autocast), HF eager attention
    fed the reference's inverted (B,1,T,T) mask vs the same weights with every self_attn replaced by AkiMMAAttention fed
    mma_segments.  Loss and the gradients of qkv_proj / o_proj / an MLP weight of both layers within bf16 tolerance."""
    import copy
    import aki_b200
    from aki_b200 import ops
    from aki_b200.model import AkiPhi3SFT, phi35_mini_config
    g = np.load(os.path.join(GOLDEN, "attn_cfg1_small.npz"))
    cfg = phi35_mini_config(num_layers=2, short_factor=g["short_factor"], long_factor=g["long_factor"])
    sft = AkiPhi3SFT(cfg, device=dev, seed=0)
    from transformers import Phi3ForCausalLM
    hf = Phi3ForCausalLM(cfg).to(dev)
    sd = {k: v for k, v in sft.lm.state_dict().items()}
    hf.load_state_dict(sd)                                 # identical state-dict keys: the drop-in contract (b1)
    B, L, N = 2, 90, 32
    lang, am = Hp.make_prompt(B, L, N, 1, pad_right=11)
    segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
    T = segs.T
    S = O.segments_ref(lang, am, N, Hp.MEDIA_ID)
    m4 = torch.from_numpy(O.expand_segments_to_4d(S)).to(dev)
    add = O.invert_4d_mask(m4, torch.bfloat16).float()     # a7: finfo(bf16).min of the embeds dtype, stored in fp32
    torch.manual_seed(7)
    embeds = (torch.randn(B, T, 3072, device=dev) * 0.5).to(torch.bfloat16)
    labels = torch.randint(3, 31000, (B, T), device=dev)
    q_end = segs.q_end.to(torch.int64)
    idx = torch.arange(T, device=dev)[None]
    labels = torch.where((idx >= q_end[:, None]) & (idx < segs.seq_len[:, None]), labels, torch.full_like(labels, -100))
    pos = torch.arange(T, device=dev)[None].expand(B, -1)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref_loss = hf(inputs_embeds=embeds, attention_mask=add, position_ids=pos, labels=labels, use_cache=False).loss
    ref_loss.backward()
    loss = sft(embeds, segs, labels)
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 2e-2 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    names = [f"model.layers.{l}.self_attn.{w}.weight" for l in range(2) for w in ("qkv_proj", "o_proj")] + \
            ["model.layers.0.mlp.down_proj.weight", "model.layers.1.mlp.gate_up_proj.weight"]
    got = dict(sft.lm.named_parameters()); ref = dict(hf.named_parameters())
    for n in names:
        a_, b_ = got[n].grad.float(), ref[n].grad.float()
        rel = float((a_ - b_).norm() / b_.norm())
        cos_ = float((a_ * b_).sum() / (a_.norm() * b_.norm()))
        assert rel < 6e-2 and cos_ > 0.998, (n, rel, cos_)


@pytest.mark.gpu
def test_trainable_splice_gradient_matches_torch_cat_reference():
    """Training path of the splice: values equal the gather kernel, gradients equal autograd through
    the reference's torch.cat construction (vlm.py:539-545)."""
    import aki_b200
    from aki_b200 import ops
    dev = torch.device("cuda", 0)
    B, L, N, E = 2, 40, 16, 64
    g = np.random.default_rng(5)
    lang = g.integers(3, 31000, size=(B, L)).astype(np.int64)
    lang[0, 3] = Hp.MEDIA_ID; lang[1, 7] = Hp.MEDIA_ID; lang[1, 20] = Hp.MEDIA_ID
    am = np.ones_like(lang)
    ids = torch.from_numpy(lang).to(dev); amt = torch.from_numpy(am).to(dev)
    emb = torch.randn(B, L, E, device=dev).to(torch.bfloat16).requires_grad_(True)
    vis = torch.randn(B, 2, N, E, device=dev).to(torch.bfloat16).requires_grad_(True)
    segs = ops.build_segments(ids, amt, N, Hp.MEDIA_ID)
    out = ops.splice_trainable(emb, vis, segs, 0.0)
    w = torch.randn_like(out)
    (out.float() * w.float()).sum().backward()
    # reference: per-sample torch.cat, right-padded with zeros
    emb_r = emb.detach().clone().requires_grad_(True); vis_r = vis.detach().clone().requires_grad_(True)
    rows = []
    for b in range(B):
        parts, k, prev = [], 0, 0
        for pos in np.where(lang[b] == Hp.MEDIA_ID)[0]:
            parts += [emb_r[b, prev:pos], vis_r[b, k]]; prev = pos + 1; k += 1
        parts.append(emb_r[b, prev:])
        row = torch.cat(parts, 0)
        rows.append(torch.cat([row, row.new_zeros(segs.T - row.shape[0], E)], 0))
    ref = torch.stack(rows)
    assert torch.equal(out.detach(), ref.detach())
    (ref.float() * w.float()).sum().backward()
    assert torch.allclose(emb.grad.float(), emb_r.grad.float(), atol=1e-6)
    assert torch.allclose(vis.grad.float(), vis_r.grad.float(), atol=1e-6)


@pytest.mark.gpu
def test_cuda_graph_decode_equals_eager_decode():
    """The graph-replayed decode step (device-resident write row / key count / position) must produce the same tokens
    and the same cache contents as the host-driven step (aki_generation.py:72-84 semantics)."""
    import aki_b200
    from aki_b200.model import AkiPhi3Runner, phi35_mini_config
    dev = torch.device("cuda", 0)
    runner = AkiPhi3Runner(phi35_mini_config(num_layers=2), device=dev, seed=0)
    B, T, n_new = 2, 70, 6
    emb = (torch.randn(B, T, 3072, generator=torch.Generator().manual_seed(1)) * 0.05).to(torch.bfloat16).to(dev)
    outs, caches = [], []
    for graphed in (False, True):
        cache = runner.new_cache(B, T + n_new + 4)
        logits = runner.prefill(emb, None, cache)
        tok = logits[:, -1].argmax(-1, keepdim=True)
        toks = [tok]
        for _ in range(n_new):
            if graphed:
                tok = runner.decode_step_graphed(tok, cache, fused=False)
            else:
                tok = runner.decode_step(tok, cache)[:, -1].argmax(-1, keepdim=True)
            toks.append(tok)
        assert cache.get_seq_length() == T + n_new
        outs.append(torch.cat(toks, 1).cpu()); caches.append(cache)
    assert torch.equal(outs[0], outs[1])
    n = T + n_new
    for l in range(2):
        assert torch.equal(caches[0].k[l][:, :, :n], caches[1].k[l][:, :, :n])
        assert torch.equal(caches[0].v[l][:, :, :n], caches[1].v[l][:, :, :n])


@pytest.mark.gpu
def test_cuda_graph_decode_uses_the_factor_set_of_the_current_position():
    """A cache allocated beyond original_max_position_embeddings (4096) with a short prompt: the eager step, the
    prefill and the reference (modeling_rope_utils.py:47-80) all use the SHORT longrope factors while the running
    position stays below 4096 -- so must the graph-replayed step (it used to pick the set from the cache capacity).
    Also: capture needs three free rows and a full cache is refused before anything is enqueued."""
    from aki_b200.model import AkiPhi3Runner, phi35_mini_config
    dev = torch.device("cuda", 0)
    short = 1.0 + np.arange(48) / 96.0
    long = 2.0 + np.arange(48) / 12.0
    runner = AkiPhi3Runner(phi35_mini_config(num_layers=2, short_factor=short, long_factor=long), device=dev, seed=0)
    B, T, n_new = 1, 40, 4
    emb = (torch.randn(B, T, 3072, generator=torch.Generator().manual_seed(2)) * 0.05).to(torch.bfloat16).to(dev)
    outs = []
    for graphed in (False, True):
        cache = runner.new_cache(B, 4200)
        logits = runner.prefill(emb, None, cache)
        tok = logits[:, -1].argmax(-1, keepdim=True)
        toks = [tok]
        for _ in range(n_new):
            tok = runner.decode_step_graphed(tok, cache, fused=False) if graphed else runner.decode_step(tok, cache)[:, -1].argmax(-1, keepdim=True)
            toks.append(tok)
        outs.append((torch.cat(toks, 1).cpu(), cache.k[1][:, :, :T + n_new].clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    # capacity: a cache with no free row refuses the step up front
    tight = runner.new_cache(B, T + 1)
    tok = runner.prefill(emb, None, tight)[:, -1].argmax(-1, keepdim=True)
    with pytest.raises(ValueError):
        runner.decode_step_graphed(tok, tight)              # needs 3 free rows to capture
    runner.decode_step(tok, tight)
    with pytest.raises(ValueError):
        runner.decode_step(tok, tight)                      # full


@pytest.mark.gpu
def test_left_padded_batched_decode_skips_the_pad_rows():
    """Batched generate pads prompts on the LEFT (aki.py:172-182); the pad rows sit at the front of the KV cache.  The
    decode steps must not attend them (the reference's 2-D mask zeroes them): per-sample kv_start from the spliced mask."""
    import aki_b200
    from aki_b200 import ops
    H, D, tcap = 32, 96, 300
    g = torch.Generator().manual_seed(5)
    q = torch.randn(2, H, D, generator=g).to(torch.bfloat16)
    kc = torch.randn(2, H, tcap, D, generator=g).to(torch.bfloat16)
    vc = torch.randn(2, H, tcap, D, generator=g).to(torch.bfloat16)
    starts, lens = [37, 0], [200, 200]
    ref = torch.cat([O.decode_attention(q[b:b + 1].float()[:, :, None], kc[b:b + 1, :, starts[b]:].float(),
                                        vc[b:b + 1, :, starts[b]:].float(), [lens[b] - starts[b]], D ** -0.5)[:, 0]
                     for b in range(2)])
    out = ops.decode_op(q.to(dev), kc.to(dev), vc.to(dev), torch.tensor(lens, dtype=torch.int32, device=dev), max(lens),
                        D ** -0.5, torch.tensor(starts, dtype=torch.int32, device=dev))
    assert float((out.float().cpu() - ref).abs().max()) < 4e-3
    cache = aki_b200.AkiKVCache(1, 2, H, D, tcap, dev)
    m = torch.ones(2, 50, dtype=torch.int64, device=dev); m[0, :37] = 0
    cache.set_key_start(m)
    assert cache.kv_start.tolist() == [37, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("B,K,N,mode", [(1, 3072, 9216, "norm"), (8, 3072, 9216, "norm"), (3, 3072, 3072, "residual"),
                                        (8, 3072, 8192, "swiglu"), (5, 8192, 3072, "residual"), (8, 3072, 32064, "norm"),
                                        (2, 3072, 3072, "plain")])
def test_skinny_linear_matches_the_layers_it_replaces(B, K, N, mode):
    """aki_mma_skinny_linear vs the HF modules it stands in for at decode size (modeling_phi3.py:49-64, 295-335):
    Phi3RMSNorm + nn.Linear, nn.Linear + residual add, gate_up_proj + SiLU gate -- same rounding points, so the difference
    is accumulation order only."""
    from aki_b200 import ops
    from transformers.models.phi3.modeling_phi3 import Phi3RMSNorm
    dev_ = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(B * 7 + N)
    x = torch.randn(B, K, generator=g).to(torch.bfloat16).to(dev_)
    w = (torch.randn((2 * N if mode == "swiglu" else N), K, generator=g) * K ** -0.5).to(torch.bfloat16).to(dev_)
    if mode == "norm":
        norm = Phi3RMSNorm(K, eps=1e-5).to(dev_).to(torch.bfloat16)
        norm.weight.data = (1 + 0.1 * torch.randn(K, generator=g)).to(torch.bfloat16).to(dev_)
        ref = torch.nn.functional.linear(norm(x), w)
        got = ops.skinny_linear(x, w, norm.weight, 1e-5)
    elif mode == "residual":
        res = torch.randn(B, N, generator=g).to(torch.bfloat16).to(dev_)
        ref = res + torch.nn.functional.linear(x, w)
        got = ops.skinny_linear(x, w, residual=res)
    elif mode == "swiglu":
        gu = torch.nn.functional.linear(x, w)
        gate, up = gu.chunk(2, dim=-1)
        ref = up * torch.nn.functional.silu(gate)
        got = ops.skinny_linear(x, w, swiglu=True)
    else:
        ref = torch.nn.functional.linear(x, w)
        got = ops.skinny_linear(x, w)
    assert got.shape == ref.shape
    d = (got.float() - ref.float()).abs().max().item()
    assert d <= 2e-2 * max(1.0, ref.float().abs().max().item()), (d, ref.float().abs().max().item())


@pytest.mark.gpu
def test_fused_decode_step_matches_the_hf_layer_path():
    """SURVEY 8 f-1 for decode: the fused step (7 launches per layer) against the step through HF's Phi3DecoderLayer
    objects (RMSNorm / Linear / MLP in ATen + cuBLAS around the same attention kernels): logits of every step, the greedy
    tokens, and the cache contents; then the graph-replayed fused step against the eager fused step."""
    from aki_b200.model import AkiPhi3Runner, phi35_mini_config
    dev_ = torch.device("cuda", 0)
    runner = AkiPhi3Runner(phi35_mini_config(num_layers=2), device=dev_, seed=0)
    B, T, n_new = 3, 50, 5
    emb = (torch.randn(B, T, 3072, generator=torch.Generator().manual_seed(3)) * 0.05).to(torch.bfloat16).to(dev_)
    caches = [runner.new_cache(B, T + n_new + 3) for _ in range(3)]
    toks = []
    for c in caches:
        toks.append(runner.prefill(emb, None, c)[:, -1].argmax(-1, keepdim=True))
    assert torch.equal(toks[0], toks[1])
    tok = toks[0]
    for step in range(n_new):
        ref = runner.decode_step(tok, caches[0]).float()
        got = runner.decode_step_fused(tok, caches[1]).float()
        scale = float(ref.abs().max())
        assert float((got - ref).abs().max()) < 3e-2 * max(scale, 1.0), (step, float((got - ref).abs().max()), scale)
        nxt_graph = runner.decode_step_graphed(tok, caches[2], fused=True)
        assert torch.equal(nxt_graph, got[:, -1].argmax(-1, keepdim=True)), step
        tok = ref[:, -1].argmax(-1, keepdim=True)
    n = T + n_new
    for l in range(2):
        a_, b_ = caches[0].k[l][:, :, :n].float(), caches[1].k[l][:, :, :n].float()
        assert float((a_ - b_).abs().max()) < 3e-2 * max(1.0, float(a_.abs().max()))
        assert torch.equal(caches[1].k[l][:, :, :n], caches[2].k[l][:, :, :n])


@pytest.mark.gpu
@pytest.mark.parametrize("M,K,with_res", [(1, 3072, False), (37, 3072, True), (5240, 3072, True), (64, 1024, True),
                                          (9, 4096, False)])
def test_add_rmsnorm_matches_the_eager_layer_code(M, K, with_res):
    """aki_mma_add_rmsnorm vs `residual + hidden_states` followed by Phi3RMSNorm.forward (modeling_phi3.py:49-64, 317-335)."""
    from transformers.models.phi3.modeling_phi3 import Phi3RMSNorm
    from aki_b200 import ops
    dev_ = torch.device("cuda", 0)
    g = torch.Generator(device=dev_).manual_seed(M + K)
    x = torch.randn(M, K, generator=g, device=dev_).to(torch.bfloat16)
    res = (torch.randn(M, K, generator=g, device=dev_) * 3).to(torch.bfloat16) if with_res else None
    norm = Phi3RMSNorm(K, eps=1e-5).to(dev_).to(torch.bfloat16)
    with torch.no_grad():
        norm.weight.copy_((1 + 0.1 * torch.randn(K, generator=g, device=dev_)).to(torch.bfloat16))
        h_ref = res + x if with_res else x
        y_ref = norm(h_ref)
    res_in = res.clone() if with_res else None
    h, y = ops.add_rmsnorm(x, norm.weight, 1e-5, residual=res_in)
    assert torch.equal(h, h_ref)
    if with_res:
        assert h.data_ptr() == res_in.data_ptr()            # the sum overwrote the residual stream in place
    # the row sum of squares is accumulated in a different order: allow one bf16 ulp on a handful of elements
    d = (y.float() - y_ref.float()).abs()
    assert float(d.max()) <= 2 ** -7 * float(y_ref.float().abs().max())
    assert float((d > 0).float().mean()) < 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("M,N", [(1, 8192), (77, 8192), (5240, 8192), (16, 64)])
def test_swiglu_matches_phi3_mlp(M, N):
    """aki_mma_swiglu vs the middle of Phi3MLP.forward: gate, up = gate_up.chunk(2); up * silu(gate) (modeling_phi3.py:295-306)."""
    from aki_b200 import ops
    dev_ = torch.device("cuda", 0)
    gu = (torch.randn(M, 2 * N, generator=torch.Generator(device=dev_).manual_seed(N + M), device=dev_) * 2).to(torch.bfloat16)
    gate, up = gu.chunk(2, dim=-1)
    ref = up * torch.nn.functional.silu(gate)
    got = ops.swiglu(gu)
    d = (got.float() - ref.float()).abs()
    assert float(d.max()) <= 2 ** -7 * float(ref.float().abs().max())
    assert float((d > 0).float().mean()) < 0.01


@pytest.mark.gpu
def test_fused_prefill_matches_the_hf_layer_path():
    """SURVEY 8 f-1 for prefill: the pass with fused residual + RMSNorm / SiLU-gate kernels against the pass through HF's
    Phi3DecoderLayer objects -- logits, caches, with image segments and a ragged batch; the caller's embeddings stay intact."""
    from aki_b200 import ops
    from aki_b200.model import AkiPhi3Runner, phi35_mini_config
    dev_ = torch.device("cuda", 0)
    runner = AkiPhi3Runner(phi35_mini_config(num_layers=3), device=dev_, seed=0)
    B, L, N = 3, 200, 144
    lang, am = Hp.make_prompt(B, L, N, 2, pad_right=31)
    segs = ops.build_segments(torch.from_numpy(lang).to(dev_), torch.from_numpy(am).to(dev_), N, Hp.MEDIA_ID)
    T = segs.T
    emb = (torch.randn(B, T, 3072, generator=torch.Generator().manual_seed(5)) * 0.05).to(torch.bfloat16).to(dev_)
    keep = emb.clone()
    c_ref, c_fused = runner.new_cache(B, T + 4), runner.new_cache(B, T + 4)
    ref = runner.prefill(emb, segs, c_ref, last_only=False, fused=False).float()
    got = runner.prefill(emb, segs, c_fused, last_only=False, fused=True).float()
    assert torch.equal(emb, keep)
    valid = torch.arange(T, device=dev_)[None] < segs.seq_len[:, None]
    err = float((got - ref)[valid].abs().max())
    assert err < 3e-2 * max(1.0, float(ref[valid].abs().max())), err
    for l in range(3):
        a_, b_ = c_ref.k[l][:, :, :T].float(), c_fused.k[l][:, :, :T].float()
        assert float((a_ - b_).abs().max()) < 3e-2 * max(1.0, float(a_.abs().max()))


@pytest.mark.gpu
def test_hf_generate_with_fused_elementwise_layers():
    """aki_b200.fuse_phi3_elementwise on a Phi3ForCausalLM driven by HF generate(): same greedy tokens and logits as the
    same model without the swap (module-level f-1 for callers that keep HF's layer objects); the swap steps aside under
    autograd (training), and unfuse restores the original forwards."""
    import copy
    import aki_b200
    from aki_b200 import ops
    cfg, model = _small_lm(seed=4)
    aki_b200.replace_phi3_attention(model)
    fused = copy.deepcopy(model)
    n_mod = aki_b200.fuse_phi3_elementwise(fused)
    assert n_mod == 2 * 2 + 1 + 2                        # per layer two norms + one MLP, plus the final norm
    assert aki_b200.fuse_phi3_elementwise(fused) == 0    # idempotent
    B, L, N, n_new = 2, 90, 32, 6
    lang, am = Hp.make_prompt(B, L, N, 2)
    segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
    T = segs.T
    torch.manual_seed(11)
    embeds = (torch.randn(B, T, 3072, device=dev) * 0.5).to(torch.bfloat16)
    outs = []
    for m in (model, fused):
        cache = aki_b200.AkiKVCache(2, B, 32, 96, t_cap=T + n_new + 1, device=dev)
        with aki_b200.mma_context(segs):
            outs.append(m.generate(inputs_embeds=embeds, attention_mask=segs.spliced_mask_2d(), max_new_tokens=n_new,
                                   do_sample=False, use_cache=True, return_dict_in_generate=True, output_logits=True,
                                   pad_token_id=0, past_key_values=cache))
    ref, got = (torch.stack(o.logits, 1).float() for o in outs)
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) < 0.03 * max(scale, 1.0)
    top2 = ref[:, 0].topk(2, dim=-1).values                 # first generated token: same prompt state in both runs
    clear = (top2[..., 0] - top2[..., 1]) > 0.05 * scale
    assert torch.equal(ref[:, 0].argmax(-1)[clear], got[:, 0].argmax(-1)[clear])
    # under autograd the original forwards run: gradients flow as before
    x = (torch.randn(1, 8, 3072, device=dev) * 0.1).to(torch.bfloat16).requires_grad_(True)
    y = fused.model.layers[0].mlp(fused.model.layers[0].post_attention_layernorm(x))
    y.float().sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad.float()).all()
    assert aki_b200.unfuse_phi3_elementwise(fused) == n_mod
    assert not hasattr(fused.model.norm, "_aki_orig_forward")


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V", [(2, 33, 512), (4, 656, 32064), (1, 5, 40)])
def test_fused_cross_entropy_matches_hf_loss(B, T, V):
    """aki_mma_cross_entropy_{fwd,bwd} vs what Phi3ForCausalLM(labels=...) computes (logits.float(), shift, CrossEntropyLoss
    with ignore_index=-100; the loss AKI.forward returns, aki.py:125-130): loss and the bf16 gradient w.r.t. the logits."""
    from aki_b200 import ops
    g = torch.Generator(device=dev).manual_seed(V + T)
    logits = (torch.randn(B, T, V, generator=g, device=dev) * 3).to(torch.bfloat16)
    labels = torch.randint(0, V, (B, T), generator=g, device=dev)
    labels[:, : T // 3] = -100                                  # prompt part ignored, as the SFT collate does
    labels[0, -1] = -100
    a = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(a[:, :-1].float().reshape(-1, V), labels[:, 1:].reshape(-1), ignore_index=-100)
    (ref * 1.7).backward()
    b = logits.clone().requires_grad_(True)
    got = ops.cross_entropy_shifted(b, labels)
    (got * 1.7).backward()
    assert abs(float(got) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert b.grad.dtype == torch.bfloat16 and b.grad.shape == a.grad.shape
    d = (b.grad.float() - a.grad.float()).abs()
    assert float(d.max()) <= 2 ** -7 * float(a.grad.float().abs().max()) + 1e-12     # one bf16 ulp of the largest entry
    assert torch.equal(b.grad[:, -1], torch.zeros_like(b.grad[:, -1]))              # the last position has no target
    assert torch.equal(b.grad[:, : T // 3 - 1], torch.zeros_like(b.grad[:, : T // 3 - 1]))


@pytest.mark.gpu
@pytest.mark.parametrize("M,K,with_a", [(5, 3072, True), (2624, 3072, True), (33, 1024, False)])
def test_add_rmsnorm_amp_forward_and_backward_match_eager_autograd(M, K, with_a):
    """Training layout (amp_bf16): fp32 residual stream + bf16 branch output -> fp32 sum, fp32 Phi3RMSNorm, bf16 cast of the
    next Linear's operand; gradients w.r.t. the stream, the branch output and the norm weight against eager autograd."""
    from transformers.models.phi3.modeling_phi3 import Phi3RMSNorm
    from aki_b200 import ops
    g = torch.Generator(device=dev).manual_seed(M + K)
    h0 = torch.randn(M, K, generator=g, device=dev) * 2
    a0 = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16) if with_a else None
    norm = Phi3RMSNorm(K, eps=1e-5).to(dev)                                   # fp32 master weight
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(K, generator=g, device=dev))
    gx = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16)        # gradient arriving at the bf16 operand
    gh = torch.randn(M, K, generator=g, device=dev) * 0.1                      # gradient arriving at the residual stream

    def run(fused):
        h = h0.clone().requires_grad_(True)
        a = a0.clone().requires_grad_(True) if with_a else None
        norm.weight.grad = None
        if fused:
            hn, x = ops.add_rmsnorm_amp(h, a, norm.weight, 1e-5)
        else:
            hn = h + a if with_a else None
            x = norm(hn if with_a else h).to(torch.bfloat16)
        outs, grads = [x], [gx]
        if with_a:
            outs.append(hn); grads.append(gh)
        torch.autograd.backward(outs, grads)
        return x.detach(), (hn.detach() if with_a else None), h.grad, (a.grad if with_a else None), norm.weight.grad.clone()

    x_r, hn_r, dh_r, da_r, dw_r = run(False)
    x_f, hn_f, dh_f, da_f, dw_f = run(True)
    if with_a:
        assert torch.equal(hn_f, hn_r)
    d = (x_f.float() - x_r.float()).abs()
    assert float(d.max()) <= 2 ** -7 * float(x_r.float().abs().max()) and float((d > 0).float().mean()) < 0.01
    rel = lambda u, v: float((u.float() - v.float()).norm() / v.float().norm())
    assert rel(dh_f, dh_r) < 1e-5
    assert rel(dw_f, dw_r) < 1e-4
    if with_a:
        assert da_f.dtype == torch.bfloat16 and rel(da_f, da_r) < 4e-3


@pytest.mark.gpu
def test_swiglu_train_backward_matches_eager_autograd():
    from aki_b200 import ops
    g = torch.Generator(device=dev).manual_seed(3)
    gu0 = (torch.randn(77, 2 * 8192, generator=g, device=dev) * 2).to(torch.bfloat16)
    go = torch.randn(77, 8192, generator=g, device=dev).to(torch.bfloat16)
    a = gu0.clone().requires_grad_(True)
    gate, up = a.chunk(2, dim=-1)
    (up * torch.nn.functional.silu(gate)).backward(go)
    b = gu0.clone().requires_grad_(True)
    ops.swiglu_train(b).backward(go)
    d = (b.grad.float() - a.grad.float()).abs()
    assert float(d.max()) <= 2 ** -6 * float(a.grad.float().abs().max())
    assert float((d > 0).float().mean()) < 0.02


@pytest.mark.gpu
def test_sft_step_with_fused_layer_kernels_matches_the_hf_layer_objects():
    """AkiPhi3SFT with the fused training-layout kernels (residual add + RMSNorm + cast, SiLU gate, cross-entropy) against
    the same model stepping through HF's Phi3DecoderLayer objects and F.cross_entropy: loss and every parameter gradient."""
    import aki_b200
    from aki_b200 import ops
    from aki_b200.model import AkiPhi3SFT, phi35_mini_config
    model = AkiPhi3SFT(phi35_mini_config(num_layers=2), device=dev, seed=1)
    B, L, N = 2, 120, 48
    lang, am = Hp.make_prompt(B, L, N, 1, pad_right=17)
    segs = ops.build_segments(torch.from_numpy(lang).to(dev), torch.from_numpy(am).to(dev), N, Hp.MEDIA_ID)
    T = segs.T
    torch.manual_seed(2)
    emb = torch.randn(B, T, 3072, device=dev) * 0.05
    labels = torch.randint(0, 32000, (B, T), device=dev); labels[:, :40] = -100
    res = []
    for fused in (False, True):
        model.fused_layers = fused; model.fused_ce = fused
        model.zero_grad(set_to_none=True)
        loss = model(emb, segs, labels)
        loss.backward()
        res.append((float(loss), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    (l0, g0), (l1, g1) = res
    assert abs(l1 - l0) < 2e-3 * abs(l0), (l0, l1)
    assert g0.keys() == g1.keys() and len(g0) > 10
    for n in g0:
        rel = float((g1[n].float() - g0[n].float()).norm() / (g0[n].float().norm() + 1e-12))
        assert rel < 3e-2, (n, rel)
