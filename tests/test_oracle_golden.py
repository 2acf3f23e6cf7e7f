"""The oracle (oracle/mma_oracle.py) against the fixtures produced by EXECUTING the reference
(oracle/gen_golden.py): masks bit-exact, labels / padding exact, attention within fp32 round-off."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mma_oracle as O


def _unpack(bits, shape):
    return np.unpackbits(bits, axis=-1)[..., : shape[-1]].reshape(shape).astype(np.int64)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "mask_reference.npz"))


def test_a1_direct_calls_bit_exact():
    d = np.load(os.path.join(GOLDEN, "mask_a1_direct.npz"))
    for k in range(int(d["n_cases"])):
        a, b, c = (int(x) for x in d[f"d{k}_args"])
        got = O.make_modality_mutual_mask(d[f"d{k}_am"], a, b, c)
        assert np.array_equal(got, d[f"d{k}_mask"].astype(np.int64)), k


def test_prepare_inputs_mask_labels_bit_exact(gold):
    media = int(gold["media_token_id"])
    for n in range(int(gold["n_cases"])):
        lang, am, N, side = gold[f"c{n}_lang"], gold[f"c{n}_am"], int(gold[f"c{n}_N"]), str(gold[f"c{n}_side"])
        ref = _unpack(gold[f"c{n}_mask_bits"], gold[f"c{n}_mask_shape"])
        got = O.prepare_inputs_for_forward(lang, am, N, media, labels=lang.copy(), padding_side=side)
        assert got.attention_mask_4d.dtype == np.int64
        assert np.array_equal(got.attention_mask_4d, ref), f"case {n}"
        assert np.array_equal(got.labels, gold[f"c{n}_labels"]), f"labels case {n}"
        assert np.array_equal(got.src_index == np.iinfo(np.int64).min, gold[f"c{n}_embed_is_pad"]), f"pad case {n}"


def test_segments_predicate_reproduces_reference_mask(gold):
    """The compact description the CUDA path consumes expands to the reference tensor bit-for-bit."""
    media = int(gold["media_token_id"])
    for n in range(int(gold["n_cases"])):
        lang, am, N = gold[f"c{n}_lang"], gold[f"c{n}_am"], int(gold[f"c{n}_N"])
        ref = _unpack(gold[f"c{n}_mask_bits"], gold[f"c{n}_mask_shape"])
        S = O.segments_ref(lang, am, N, media)
        assert np.array_equal(O.expand_segments_to_4d(S), ref), f"case {n}"
        assert O.count_allowed(S) == int(ref.sum()), f"nnz case {n}"


def test_multi_image_generalisation_reduces_to_reference(gold):
    media = int(gold["media_token_id"])
    for n in range(int(gold["n_cases"])):
        lang, am, N, side = gold[f"c{n}_lang"], gold[f"c{n}_am"], int(gold[f"c{n}_N"]), str(gold[f"c{n}_side"])
        ref = _unpack(gold[f"c{n}_mask_bits"], gold[f"c{n}_mask_shape"])
        for variant in ("contiguous", "text_only"):
            got = O.prepare_inputs_for_forward(lang, am, N, media, padding_side=side, multi_image=variant)
            assert np.array_equal(got.attention_mask_4d, ref), (n, variant)


def test_reference_raises_on_second_image():
    lang = np.array([[1, 32012, 5, 32012, 32001, 7]])
    with pytest.raises(RuntimeError):
        O.prepare_inputs_for_forward(lang, np.ones_like(lang), 3, 32012)
    got = O.prepare_inputs_for_forward(lang, np.ones_like(lang), 3, 32012, multi_image="contiguous")
    S = O.segments_ref(lang, np.ones_like(lang), 3, 32012)
    assert np.array_equal(got.attention_mask_4d, O.expand_segments_to_4d(S))
    m = got.attention_mask_4d[0, 0]
    # image 1 = rows 1..3, text 4, image 2 = rows 5..7, <|assistant|> at 8 -> q_end 9
    assert m[1, 4] == 1 and m[1, 5] == 1 and m[1, 8] == 1 and m[1, 9] == 0 and m[1, 2] == 0
    assert m[5, 8] == 1 and m[5, 6] == 0 and m[4, 5] == 0
    strict = O.prepare_inputs_for_forward(lang, np.ones_like(lang), 3, 32012, multi_image="text_only")
    assert strict.attention_mask_4d[0, 0, 1, 5] == 0 and strict.attention_mask_4d[0, 0, 1, 4] == 1


def test_attention_module_matches_installed_phi3_eager():
    g = np.load(os.path.join(GOLDEN, "attn_cfg1_small.npz"))
    torch.manual_seed(0)
    w_qkv = torch.empty(9216, 3072); w_o = torch.empty(3072, 3072)
    # same parameter order / init as gen_golden.py: o_proj is registered first in Phi3Attention
    torch.nn.init.normal_(w_o, std=0.02) if False else None
    from transformers import Phi3Config  # noqa: F401  (only to mirror RNG consumption below)
    torch.manual_seed(0)
    lin_o = torch.nn.Linear(3072, 3072, bias=False); lin_qkv = torch.nn.Linear(3072, 9216, bias=False)
    torch.nn.init.normal_(lin_o.weight, std=0.02); torch.nn.init.normal_(lin_qkv.weight, std=0.02)
    w_o, w_qkv = lin_o.weight.detach(), lin_qkv.weight.detach()
    assert np.allclose(w_qkv[::257, ::31].numpy(), g["w_qkv_sample"]) and np.allclose(w_o[::129, ::29].numpy(), g["w_o_sample"])
    T = int(g["T"]); N = int(g["N"])
    S = O.segments_ref(g["lang"], np.ones_like(g["lang"]), N, 32012)
    m4 = _unpack(g["mask_bits"], g["mask_shape"])
    assert np.array_equal(O.expand_segments_to_4d(S), m4)
    add = O.invert_4d_mask(torch.from_numpy(m4), torch.float32)
    af = O.longrope_attention_factor(131072, 4096)
    assert abs(af - float(g["attention_scaling"])) < 1e-12
    for tag in ("short", "long"):
        pos0 = int(g[f"{tag}_pos0"])
        pos = torch.arange(pos0, pos0 + T)[None]
        ext = O.select_ext_factors(pos, g["short_factor"], g["long_factor"], 4096)
        inv = O.longrope_inv_freq(96, 10000.0, ext)
        assert np.allclose(inv.numpy(), g[f"{tag}_inv_freq"], rtol=1e-6)
        cos, sin = O.rope_cos_sin(pos, inv, af)
        assert np.allclose(cos.numpy()[:, ::7], g[f"{tag}_cos"], atol=2e-4) and np.allclose(sin.numpy()[:, ::7], g[f"{tag}_sin"], atol=2e-4)
        torch.manual_seed(1)
        hidden = torch.randn(1, T, 3072)
        out, (k, v) = O.attention_module_forward(hidden, w_qkv, w_o, cos, sin, add)
        assert k.shape == (1, 32, T, 96)
        ref = g[f"{tag}_out"]
        err = np.abs(out.numpy()[:, :, ::16] - ref).max()
        assert err < 5e-4 * max(1.0, np.abs(ref).max()), (tag, err)
        assert abs(float(out.double().norm()) - float(g[f"{tag}_out_norm"])) < 1e-3 * float(g[f"{tag}_out_norm"])
