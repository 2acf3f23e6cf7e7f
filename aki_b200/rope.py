"""Host logic of Phi-3 longrope (third-party: microsoft/Phi-3.5-mini-instruct remote code; installed equivalent
transformers/modeling_rope_utils.py:462-547 and models/phi3/modeling_phi3.py:67-131).  The 48-entry short/long
factor arrays live in the hub config.json, so they are inputs here."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from . import ops


def longrope_attention_factor(max_position_embeddings: int, original_max_position_embeddings: int) -> float:
    factor = max_position_embeddings / original_max_position_embeddings
    if factor <= 1.0:
        return 1.0
    return math.sqrt(1 + math.log(factor) / math.log(original_max_position_embeddings))


class LongRope:
    """inv_freq selection + device tables.  `tables(position_ids)` -> (cos, sin) each (B,T,head_dim/2) fp32."""

    def __init__(self, head_dim: int = 96, rope_theta: float = 10000.0, short_factor: Optional[Sequence[float]] = None,
                 long_factor: Optional[Sequence[float]] = None, max_position_embeddings: int = 131072,
                 original_max_position_embeddings: int = 4096, device="cuda"):
        half = head_dim // 2
        short = torch.ones(half) if short_factor is None else torch.tensor(list(short_factor), dtype=torch.float32)
        long = torch.ones(half) if long_factor is None else torch.tensor(list(long_factor), dtype=torch.float32)
        shape = torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim
        self.inv_freq_short = (1.0 / (short * rope_theta ** shape)).to(device)
        self.inv_freq_long = (1.0 / (long * rope_theta ** shape)).to(device)
        self.original_max = original_max_position_embeddings
        self.attention_factor = longrope_attention_factor(max_position_embeddings, original_max_position_embeddings)

    def tables(self, position_ids: torch.Tensor, max_position: Optional[int] = None):
        """long factors iff max(position_ids)+1 > original_max (modeling_rope_utils.py:47-80).  Pass max_position
        to avoid the device->host read."""
        if max_position is None:
            max_position = int(position_ids.max().item())
        inv = self.inv_freq_long if max_position + 1 > self.original_max else self.inv_freq_short
        return ops.rope_table(position_ids, inv, self.attention_factor)


def half_tables_from_hf(position_embeddings):
    """HF hands every layer (cos, sin) of shape (B,T,head_dim) in the activation dtype with the two halves
    duplicated (modeling_phi3.py:124-131); the kernels want the fp32 first half."""
    cos, sin = position_embeddings
    half = cos.shape[-1] // 2
    return cos[..., :half].float().contiguous(), sin[..., :half].float().contiguous()
