"""Preallocated KV cache honouring the HF past_key_values contract the reference relies on
(codes/open_flamingo/src/vlm.py:463-468 reads past_key_values[0][0].shape[2];
codes/open_flamingo/src/aki_generation.py:45-47,80): per layer (K, V), each (B, H, T_kv, D), K stored post-RoPE,
appended along dim 2.  Instead of DynamicCache's torch.cat per step, prefill and decode write rows in place into
(B, H, t_cap, D) buffers (layout chosen so one (b,h) stream is contiguous for TMA and for the decode kernel)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from transformers.cache_utils import Cache, CacheLayerMixin


class _AkiCacheLayer(CacheLayerMixin):
    """One layer's view of the preallocated buffers, so that AkiKVCache is a real transformers `Cache` (HF generate,
    masking utilities and model code only ever talk to `cache.layers[i]` / the Cache methods built on them)."""
    is_compileable = False
    is_sliding = False

    def __init__(self, owner: "AkiKVCache", idx: int):
        self._owner, self._idx = owner, idx
        super().__init__()
        self.is_initialized = True

    # the base class assigns keys / values = None in __init__; here they are live views of the owner's buffers
    keys = property(lambda self: self._owner.k[self._idx][:, :, :self._owner._len[self._idx]], lambda self, v: None)
    values = property(lambda self: self._owner.v[self._idx][:, :, :self._owner._len[self._idx]], lambda self, v: None)

    @property
    def device(self):
        return self._owner.k[self._idx].device

    def lazy_initialization(self, key_states, value_states) -> None:
        return None

    def update(self, key_states, value_states, *args, **kwargs):
        return self._owner._append(self._idx, key_states, value_states)

    def get_mask_sizes(self, query_length, *args) -> Tuple[int, int]:
        q = int(query_length.shape[0]) if isinstance(query_length, torch.Tensor) else int(query_length)
        return self._owner._len[self._idx] + q, 0

    def get_seq_length(self, *args) -> int:
        return self._owner._len[self._idx]

    def get_max_cache_shape(self) -> int:
        return self._owner.t_cap

    def reset(self) -> None:
        self._owner._len[self._idx] = 0

    def offload(self):
        raise NotImplementedError("AkiKVCache lives in HBM (180 GB per B200): offloading is not supported")

    prefetch = offload

    def reorder_cache(self, beam_idx) -> None:
        raise NotImplementedError("beam search is not part of the reference path (AKI.generate pops num_beams, aki.py:160)")


class AkiKVCache(Cache):
    def __init__(self, num_layers: int, batch: int, num_heads: int, head_dim: int, t_cap: int, device,
                 dtype=torch.bfloat16):
        super().__init__(layers=[_AkiCacheLayer(self, i) for i in range(num_layers)])
        self.t_cap = int(t_cap)
        self.k: List[torch.Tensor] = [torch.zeros(batch, num_heads, t_cap, head_dim, dtype=dtype, device=device)
                                      for _ in range(num_layers)]
        self.v: List[torch.Tensor] = [torch.zeros_like(self.k[0]) for _ in range(num_layers)]
        self._len = [0] * num_layers
        self.kv_len = torch.zeros(batch, dtype=torch.int32, device=device)   # device copy for the decode kernel
        # first visible key of every sequence: the leading pad rows of a left-padded prompt (AKI.generate pads on the
        # left, aki.py:172-182) stay in the cache but must not be attended by the decode steps
        self.kv_start = torch.zeros(batch, dtype=torch.int32, device=device)
        # CUDA-graph decode: the step reads the write row / key count from these device tensors instead of host ints
        self.past_dev = torch.zeros(batch, dtype=torch.int32, device=device)
        self.device_driven = False

    # ---- HF Cache surface --------------------------------------------------------------------------
    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self._len[layer_idx]

    def get_max_cache_shape(self, layer_idx: int = 0) -> int:
        return self.t_cap

    def get_mask_sizes(self, query_length, layer_idx: int = 0):
        return self.layers[layer_idx].get_mask_sizes(query_length)

    def __len__(self) -> int:
        return len(self.k)

    def __getitem__(self, layer_idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        n = self._len[layer_idx]
        return self.k[layer_idx][:, :, :n], self.v[layer_idx][:, :, :n]

    def __iter__(self):
        for i in range(len(self.k)):
            yield self[i]

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor, layer_idx: int, *args, **kwargs):
        """DynamicCache-compatible append of already rotated (B,H,T,D) states (used by foreign callers; the
        drop-in module writes through reserve()/commit() instead so RoPE and the copy are one kernel)."""
        return self._append(layer_idx, key_states, value_states)

    def _append(self, layer_idx: int, key_states: torch.Tensor, value_states: torch.Tensor):
        n, t = self._len[layer_idx], key_states.shape[2]
        self.reserve(layer_idx, t)
        self.k[layer_idx][:, :, n:n + t].copy_(key_states)
        self.v[layer_idx][:, :, n:n + t].copy_(value_states)
        self.commit(layer_idx, t)
        return self[layer_idx]

    # ---- in-place protocol used by AkiMMAAttention -------------------------------------------------
    def set_key_start(self, mask_2d: Optional[torch.Tensor]) -> None:
        """Number of leading invalid keys per sequence from the spliced 2-D mask (device arithmetic, no host read)."""
        if mask_2d is None:
            self.kv_start.zero_()
        else:
            self.kv_start.copy_((mask_2d.to(torch.int32).cumsum(1) == 0).sum(1).to(torch.int32))

    def reserve(self, layer_idx: int, t: int) -> int:
        n = self._len[layer_idx]
        self._check(n + t)
        if layer_idx == 0:            # one tiny fill per step; every layer then reads the same device scalar(s)
            self.kv_len.fill_(n + t)
        return n

    def commit(self, layer_idx: int, t: int) -> None:
        self._len[layer_idx] += t

    def _check(self, need: int) -> None:
        if need > self.t_cap:
            raise ValueError(f"KV cache capacity {self.t_cap} exceeded (need {need})")

    def advance_host(self, t: int = 1) -> None:
        """Host-side bookkeeping after a device-driven (graph-replayed) step appended t rows to every layer."""
        self._check(self._len[0] + t)
        self._len = [n + t for n in self._len]

    def reset(self) -> None:
        """Forget the contents (buffers are reused; nothing is freed)."""
        self._len = [0] * len(self.k)
        self.kv_len.zero_()
        self.kv_start.zero_()
        self.past_dev.zero_()

    def to_legacy_cache(self):
        return tuple(self[i] for i in range(len(self.k)))
