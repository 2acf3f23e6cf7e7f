"""Inference-time fusion of the element-wise work of HF's Phi-3 decoder layers around the attention op ("next" row f-1 of
SURVEY 8) for callers that keep HF's own layer objects -- the reference drives `Phi3ForCausalLM` through HF `generate`
(codes/open_flamingo/src/aki.py:184-191).  `fuse_phi3_elementwise(model)` rebinds, per instance,
  Phi3RMSNorm.forward -> aki_mma_add_rmsnorm            (modeling_phi3.py:49-64: 8 ATen kernels -> 1)
  Phi3MLP.forward     -> gate_up_proj, aki_mma_swiglu, down_proj   (modeling_phi3.py:295-306: chunk + silu + mul -> 1)
with the same rounding points.  The fused path is taken only when autograd is off and the input is a bf16 CUDA tensor;
otherwise the module's original forward runs (training, fp32, CPU), so the swap is safe to leave in place."""
from __future__ import annotations

import types

import torch

from . import ops


def _fusable(x: torch.Tensor, width: int) -> bool:
    return (not torch.is_grad_enabled()) and x.is_cuda and x.dtype == torch.bfloat16 and width % 256 == 0 and width <= 4096


def _rmsnorm_forward(self, hidden_states):
    if _fusable(hidden_states, hidden_states.shape[-1]) and self.weight.dtype == torch.bfloat16:
        return ops.add_rmsnorm(hidden_states, self.weight, self.variance_epsilon)[1]
    return self._aki_orig_forward(hidden_states)


def _mlp_forward(self, hidden_states):
    if _fusable(hidden_states, hidden_states.shape[-1]) and self.gate_up_proj.weight.dtype == torch.bfloat16 \
            and getattr(self.config, "hidden_act", "silu") == "silu":
        return self.down_proj(ops.swiglu(self.gate_up_proj(hidden_states)))
    return self._aki_orig_forward(hidden_states)


def fuse_phi3_elementwise(model: torch.nn.Module) -> int:
    """Rebinds the forward of every Phi3RMSNorm / Phi3MLP under `model` (idempotent).  Returns the number of modules
    swapped.  `unfuse_phi3_elementwise` restores them."""
    from transformers.models.phi3.modeling_phi3 import Phi3MLP, Phi3RMSNorm
    n = 0
    for m in model.modules():
        if hasattr(m, "_aki_orig_forward"):
            continue
        if isinstance(m, Phi3RMSNorm):
            m._aki_orig_forward = m.forward
            m.forward = types.MethodType(_rmsnorm_forward, m)
            n += 1
        elif isinstance(m, Phi3MLP):
            m._aki_orig_forward = m.forward
            m.forward = types.MethodType(_mlp_forward, m)
            n += 1
    return n


def unfuse_phi3_elementwise(model: torch.nn.Module) -> int:
    n = 0
    for m in model.modules():
        if hasattr(m, "_aki_orig_forward"):
            m.forward = m._aki_orig_forward
            del m._aki_orig_forward
            n += 1
    return n
