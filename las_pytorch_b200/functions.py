"""Mirrors of the two tensor helpers on the hot path, utils/functions.py:54-77 of the reference.

Inside the decoder loop neither is needed any more (the <sos> one-hot and the per-step psi projection are done
by the kernels); they are kept with the reference's names and semantics for callers that use them directly.
"""
from __future__ import annotations

import torch


def CreateOnehotVariable(input_x, encoding_dim=63):
    """[B, T] indices -> [B, T, encoding_dim] one-hot of the input's dtype/device (utils/functions.py:54-63).

    Built on the input's own device (the reference builds it on the CPU and copies)."""
    idx = input_x.detach().unsqueeze(2).to(torch.int64)
    onehot = torch.zeros(input_x.size(0), input_x.size(1), encoding_dim, dtype=input_x.dtype, device=input_x.device)
    return onehot.scatter_(-1, idx, 1)


def TimeDistributed(input_module, input_x):
    """Apply `input_module` to every timestep of [B, T, F] (utils/functions.py:72-77)."""
    batch_size, time_steps = input_x.size(0), input_x.size(1)
    out = input_module(input_x.contiguous().view(-1, input_x.size(-1)))
    return out.view(batch_size, time_steps, -1)
