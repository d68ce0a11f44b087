"""Global gradient-norm clipping — replaces `nn.utils.clip_grad_norm_(self.parameters, max_norm)` at
/root/reference/modules/trainer/GenericTrainer.py:712-713.

Same semantics (L2 norm over all gradients, coef = max_norm / (norm + 1e-6) clamped to 1, gradients
scaled in place, the total norm returned as a 0-dim tensor) in three launches for ANY number of
tensors: chunked sum-of-squares, a deterministic final reduction, and the scale.  The patched
optimizer can skip the third launch entirely (patch_adamw(..., clip_grad_norm=...)).
"""
from __future__ import annotations

import torch

from ... import ops
from ... import _native as N

_tables: dict = {}


@torch.no_grad()
def clip_grad_norm_(parameters, max_norm: float, grad_prescale: float = 1.0) -> torch.Tensor:
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return torch.tensor(0.0)
    N.require_cuda(*params)
    key = tuple((id(p), p.data_ptr(), p.numel()) for p in params)     # ids are recycled: the address and size pin the tensor
    table = _tables.get(key)
    if table is None:
        if len(_tables) > 8:
            _tables.clear()
        table = ops.OptimTable([dict(p=p.data, g=p.grad) for p in params], params[0].device)
        _tables[key] = table
    table.set_grad_pointers([p.grad for p in params])
    table.upload()
    norm_coef = ops.grad_norm_multi(table, max_norm, grad_prescale)
    ops.grad_scale_multi(table, norm_coef[1:2])
    return norm_coef[0].clone()
