"""patch_adamw — drop-in for /root/reference/modules/util/optimizer/adamw_extensions.py:202-205.

Same contract: rebinds `optimizer.step(closure=None)`, adds `optimizer.step_parameter(p, group, i)`
and `optimizer.stochastic_rounding`.  The arithmetic (decoupled decay, lerp, addcmul, sqrt/div/add
denominators, addcdiv, bf16 stochastic rounding by the int32 bit trick) is reproduced rounding by
rounding inside ONE multi-tensor CUDA launch (onetrainer_b200/csrc/optim.cu) instead of ~13 ATen
launches per tensor.  Extra keyword arguments opt into fusions the reference cannot express:
`ema` + `ema_step_fn` (EMA update in the same launch) and `clip_grad_norm` (global-norm clip folded in).
"""
from __future__ import annotations

from torch.optim import AdamW

from . import fused_adam


def patch_adamw(optimizer: AdamW, stochastic_rounding: bool, *, ema=None, ema_step_fn=None,
                clip_grad_norm: float | None = None):
    return fused_adam.install(optimizer, stochastic_rounding, decoupled=True, ema=ema, ema_step_fn=ema_step_fn,
                              clip_grad_norm=clip_grad_norm)
