"""patch_adam — drop-in for /root/reference/modules/util/optimizer/adam_extensions.py:203-206.

Identical to patch_adamw except for the weight decay: Adam adds `weight_decay * p` to the gradient
(L2, adam_extensions.py:74-75) where AdamW scales the parameter (`p *= 1 - lr*wd`).
"""
from __future__ import annotations

from torch.optim import Adam

from . import fused_adam


def patch_adam(optimizer: Adam, stochastic_rounding: bool, *, ema=None, ema_step_fn=None,
               clip_grad_norm: float | None = None):
    return fused_adam.install(optimizer, stochastic_rounding, decoupled=False, ema=ema, ema_step_fn=ema_step_fn,
                              clip_grad_norm=clip_grad_norm)
