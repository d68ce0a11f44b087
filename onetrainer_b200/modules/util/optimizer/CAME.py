"""CAME — drop-in for /root/reference/modules/util/optimizer/CAME.py (class CAME, :12-198): same constructor, defaults,
param-group keys, state entries, `step(closure=None)` and `step_parameter(p, group, i)`.

The step (:80-176 — factored second moment, update clipping, first moment, confidence-guided instability factors, weight
decay and the update applied through add_stochastic_ for bf16 parameters) is ONE multi-tensor launch
(otb_came_multi, csrc/optim_factored.cu) instead of ~45 ATen launches and full-size fp32 temporaries per tensor.
Stochastic-rounding bits come from this package's Philox stream keyed by (torch.initial_seed(), call counter, global element
index) instead of torch.randint_like's global-generator draws.
"""
from __future__ import annotations

import torch
import torch.optim

from . import fused_factored


class CAME(torch.optim.Optimizer):
    def __init__(self, params, lr=None, eps=(1e-30, 1e-16), clip_threshold=1.0, betas=(0.9, 0.999, 0.9999), weight_decay=0.0,
                 stochastic_rounding=False):
        assert lr > 0.
        assert all(0. <= beta <= 1. for beta in betas)
        defaults = {"lr": lr, "eps": eps, "clip_threshold": clip_threshold, "betas": betas, "weight_decay": weight_decay}
        self.stochastic_rounding = stochastic_rounding
        super().__init__(params, defaults)
        self.otb_driver = fused_factored.FactoredDriver(self, "came")
        self.register_load_state_dict_post_hook(lambda opt: opt.otb_driver._tables.clear())

    @property
    def supports_memory_efficient_fp16(self):
        return True

    @property
    def supports_flat_params(self):
        return False

    def step_parameter(self, p, group, i):
        self.otb_driver.step_parameter(p, group, i)

    def step(self, closure=None):
        return self.otb_driver.step(closure)
