"""patch_adafactor — drop-in for /root/reference/modules/util/optimizer/adafactor_extensions.py:115-118.

Same contract: rebinds `optimizer.step(closure=None)` of a `transformers.Adafactor`, adds
`optimizer.step_parameter(p, group, i)` and `optimizer.stochastic_rounding`.  The step itself (:17-93 — factored /
unfactored second moment, relative step / warm-up / fixed lr, parameter scaling by RMS(p), update clipping, optional
first moment, weight decay) is ONE multi-tensor launch (otb_adafactor_multi, csrc/optim_factored.cu).

One reference quirk is kept on purpose: for bf16 parameters the reference calls copy_stochastic_ and then, unconditionally,
`p.copy_(p_data_fp32)` (:88-91), so its "stochastic rounding" is overwritten by round-to-nearest.  Parity means doing the
same: `stochastic_rounding` is recorded on the optimizer and has no effect on the result, exactly as in the reference.
"""
from __future__ import annotations

from . import fused_factored


def patch_adafactor(optimizer, stochastic_rounding: bool):
    for group in optimizer.param_groups:
        for key in ("eps", "clip_threshold", "decay_rate", "beta1", "weight_decay", "scale_parameter", "relative_step", "warmup_init"):
            if key not in group:
                raise TypeError(f"patch_adafactor expects a transformers.Adafactor optimizer (param group has no '{key}')")
    driver = fused_factored.install(optimizer, "adafactor", stochastic_rounding)
    optimizer.register_load_state_dict_post_hook(lambda opt: driver._tables.clear())
    return driver
