"""`init_model_parameters` — mirror of /root/reference/modules/util/optimizer_util.py:50-65: hangs the parameter
collection, the optimizer, the EMA and the group-name mapping on the model, consuming the saved state dicts once.
EMA is created first here so that its update can ride in the optimizer's fused launch."""
from __future__ import annotations

from . import create
from .NamedParameterGroup import NamedParameterGroupCollection


def init_model_parameters(model, parameters: NamedParameterGroupCollection, train_device=None, *, fused_ema: bool = True,
                          fused_clip: bool = True):
    config = model.train_config
    model.parameters = parameters
    model.ema = create.create_ema(parameters.parameters(), getattr(model, "ema_state_dict", None), config)
    model.ema_state_dict = None
    progress = getattr(model, "train_progress", None)
    fold_ema = model.ema is not None and fused_ema and progress is not None
    model.optimizer = create.create_optimizer(
        parameters, getattr(model, "optimizer_state_dict", None), config,
        ema=model.ema if fold_ema else None,
        ema_step_fn=(lambda: progress.global_step // config.gradient_accumulation_steps) if fold_ema else None,
        clip_grad_norm=config.clip_grad_norm if fused_clip else None)
    model.optimizer_state_dict = None
    model.param_group_mapping = parameters.unique_name_mapping
