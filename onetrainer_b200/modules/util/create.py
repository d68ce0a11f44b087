"""Optimizer and EMA construction for the fused step.

Mirror of the ADAM / ADAMW / ADAFACTOR / CAME branches of `create_optimizer` (/root/reference/modules/util/create.py:481-534,
914-951), of the saved-state hand-over that follows them (:1038-1084) and of `create_ema` (:1089-1111): same argument
meaning, same defaults, same RuntimeErrors.  Differences, all loud: every other optimizer raises NotImplementedError (no
fused step on this path); the optimizer is always patched, because the fused launch IS the step here; EMA on the CPU raises
(the fused EMA lives next to the parameters in HBM).
"""
from __future__ import annotations

import torch

from ..module.EMAModule import EMAModuleWrapper
from .enum import enum_name
from .NamedParameterGroup import NamedParameterGroupCollection
from .optimizer.adam_extensions import patch_adam
from .optimizer.adamw_extensions import patch_adamw


def _or(value, default):
    return default if value is None else value


def fit_state_dict_to_groups(state_dict: dict, optimizer, parameter_group_collection: NamedParameterGroupCollection, config) -> dict:
    """A saved optimizer state made loadable into THIS run's parameter groups — what the reference does between building the
    optimizer and `optimizer.load_state_dict` (create.py:1038-1084).  A backup carries `param_group_mapping` (the groups'
    unique names, InternalModelSaverMixin.py:19-26) and `param_group_optimizer_mapping` (the optimizer each was trained with):
    a group of this run takes over the saved state of the group with the same name if the optimizer type still matches —
    its per-parameter states renumbered in the new order, its learning rates replaced by this run's — and starts empty
    otherwise, so groups may be added, dropped or reordered between runs.  Without the mapping (old backups) only the
    learning rates are refreshed, group by group.  The dict is edited in place and returned, as the reference edits it."""
    if "param_group_mapping" not in state_dict:
        for saved, mine in zip(state_dict["param_groups"], parameter_group_collection.parameters_for_optimizer(config)):
            saved["lr"], saved["initial_lr"] = mine["lr"], mine["initial_lr"]
        return state_dict
    saved_state, saved_groups = state_dict["state"], state_dict["param_groups"]
    saved_names, saved_kinds = state_dict["param_group_mapping"], state_dict["param_group_optimizer_mapping"]
    fresh_groups = optimizer.state_dict()["param_groups"]
    kind = str(config.optimizer.optimizer)
    state, groups, next_index = {}, [], 0
    for fresh, name in zip(fresh_groups, parameter_group_collection.unique_name_mapping):
        at = saved_names.index(name) if name in saved_names else -1
        if at >= 0 and saved_kinds[at] == kind:
            group = saved_groups[at]
            for slot, saved_index in enumerate(group["params"]):
                if saved_index in saved_state:
                    state[next_index] = saved_state[saved_index]
                group["params"][slot] = next_index
                next_index += 1
            group["lr"], group["initial_lr"] = fresh["lr"], fresh["initial_lr"]
        else:
            group = fresh
            group["params"][:] = range(next_index, next_index + len(group["params"]))
            next_index += len(group["params"])
        groups.append(group)
    state_dict["state"], state_dict["param_groups"] = state, groups
    return state_dict


def create_optimizer(parameter_group_collection: NamedParameterGroupCollection, state_dict: dict | None, config,
                     ema: EMAModuleWrapper | None = None, ema_step_fn=None, clip_grad_norm: float | None = None):
    """torch.optim.AdamW / Adam over `parameters_for_optimizer(config)`, step rebound to the fused launch.
    `ema` / `ema_step_fn` / `clip_grad_norm` optionally fold EMA and the global-norm clip into that launch."""
    oc = config.optimizer
    name = enum_name(oc.optimizer)
    if name in ("ADAFACTOR", "CAME"):
        return _create_factored(name, parameter_group_collection, state_dict, config, ema, clip_grad_norm)
    if name not in ("ADAMW", "ADAM"):
        raise NotImplementedError(f"optimizer {name} has no fused step on the B200 path (AdamW, Adam, Adafactor and CAME do)")
    if oc.stochastic_rounding and (oc.fused or oc.foreach):
        raise RuntimeError('"stochastic_rounding" is only allowed when "fused" and "foreach" are disabled')
    if oc.fused_back_pass and (oc.fused or oc.foreach):
        raise RuntimeError('"fused_back_pass" is only allowed when "fused" and "foreach" are disabled')
    cls, patch, default_decay = (torch.optim.AdamW, patch_adamw, 1e-2) if name == "ADAMW" else (torch.optim.Adam, patch_adam, 0)
    optimizer = cls(
        params=parameter_group_collection.parameters_for_optimizer(config), lr=config.learning_rate,
        betas=(_or(oc.beta1, 0.9), _or(oc.beta2, 0.999)), weight_decay=_or(oc.weight_decay, default_decay), eps=_or(oc.eps, 1e-8),
        amsgrad=_or(oc.amsgrad, False), foreach=_or(oc.foreach, False), maximize=_or(oc.maximize, False),
        capturable=_or(oc.capturable, False), differentiable=_or(oc.differentiable, False), fused=_or(oc.fused, False))
    patch(optimizer, oc.stochastic_rounding, ema=ema, ema_step_fn=ema_step_fn, clip_grad_norm=clip_grad_norm)
    if state_dict is not None:
        optimizer.load_state_dict(fit_state_dict_to_groups(state_dict, optimizer, parameter_group_collection, config))
    return optimizer


def _create_factored(name: str, parameter_group_collection, state_dict, config, ema, clip_grad_norm):
    """ADAFACTOR (create.py:914-936: transformers' Adafactor + patch_adafactor) and CAME (:939-951), same arguments and
    defaults.  EMA and the global-norm clip are NOT folded into these launches (`ema` / `clip_grad_norm` are ignored): the
    trainer's own EMAModuleWrapper.step and clip_grad_norm_ calls do them, one multi-tensor launch each."""
    oc = config.optimizer
    parameters = parameter_group_collection.parameters_for_optimizer(config)
    if name == "ADAFACTOR":
        from transformers.optimization import Adafactor
        from .optimizer.adafactor_extensions import patch_adafactor
        if oc.relative_step:
            for parameter in parameters:
                if isinstance(parameter, dict) and "lr" in parameter:
                    parameter.pop("lr")
        optimizer = Adafactor(
            params=parameters, lr=None if oc.relative_step is True else config.learning_rate,
            eps=(_or(oc.eps, 1e-30), _or(oc.eps2, 1e-3)), clip_threshold=_or(oc.clip_threshold, 1.0), decay_rate=_or(oc.decay_rate, -0.8),
            beta1=oc.beta1, weight_decay=_or(oc.weight_decay, 0.0), scale_parameter=_or(oc.scale_parameter, True),
            relative_step=_or(oc.relative_step, True), warmup_init=_or(oc.warmup_init, False))
        patch_adafactor(optimizer, oc.stochastic_rounding)
    else:
        from .optimizer.CAME import CAME
        optimizer = CAME(params=parameters, lr=config.learning_rate, eps=(_or(oc.eps, 1e-30), _or(oc.eps2, 1e-16)),
                         betas=(_or(oc.beta1, 0.9), _or(oc.beta2, 0.999), _or(oc.beta3, 0.9999)), weight_decay=_or(oc.weight_decay, 0),
                         stochastic_rounding=oc.stochastic_rounding)
    if state_dict is not None:
        optimizer.load_state_dict(fit_state_dict_to_groups(state_dict, optimizer, parameter_group_collection, config))
    return optimizer


def create_ema(parameters, state_dict: dict | None, config) -> EMAModuleWrapper | None:
    mode = enum_name(config.ema)
    if mode == "OFF":
        return None
    if mode == "CPU":
        raise NotImplementedError("EMA on the CPU is outside the B200 path: the fused EMA update reads the parameters in HBM")
    ema = EMAModuleWrapper(parameters=parameters, decay=config.ema_decay, update_step_interval=config.ema_update_step_interval,
                           device=torch.device(config.train_device))
    if state_dict is not None:
        ema.load_state_dict(state_dict)
    return ema
