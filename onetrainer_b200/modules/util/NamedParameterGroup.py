"""Named parameter groups — the optimizer's view of the trainable tensors.

Mirror of /root/reference/modules/util/NamedParameterGroup.py:10-69 (same class and method names, same dict layout
handed to torch.optim).  One deliberate extension: under data parallelism the BATCH / BOTH learning-rate
scalers see the GLOBAL batch (`batch_size x world_size`), so an N-GPU run at per-GPU batch b gets the learning rate of
the 1-GPU run at batch N*b that it reproduces; with one process this is the reference's arithmetic exactly.
"""
from __future__ import annotations

from collections.abc import Iterable
from functools import cached_property

from torch.nn import Parameter

from ... import dp
from .enum import enum_name


class NamedParameterGroup:
    def __init__(self, unique_name: str, parameters: Iterable[Parameter], learning_rate: float | None,
                 display_name: str | None = None):
        self.unique_name = unique_name
        self.display_name = unique_name if display_name is None else display_name
        self.parameters = list(parameters)
        self.learning_rate = learning_rate


class NamedParameterGroupCollection:
    def __init__(self):
        self._groups: list[NamedParameterGroup] = []

    def add_group(self, group: NamedParameterGroup):
        self._groups.append(group)

    def parameters(self) -> list[Parameter]:
        """Flat list in group order — the order the optimizer table, the EMA list and the gradient arena share."""
        return [p for group in self._groups for p in group.parameters]

    @staticmethod
    def learning_rate_scale(config) -> float:
        """sqrt(batch x accumulation) for the parts the scaler names (reference :40-52)."""
        scaler = enum_name(config.learning_rate_scaler)
        batch = config.batch_size * dp.world_size() if scaler in ("BATCH", "BOTH") else 1
        accumulation = config.gradient_accumulation_steps if scaler in ("GRADIENT_ACCUMULATION", "BOTH") else 1
        return (batch * accumulation) ** 0.5

    def parameters_for_optimizer(self, config) -> list[dict]:
        scale = self.learning_rate_scale(config)
        out = []
        for group in self._groups:
            lr = (config.learning_rate if group.learning_rate is None else group.learning_rate) * scale
            out.append({"params": list(group.parameters), "lr": lr, "initial_lr": lr})
        return out

    @cached_property
    def unique_name_mapping(self) -> list[str]:
        return [group.unique_name for group in self._groups]

    @cached_property
    def display_name_mapping(self) -> list[str]:
        return [group.display_name for group in self._groups]
