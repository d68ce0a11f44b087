"""Named parameter groups — the optimizer's view of the trainable tensors.

Same public surface as /root/reference/modules/util/NamedParameterGroup.py:10-69 (class names, `add_group`,
`parameters`, `parameters_for_optimizer`, the two name mappings, the `{"params", "lr", "initial_lr"}` dicts handed to
torch.optim), written for this package.  One deliberate extension: under data parallelism the BATCH /
BOTH learning-rate scalers see the GLOBAL batch (`batch_size x world_size`), so an N-GPU run at per-GPU batch b gets
the learning rate of the 1-GPU run at batch N*b that it reproduces; with one process the arithmetic is the reference's.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from torch.nn import Parameter

from ... import dp
from .enum import enum_name

_SCALES_WITH_BATCH = ("BATCH", "BOTH")
_SCALES_WITH_ACCUMULATION = ("GRADIENT_ACCUMULATION", "BOTH")


@dataclass
class NamedParameterGroup:
    unique_name: str
    parameters: list[Parameter]
    learning_rate: float | None               # None: take the config's
    display_name: str | None = None

    def __post_init__(self):
        self.parameters = list(self.parameters)
        if self.display_name is None:
            self.display_name = self.unique_name


@dataclass
class NamedParameterGroupCollection:
    _groups: list[NamedParameterGroup] = field(default_factory=list)

    def add_group(self, group: NamedParameterGroup) -> None:
        self._groups.append(group)

    def parameters(self) -> list[Parameter]:
        """Flat, in group order: the order the optimizer table, the EMA list and the gradient arena share."""
        flat: list[Parameter] = []
        for g in self._groups:
            flat.extend(g.parameters)
        return flat

    @staticmethod
    def learning_rate_scale(config) -> float:
        """sqrt of (batch, where the scaler names it) x (accumulation steps, likewise) — reference :40-52."""
        which = enum_name(config.learning_rate_scaler)
        samples = 1
        if which in _SCALES_WITH_BATCH:
            samples *= config.batch_size * dp.world_size()
        if which in _SCALES_WITH_ACCUMULATION:
            samples *= config.gradient_accumulation_steps
        return samples ** 0.5              # `** 0.5`, not math.sqrt: the reference's pow, to the last bit

    def parameters_for_optimizer(self, config) -> list[dict]:
        scale = self.learning_rate_scale(config)

        def entry(g: NamedParameterGroup) -> dict:
            rate = (g.learning_rate if g.learning_rate is not None else config.learning_rate) * scale
            return {"params": list(g.parameters), "lr": rate, "initial_lr": rate}

        return [entry(g) for g in self._groups]

    # attribute-style access like the reference's cached properties (read by the savers: param_group_mapping)
    @property
    def unique_name_mapping(self) -> list[str]:
        return [g.unique_name for g in self._groups]

    @property
    def display_name_mapping(self) -> list[str]:
        return [g.display_name for g in self._groups]
