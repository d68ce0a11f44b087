"""EMAModuleWrapper — drop-in for /root/reference/modules/module/EMAModule.py:6-86.

Same constructor, `step`, `get_current_decay`, `copy_ema_to`, `copy_temp_to`, `to`,
`state_dict`/`load_state_dict` (keys "decay", "ema_parameters": list[Tensor] aligned with the
parameter list — backups stay interchangeable).  `step` is ONE multi-tensor launch
(otb_ema_multi) instead of 3 launches + 2 temporaries per tensor; when the wrapper is attached to a
patched optimizer (patch_adamw(..., ema=...)) the update already happened inside the optimizer
launch and `step` only verifies the step number.  The three roundings of
`ema += (1-d) * (p - ema)` in the EMA dtype are reproduced exactly.
"""
from __future__ import annotations

from collections.abc import Iterable

import torch

from ... import ops
from ... import _native as N


class EMAModuleWrapper:
    def __init__(self, parameters: Iterable[torch.nn.Parameter], decay: float = 0.9999,
                 update_step_interval: int = 1, device: torch.device | None = None):
        parameters = list(parameters)
        self.ema_parameters = [p.clone().detach().to(device) for p in parameters]
        self.temp_stored_parameters = None
        self.decay = decay
        self.update_step_interval = update_step_interval
        self.device = device
        self._params = parameters           # for the fused-optimizer lookup
        self._table = None
        self._table_key = None
        self._driver = None
        self._fused_step = None             # optimization_step already applied inside the optimizer launch

    def get_current_decay(self, optimization_step) -> float:
        """Warm-up schedule of the reference (:31-35): (1 + s) / (10 + s), capped by the configured decay."""
        warm_up = (optimization_step + 1) / (optimization_step + 10)
        return warm_up if warm_up < self.decay else self.decay

    # ---- fusion with the patched optimizer ------------------------------------------------
    def attach_fused_driver(self, driver) -> None:
        self._driver = driver

    def ema_lookup(self) -> dict:
        """id(parameter) -> its EMA tensor, for parameters whose EMA lives on the same device."""
        return {id(p): e for p, e in zip(self._params, self.ema_parameters)
                if p.requires_grad and e.device == p.device and e.dtype == p.dtype}

    def fused_schedule(self, optimization_step: int) -> tuple[bool, float]:
        do = (optimization_step + 1) % self.update_step_interval == 0
        self._fused_step = optimization_step
        return do, 1 - self.get_current_decay(optimization_step)

    # ---- reference API --------------------------------------------------------------------
    @torch.no_grad()
    def step(self, parameters: Iterable[torch.nn.Parameter], optimization_step) -> None:
        parameters = list(parameters)
        covered: set = set()
        if self._fused_step is not None and self._fused_step == optimization_step:
            # the optimizer launch already updated the averages of the tensors it stepped; whatever it did not cover
            # (no gradient this step, an EMA kept in another dtype) still gets the reference's update below (:37-54
            # updates EVERY requires_grad parameter)
            self._fused_step = None
            covered = self._driver.ema_covered_ids() if self._driver is not None else set()
        one_minus_decay = 1 - self.get_current_decay(optimization_step)
        if (optimization_step + 1) % self.update_step_interval != 0:
            return
        pairs = [(e, p) for e, p in zip(self.ema_parameters, parameters, strict=True) if p.requires_grad and id(p) not in covered]
        if not pairs:
            return
        same = [(e, p) for e, p in pairs if e.device == p.device and e.dtype == p.dtype]
        if same:
            N.require_cuda(*[p for _, p in same])
            key = tuple((e.data_ptr(), p.data_ptr()) for e, p in same)
            if self._table is None or key != self._table_key:
                self._table = ops.OptimTable([dict(p=p.data, ema=e) for e, p in same], same[0][1].device)
                self._table.upload()
                self._table_key = key
            ops.ema_multi(self._table, one_minus_decay)
        for e, p in pairs:
            if e.device != p.device or e.dtype != p.dtype:
                # EMA kept on another device (EMAMode.CPU) or in another dtype: the reference's three ops, in e's dtype
                delta = p.detach().to(device=e.device, dtype=e.dtype)
                delta.sub_(e)
                delta.mul_(one_minus_decay)
                e.add_(delta)

    def to(self, device: torch.device = None, dtype: torch.dtype = None) -> None:
        """Move (and for floating tensors, cast) the averages; integer buffers keep their dtype (:56-61)."""
        moved = []
        for avg in self.ema_parameters:
            kwargs = {"device": device}
            if avg.is_floating_point():
                kwargs["dtype"] = dtype
            moved.append(avg.to(**kwargs))
        self.ema_parameters, self.device = moved, device
        self._table = None                   # pointer tables refer to the old storage
        if self._driver is not None:
            self._driver.table = None

    def _write_into(self, sources, parameters) -> None:
        live = list(parameters)
        for src, dst in zip(sources, live, strict=True):
            dst.data.copy_(src.to(dst.device).data)
        # `.data.copy_` leaves Parameter._version alone, and the fused LoRA path keys its bf16 shadows on it: without the
        # bump a forward after the swap would still run the old weights (sampling "with EMA" on non-EMA weights)
        from ..util.optimizer.fused_adam import _bump_versions
        _bump_versions(live)

    def copy_ema_to(self, parameters: Iterable[torch.nn.Parameter], store_temp: bool = True) -> None:
        """Swap the averages in for sampling / saving (:63-69), optionally parking the live weights on the host."""
        live = list(parameters)
        self.temp_stored_parameters = [w.detach().cpu() for w in live] if store_temp else self.temp_stored_parameters
        self._write_into(self.ema_parameters, live)

    def copy_temp_to(self, parameters: Iterable[torch.nn.Parameter]) -> None:
        """Put the parked live weights back (:71-75)."""
        self._write_into(self.temp_stored_parameters, parameters)
        self.temp_stored_parameters = None

    def state_dict(self) -> dict:
        return dict(decay=self.decay, ema_parameters=self.ema_parameters)

    def load_state_dict(self, state_dict: dict) -> None:
        """A configured decay wins over the saved one, as in the reference (:80-86)."""
        if not self.decay:
            self.decay = state_dict.get("decay", self.decay)
        self.ema_parameters = state_dict.get("ema_parameters")
        self.to(self.device)
