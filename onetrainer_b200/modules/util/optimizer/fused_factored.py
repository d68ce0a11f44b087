"""Shared implementation behind patch_adafactor and CAME: one multi-tensor launch per optimizer step.

Reference: step_adafactor_parameter (/root/reference/modules/util/optimizer/adafactor_extensions.py:17-93, over
transformers' Adafactor helpers `_get_options`, `_get_lr`, `_rms`, `_approx_sq_grad`) and CAME.step_parameter
(/root/reference/modules/util/optimizer/CAME.py:80-176).  There, every tensor costs ~25 / ~45 ATen launches with
full-size fp32 temporaries in a python loop, and Adafactor's `max(eps, RMS)` on a device tensor is a host sync per tensor;
here `optimizer.step()` is ONE launch of otb_adafactor_multi / otb_came_multi (csrc/optim_factored.cu: one CTA per
tensor, fixed-order reductions) and `step_parameter` (fused-back-pass mode) one single-tensor launch.

Checkpoint compatibility: `optimizer.state[p]` keeps the reference's entries with the reference's shapes and dtypes —
'step' (python int), 'RMS' (0-dim fp32 tensor; the reference stores the tensor `_rms` returns), 'exp_avg',
'exp_avg_sq_row' / 'exp_avg_sq_col' or 'exp_avg_sq' (fp32), and for CAME 'exp_avg_res_row' / 'exp_avg_res_col'.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .... import ops
from .... import _native as N
from .fused_adam import _bump_versions, bind_to_optimizer

_DT = np.dtype(N.FactoredTensor)


def _shape3(shape) -> tuple[int, int, int, bool]:
    """(batch, rows, cols, factored): `_get_options` factors every tensor with >= 2 dimensions over its last two."""
    if len(shape) >= 2:
        return int(math.prod(shape[:-2])), int(shape[-2]), int(shape[-1]), True
    n = int(math.prod(shape)) if len(shape) else 1
    return 1, 1, n, False


class FactoredDriver:
    """Owns the device table of one Adafactor / CAME optimizer and launches the kernel."""
    SMALL_NUMEL = 1 << 17        # up to 128 k elements (every rank <= 64 LoRA factor of SD / SDXL / FLUX): one CTA per tensor
    LARGE_CLUSTER = 8            # larger tensors (full fine-tuning): a cluster of 8 CTAs per tensor

    def __init__(self, optimizer, kind: str):
        assert kind in ("adafactor", "came")
        self.opt, self.kind = optimizer, kind
        self.step_count = 0              # CAME stochastic-rounding stream position
        self._tables = {}                # key -> (host table, device table, scratch, params)
        self.post_step_hooks = []
        self.clip_max_norm = None        # the global-norm clip is not folded into these launches (trainer calls clip_grad_norm_)
        self.grad_prescale = 1.0

    # ---- state -------------------------------------------------------------------------------
    def _init_state(self, p, group) -> dict:
        st = self.opt.state[p]
        if len(st) == 0:
            if p.grad.is_sparse:
                raise RuntimeError(f"{'Adafactor' if self.kind == 'adafactor' else 'CAME'} does not support sparse gradients.")
            batch, rows, cols, factored = _shape3(p.shape)
            f32 = dict(dtype=torch.float32, device=p.device)
            st["step"] = 0
            if self.kind == "came" or group.get("beta1") is not None:
                st["exp_avg"] = torch.zeros(p.shape, **f32)
            if factored:
                st["exp_avg_sq_row"] = torch.zeros(p.shape[:-1], **f32)
                st["exp_avg_sq_col"] = torch.zeros(p.shape[:-2] + p.shape[-1:], **f32)
                if self.kind == "came":
                    st["exp_avg_res_row"] = torch.zeros(p.shape[:-1], **f32)
                    st["exp_avg_res_col"] = torch.zeros(p.shape[:-2] + p.shape[-1:], **f32)
            else:
                st["exp_avg_sq"] = torch.zeros(p.shape, **f32)
            st["RMS"] = torch.zeros((), **f32)
        elif not isinstance(st.get("RMS"), torch.Tensor) or st["RMS"].device != p.device:
            st["RMS"] = torch.zeros((), dtype=torch.float32, device=p.device)          # a loaded state carries a number here
        return st

    def _table(self, active):
        key = tuple((id(p), p.data_ptr(), p.dtype) for p, _ in active) + tuple(self.opt.state[p]["RMS"].data_ptr() for p, _ in active)
        entry = self._tables.get(key)
        if entry is None:
            if len(self._tables) > 4:
                self._tables.clear()
            host = np.zeros(len(active), dtype=_DT)
            rng_base = scratch = 0
            for i, (p, group) in enumerate(active):
                if p.dtype not in (torch.float32, torch.bfloat16, torch.float16):
                    raise NotImplementedError(f"parameter dtype {p.dtype} is not supported by the fused {self.kind} step")
                if not p.is_contiguous():
                    raise ValueError("the fused optimizers need contiguous parameters")
                st = self.opt.state[p]
                batch, rows, cols, factored = _shape3(p.shape)
                h = host[i]
                h["p"] = p.data_ptr()
                for name, k in (("exp_avg", "exp_avg"), ("sq_row", "exp_avg_sq_row"), ("sq_col", "exp_avg_sq_col"), ("sq", "exp_avg_sq"),
                                ("res_row", "exp_avg_res_row"), ("res_col", "exp_avg_res_col"), ("rms_out", "RMS")):
                    t = st.get(k)
                    if t is not None:
                        if t.dtype != torch.float32 or not t.is_contiguous() or t.device != p.device:
                            raise NotImplementedError(f"optimizer state '{k}' must be a contiguous fp32 tensor on the parameter's device")
                        h[name] = t.data_ptr()
                h["numel"], h["rng_base"], h["scratch_offset"] = p.numel(), rng_base, scratch
                h["batch"], h["rows"], h["cols"], h["p_dtype"] = batch, rows, cols, N.dtype_code(p.dtype)
                rng_base += (p.numel() + 7) // 8 * 8
                scratch += batch
            dev = active[0][0].device
            entry = dict(host=host, dev=torch.empty((max(len(active), 1) * _DT.itemsize,), dtype=torch.uint8, device=dev),
                         scratch=torch.empty((max(scratch, 1),), dtype=torch.float32, device=dev),
                         pinned=torch.empty((max(len(active), 1) * _DT.itemsize,), dtype=torch.uint8).pin_memory(), event=None)
            self._tables[key] = entry
        return entry

    def _fill(self, entry, active) -> None:
        host = entry["host"]
        sr = bool(getattr(self.opt, "stochastic_rounding", False))
        for i, (p, group) in enumerate(active):
            st = self.opt.state[p]
            st["step"] += 1
            step = st["step"]
            _, _, _, factored = _shape3(p.shape)
            h = host[i]
            h["g"] = p.grad.data_ptr()
            flags = N.FACTORED_FACTORED if factored else 0
            if self.kind == "adafactor":
                # transformers' Adafactor._get_lr, host part (the parameter scale max(eps1, RMS(p)) is applied on the device)
                step_size = group["lr"]
                if group["relative_step"]:
                    min_step = 1e-6 * step if group["warmup_init"] else 1e-2
                    step_size = min(min_step, 1.0 / math.sqrt(step))
                if step_size is None:
                    raise ValueError("Adafactor needs lr when relative_step is off")
                if group["beta1"] is not None:
                    flags |= N.FACTORED_FIRST_MOMENT
                if group["scale_parameter"]:
                    flags |= N.FACTORED_SCALE_PARAMETER
                b1 = group["beta1"] if group["beta1"] is not None else 0.0
                b2 = 1.0 - math.pow(step, group["decay_rate"])                 # beta2t
                h["beta1"], h["beta2"], h["one_minus_beta1"], h["one_minus_beta2"] = b1, b2, 1.0 - b1, 1.0 - b2
                h["lr"] = step_size
            else:
                flags |= N.FACTORED_FIRST_MOMENT | (N.FACTORED_SR if sr else 0)
                h["beta1"], h["beta2"], h["beta3"] = group["betas"]
                h["one_minus_beta1"], h["one_minus_beta2"], h["one_minus_beta3"] = (1.0 - b for b in group["betas"])     # in double, as python does
                h["lr"] = group["lr"]
            h["eps0"], h["eps1"] = group["eps"]
            h["clip_threshold"], h["weight_decay"], h["flags"] = group["clip_threshold"], group["weight_decay"], flags
            if p.grad.dtype != p.dtype or not p.grad.is_contiguous():
                raise ValueError("the fused optimizers need a contiguous gradient in the parameter's dtype")

    def _launch(self, entry, n: int) -> None:
        raw = entry["host"].view(np.uint8).reshape(-1)
        if entry["event"] is not None:
            entry["event"].synchronize()             # the copy that last read the pinned staging buffer
        entry["pinned"][: raw.size].numpy()[:] = raw
        entry["dev"][: raw.size].copy_(entry["pinned"][: raw.size], non_blocking=True)
        entry["event"] = entry["event"] or torch.cuda.Event()
        entry["event"].record()
        self.step_count += 1
        wd_calls = 2 if any(g["weight_decay"] != 0 for g in self.opt.param_groups) else 1
        fn = N.lib().otb_adafactor_multi if self.kind == "adafactor" else N.lib().otb_came_multi
        n_small = int((entry["host"]["numel"][:n] <= self.SMALL_NUMEL).sum())          # the table is sorted: small tensors first
        for first, count, cluster in ((0, n_small, 1), (n_small, n - n_small, self.LARGE_CLUSTER)):
            if count == 0:
                continue
            a = N.FactoredLaunch(tensors=entry["dev"].data_ptr() + first * _DT.itemsize, scratch=entry["scratch"].data_ptr(), n_tensors=count,
                                 cluster=cluster, seed=self._sr_seed() & (2 ** 64 - 1), offset=(self.step_count - 1) * wd_calls + 1)
            N.check(fn(C.byref(a), ops._stream()), f"otb_{self.kind}_multi")
            ops._count()

    @staticmethod
    def _sr_seed() -> int:
        """Key of CAME's stochastic-rounding stream: torch's seed — and under data parallelism rank 0's, so that the bf16
        replicas of a parameter round identically whatever per-rank seeding the application uses (as fused_adam.sr_seed)."""
        from .... import dp
        return dp.replicated_seed() if dp.world_size() > 1 else torch.initial_seed()

    # ---- entry points ------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        active = [(p, g) for g in self.opt.param_groups for p in g["params"] if p.grad is not None]
        if active:
            # size classes: LoRA-sized tensors get one CTA each, large matrices a cluster of 8 (stable order inside a class)
            active = [a for a in active if a[0].numel() <= self.SMALL_NUMEL] + [a for a in active if a[0].numel() > self.SMALL_NUMEL]
            N.require_cuda(*[p for p, _ in active])
            for p, g in active:
                self._init_state(p, g)
            entry = self._table(active)
            self._fill(entry, active)
            self._launch(entry, len(active))
            _bump_versions([p for p, _ in active])
            for hook in self.post_step_hooks:
                hook()
        return loss

    @torch.no_grad()
    def step_parameter(self, p, group, i):
        if p.grad is None:
            return
        N.require_cuda(p)
        self._init_state(p, group)
        active = [(p, group)]
        entry = self._table(active)
        self._fill(entry, active)
        self._launch(entry, 1)
        _bump_versions([p])


def install(optimizer, kind: str, stochastic_rounding: bool) -> FactoredDriver:
    optimizer.stochastic_rounding = stochastic_rounding
    driver = FactoredDriver(optimizer, kind)
    optimizer.otb_driver = driver
    bind_to_optimizer(optimizer)        # bound to the optimizer, so that an LR scheduler's step wrapper finds the driver
    return driver
