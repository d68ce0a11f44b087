"""Shared implementation behind patch_adamw / patch_adam: one multi-tensor launch per step.

The reference runs ~13 ATen launches plus fp32/int32 temporaries PER PARAMETER TENSOR in a python
loop (~20-25 k launches per step for SDXL); here `optimizer.step()` is:
  [optional] otb_grad_norm_multi   (2 launches, deterministic)   — clip_grad_norm_ fused in
  otb_adamw_sr_ema_multi           (1 launch)                    — AdamW/Adam + SR + EMA + clip scale
and `step_parameter` (fused-back-pass mode, GenericTrainer.py:531-558) is one single-tensor launch.

Checkpoint compatibility: `optimizer.state[p]` keeps exactly the reference's
entries — 'step' (CPU fp32 scalar), 'exp_avg', 'exp_avg_sq' (zeros_like(p)) — as real tensors.
"""
from __future__ import annotations

import torch

from .... import ops, step_state
from .... import _native as N


def _check_group(group: dict) -> None:
    if group.get("amsgrad"):
        raise NotImplementedError("amsgrad is not implemented on the B200 path (default off in the reference)")
    if group.get("capturable") or group.get("differentiable"):
        raise NotImplementedError("capturable/differentiable AdamW is not implemented on the B200 path")


def _init_state(state: dict, p: torch.Tensor, group: dict) -> None:
    """Lazy state initialisation, same tensors as adamw_extensions.py:28-48."""
    _check_group(group)
    state["step"] = torch.tensor(0.0, dtype=torch.float32)
    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)


def _bump_versions(params) -> None:
    """The kernels write parameters behind torch's back; bump the autograd version counters so
    everything keyed on `_version` (the LoRA bf16 shadows, saved-tensor checks) sees the update."""
    try:
        torch.autograd.graph.increment_version(params)
    except TypeError:
        for p in params:
            torch.autograd.graph.increment_version(p)


class FusedAdamDriver:
    """Owns the device pointer table of one optimizer and launches the kernels."""

    def __init__(self, optimizer, decoupled_weight_decay: bool):
        self.opt = optimizer
        self.decoupled = decoupled_weight_decay
        self.flavor = N.FLAVOR_TORCH_CUDA
        self.table: ops.OptimTable | None = None
        self.table_key = None
        self.ema = None                 # EMAModuleWrapper fused into the step, or None
        self.ema_step_fn = None         # () -> optimization_step the trainer will pass to ema.step
        self.clip_max_norm: float | None = None
        self.grad_prescale = 1.0        # 1/world_size under data parallel
        self.norm_partials = None       # squared-norm partials of ALL active gradients, when the exchange kernel made them
        self.post_step_hooks = []       # e.g. LoRAModuleWrapper.refresh_shadows
        self.rng_base_of = None         # id(p) -> global element offset in the SR stream (dp.ShardedExchange), else table order
        self.peer_store = None          # _native.PeerStore: replicate every parameter store into the peers' arenas (K-z)
        self.step_count = 0             # SR Philox offset (persisted as state_dict()["otb_sr_step_count"])
        self._sr_seed, self._sr_seed_src = None, None
        self.last_norm_coef: torch.Tensor | None = None
        self._single = {}               # id(p) -> one-entry table for step_parameter
        self._frozen_active = None
        self._last_active = []
        self._dyn = (0, 0, 0.0, False)
        self._prescale_dev, self._prescale_val = None, None
        self._ema_covered: set = set()  # id(p) of the tensors whose EMA rides in the table's launch

    # -- table construction ---------------------------------------------------------------
    def _active(self):
        out = []
        for group in self.opt.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    out.append((p, group))
        return out

    def _entries(self, active):
        ema_of = {}
        if self.ema is not None:
            ema_of = self.ema.ema_lookup()
        entries = []
        for p, group in active:
            if p.grad.is_sparse:
                raise RuntimeError("AdamW does not support sparse gradients")
            st = self.opt.state[p]
            if len(st) == 0:
                _init_state(st, p, group)
            entries.append(dict(p=p.data, g=p.grad, exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"],
                                ema=ema_of.get(id(p)), rng_base=None if self.rng_base_of is None else self.rng_base_of[id(p)]))
        self._ema_covered = {id(p) for p, _ in active if id(p) in ema_of}
        return entries

    def ema_covered_ids(self) -> set:
        """Parameters whose EMA update happened inside the last optimizer launch (EMAModuleWrapper.step does the rest)."""
        return self._ema_covered

    def invalidate(self) -> None:
        """Forget every cached pointer table (optimizer.load_state_dict replaced the state tensors)."""
        if self._frozen_active is not None:
            raise RuntimeError("optimizer state was replaced while a captured CUDA graph still points at the old tensors; "
                               "re-capture the step after load_state_dict()")
        self.table, self.table_key = None, None
        self._single.clear()

    def _ensure_table(self, active):
        # the key covers everything the device table points at: parameters AND their state tensors (load_state_dict
        # replaces exp_avg / exp_avg_sq / step; a stale table would keep stepping the old moments)
        st = self.opt.state
        for p, group in active:
            if len(st[p]) == 0:
                _init_state(st[p], p, group)
        key = tuple((id(p), p.data_ptr(), p.dtype, st[p]["exp_avg"].data_ptr(), st[p]["exp_avg_sq"].data_ptr())
                    for p, _ in active) + (id(self.ema),)
        if self.table is None or key != self.table_key:
            if active and not active[0][0].is_cuda:
                N.require_cuda(active[0][0])
            self.table = ops.OptimTable(self._entries(active), active[0][0].device)
            self.table_key = key
            self._prepare_static(self.table, active)
        return self.table

    def _prepare_static(self, table: ops.OptimTable, active) -> None:
        """Per-table constants: flags, and the step counters re-homed into ONE cpu tensor so that
        `step_t += 1` for every parameter (adamw_extensions.py:80) is a single host op.  Each
        state['step'] stays a real 0-dim fp32 CPU tensor (a view), as checkpoints expect."""
        sr = bool(getattr(self.opt, "stochastic_rounding", False))
        h = table.host
        steps = torch.stack([self.opt.state[p]["step"].detach().to("cpu", torch.float32).reshape(()) for p, _ in active]) \
            if active else torch.zeros(0)
        for i, (p, group) in enumerate(active):
            self.opt.state[p]["step"] = steps[i]
            flags = int(h["flags"][i]) & N.OPT_FLAG_ALIGNED16
            if sr and p.dtype == torch.bfloat16:
                flags |= N.OPT_FLAG_SR
            if group.get("maximize"):
                flags |= N.OPT_FLAG_MAXIMIZE
            h["flags"][i] = flags
        table.steps_flat = steps

    def _fill_scalars(self, table: ops.OptimTable, active) -> None:
        table.steps_flat += 1                                   # step_t += 1 for every tensor
        steps = table.steps_flat.tolist()
        h = table.host
        i, n = 0, len(active)
        while i < n:                                            # runs of equal (group, step) share scalars
            group, s = active[i][1], steps[i]
            j = i + 1
            while j < n and active[j][1] is group and steps[j] == s:
                j += 1
            beta1, beta2 = group["betas"]
            sc = ops.adam_scalars(float(group["lr"]), beta1, beta2, group["eps"], group["weight_decay"], s, self.decoupled)
            for k, v in sc.items():
                h[k][i:j] = v
            i = j

    def sr_seed(self) -> int:
        """Key of the stochastic-rounding stream.  Under data parallelism every rank must draw the SAME bits or the bf16
        replicas drift apart, whatever per-rank seeding the application uses: rank 0's seed is broadcast once."""
        if self._sr_seed is None:
            seed = torch.initial_seed()
            from .... import dp
            if dp.world_size() > 1:
                import torch.distributed as dist
                dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
                t = torch.tensor([seed & (2 ** 63 - 1)], dtype=torch.int64, device=dev)
                dist.broadcast(t, src=0)
                seed = int(t.item())
            self._sr_seed = seed
            self._sr_seed_src = torch.initial_seed()
        elif self._sr_seed_src != torch.initial_seed():      # the application re-seeded torch: follow it (single process)
            self._sr_seed = None
            return self.sr_seed()
        return self._sr_seed

    # -- entry points ---------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.host_prepare():
            self.device_step()
        return loss

    def host_prepare(self) -> bool:
        """Everything `step()` does on the HOST: state init, step counters, per-tensor scalars into the
        pinned table, the per-step dynamic words.  Under a captured CUDA graph this runs before every
        replay; the graph itself contains only `device_step()`."""
        for group in self.opt.param_groups:
            _check_group(group)
        if self._frozen_active is not None:
            active = self._frozen_active          # graph mode: same tensors every step, gradients live in static memory
        else:
            active = self._active()
            if not active:
                return False
            table = self._ensure_table(active)
            table.set_grad_pointers([p.grad for p, _ in active])
        table = self.table
        self._fill_scalars(table, active)
        do_ema, omd = False, 0.0
        if self.ema is not None and self.ema_step_fn is not None:
            do_ema, omd = self.ema.fused_schedule(int(self.ema_step_fn()))
        self.step_count += 1
        self._dyn = (self.sr_seed(), self.step_count, omd, do_ema)
        state = step_state.active()
        if state is not None:
            state.set_optim(*self._dyn)
        self._last_active = active
        return True

    @torch.no_grad()
    def device_step(self, upload: bool = True) -> None:
        """The DEVICE side of `step()`: table upload, [grad-norm], fused clip+AdamW(+SR)(+EMA).  A captured step passes
        `upload=False` and uploads the table before each replay instead (`upload_table`)."""
        table = self.table
        if upload:
            table.upload()
        clip = None
        if self.clip_max_norm is not None:
            if self.norm_partials is not None:     # the data-parallel exchange kernel already accumulated the squared norm
                self.last_norm_coef = ops.grad_norm_from_partials(table, self.norm_partials, self.clip_max_norm, self.grad_prescale)
            else:
                self.last_norm_coef = ops.grad_norm_multi(table, self.clip_max_norm, self.grad_prescale)
            clip = self.last_norm_coef[1:2]
        elif self.grad_prescale != 1.0:                         # data parallel without clipping: g * 1/n only
            if self._prescale_dev is None or self._prescale_val != self.grad_prescale:
                self._prescale_dev = torch.full((1,), self.grad_prescale, dtype=torch.float32, device=table.device)
                self._prescale_val = self.grad_prescale
            clip = self._prescale_dev
        seed, offset, omd, do_ema = self._dyn
        state = step_state.active()
        ops.adamw_sr_ema_multi(table, flavor=self.flavor, seed=seed, offset=offset, ema_one_minus_decay=omd, do_ema=do_ema,
                               clip_coef=clip, dyn=state.dyn_ptr if state is not None else None,
                               peer_store=self.peer_store if self.flavor == N.FLAVOR_TORCH_CUDA else None)
        _bump_versions([p for p, _ in self._last_active])
        for hook in self.post_step_hooks:
            hook()

    def upload_table(self) -> None:
        """Graph mode: the per-tensor scalars written by host_prepare() go to the device before the replay."""
        self.table.upload()

    def freeze_for_graph(self) -> None:
        """Pin the current parameter/gradient set: later host_prepare() calls skip rediscovery."""
        self._frozen_active = self._last_active

    @torch.no_grad()
    def step_parameter(self, p, group, i):
        """Per-parameter step used from post-accumulate-grad hooks (fused back pass)."""
        if p.grad is None:
            return
        if p.grad.is_sparse:
            raise RuntimeError("AdamW does not support sparse gradients")
        N.require_cuda(p)
        st = self.opt.state[p]
        if len(st) == 0:
            _init_state(st, p, group)
        key = (id(p), p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr())
        tab = self._single.get(key)
        if tab is None:
            tab = ops.OptimTable([dict(p=p.data, g=p.grad, exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"])], p.device)
            self._prepare_static(tab, [(p, group)])
            self._single[key] = tab
        tab.set_grad_pointers([p.grad])
        self._fill_scalars(tab, [(p, group)])
        tab.upload()
        self.step_count += 1
        ops.adamw_sr_ema_multi(tab, flavor=self.flavor, seed=self.sr_seed(), offset=self.step_count)
        _bump_versions([p])


def _opt_step(self, closure=None):
    return self.otb_driver.step(closure)


def _opt_step_parameter(self, p, group, i):
    return self.otb_driver.step_parameter(p, group, i)


def bind_to_optimizer(optimizer) -> None:
    """`optimizer.step` / `optimizer.step_parameter` as methods BOUND TO THE OPTIMIZER, the way the reference binds them
    (`step_adamw.__get__(optimizer, AdamW)`, adamw_extensions.py:202-205).  torch's LR schedulers wrap `optimizer.step` by
    taking its `__func__` and re-binding it to the optimizer (lr_scheduler.py: patch_track_step_called) — the reference
    creates its LambdaLR right after the optimizer (create.py), so a method bound to the driver would be called with the
    optimizer as `self`."""
    import types
    optimizer.step = types.MethodType(_opt_step, optimizer)
    optimizer.step_parameter = types.MethodType(_opt_step_parameter, optimizer)


def install(optimizer, stochastic_rounding: bool, decoupled: bool, ema=None, ema_step_fn=None,
            clip_grad_norm: float | None = None):
    optimizer.stochastic_rounding = stochastic_rounding
    driver = FusedAdamDriver(optimizer, decoupled)
    driver.ema, driver.ema_step_fn, driver.clip_max_norm = ema, ema_step_fn, clip_grad_norm
    if ema is not None:
        ema.attach_fused_driver(driver)
    optimizer.otb_driver = driver
    bind_to_optimizer(optimizer)

    # resume: the SR stream position travels with the optimizer state (a resumed run must not replay offsets 1, 2, ...),
    # and a load replaces the state tensors the cached pointer tables refer to
    def _save(opt, state_dict):
        state_dict["otb_sr_step_count"] = driver.step_count
        return state_dict

    def _pre_load(opt, state_dict):
        if "otb_sr_step_count" in state_dict:
            state_dict = dict(state_dict)
            driver.step_count = int(state_dict.pop("otb_sr_step_count"))
            return state_dict
        return None

    def _post_load(opt):
        driver.invalidate()

    optimizer.register_state_dict_post_hook(_save)
    optimizer.register_load_state_dict_pre_hook(_pre_load)
    optimizer.register_load_state_dict_post_hook(_post_load)
    return driver
