"""LoRA adapters on the B200 fused kernel — drop-in for the LoRA part of
/root/reference/modules/module/LoRAModule.py (PeftBase :17-206, LoRAModule :283-331,
LoRAModuleWrapper :427-587; DoRAModule :333-419 — SURVEY.md §8f rank 4 — on K-a with fused weight-space kernels; LoHa raises).

Seam contract kept: `klass(prefix, orig_module, rank, alpha)`;
`forward(x, *args, **kwargs)`; `hook_to_module()/remove_hook_from_module()` swapping the wrapped
module's forward/train/eval; `lora_down` / `lora_up` sub-modules whose `.weight` are real
nn.Parameters (optimizer, EMA, clipping and savers hold references to those objects); buffer
`alpha`; `dropout.p`; `state_dict(prefix=module.prefix)` keys `{prefix}lora_down.weight`,
`{prefix}lora_up.weight`, `{prefix}alpha`; prefix-filtered `load_state_dict`; `make_dummy()`.

What changes underneath: instead of base GEMM + down GEMM + dropout + up GEMM + scale + add
(3 GEMMs and 2-3 elementwise launches forward, ~7 GEMMs backward, the skinny ones at a few % of
tensor peak) every hooked leaf is
    forward : 1 launch  otb_lora_linear   (tcgen05; T = s*x*A^T rides in the same UMMA as x*W^T)
    backward: 2 launches otb_lora_linear (dX, with the transposed bf16 shadows) + otb_lora_grad_ab (dA, dB)
bf16 shadows of A/B (and their transposes) and of the frozen W^T are kept resident in HBM — on a
180 GB part the duplicated frozen weights are cheap and remove every per-call cast/transposition.
"""
from __future__ import annotations

import copy
import math
from abc import abstractmethod
from collections.abc import Mapping
from typing import Any

import torch
from torch import Tensor, nn
from torch.nn import Conv2d, Dropout, Linear, Parameter

from ... import ops
from ..util.enum import enum_name


# ---------------------------------------------------------------------------------------
# gradient arena + bf16 shadows shared by all leaves of one wrapper
# ---------------------------------------------------------------------------------------
class GradArena:
    """One flat fp32 buffer holding dA/dB of every leaf.

    The split-M gradient kernel accumulates with fp32 reductions, so its outputs must start at
    zero; one memset per step replaces two per leaf.  The views are handed to autograd as the
    gradients, which makes `p.grad` of every LoRA tensor a slice of ONE buffer — the data-parallel
    allreduce then runs in place on the arena with no gather/scatter (onetrainer_b200/dp.py).
    """

    allocator = None        # (numel, dtype, device) -> zeroed flat tensor; dp.symmetric_zeros under data parallelism

    def __init__(self, shapes: list[tuple[int, ...]], device):
        total = 0
        self.offsets = []
        for s in shapes:
            n = math.prod(s)
            self.offsets.append((total, n, s))
            total += (n + 3) // 4 * 4           # keep every slice 16-byte aligned
        total = max(total, 4)
        self.flat = GradArena.allocator(total, torch.float32, device) if GradArena.allocator is not None \
            else torch.zeros((total,), dtype=torch.float32, device=device)
        self.generation = 0

    def view(self, i: int) -> Tensor:
        off, n, s = self.offsets[i]
        return self.flat[off:off + n].view(s)

    def start_new_accumulation(self) -> None:
        self.flat.zero_()
        self.generation += 1


class _LoraLinearFn(torch.autograd.Function):
    """y = base(x) + drop(s * x A^T) B^T with the fused kernels.  Saves only x and Ts (M x r)."""

    @staticmethod
    def forward(ctx, x2: Tensor, a_param: Tensor, b_param: Tensor, mod: "LoRAModule"):
        sh = mod._operands()
        p = mod.dropout.p if mod.training else 0.0
        seed = offset = 0
        if p > 0.0:
            gen = torch.cuda.default_generators[x2.device.index]
            seed, offset = gen.initial_seed(), gen.get_offset()
            gen.set_offset(offset + 4)      # checkpoint recompute restores this state -> same mask
        y, ts = ops.lora_linear(x2, sh["w"], sh["a"], sh["b"], sh["bias"], mod._scale(), dropout_p=p,
                                seed=seed, offset=offset, want_t=any(ctx.needs_input_grad[:3]), static_weights=sh["static"])
        ctx.mod, ctx.p, ctx.key = mod, p, (seed, offset)
        ctx.save_for_backward(x2, ts)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        mod: LoRAModule = ctx.mod
        x2, ts = ctx.saved_tensors
        sh = mod._operands()
        dy = dy.contiguous()
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        need_a, need_b = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        mod._last_input_grad = ctx.needs_input_grad[0]      # read by bench.py to replay the step's launches
        # dX = dY W + drop(s * dY B) A  — same kernel, transposed shadows; also parks Gs = drop(s * dY B)
        dx = gs = None
        if ctx.needs_input_grad[0]:
            dx, gs = ops.lora_linear(dy, sh["wt"], sh["bt"], sh["at"], None, mod._scale(), dropout_p=ctx.p,
                                     seed=ctx.key[0], offset=ctx.key[1], want_t=True, static_weights=sh["static"])
        elif need_a:
            gs = _gs_only([dy], [sh], [mod._scale()], ctx.p, [ctx.key])[0]
        dA, dB, ret_a, ret_b = mod._grad_buffers(need_a, need_b)
        if dA is not None or dB is not None:
            _launch_grad_ab(mod, x2, dy, gs, ts, dA, dB, ret_a, ret_b)
        if ret_a is not None and ret_a.dtype != mod.lora_down.weight.dtype:
            ret_a = ret_a.to(mod.lora_down.weight.dtype)
        if ret_b is not None and ret_b.dtype != mod.lora_up.weight.dtype:
            ret_b = ret_b.to(mod.lora_up.weight.dtype)
        return dx, ret_a, ret_b, None


class _UnfoldRows(torch.autograd.Function):
    """im2col for a k x k conv leaf: one bf16 row per output pixel, columns (c, ki, kj).  Its adjoint (col2im) adds up to
    kh*kw overlapping patch gradients per input pixel; torch's own backward would do that in bf16, one rounding per
    addend — here the sum is formed in fp32 and rounded ONCE, like the fp32-accumulating dgrad of the reference's conv."""

    @staticmethod
    def forward(ctx, x: Tensor, kernel_size, dilation, padding, stride):
        ctx.geom = (tuple(x.shape[-2:]), kernel_size, dilation, padding, stride)
        cols = nn.functional.unfold(x, kernel_size, dilation, padding, stride)
        return cols.transpose(1, 2).reshape(-1, cols.shape[1])

    @staticmethod
    def backward(ctx, drows: Tensor):
        hw, kernel_size, dilation, padding, stride = ctx.geom
        L = ((hw[0] + 2 * padding[0] - dilation[0] * (kernel_size[0] - 1) - 1) // stride[0] + 1) * \
            ((hw[1] + 2 * padding[1] - dilation[1] * (kernel_size[1] - 1) - 1) // stride[1] + 1)
        dcols = drows.reshape(-1, L, drows.shape[1]).transpose(1, 2).float()
        dx = nn.functional.fold(dcols, hw, kernel_size, dilation, padding, stride)
        return dx.to(drows.dtype), None, None, None, None


NO_DX_ROWS = 64


def _gs_only(dys: list, shs: list, scales: list, p: float, keys: list) -> list:
    """Gs = drop(s * dY B) of leaves whose input takes no gradient (the K/V projections of the text context, the first
    layer after the latents): autograd would not form dX there and neither does this path.  K-a produces T alongside
    any output tile, so the launch is cut down to the first NO_DX_ROWS output columns of dX (discarded) and streams
    dY once; the full dX GEMM of such a leaf (M x K x N) is never issued."""
    rows = min(NO_DX_ROWS, shs[0]["wt"].shape[0])
    _, gss = ops.lora_linear_grouped(dys, [sh["wt"][:rows] for sh in shs], [sh["bt"] for sh in shs], [sh["at"][:rows] for sh in shs],
                                     [None] * len(dys), scales, dropout_p=p, keys=keys, want_t=True,
                                     static_weights=all(sh.get("static", False) for sh in shs))
    return gss


class _LoraGroupFn(torch.autograd.Function):
    """The same contract as _LoraLinearFn for up to three sibling leaves that share their input (the q/k/v projections
    of a self-attention, the k/v of a cross-attention): ONE grouped launch forward, ONE for the dX GEMMs."""

    @staticmethod
    def forward(ctx, mods: tuple, x2: Tensor, *ab: Tensor):
        ctx.set_materialize_grads(False)             # an output nobody consumed arrives as None in backward
        ops_ = [m._operands() for m in mods]
        p = mods[0].dropout.p if mods[0].training else 0.0
        keys = [(0, 0)] * len(mods)
        if p > 0.0:
            gen = torch.cuda.default_generators[x2.device.index]
            keys = []
            for _ in mods:
                seed, offset = gen.initial_seed(), gen.get_offset()
                gen.set_offset(offset + 4)
                keys.append((seed, offset))
        ys, ts = ops.lora_linear_grouped([x2] * len(mods), [o["w"] for o in ops_], [o["a"] for o in ops_], [o["b"] for o in ops_],
                                         [o["bias"] for o in ops_], [m._scale() for m in mods], dropout_p=p, keys=keys,
                                         want_t=any(ctx.needs_input_grad[1:]), static_weights=all(o["static"] for o in ops_))
        ctx.mods, ctx.p, ctx.keys = mods, p, keys
        ctx.save_for_backward(x2, *[t for t in ts if t is not None])
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        mods = ctx.mods
        x2, *ts = ctx.saved_tensors
        live = [i for i, dy in enumerate(dys) if dy is not None]
        grads = [None] * (2 * len(mods))
        dx = None
        if live:
            dyc = []
            for i in live:
                dy = dys[i].contiguous()
                dyc.append(dy if dy.dtype == torch.bfloat16 else dy.to(torch.bfloat16))
            ops_ = [mods[i]._operands() for i in live]
            dxs = None
            if ctx.needs_input_grad[1]:
                dxs, gss = ops.lora_linear_grouped(dyc, [o["wt"] for o in ops_], [o["bt"] for o in ops_], [o["at"] for o in ops_],
                                                   [None] * len(live), [mods[i]._scale() for i in live], dropout_p=ctx.p,
                                                   keys=[ctx.keys[i] for i in live], want_t=True, static_weights=all(o["static"] for o in ops_))
            elif any(ctx.needs_input_grad[2 + 2 * i] for i in live):
                gss = _gs_only(dyc, ops_, [mods[i]._scale() for i in live], ctx.p, [ctx.keys[i] for i in live])
            else:
                gss = [None] * len(live)
            for j, i in enumerate(live):
                mod = mods[i]
                mod._last_input_grad = ctx.needs_input_grad[1]
                need_a, need_b = ctx.needs_input_grad[2 + 2 * i], ctx.needs_input_grad[3 + 2 * i]
                dA, dB, ret_a, ret_b = mod._grad_buffers(need_a, need_b)
                if dA is not None or dB is not None:
                    _launch_grad_ab(mod, x2, dyc[j], gss[j], ts[i], dA, dB, ret_a, ret_b)
                if ret_a is not None and ret_a.dtype != mod.lora_down.weight.dtype:
                    ret_a = ret_a.to(mod.lora_down.weight.dtype)
                if ret_b is not None and ret_b.dtype != mod.lora_up.weight.dtype:
                    ret_b = ret_b.to(mod.lora_up.weight.dtype)
                grads[2 * i], grads[2 * i + 1] = ret_a, ret_b
            if dxs is not None:
                dx = dxs[0]
                for d in dxs[1:]:
                    dx = dx + d
        return (None, dx, *grads)


def _launch_grad_ab(mod, x2, dy, gs, ts, dA, dB, ret_a, ret_b) -> None:
    """dA / dB of one leaf: queued for the batched launch (ops.GradQueue) unless something will READ the result before
    the backward pass ends.  Autograd only ADOPTS the arena slices (or gets None when the kernel accumulates in place);
    a temporary (someone replaced p.grad, or the leaf is used twice in one graph) is added by autograd right away, and
    a post-accumulate-grad hook (optimizer-in-backward, GenericTrainer.py:550-557) steps the parameter right away — in
    both cases the kernel runs now."""
    arena = mod._wrapper_slot[0].grad_arena if mod._wrapper_slot is not None else None

    def outside(t) -> bool:
        return t is not None and (arena is None or not arena.data_ptr() <= t.data_ptr() < arena.data_ptr() + arena.numel() * arena.element_size())
    hooked = any(getattr(p, "_post_accumulate_grad_hooks", None) for p in (mod.lora_down.weight, mod.lora_up.weight))
    if outside(dA) or outside(dB) or hooked:
        # whoever reads this result now (autograd adding a second use of the leaf to the first, an optimizer hook) may also
        # read slices that are still in the queue: launch those first
        ops.grad_queue.flush_all()
        ops.lora_grad_ab(x2, dy, gs, ts, dA, dB)
    else:
        ops.grad_queue.push((x2, dy, gs, ts, dA, dB), dy.device)


class SiblingGroup:
    """Leaves that the wrapped network calls back to back on the SAME input: {to_q, to_k, to_v} of a self-attention,
    {to_k, to_v} of a cross-attention, {add_q_proj, add_k_proj, add_v_proj} of a joint attention.

    The plugin seam is one forward per leaf, so the batching is speculative: the first member called with input x
    computes every member on x in one grouped launch and parks the siblings' outputs; a sibling called next with the
    same tensor (same storage, same version, same shape) takes its parked output.  A sibling called with anything else
    proves the guess wrong: the group switches itself off for good and its members go back to single launches."""

    def __init__(self, members: list["LoRAModule"]):
        self.members = members
        self.enabled = True
        self._key = None
        self._parked: dict[int, Tensor] = {}
        self._drew_keys = False
        for i, m in enumerate(members):
            m._group = (self, i)

    @staticmethod
    def _key_of(x2: Tensor):
        return (x2.data_ptr(), x2._version, tuple(x2.shape), x2.dtype)

    def forward(self, idx: int, x_rows: Tensor) -> Tensor | None:
        """The member's output from a grouped launch, or None when the caller should launch singly.  `x_rows` is the
        member's input as the network passed it (flattened to rows, BEFORE the bf16 cast, so siblings fed the same fp32
        tensor are recognised and the cast happens once)."""
        key = self._key_of(x_rows)
        if idx in self._parked:
            if key == self._key:
                y = self._parked.pop(idx)
                if not self._parked:
                    self._key = None
                return y
            # same member, different input: not siblings after all.  Give back the dropout keys drawn for the members
            # that never consumed their output, so the single launches that follow draw what they would have drawn.
            if len(self.members) == 3 and idx == 1:
                # q read one tensor, k another: a cross-attention whose context happens to have the query's width.
                # k and v still share their input: carry on as a pair.
                self.members[0]._group = None
                self.members = self.members[1:]
                for i, m in enumerate(self.members):
                    m._group = (self, i)
            else:
                self.enabled = False
            if self._drew_keys:
                gen = torch.cuda.default_generators[x_rows.device.index]
                gen.set_offset(gen.get_offset() - 4 * len(self._parked))
            self._parked.clear()
            self._key = None
            return None
        if idx != 0:
            # only the first member starts a grouped launch (the network calls it first); a sibling that finds nothing
            # parked — its leader ran singly, or was not called in this pass — launches singly.  This also keeps a stale
            # parked output (a pass that ended early) from ever being served: the leader's next call replaces it.
            return None
        mods = tuple(self.members)
        x2 = x_rows if x_rows.dtype == torch.bfloat16 else x_rows.to(torch.bfloat16)
        x2 = x2.contiguous()
        ab = []
        for m in mods:
            ab += [m.lora_down.weight, m.lora_up.weight]
        self._drew_keys = mods[0].training and mods[0].dropout.p > 0.0
        ys = _LoraGroupFn.apply(mods, x2, *ab)
        self._key = key
        self._parked = {i: y for i, y in enumerate(ys) if i != idx}
        for m in mods:
            m._last_rows = x2.shape[0]
        return ys[idx]


SIBLING_SETS = (("to_q", "to_k", "to_v"), ("add_q_proj", "add_k_proj", "add_v_proj"))
CONTEXT_LEAVES = ("attn2.to_k", "attn2.to_v")       # diffusers BasicTransformerBlock: cross-attention over the text context
CONTEXT_ARG = ("encoder_hidden_states", 2)          # UNet2DConditionModel.forward(sample, timestep, encoder_hidden_states, ...)


class ContextLaunch:
    """The K / V projections of the text context, for EVERY cross-attention of the wrapped network, in a few wide launches.

    All `attn2.to_k` / `attn2.to_v` leaves of a UNet read the same tensor — the `encoder_hidden_states` handed to the
    network's forward — and there are many of them (SDXL: 140, M = 77 x batch rows each): as leaf-by-leaf launches they
    are the worst-shaped GEMMs of the step (a long serial K over two or three row tiles, 18 us per k/v pair).  The
    wrapper therefore hooks the NETWORK's forward (it already owns `orig_module`; the seam stays one forward per leaf):
    when the network is called, every context leaf of matching width is computed on that tensor, 32 leaves per launch
    (otb_lora_linear_grouped, wide form), and the outputs are parked on the leaves.  A leaf later called with that very
    tensor (same storage, version, shape — the input is kept alive so the address cannot be recycled) returns its
    parked output — also when a checkpointed block is re-run in backward; called with anything else it computes as
    usual.  Backward is one `Gs` launch per 32 leaves (the context takes no gradient, so there is no dX) plus the
    queued dA / dB.  OTB_CTX_GROUP=0 switches it off."""

    def __init__(self, members: list["LoRAModule"]):
        self.members = members

    def run(self, ctx_rows: Tensor) -> None:
        key = SiblingGroup._key_of(ctx_rows)
        x2 = ctx_rows if ctx_rows.dtype == torch.bfloat16 else ctx_rows.to(torch.bfloat16)
        x2 = x2.contiguous()
        for lo in range(0, len(self.members), ops.N.LORA_GROUP_MAX):
            mods = tuple(self.members[lo:lo + ops.N.LORA_GROUP_MAX])
            ab = []
            for m in mods:
                ab += [m.lora_down.weight, m.lora_up.weight]
            ys = _LoraGroupFn.apply(mods, x2, *ab)
            for m, y in zip(mods, ys):
                m._ctx_parked = (key, ctx_rows, y)




# ---------------------------------------------------------------------------------------
# PeftBase
# ---------------------------------------------------------------------------------------
class PeftBase(nn.Module):
    """Hooking / state-dict plumbing shared by adapter types (reference PeftBase, :17-206)."""
    is_applied: bool
    orig_forward: Any | None
    orig_eval: Any | None
    orig_train: Any | None
    _orig_module: list[nn.Module] | None   # in a list so nn.Module does not register it
    prefix: str
    layer_kwargs: dict
    _initialized: bool

    def __init__(self, prefix: str, orig_module: nn.Module | None):
        super().__init__()
        self.prefix = prefix + "."
        self._orig_module = [orig_module] if orig_module else None
        self.is_applied = False
        self.layer_kwargs = {}
        self._initialized = False
        self.orig_forward = self.orig_train = self.orig_eval = None
        if orig_module is not None:
            if isinstance(orig_module, nn.Linear):
                self.shape = tuple(orig_module.weight.shape)
            elif isinstance(orig_module, nn.Conv2d):
                self.shape = tuple(orig_module.weight.shape)
                for k in ("stride", "padding", "dilation", "groups"):
                    self.layer_kwargs.setdefault(k, getattr(orig_module, k))
            else:
                raise NotImplementedError("Only Linear and Conv2d are supported layers.")

    @property
    def orig_module(self) -> nn.Module:
        assert self._orig_module is not None
        return self._orig_module[0]

    def hook_to_module(self):
        if self.is_applied:
            return
        om = self.orig_module
        self.orig_forward, self.orig_train, self.orig_eval = om.forward, om.train, om.eval
        om.forward, om.train, om.eval = self.forward, self._wrap_train, self._wrap_eval
        self.is_applied = True

    def remove_hook_from_module(self):
        assert self.orig_forward is not None
        if not self.is_applied:
            return
        om = self.orig_module
        om.forward, om.train, om.eval = self.orig_forward, self.orig_train, self.orig_eval
        self.is_applied = False

    def _wrap_train(self, mode=True):
        self.orig_train(mode)
        self.train(mode)

    def _wrap_eval(self):
        self.orig_eval()
        self.eval()

    def make_weight(self, A: Tensor, B: Tensor):
        """Full-size delta weight B @ A for Linear and Conv2d alike (reference :76-91)."""
        return (B.view(B.size(0), -1) @ A.view(A.size(0), -1)).view(self.shape)

    def check_initialized(self):
        if not self._initialized:
            raise RuntimeError(f"Module {self.prefix} is not initialized.")
        assert self.orig_forward is not None
        assert self.orig_module is not None

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True, assign: bool = False):
        own = {k.removeprefix(self.prefix): v for k, v in state_dict.items() if k.startswith(self.prefix)}
        return super().load_state_dict(own, strict, assign)

    @abstractmethod
    def initialize_weights(self):
        pass

    @abstractmethod
    def apply_to_module(self):
        pass

    @abstractmethod
    def extract_from_module(self, base_module: nn.Module):
        pass

    def create_layer(self) -> tuple[nn.Module, nn.Module]:
        """Down/up modules for the wrapped layer type; uninitialised (reference :125-160)."""
        om = self.orig_module
        device = om.weight.device
        if isinstance(om, nn.Linear):
            down = nn.Linear(om.in_features, self.rank, bias=False, device=device)
            up = nn.Linear(self.rank, om.out_features, bias=False, device=device)
        elif isinstance(om, nn.Conv2d):
            down = Conv2d(om.in_channels, self.rank, om.kernel_size, om.stride, om.padding, dilation=om.dilation,
                          bias=False, device=device)
            up = Conv2d(self.rank, om.out_channels // om.groups, (1, 1), 1, bias=False, device=device)
        else:
            raise NotImplementedError("Only Linear and Conv2d are supported layers.")
        return down, up

    @classmethod
    def make_dummy(cls):
        """The adapter type's placeholder class for state-dict keys that match no trainable layer (reference :162-206):
        `Placeholder(prefix, None, *args)` parks a loaded dict and hands it back on save, and refuses everything else."""
        return type(f"Dummy{cls.__name__}", (_KeyHolder, cls), {})


def _refuse(what: str):
    def method(self, *args, **kwargs):
        raise NotImplementedError(f"{what}() should never be called on a dummy module.")
    method.__name__ = what
    return method


class _KeyHolder:
    """Mixed in FRONT of an adapter class by PeftBase.make_dummy: keeps unknown keys alive across load -> save so that a
    LoRA file with layers this network does not have round-trips unchanged; it never hooks, trains or computes."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._parked: dict = {}

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True, assign: bool = False):
        self._parked = copy.deepcopy(state_dict)
        self._initialized = True
        return nn.modules.module._IncompatibleKeys([], [])

    def state_dict(self, *args, **kwargs):  # type: ignore[override]
        if not self._initialized:
            raise RuntimeError("A state dict must be loaded before one can be returned.")
        return self._parked

    forward = _refuse("forward")
    initialize_weights = _refuse("initialize_weights")
    hook_to_module = _refuse("hook_to_module")
    remove_hook_from_module = _refuse("remove_hook_from_module")
    apply_to_module = _refuse("apply_to_module")
    extract_from_module = _refuse("extract_from_module")


# ---------------------------------------------------------------------------------------
# LoRAModule
# ---------------------------------------------------------------------------------------
class LoRAModule(PeftBase):
    lora_down: nn.Module | None
    lora_up: nn.Module | None
    rank: int
    alpha: torch.Tensor
    dropout: Dropout

    def __init__(self, prefix: str, orig_module: nn.Module | None, rank: int, alpha: float):
        super().__init__(prefix, orig_module)
        self.rank = rank
        self.dropout = Dropout(0)
        self.register_buffer("alpha", torch.tensor(alpha))
        self.lora_down = None
        self.lora_up = None
        # fused-path state (not part of the state dict)
        self._wrapper_slot = None       # (wrapper, index) when owned by a LoRAModuleWrapper
        self._own_shadow = None         # ShadowTable of [A, B] for a stand-alone module
        self._own_versions = None
        self._base_cache = None         # (key, w_bf16, wt_bf16)
        self._scale_cache = None
        self._ctx_parked = None         # output of the wrapper's wide context launch for this leaf (ContextLaunch)
        if orig_module is not None:
            self._check_supported(orig_module)
            self.initialize_weights()
            self.alpha = self.alpha.to(orig_module.weight.device)
        self.alpha.requires_grad_(False)

    @staticmethod
    def _check_supported(om: nn.Module) -> None:
        # fail loudly instead of silently falling back: quantised bases and
        # general convolutions have no fused kernel yet
        if type(om) not in (nn.Linear, nn.Conv2d) and type(om).__name__ in ("LinearNf4", "LinearFp8", "Linear4bit", "Linear8bitLt"):
            raise NotImplementedError(f"quantized base layer {type(om).__name__} is not supported by the fused LoRA kernel")
        if isinstance(om, nn.Conv2d):
            # a padding MODE only matters where something is padded: `force_circular_padding` (conv_util.py:6-9, applied to
            # the whole UNet before the adapters exist) also flips the 1x1 projections, whose padding is 0
            unpadded = not isinstance(om.padding, str) and all(int(p) == 0 for p in om.padding)
            if om.groups != 1 or isinstance(om.padding, str) or (om.padding_mode != "zeros" and not unpadded):
                raise NotImplementedError(
                    "fused LoRA supports Linear and groups-1, zero-padded Conv2d leaves; "
                    f"got Conv2d(kernel={om.kernel_size}, padding={om.padding}, padding_mode={om.padding_mode}, groups={om.groups})")

    @staticmethod
    def _is_pointwise(om: nn.Conv2d) -> bool:
        return om.kernel_size == (1, 1) and om.stride == (1, 1) and tuple(om.padding) == (0, 0)

    def initialize_weights(self):
        self._initialized = True
        self.lora_down, self.lora_up = self.create_layer()
        nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_up.weight)

    def check_initialized(self):
        super().check_initialized()
        assert self.lora_down is not None
        assert self.lora_up is not None

    # ---- operands of the fused kernels --------------------------------------------------
    def _scale(self) -> float:
        # alpha is a buffer; reading it back every call would be a device sync, so cache by version
        key = (self.alpha._version, self.alpha.data_ptr(), self.rank)
        if self._scale_cache is None or self._scale_cache[0] != key:
            self._scale_cache = (key, float(self.alpha) / self.rank)
        return self._scale_cache[1]

    def _base(self):
        om = self.orig_module
        w = om.weight
        key = (w.data_ptr(), w._version, w.dtype)
        if self._base_cache is None or self._base_cache[0] != key:
            self._rebuilt = True
            if w.dtype not in (torch.bfloat16, torch.float32, torch.float16):
                raise NotImplementedError(f"base weight dtype {w.dtype} is not supported by the fused LoRA kernel")
            w2 = w.detach().reshape(w.shape[0], -1)
            wb = w2 if w2.dtype == torch.bfloat16 and w2.is_contiguous() else w2.to(torch.bfloat16).contiguous()
            self._base_cache = (key, wb, wb.t().contiguous())
        return self._base_cache[1], self._base_cache[2]

    def _operands(self) -> dict:
        """The bf16 operands of the fused kernels.  'static' says that none of them was (re)written by a launch issued
        during THIS call — i.e. the launch preceding the GEMM on the stream is not their producer, and the kernel may
        fetch weight tiles ahead of its stream dependency (OTB_LINEAR_STATIC_WEIGHTS)."""
        self._rebuilt = False
        w, wt = self._base()
        if self._wrapper_slot is not None:
            wrapper, idx = self._wrapper_slot
            a, at, b, bt = wrapper._shadows_for(idx, self)
        else:
            A, B = self.lora_down.weight, self.lora_up.weight
            vers = (A._version, B._version, A.data_ptr(), B.data_ptr(), A.dtype)
            if self._own_shadow is None or self._own_versions != vers:
                self._rebuilt = True
                if self._own_shadow is None or self._own_versions[2:] != vers[2:]:
                    self._own_shadow = ops.ShadowTable([A.detach().view(A.shape[0], -1), B.detach().view(B.shape[0], -1)], A.device)
                self._own_shadow.refresh()
                self._own_versions = vers
            (a, b), (at, bt) = self._own_shadow.views, self._own_shadow.views_t
        return dict(w=w, wt=wt, a=a, at=at, b=b, bt=bt, bias=self.orig_module.bias, static=not self._rebuilt)

    def _grad_buffers(self, need_a: bool, need_b: bool):
        """(dA accumulate target, dB accumulate target, value to return for A, for B)."""
        A, B = self.lora_down.weight, self.lora_up.weight
        if self._wrapper_slot is not None:
            wrapper, idx = self._wrapper_slot
            return wrapper._grad_views(idx, A, B, need_a, need_b)
        dA = torch.zeros((A.shape[0], A[0].numel()), dtype=torch.float32, device=A.device) if need_a else None
        dB = torch.zeros((B.shape[0], B[0].numel()), dtype=torch.float32, device=B.device) if need_b else None
        return dA, dB, (dA.view(A.shape) if need_a else None), (dB.view(B.shape) if need_b else None)

    # ---- forward ------------------------------------------------------------------------
    def forward(self, x, *args, **kwargs):
        self.check_initialized()
        if not x.is_cuda:
            raise RuntimeError("the fused LoRA path runs on sm_100a CUDA tensors only; there is no CPU fallback")
        out_dtype = x.dtype
        if torch.is_autocast_enabled("cuda"):
            out_dtype = torch.get_autocast_dtype("cuda")
        if out_dtype != torch.bfloat16:
            raise NotImplementedError(
                f"the fused LoRA kernel computes in bf16 (tcgen05 kind::f16); got compute dtype {out_dtype}. "
                "Run under bf16 autocast or with a bf16 train dtype (every B200 config in BASELINE.json is bf16).")
        om = self.orig_module
        A, B = self.lora_down.weight, self.lora_up.weight
        if isinstance(om, nn.Conv2d):
            b_, c_, h_, w_ = x.shape
            if self._is_pointwise(om):
                x2 = x.permute(0, 2, 3, 1).reshape(-1, c_)
            else:
                # k x k leaf (create_layer, reference :142-155): one row per output pixel, columns (c, ki, kj) in the
                # order of weight.reshape(Cout, -1); base and down then share the row matrix like a Linear leaf.
                # The rows are formed in bf16 (the kernel's operand type), so the 9x copy costs 2 B per element.
                (kh, kw), (sh, sw), (ph, pw), (dh, dw) = om.kernel_size, om.stride, om.padding, om.dilation
                h_ = (h_ + 2 * ph - dh * (kh - 1) - 1) // sh + 1
                w_ = (w_ + 2 * pw - dw * (kw - 1) - 1) // sw + 1
                x2 = _UnfoldRows.apply(x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16), om.kernel_size, om.dilation,
                                       tuple(om.padding), om.stride)
        else:
            x2 = x.reshape(-1, x.shape[-1])
        self._last_rows = x2.shape[0]          # read by bench.py to replay the step's GEMM shapes
        y2 = None
        parked = self._ctx_parked               # (key, input kept alive, output) from the wrapper's context launch
        if parked is not None and parked[0] == SiblingGroup._key_of(x2):
            y2 = parked[2]
        grp = getattr(self, "_group", None)
        if y2 is None and grp is not None and grp[0].enabled and not isinstance(om, nn.Conv2d):
            y2 = grp[0].forward(grp[1], x2)
        if y2 is None:
            if x2.dtype != torch.bfloat16:
                x2 = x2.to(torch.bfloat16)
            y2 = _LoraLinearFn.apply(x2.contiguous(), A, B, self)
        if isinstance(om, nn.Conv2d):
            return y2.view(b_, h_, w_, -1).permute(0, 3, 1, 2)
        return y2.view(*x.shape[:-1], y2.shape[-1])

    def apply_to_module(self):
        pass  # not implemented in the reference either (:325-327)

    def extract_from_module(self, base_module: nn.Module):
        pass  # not implemented in the reference either (:329-331)


# ---------------------------------------------------------------------------------------
# DoRA
# ---------------------------------------------------------------------------------------
class _DoraFn(torch.autograd.Function):
    """y = x WP^T + bias with WP = dora_scale * (V / ||V||), V = W + (B A) * alpha / rank  (DoRAModule.forward,
    reference :385-419) and its backward with the norm held constant (:396-400).

    Activation-sized work on K-a (rank 0): y = x WP^T forward, dX = dY WP backward, on the bf16 WP / WP^T the compose
    kernel emits.  Weight-sized work: B A, G = dY^T x (the bf16 GEMM autocast gives the reference), dV A^T and B^T dV are
    library GEMMs; norms, WP, dV and d(dora_scale) are otb_dora_compose / otb_dora_backward."""

    @staticmethod
    def forward(ctx, x2: Tensor, a_param: Tensor, b_param: Tensor, m_param: Tensor, mod: "DoRAModule"):
        om = mod.orig_module
        w2 = om.weight.detach().reshape(om.weight.shape[0], -1)
        A2, B2 = a_param.detach().reshape(a_param.shape[0], -1).float(), b_param.detach().reshape(b_param.shape[0], -1).float()
        with torch.autocast("cuda", enabled=False):       # weight space stays fp32, as the reference builds WP in A's dtype
            ba = B2 @ A2
        gsz = 1 if isinstance(om, nn.Linear) else om.kernel_size[0] * om.kernel_size[1]
        eps = torch.finfo(torch.float32).eps if mod.norm_epsilon else 0.0
        wp, wpt, norm = ops.dora_compose(w2.contiguous(), ba, m_param, mod._scale(), eps=eps, group_size=gsz, out_axis=mod.decompose_output_axis,
                                         want_transposed=ctx.needs_input_grad[0])
        y, _ = ops.lora_linear(x2, wp, None, None, om.bias, 0.0, want_t=False)
        ctx.mod, ctx.gsz = mod, gsz
        ctx.save_for_backward(x2, wpt, norm)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        mod: DoRAModule = ctx.mod
        x2, wpt, norm = ctx.saved_tensors
        om = mod.orig_module
        dy = dy.contiguous()
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        dx = None
        if ctx.needs_input_grad[0]:
            dx, _ = ops.lora_linear(dy, wpt, None, None, None, 0.0, want_t=False)
        dA = dB = dm = None
        if any(ctx.needs_input_grad[1:4]):
            A, B, m = mod.lora_down.weight, mod.lora_up.weight, mod.dora_scale
            A2, B2 = A.detach().reshape(A.shape[0], -1).float(), B.detach().reshape(B.shape[0], -1).float()
            w2 = om.weight.detach().reshape(om.weight.shape[0], -1).contiguous()
            G = dy.t() @ x2                                           # dL/dWP, bf16 as the reference's autocast GEMM leaves it
            s = mod._scale()
            with torch.autocast("cuda", enabled=False):
                dV, dmf = ops.dora_backward(G.contiguous(), w2, B2 @ A2, m, norm, s, group_size=ctx.gsz, out_axis=mod.decompose_output_axis)
                if ctx.needs_input_grad[1]:
                    dA = ((B2.t() @ dV) * s).reshape(A.shape).to(A.dtype)
                if ctx.needs_input_grad[2]:
                    dB = ((dV @ A2.t()) * s).reshape(B.shape).to(B.dtype)
            if ctx.needs_input_grad[3]:
                dm = dmf.reshape(m.shape).to(m.dtype)
        return dx, dA, dB, dm, None


class DoRAModule(LoRAModule):
    """Weight-decomposed low-rank adaptation (reference DoRAModule, :333-419): same constructor keywords
    (`norm_epsilon`, `decompose_output_axis`, `train_device`), same `dora_scale` parameter shape and state-dict key."""
    dora_num_dims: int
    dora_scale: Tensor | None
    norm_epsilon: bool
    decompose_output_axis: bool

    def __init__(self, *args, **kwargs):
        self.dora_scale = None
        self.norm_epsilon = kwargs.pop("norm_epsilon", False)
        self.decompose_output_axis = kwargs.pop("decompose_output_axis", False)
        self.train_device = kwargs.pop("train_device")
        super().__init__(*args, **kwargs)

    def initialize_weights(self):
        super().initialize_weights()
        w = self.orig_module.weight.detach().to(device=self.train_device, dtype=torch.float)
        self.dora_num_dims = w.dim() - 1
        if self.decompose_output_axis:
            scale = torch.norm(w.reshape(w.shape[0], -1), dim=1, keepdim=True).reshape(w.shape[0], *[1] * self.dora_num_dims)
        else:
            scale = torch.norm(w.transpose(1, 0).reshape(w.shape[1], -1), dim=1, keepdim=True) \
                .reshape(w.shape[1], *[1] * self.dora_num_dims).transpose(1, 0)
        self.dora_scale = nn.Parameter(scale.contiguous().to(device=self.orig_module.weight.device))

    def check_initialized(self):
        super().check_initialized()
        assert self.dora_scale is not None

    def forward(self, x, *args, **kwargs):
        self.check_initialized()
        if not x.is_cuda:
            raise RuntimeError("the fused DoRA path runs on sm_100a CUDA tensors only; there is no CPU fallback")
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        if out_dtype != torch.bfloat16:
            raise NotImplementedError(f"the fused DoRA kernel computes in bf16 (tcgen05 kind::f16); got compute dtype {out_dtype}")
        om = self.orig_module
        x = self.dropout(x)                      # the DoRA code base drops out the INPUT (reference :413-416)
        if isinstance(om, nn.Conv2d):
            b_, c_, h_, w_ = x.shape
            if self._is_pointwise(om):
                x2 = x.permute(0, 2, 3, 1).reshape(-1, c_)
            else:
                (kh, kw), (sh, sw), (ph, pw), (dh, dw) = om.kernel_size, om.stride, om.padding, om.dilation
                h_ = (h_ + 2 * ph - dh * (kh - 1) - 1) // sh + 1
                w_ = (w_ + 2 * pw - dw * (kw - 1) - 1) // sw + 1
                x2 = _UnfoldRows.apply(x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16), om.kernel_size, om.dilation, tuple(om.padding), om.stride)
        else:
            x2 = x.reshape(-1, x.shape[-1])
        self._last_rows = x2.shape[0]
        if x2.dtype != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)
        y2 = _DoraFn.apply(x2.contiguous(), self.lora_down.weight, self.lora_up.weight, self.dora_scale, self)
        if isinstance(om, nn.Conv2d):
            return y2.view(b_, h_, w_, -1).permute(0, 3, 1, 2)
        return y2.view(*x.shape[:-1], y2.shape[-1])


DummyLoRAModule = LoRAModule.make_dummy()
DummyDoRAModule = DoRAModule.make_dummy()


# ---------------------------------------------------------------------------------------
# LoRAModuleWrapper
# ---------------------------------------------------------------------------------------
class LoRAModuleWrapper:
    """Builds one adapter per Linear/Conv2d whose name matches the filter (reference :427-587)."""
    orig_module: nn.Module
    rank: int
    alpha: float
    lora_modules: dict[str, PeftBase]

    def __init__(self, orig_module: nn.Module | None, prefix: str, config, module_filter: list[str] = None):
        self.orig_module = orig_module
        self.prefix = prefix
        self.peft_type = config.peft_type
        self.rank = config.lora_rank
        self.alpha = config.lora_alpha
        self.module_filter = [x.strip() for x in module_filter] if module_filter is not None else []
        if enum_name(self.peft_type) != "LORA":
            raise NotImplementedError("LoRA and DoRA are implemented on the B200 path; LoHa is not (SURVEY.md §2.1)")
        self.additional_args = [self.rank, self.alpha]
        if getattr(config, "lora_decompose", False):          # reference :448-457
            self.klass, self.dummy_klass = DoRAModule, DummyDoRAModule
            self.additional_kwargs = {"norm_epsilon": getattr(config, "lora_decompose_norm_epsilon", False),
                                      "decompose_output_axis": getattr(config, "lora_decompose_output_axis", False),
                                      "train_device": torch.device(config.train_device)}
        else:
            self.klass, self.dummy_klass = LoRAModule, DummyLoRAModule
            self.additional_kwargs = {}
        self.lora_modules = self.__create_modules(orig_module)
        self._shadow = None
        self._shadow_key = None
        self._arena = None
        self._arena_key = None
        self._written_gen = {}
        self._written_pass = {}
        self._net_hook = None
        self.context_launches: list[ContextLaunch] = []

    def __create_modules(self, orig_module: nn.Module | None) -> dict[str, PeftBase]:
        found = {}
        if orig_module is not None:
            for name, child in orig_module.named_modules():
                if len(self.module_filter) == 0 or any(f in name for f in self.module_filter):
                    if isinstance(child, Linear | Conv2d):
                        found[name] = self.klass(self.prefix + "." + name, child, *self.additional_args, **self.additional_kwargs)
        return found

    # ---- fused-path services for the leaves ---------------------------------------------
    def _trainable(self) -> list[LoRAModule]:
        return [m for m in self.lora_modules.values() if isinstance(m, LoRAModule) and not isinstance(m, self.dummy_klass)]

    def _bind(self) -> None:
        for i, m in enumerate(self._trainable()):
            m._wrapper_slot = (self, i)
        self._build_context_launches()
        self._build_sibling_groups()

    def _build_context_launches(self) -> None:
        """One ContextLaunch per shape class of `attn2.to_k` / `attn2.to_v` leaves (SDXL: 2048 -> 1280 and 2048 -> 640)."""
        import os
        self.context_launches: list[ContextLaunch] = []
        for m in self._trainable():
            m._ctx_parked = None
        if os.environ.get("OTB_CTX_GROUP", "1") == "0" or self.klass is not LoRAModule:      # DoRA leaves compose their own weight
            return
        classes: dict[tuple, list[LoRAModule]] = {}
        for name, m in self.lora_modules.items():
            if isinstance(m, LoRAModule) and not isinstance(m, self.dummy_klass) and isinstance(m.orig_module, nn.Linear) \
                    and name.endswith(CONTEXT_LEAVES):
                om = m.orig_module
                k = (om.in_features, om.out_features, m.rank, float(m.alpha), m.lora_down.weight.dtype, om.bias is None)
                classes.setdefault(k, []).append(m)
        self.context_launches = [ContextLaunch(ms) for ms in classes.values() if len(ms) >= 4]

    def _on_network_forward(self, module, args, kwargs) -> None:
        """Forward pre-hook of the wrapped network: start the wide context launches on `encoder_hidden_states`."""
        if not self.context_launches:
            return
        ctx = kwargs.get(CONTEXT_ARG[0])
        if ctx is None and len(args) > CONTEXT_ARG[1]:
            ctx = args[CONTEXT_ARG[1]]
        if not isinstance(ctx, Tensor) or not ctx.is_cuda or ctx.dim() < 2:
            return
        compute = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else ctx.dtype
        if compute != torch.bfloat16:
            return                              # the leaves will raise their own error for a non-bf16 compute dtype
        rows = ctx.reshape(-1, ctx.shape[-1])
        for launch in self.context_launches:
            if launch.members[0].orig_module.in_features == rows.shape[1]:
                launch.run(rows)

    def _build_sibling_groups(self) -> None:
        """Group the q/k/v-style projections that share a parent module and a shape (SiblingGroup).  With equal in_features
        all three are taken to read the same tensor (self- / joint attention); otherwise only k and v (cross-attention)."""
        import os
        self.sibling_groups: list[SiblingGroup] = []
        for m in self._trainable():
            m._group = None
        if os.environ.get("OTB_GROUP_SIBLINGS", "1") == "0" or self.klass is not LoRAModule:
            return
        by_parent: dict[str, dict[str, LoRAModule]] = {}
        in_context_launch = {id(m) for cl in self.context_launches for m in cl.members}      # those are batched network-wide
        for name, m in self.lora_modules.items():
            if id(m) in in_context_launch:
                continue
            if isinstance(m, LoRAModule) and not isinstance(m, self.dummy_klass) and isinstance(m.orig_module, nn.Linear) and "." in name:
                parent, leaf = name.rsplit(".", 1)
                by_parent.setdefault(parent, {})[leaf] = m

        def same(ms):
            a = ms[0]
            return all(b.orig_module.in_features == a.orig_module.in_features and b.orig_module.out_features == a.orig_module.out_features
                       and b.rank == a.rank and float(b.alpha) == float(a.alpha) and b.lora_down.weight.dtype == a.lora_down.weight.dtype
                       for b in ms[1:])
        for leaves in by_parent.values():
            for names in SIBLING_SETS:
                ms = [leaves[n] for n in names if n in leaves]
                if len(ms) == 3 and same(ms):
                    self.sibling_groups.append(SiblingGroup(ms))
                elif len(ms) == 3 and same(ms[1:]):
                    self.sibling_groups.append(SiblingGroup(ms[1:]))
                elif len(ms) == 2 and names[0] not in leaves and same(ms):
                    self.sibling_groups.append(SiblingGroup(ms))

    def refresh_shadows(self) -> None:
        """Re-emit the bf16 (and transposed) shadows of every A/B in ONE launch.  Called after each
        optimizer step (post-step hook) and whenever a leaf sees its parameters changed."""
        mods = self._trainable()
        if not mods or self.klass is not LoRAModule:       # DoRA reads its fp32 masters directly
            return
        key = tuple((m.lora_down.weight.data_ptr(), m.lora_up.weight.data_ptr(), m.lora_down.weight.dtype) for m in mods)
        if self._shadow is None or key != self._shadow_key:
            masters = []
            for m in mods:
                A, B = m.lora_down.weight.detach(), m.lora_up.weight.detach()
                masters += [A.view(A.shape[0], -1), B.view(B.shape[0], -1)]
            self._shadow = ops.ShadowTable(masters, masters[0].device)
            self._shadow_key = key
        self._shadow.refresh()
        self._versions = [(m.lora_down.weight._version, m.lora_up.weight._version) for m in mods]

    def _shadows_for(self, idx: int, m: LoRAModule):
        A, B = m.lora_down.weight, m.lora_up.weight
        stale = (self._shadow is None or self._versions[idx] != (A._version, B._version)
                 or self._shadow_key[idx][:2] != (A.data_ptr(), B.data_ptr()))
        if stale:
            self.refresh_shadows()
            m._rebuilt = True
        v, vt = self._shadow.views, self._shadow.views_t
        return v[2 * idx], vt[2 * idx], v[2 * idx + 1], vt[2 * idx + 1]

    def mark_parameters_updated(self) -> None:
        """Post-step hook for optimizers that write parameters outside torch (our fused AdamW)."""
        self.refresh_shadows()

    def _grad_views(self, idx: int, A: Tensor, B: Tensor, need_a: bool, need_b: bool):
        if self._arena is None or self._arena_key != A.device:
            mods = self._trainable()
            shapes = []
            for m in mods:
                a, b = m.lora_down.weight, m.lora_up.weight
                shapes += [(a.shape[0], a[0].numel()), (b.shape[0], b[0].numel())]
            self._arena = GradArena(shapes, A.device)
            self._arena_key = A.device
            self._written_gen = {}
            self._written_pass = {}
        arena = self._arena
        va, vb = arena.view(2 * idx), arena.view(2 * idx + 1)

        def state(p: Tensor, v: Tensor) -> str:
            if p.grad is None:
                return "fresh"                  # first accumulation of this step: hand the arena slice to autograd
            return "inplace" if p.grad.data_ptr() == v.data_ptr() else "foreign"

        sa, sb = state(A, va) if need_a else "skip", state(B, vb) if need_b else "skip"
        # One backward pass may reach the same leaf twice (a shared module, two forwards before one backward): autograd
        # defers AccumulateGrad, so p.grad is still None the second time and the slice handed out the first time is
        # still pending.  Do not mistake that for a new step (which would zero every gradient gathered so far) and do not
        # hand the same slice out twice (autograd would add it to itself): the second use goes through a temporary and
        # plain autograd accumulation, as the reference's per-use gradients do.
        pass_id = torch._C._current_graph_task_id()
        reused = pass_id >= 0 and self._written_pass.get(idx) == pass_id
        if reused:
            sa = "foreign" if sa == "fresh" else sa
            sb = "foreign" if sb == "fresh" else sb
        elif "fresh" in (sa, sb) and self._written_gen.get(idx) == arena.generation:
            arena.start_new_accumulation()      # first leaf of a new step: one memset for everyone
        self._written_gen[idx] = arena.generation
        self._written_pass[idx] = pass_id

        def pick(st: str, p: Tensor, v: Tensor):
            if st == "skip":
                return None, None
            if st == "fresh":
                return v, v.view(p.shape)
            if st == "inplace":                 # p.grad already IS the slice (gradient accumulation): add in place
                return v, None
            tmp = torch.zeros_like(v)           # someone replaced p.grad: plain autograd accumulation
            return tmp, tmp.view(p.shape)

        dA, ret_a = pick(sa, A, va)
        dB, ret_b = pick(sb, B, vb)
        return dA, dB, ret_a, ret_b

    @property
    def grad_arena(self) -> Tensor | None:
        return None if self._arena is None else self._arena.flat

    # ---- reference API ------------------------------------------------------------------
    def requires_grad_(self, requires_grad: bool):
        for m in self.lora_modules.values():
            m.requires_grad_(requires_grad)

    def parameters(self) -> list[Parameter]:
        out = []
        for m in self.lora_modules.values():
            out += m.parameters()
        return out

    def to(self, device: torch.device = None, dtype: torch.dtype = None) -> "LoRAModuleWrapper":
        for m in self.lora_modules.values():
            m.to(device, dtype)
        self._shadow = None
        self._arena = None
        return self

    def load_state_dict(self, state_dict: dict[str, Tensor]):
        state_dict = {k: v for k, v in state_dict.items() if k.startswith(self.prefix)}
        for name, m in self.lora_modules.items():
            try:
                m.load_state_dict(state_dict)
            except RuntimeError:  # noqa: PERF203
                print(f"Missing key for {name}; initializing it to zero.")
        leftover = set(state_dict) - set(self.state_dict())
        for name in leftover:
            if name.endswith(".alpha"):
                prefix = name.removesuffix(".alpha")
                dummy = self.dummy_klass(prefix, None, *self.additional_args, **self.additional_kwargs)
                dummy.load_state_dict(state_dict)
                self.lora_modules[prefix] = dummy
        self._shadow = None

    def state_dict(self) -> dict:
        out = {}
        for m in self.lora_modules.values():
            out |= m.state_dict(prefix=m.prefix)
        return out

    def modules(self) -> list[nn.Module]:
        out = []
        for m in self.lora_modules.values():
            out += m.modules()
        return out

    def hook_to_module(self):
        self._bind()
        for m in self.lora_modules.values():
            m.hook_to_module()
        if self._net_hook is None and isinstance(self.orig_module, nn.Module):
            self._net_hook = self.orig_module.register_forward_pre_hook(self._on_network_forward, with_kwargs=True)

    def remove_hook_from_module(self):
        for m in self.lora_modules.values():
            m.remove_hook_from_module()
        if self._net_hook is not None:
            self._net_hook.remove()
            self._net_hook = None
        for m in self._trainable():
            m._ctx_parked = None

    def launch_units(self) -> list[list["LoRAModule"]]:
        """The leaves grouped as the step launches them: context launches (32 per launch), sibling groups, single leaves —
        in module discovery order.  bench.py replays exactly these to time K-a."""
        units, taken = [], set()
        for cl in getattr(self, "context_launches", []):
            for lo in range(0, len(cl.members), ops.N.LORA_GROUP_MAX):
                chunk = cl.members[lo:lo + ops.N.LORA_GROUP_MAX]
                units.append(list(chunk))
                taken.update(id(m) for m in chunk)
        for grp in getattr(self, "sibling_groups", []):
            if grp.enabled and not any(id(m) in taken for m in grp.members):
                units.append(list(grp.members))
                taken.update(id(m) for m in grp.members)
        for m in self._trainable():
            if id(m) not in taken:
                units.append([m])
        return units

    def apply_to_module(self):
        for m in self.lora_modules.values():
            m.apply_to_module()

    def extract_from_module(self, base_module: nn.Module):
        for m in self.lora_modules.values():
            m.extract_from_module(base_module)

    def prune(self):
        self.lora_modules = {k: v for k, v in self.lora_modules.items() if not isinstance(v, self.dummy_klass)}
        self._shadow = self._arena = None
        self._bind()

    def set_dropout(self, dropout_probability: float):
        if dropout_probability < 0 or dropout_probability > 1:
            raise ValueError("Dropout probability must be in [0, 1]")
        for m in self.lora_modules.values():
            m.dropout.p = dropout_probability
