"""Functional wrappers over the C ABI: tensors in, kernels launched on the current stream.

These are the only call sites of libotb200 in the package; the reference-facing classes
(modules/..., util/...) are written on top of them.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _native as N

_OPT_DTYPE = np.dtype(N.OptimTensor)
_SHADOW_DTYPE = np.dtype(N.ShadowTensor)

CHUNK_ELEMS = 16384          # elements of one tensor handled by one CTA of the multi-tensor kernels

# number of libotb200 kernels launched by this process (bench.py reports it as `gpu_launches`)
LAUNCHES = 0


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class PinnedRing:
    """Slot bookkeeping for pinned staging buffers that feed asynchronous H2D copies."""
    SLOTS = 3

    def __init__(self):
        self._events = [None] * self.SLOTS
        self._next = 0

    def acquire(self) -> int:
        slot = self._next
        self._next = (slot + 1) % self.SLOTS
        ev = self._events[slot]
        if ev is not None:
            ev.synchronize()            # the copy that read this slot SLOTS uploads ago (normally long done)
        return slot

    def release(self, slot: int) -> None:
        ev = self._events[slot]
        if ev is None:
            ev = self._events[slot] = torch.cuda.Event()
        ev.record()                     # on the current stream, right after the copy


# --------------------------------------------------------------------------------------
# deferred, batched weight gradients
# --------------------------------------------------------------------------------------
class GradQueue:
    """Queues the dA / dB contractions of the leaves of one backward pass and flushes them in batches (K-a').

    dA / dB of a leaf feed nothing until the optimizer, but one launch per leaf (722 per SDXL step) is a 10 us,
    latency-bound streaming pass.  A leaf's backward therefore only RECORDS its operands here; every `FLUSH_LEAVES`
    leaves — and, through an autograd-engine callback, when the backward pass ends — the queue is handed to
    otb_lora_grad_batched, one persistent launch per 32 leaves that keeps every SM streaming.  When `.backward()`
    returns, every gradient has been launched on the caller's stream, exactly as before.  The queue holds references
    to the operands until the flush, which bounds their extra lifetime to one batch.  Works the same under CUDA-graph
    capture.  OTB_GRAD_BATCH=1 launches every leaf by itself (A/B runs)."""
    FLUSH_LEAVES = 96        # three chained launches per flush: only the last one's tail is exposed

    def __init__(self):
        import os
        self.FLUSH_LEAVES = max(1, int(os.environ.get("OTB_GRAD_BATCH", self.FLUSH_LEAVES)))
        self._pending: dict[int, list] = {}
        self._task: dict[int, int] = {}       # device -> id of the backward pass whose end-of-pass flush covers `_pending`

    def push(self, leaf: tuple, device: torch.device) -> None:
        """leaf = (x, dy, gs, ts, dA, dB); launched at the latest when the current backward pass ends."""
        task = torch._C._current_graph_task_id()
        if self.FLUSH_LEAVES == 1 or task < 0:
            lora_grad_batched([leaf])           # outside a backward pass nobody would flush the queue
            return
        idx = device.index if device.index is not None else torch.cuda.current_device()
        held = self._pending.setdefault(idx, [])
        if held and self._task.get(idx) != task:
            # leaves of ANOTHER backward pass are still queued: that pass raised before its end-of-pass callback ran (the
            # engine drops the callbacks of a failed pass), or it is the outer pass of a re-entrant backward.  Launch them
            # now — their operands are still referenced — so that this pass's callback only ever covers its own leaves
            # and nothing stays queued past the pass that registered for it.
            self.flush(idx)
            held = self._pending[idx]
        held.append(leaf)
        if len(held) == 1:
            # first pending leaf of this pass since the last flush: make sure the pass ends with one
            self._task[idx] = task
            torch.autograd.Variable._execution_engine.queue_callback(self.flush_all)
        if len(held) >= self.FLUSH_LEAVES:
            self.flush(idx)

    def flush(self, idx: int) -> None:
        held = self._pending.get(idx)
        if held:
            self._pending[idx] = []
            with torch.cuda.device(idx):
                lora_grad_batched(held)

    def flush_all(self) -> None:
        for idx in list(self._pending):
            self.flush(idx)


grad_queue = GradQueue()


def _aligned16(*tensors) -> bool:
    return all(t is None or (t.data_ptr() & 15) == 0 for t in tensors)


# --------------------------------------------------------------------------------------
# K-a
# --------------------------------------------------------------------------------------
def lora_linear(x: torch.Tensor, w: torch.Tensor, down: torch.Tensor | None, up: torch.Tensor | None,
                bias: torch.Tensor | None, scale: float, *, dropout_p: float = 0.0, seed: int = 0,
                offset: int = 0, want_t: bool = True, tile_n: int = 0, stages: int = 0, static_weights: bool = False):
    """y = x w^T + bias + drop(scale * x down^T)_bf16 up^T  on tcgen05.  Returns (y, t) with
    t = drop(scale * x down^T) in bf16 ([M, R]) or None.  x:[M,K] w:[N,K] down:[R,K] up:[N,R], bf16.
    `static_weights`: w / down / up / bias were not written by the launch preceding this one on the stream
    (OTB_LINEAR_STATIC_WEIGHTS: the kernel may fetch weight tiles before that launch has finished)."""
    N.require_cuda(x, w, down, up, bias)
    for name, t in (("x", x), ("w", w), ("down", down), ("up", up)):
        if t is not None and (t.dtype != torch.bfloat16 or not t.is_contiguous() or t.dim() != 2):
            raise ValueError(f"lora_linear: {name} must be a contiguous 2-D bf16 tensor")
    M, K = x.shape
    Nn = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"lora_linear: w is {tuple(w.shape)}, expected [N, {K}]")
    R = 0
    if down is not None:
        R = down.shape[0]
        if down.shape[1] != K or up is None or tuple(up.shape) != (Nn, R):
            raise ValueError("lora_linear: down must be [R, K] and up [N, R]")
    y = torch.empty((M, Nn), dtype=torch.bfloat16, device=x.device)
    t = torch.empty((M, R), dtype=torch.bfloat16, device=x.device) if (want_t and R > 0) else None
    if M == 0:
        return y, t
    a = N.LoraLinearArgs(
        x=x.data_ptr(), w=w.data_ptr(), down=N.ptr(down), up=N.ptr(up), bias=N.ptr(bias), y=y.data_ptr(),
        t_out=N.ptr(t), M=M, N=Nn, K=K, R=R,
        bias_dtype=N.dtype_code(bias.dtype) if bias is not None else N.F32,
        scale=float(scale), dropout_p=float(dropout_p), seed=int(seed) & (2**64 - 1), offset=int(offset) & (2**64 - 1),
        tile_n=int(tile_n), stages=int(stages), flags=N.LINEAR_STATIC_WEIGHTS if static_weights else 0)
    N.check(N.lib().otb_lora_linear(C.byref(a), _stream()), "otb_lora_linear")
    _count()
    return y, t


def lora_linear_grouped(xs: list[torch.Tensor], ws: list[torch.Tensor], downs: list[torch.Tensor], ups: list[torch.Tensor],
                        biases: list[torch.Tensor | None], scales: list[float], *, dropout_p: float = 0.0,
                        keys: list[tuple[int, int]] | None = None, want_t: bool = True, tile_n: int = 0, static_weights: bool = False):
    """`len(xs)` (1..32) same-shape lora_linear problems as ONE launch (otb_lora_linear_grouped): the q/k/v projections of
    a self-attention (pass the same x three times), the k/v projections of a cross-attention, their dX GEMMs — or the
    K/V projections of the text context of up to 16 cross-attentions at once.  Returns ([y_i], [t_i]); bit-identical to
    the single calls."""
    g = len(xs)
    if not (1 <= g <= N.LORA_GROUP_MAX and len(ws) == len(downs) == len(ups) == len(biases) == len(scales) == g):
        raise ValueError(f"lora_linear_grouped: 1..{N.LORA_GROUP_MAX} members, one entry per member in every list")
    M, K = xs[0].shape
    Nn, R = ws[0].shape[0], downs[0].shape[0]
    arr = (N.LoraLinearArgs * g)()
    ys, ts = [], []
    for i in range(g):
        x, w, down, up, bias = xs[i], ws[i], downs[i], ups[i], biases[i]
        N.require_cuda(x, w, down, up, bias)
        for name, t in (("x", x), ("w", w), ("down", down), ("up", up)):
            if t.dtype != torch.bfloat16 or not t.is_contiguous() or t.dim() != 2:
                raise ValueError(f"lora_linear_grouped: {name}[{i}] must be a contiguous 2-D bf16 tensor")
        if tuple(x.shape) != (M, K) or tuple(w.shape) != (Nn, K) or tuple(down.shape) != (R, K) or tuple(up.shape) != (Nn, R):
            raise ValueError("lora_linear_grouped: members must share M, N, K and rank")
        y = torch.empty((M, Nn), dtype=torch.bfloat16, device=x.device)
        t = torch.empty((M, R), dtype=torch.bfloat16, device=x.device) if want_t else None
        seed, offset = keys[i] if keys is not None else (0, 0)
        arr[i] = N.LoraLinearArgs(
            x=x.data_ptr(), w=w.data_ptr(), down=down.data_ptr(), up=up.data_ptr(), bias=N.ptr(bias), y=y.data_ptr(), t_out=N.ptr(t),
            M=M, N=Nn, K=K, R=R, bias_dtype=N.dtype_code(bias.dtype) if bias is not None else N.F32, scale=float(scales[i]),
            dropout_p=float(dropout_p), seed=int(seed) & (2**64 - 1), offset=int(offset) & (2**64 - 1), tile_n=int(tile_n), stages=0,
            flags=N.LINEAR_STATIC_WEIGHTS if static_weights else 0)
        ys.append(y)
        ts.append(t)
    if M > 0:
        N.check(N.lib().otb_lora_linear_grouped(arr, g, _stream()), "otb_lora_linear_grouped")
        _count()
    return ys, ts


class _GradWorkspace:
    """Per-device scratch of the batched weight-gradient kernel, cut into fixed SLOTS (one per chained launch of a flush).
    Every slot starts with 64 KB of arrival counters — zero before the first use, left zero by every launch — followed
    by the partial tiles of split problems.  The slot size and count only ever grow, and growing allocates a new, zeroed
    buffer: a counter block therefore never lands on memory that held partial sums.  (The old buffer stays alive until
    every launch that used it has been ordered before the new one's first use — stream order on the one stream using it.)"""

    def __init__(self):
        self._bufs = {}

    def get(self, device, slots: int, slot_bytes: int) -> tuple[torch.Tensor, int]:
        key = str(device)
        buf, n, size = self._bufs.get(key, (None, 0, 0))
        if buf is None or n < slots or size < slot_bytes:
            n = max(n, slots, 3)
            size = max(size, (max(slot_bytes, 4 << 20) * 3 // 2 + 255) // 256 * 256)
            buf = torch.zeros((n * size,), dtype=torch.uint8, device=device)
            self._bufs[key] = (buf, n, size)
        return buf, size


_grad_ws = _GradWorkspace()


def lora_grad_batched(leaves: list[tuple]) -> None:
    """dA[R,K] += gs^T x ; dB[N,R] += dy^T ts for a LIST of leaves (x, dy, gs, ts, dA, dB) in as few launches as possible
    (otb_lora_grad_batched: up to 32 leaves each).  Fixed-order fp32 sums: bit-identical from run to run."""
    if not leaves:
        return
    arr = (N.LoraGradArgs * len(leaves))()
    for i, (x, dy, gs, ts, dA, dB) in enumerate(leaves):
        N.require_cuda(x, dy, gs, ts, dA, dB)
        for t in (dA, dB):
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise ValueError("lora_grad: dA / dB must be contiguous fp32")
        for t in (x, dy, gs, ts):
            if t is not None and (t.dtype != torch.bfloat16 or not t.is_contiguous() or t.dim() != 2):
                raise ValueError("lora_grad: x / dy / gs / ts must be contiguous 2-D bf16 tensors")
        M = x.shape[0] if x is not None else dy.shape[0]
        R = gs.shape[1] if gs is not None else ts.shape[1]
        arr[i] = N.LoraGradArgs(x=N.ptr(x), dy=N.ptr(dy), gs=N.ptr(gs), ts=N.ptr(ts), dA=N.ptr(dA), dB=N.ptr(dB),
                                M=M, N=dy.shape[1] if dy is not None else 0, K=x.shape[1] if x is not None else 0, R=R)
    dev = next(t for t in leaves[0] if t is not None).device
    lib = N.lib()
    chunks = []
    for lo in range(0, len(leaves), N.GRAD_BATCH_MAX):
        n = min(N.GRAD_BATCH_MAX, len(leaves) - lo)
        sub = C.cast(C.byref(arr, lo * C.sizeof(N.LoraGradArgs)), C.POINTER(N.LoraGradArgs))
        need = lib.otb_lora_grad_workspace_bytes(sub, n)
        if need < 0:
            N.check(lib.otb_lora_grad_batched(sub, n, None, 0, 0, _stream()), "otb_lora_grad_batched")      # reports the bad argument
        chunks.append((sub, n, (int(need) + 255) // 256 * 256))
    # the launches of one flush are independent of each other (different leaves, different outputs): each gets its own
    # slice of the workspace and all but the first are chained, so that they overlap instead of each ending in its own tail
    ws, slot = _grad_ws.get(dev, len(chunks), max(c[2] for c in chunks))
    for i, (sub, n, need) in enumerate(chunks):
        N.check(lib.otb_lora_grad_batched(sub, n, ws.data_ptr() + i * slot, slot, 1 if i > 0 else 0, _stream()), "otb_lora_grad_batched")
        _count()


def lora_grad_ab(x, dy, gs, ts, dA: torch.Tensor | None, dB: torch.Tensor | None) -> None:
    """dA[R,K] += gs^T x ; dB[N,R] += dy^T ts  for one leaf (a batch of one)."""
    if dA is None and dB is None:
        return
    lora_grad_batched([(x, dy, gs, ts, dA, dB)])


# --------------------------------------------------------------------------------------
# DoRA: weight-space kernels
# --------------------------------------------------------------------------------------
def _dora_args(w2, ba, dora_scale, norm, scratch, scale, eps, group_size, out_axis, **kw) -> "N.DoraArgs":
    Nn, J = w2.shape
    return N.DoraArgs(w=w2.data_ptr(), ba=ba.data_ptr(), dora_scale=dora_scale.data_ptr(), norm=norm.data_ptr(), scratch=scratch.data_ptr(),
                      N=Nn, J=J, group_size=int(group_size), out_axis=1 if out_axis else 0, w_dtype=N.dtype_code(w2.dtype),
                      scale=float(scale), eps=float(eps), **{k: N.ptr(v) for k, v in kw.items()})


def dora_compose(w2: torch.Tensor, ba: torch.Tensor, dora_scale: torch.Tensor, scale: float, *, eps: float = 0.0, group_size: int = 1,
                 out_axis: bool = False, want_transposed: bool = True):
    """WP = dora_scale * (V / ||V||), V = w2 + ba * scale (DoRAModule.forward, LoRAModule.py:389-412): returns
    (wp [N, J] bf16, wp^T [J, N] bf16 or None, norm fp32).  w2: [N, J] base weight; ba: [N, J] fp32; dora_scale flat fp32."""
    N.require_cuda(w2, ba, dora_scale)
    Nn, J = w2.shape
    if not (w2.is_contiguous() and ba.is_contiguous() and ba.dtype == torch.float32 and tuple(ba.shape) == (Nn, J)):
        raise ValueError("dora_compose: w2 and ba must be contiguous [N, J], ba fp32")
    groups = Nn if out_axis else J // group_size
    m = dora_scale.detach().reshape(-1)
    if m.numel() != groups or m.dtype != torch.float32 or not m.is_contiguous():
        raise ValueError(f"dora_compose: dora_scale must be {groups} contiguous fp32 values")
    dev = w2.device
    norm = torch.empty((groups,), dtype=torch.float32, device=dev)
    scratch = torch.empty((max(Nn, 64 * J),), dtype=torch.float32, device=dev)
    wp = torch.empty((Nn, J), dtype=torch.bfloat16, device=dev)
    wpt = torch.empty((J, Nn), dtype=torch.bfloat16, device=dev) if want_transposed else None
    a = _dora_args(w2, ba, m, norm, scratch, scale, eps, group_size, out_axis, wp=wp, wpt=wpt)
    N.check(N.lib().otb_dora_compose(C.byref(a), _stream()), "otb_dora_compose")
    _count(3)
    return wp, wpt, norm


def dora_backward(G: torch.Tensor, w2: torch.Tensor, ba: torch.Tensor, dora_scale: torch.Tensor, norm: torch.Tensor, scale: float, *,
                  group_size: int = 1, out_axis: bool = False):
    """(dV [N, J] fp32, d_dora_scale flat fp32) from G = dL/dWP (bf16), with the norm held constant (LoRAModule.py:396-400)."""
    N.require_cuda(G, w2, ba, dora_scale, norm)
    Nn, J = w2.shape
    if G.dtype != torch.bfloat16 or not G.is_contiguous() or tuple(G.shape) != (Nn, J):
        raise ValueError("dora_backward: G must be a contiguous bf16 [N, J] tensor")
    dev = w2.device
    m = dora_scale.detach().reshape(-1)
    dV = torch.empty((Nn, J), dtype=torch.float32, device=dev)
    dm = torch.empty_like(m)
    scratch = torch.empty((max(Nn, 64 * J),), dtype=torch.float32, device=dev)
    a = _dora_args(w2, ba, m, norm, scratch, scale, 0.0, group_size, out_axis, G=G, dV=dV, d_scale=dm)
    N.check(N.lib().otb_dora_backward(C.byref(a), _stream()), "otb_dora_backward")
    _count(2)
    return dV, dm


# --------------------------------------------------------------------------------------
# K-b1
# --------------------------------------------------------------------------------------
def noise_addnoise(*, latent=None, noise_in=None, coef_x0=None, coef_noise=None, timestep=None,
                   B: int, C_: int, hw: int, latent_dtype: torch.dtype, device,
                   want_x0=False, want_noise=False, want_noisy=False, cast_dtype: torch.dtype | None = None,
                   target: int = N.TARGET_NONE, ts_mode: int = N.TS_GIVEN, num_train_timesteps: int = 1000,
                   generate_noise: bool = True, do_add_noise: bool = True,
                   latent_scale: float = 1.0, latent_shift: float = 0.0,
                   offset_noise_weight: float = 0.0, perturbation_noise_weight: float = 0.0,
                   min_timestep: int = 0, max_timestep: int = 1000, noising_weight: float = 0.0,
                   noising_bias: float = 0.0, shift: float = 1.0, seed: int = 0, offset: int = 0,
                   sample_base: int = 0, want_sigma: bool = False, shape=None, rng_dev: torch.Tensor | None = None,
                   pack_w: int = 0, ts_cdf: torch.Tensor | None = None) -> dict:
    """One launch of the pre-network half of the step kernel; returns the requested tensors."""
    N.require_cuda(latent, noise_in, coef_x0, coef_noise, timestep, ts_cdf)
    if ts_mode == N.TS_TABLE and (ts_cdf is None or ts_cdf.dtype != torch.float32 or not ts_cdf.is_contiguous()
                                  or ts_cdf.numel() != max_timestep - min_timestep):
        raise ValueError("noise_addnoise: TS_TABLE needs a contiguous fp32 ts_cdf of max_timestep - min_timestep entries")
    shape = tuple(shape) if shape is not None else (B, C_ * hw)
    ldt = N.dtype_code(latent_dtype)

    def new(dt):
        return torch.empty(shape, dtype=dt, device=device)

    out = {}
    x0 = new(latent_dtype) if want_x0 else None
    noise = new(latent_dtype) if want_noise else None
    noisy = new(latent_dtype) if want_noisy else None
    noisy_cast = new(cast_dtype) if cast_dtype is not None else None
    if pack_w and noisy_cast is not None:          # FLUX packed layout [B, (H/2)(W/2), 4C] (FluxModel.pack_latents)
        noisy_cast = noisy_cast.view(B, hw // 4, 4 * C_)
    tgt = new(latent_dtype) if target != N.TARGET_NONE else None
    if ts_mode == N.TS_GIVEN:
        if timestep is None:
            raise ValueError("noise_addnoise: ts_mode GIVEN needs timesteps")
        ts = timestep.to(dtype=torch.int32).contiguous()
    else:
        ts = torch.empty((B,), dtype=torch.int32, device=device)
    sigma = torch.empty((B,), dtype=torch.float32, device=device) if want_sigma else None
    T = int(num_train_timesteps)
    a = N.NoiseArgs(
        latent=N.ptr(latent), noise_in=N.ptr(noise_in), x0_out=N.ptr(x0), noise_out=N.ptr(noise), noisy_out=N.ptr(noisy),
        noisy_cast_out=N.ptr(noisy_cast), target_out=N.ptr(tgt), timestep=ts.data_ptr(), sigma_out=N.ptr(sigma),
        coef_x0=N.ptr(coef_x0), coef_noise=N.ptr(coef_noise),
        B=B, C=C_, hw=hw, latent_dtype=ldt,
        cast_dtype=N.dtype_code(cast_dtype) if cast_dtype is not None else ldt,
        target=target, ts_mode=ts_mode, num_train_timesteps=T,
        generate_noise=1 if generate_noise else 0, do_add_noise=1 if do_add_noise else 0,
        latent_scale=float(latent_scale), latent_shift=float(latent_shift),
        offset_noise_weight=float(offset_noise_weight), perturbation_noise_weight=float(perturbation_noise_weight),
        min_timestep=int(min_timestep), max_timestep=int(max_timestep),
        noising_weight=float(noising_weight), noising_bias=float(noising_bias),
        shift_c1=float(T * shift), shift_c2=float(shift - 1),
        seed=int(seed) & (2**64 - 1), offset=int(offset) & (2**64 - 1), sample_base=int(sample_base), rng_dev=N.ptr(rng_dev),
        pack_w=int(pack_w), ts_cdf=N.ptr(ts_cdf), ts_cdf_len=0 if ts_cdf is None else ts_cdf.numel())
    N.check(N.lib().otb_noise_addnoise(C.byref(a), _stream()), "otb_noise_addnoise")
    _count()
    out.update(x0=x0, noise=noise, noisy=noisy, noisy_cast=noisy_cast, target=tgt, timestep=ts, sigma=sigma)
    return out


# --------------------------------------------------------------------------------------
# K-b2
# --------------------------------------------------------------------------------------
class LossWorkspace:
    """Per-device scratch of the loss kernel (partials, arrival counters, mask statistics)."""

    def __init__(self):
        self._bufs = {}

    def get(self, device, B: int, bps: int):
        key = (str(device), B, bps)
        if key not in self._bufs:
            self._bufs[key] = (
                torch.empty((B * bps,), dtype=torch.float32, device=device),
                torch.zeros((B,), dtype=torch.int32, device=device),
                torch.empty((B * 2,), dtype=torch.float32, device=device),
            )
        return self._bufs[key]


_loss_ws = LossWorkspace()


def loss_fwd_bwd(predicted, target, *, prior_target=None, mask=None, loss_weight=None, timestep=None,
                 sqrt_ac=None, sqrt_1mac=None, sigmas=None, masked=False, normalize_masked_area_loss=False,
                 loss_weight_fn=N.LW_CONSTANT, v_prediction=False, num_train_timesteps=1000,
                 mse_strength=1.0, mae_strength=0.0, unmasked_weight=0.1, masked_prior_preservation_weight=0.0,
                 loss_scale=1.0, loss_weight_strength=5.0, upstream=1.0, want_grad=True, pred_pack_w: int = 0):
    """Per-sample losses [B] (fp32) and d(sum_b upstream*losses[b])/d(predicted) in one pass.  With `pred_pack_w` = W the
    prediction (and its gradient) are in FLUX's packed layout [B, (H/2)(W/2), 4C]; the target stays [B, C, H, W]."""
    N.require_cuda(predicted, target, prior_target, mask, timestep, sqrt_ac, sqrt_1mac, sigmas)
    if pred_pack_w:
        if target.dim() != 4 or predicted.dim() != 3 or predicted.numel() != target.numel() or target.shape[-1] != pred_pack_w \
                or tuple(predicted.shape[1:]) != (target.shape[2] * target.shape[3] // 4, 4 * target.shape[1]):
            raise ValueError(f"loss_fwd_bwd: packed prediction {tuple(predicted.shape)} does not match target {tuple(target.shape)}")
    elif predicted.shape != target.shape:
        raise ValueError("loss_fwd_bwd: predicted / target shapes differ")
    predicted = predicted.contiguous()
    target = target.contiguous()
    B = target.shape[0]
    Cc = target.shape[1] if target.dim() > 2 else 1
    hw = target[0].numel() // Cc
    dev = predicted.device
    if prior_target is not None:
        prior_target = prior_target.to(dtype=target.dtype).contiguous()
    if mask is not None:
        mask = mask.contiguous()
        if mask.numel() != B * hw:
            raise ValueError(f"loss_fwd_bwd: mask has {mask.numel()} elements, expected B*H*W = {B * hw}")
    if loss_weight is not None:
        loss_weight = loss_weight.to(device=dev, dtype=torch.float32).contiguous()
    if timestep is not None:
        timestep = timestep.to(dtype=torch.int32).contiguous()
        if timestep.numel() == 1 and B > 1:
            timestep = timestep.expand(B).contiguous()
    bps = N.lib().otb_loss_blocks_per_sample(B, Cc, hw)
    partials, counters, mask_stats = _loss_ws.get(dev, B, bps)
    losses = torch.empty((B,), dtype=torch.float32, device=dev)
    dpred = torch.empty_like(predicted) if want_grad else None
    if not _aligned16(predicted, target, prior_target, mask, dpred):
        raise ValueError("loss_fwd_bwd: tensors must be 16-byte aligned")
    a = N.LossArgs(
        predicted=predicted.data_ptr(), target=target.data_ptr(), prior_target=N.ptr(prior_target), mask=N.ptr(mask),
        loss_weight=N.ptr(loss_weight), timestep=N.ptr(timestep), sqrt_ac=N.ptr(sqrt_ac), sqrt_1mac=N.ptr(sqrt_1mac),
        sigmas=N.ptr(sigmas), losses=losses.data_ptr(), dpred=N.ptr(dpred), partials=partials.data_ptr(),
        counters=counters.data_ptr(), mask_stats=mask_stats.data_ptr(),
        B=B, C=Cc, hw=hw, pred_dtype=N.dtype_code(predicted.dtype), target_dtype=N.dtype_code(target.dtype),
        mask_dtype=N.dtype_code(mask.dtype) if mask is not None else N.F32,
        masked=1 if masked else 0, normalize_masked_area_loss=1 if normalize_masked_area_loss else 0,
        loss_weight_fn=int(loss_weight_fn), v_prediction=1 if v_prediction else 0,
        num_train_timesteps=int(num_train_timesteps), blocks_per_sample=bps,
        mse_strength=float(mse_strength), mae_strength=float(mae_strength), unmasked_weight=float(unmasked_weight),
        masked_prior_preservation_weight=float(masked_prior_preservation_weight), loss_scale=float(loss_scale),
        loss_weight_strength=float(loss_weight_strength), upstream=float(upstream), pred_pack_w=int(pred_pack_w))
    N.check(N.lib().otb_loss_fwd_bwd(C.byref(a), _stream()), "otb_loss_fwd_bwd")
    _count(2 if (normalize_masked_area_loss and mask is not None) else 1)
    return losses, dpred


def loss_bwd_rescale(dpred: torch.Tensor, grad_out: torch.Tensor, assumed: float) -> None:
    B = dpred.shape[0]
    go = grad_out.to(dtype=torch.float32).contiguous()
    N.check(N.lib().otb_loss_bwd_rescale(dpred.data_ptr(), N.dtype_code(dpred.dtype), go.data_ptr(), float(assumed),
                                         B, dpred[0].numel(), _stream()), "otb_loss_bwd_rescale")
    _count()


# --------------------------------------------------------------------------------------
# K-c: multi-tensor tables
# --------------------------------------------------------------------------------------
class OptimTable:
    """Device-resident pointer table for the multi-tensor kernels.

    `entries` is a list of dicts with tensors p, g, exp_avg, exp_avg_sq, ema, shadow (any may be
    None) — all of one tensor sharing a dtype.  The chunk map is static; per-step scalars and
    gradient pointers are refreshed on the host (numpy, vectorised) and re-uploaded with one
    pinned async copy.
    """

    def __init__(self, entries: list[dict], device, chunk_elems: int = CHUNK_ELEMS):
        self.device = torch.device(device)
        self.n = len(entries)
        self.chunk_elems = chunk_elems
        self.host = np.zeros(self.n, dtype=_OPT_DTYPE)
        # ring of pinned staging buffers: an async H2D copy reads its source when the GPU gets to it, by which time a
        # host running ahead may already be writing the NEXT step's scalars — each upload takes the next slot and a
        # slot is rewritten only after the copy that last read it has completed
        nbytes = max(self.n, 1) * _OPT_DTYPE.itemsize
        self._pinned = [torch.empty((nbytes,), dtype=torch.uint8).pin_memory() for _ in range(PinnedRing.SLOTS)] \
            if torch.cuda.is_available() else None
        self._ring = PinnedRing()
        self.dev = torch.empty((max(self.n, 1) * _OPT_DTYPE.itemsize,), dtype=torch.uint8, device=self.device)
        chunk_tensor, chunk_index = [], []
        rng_base = 0
        self.numels = []
        for i, e in enumerate(entries):
            p = e["p"]
            n = p.numel()
            self.numels.append(n)
            h = self.host[i]
            h["numel"] = n
            # position in the stochastic-rounding stream: by default the running offset inside this table; a sharded
            # optimizer (dp.ShardedExchange) passes the GLOBAL offset so that its bits equal the replicated run's
            h["rng_base"] = e.get("rng_base", rng_base) if e.get("rng_base") is not None else rng_base
            rng_base += (n + 7) // 8 * 8
            h["p_dtype"] = N.dtype_code(p.dtype)
            nch = (n + chunk_elems - 1) // chunk_elems
            chunk_tensor += [i] * nch
            chunk_index += list(range(nch))
        self.total_rng = rng_base
        self.n_chunks = len(chunk_tensor)
        self.chunk_tensor = torch.tensor(chunk_tensor, dtype=torch.int32).to(self.device)
        self.chunk_index = torch.tensor(chunk_index, dtype=torch.int32).to(self.device)
        self.partials = torch.empty((max(self.n_chunks, 1),), dtype=torch.float32, device=self.device)
        self.norm_coef = torch.zeros((2,), dtype=torch.float32, device=self.device)
        self.set_pointers(entries)

    def set_pointers(self, entries: list[dict]) -> None:
        h = self.host
        for i, e in enumerate(entries):
            ptrs = [e.get(k) for k in ("p", "g", "exp_avg", "exp_avg_sq", "ema", "shadow")]
            for name, t in zip(("p", "g", "exp_avg", "exp_avg_sq", "ema", "shadow_bf16"), ptrs):
                if t is not None:
                    if not t.is_contiguous():
                        raise ValueError("multi-tensor kernels need contiguous tensors")
                    if name not in ("shadow_bf16",) and t.dtype != e["p"].dtype:
                        raise NotImplementedError(
                            f"optimizer/EMA tensor '{name}' has dtype {t.dtype} but the parameter is {e['p'].dtype}")
                    if t.numel() != e["p"].numel():
                        raise ValueError(f"'{name}' has {t.numel()} elements, parameter has {e['p'].numel()}")
                h[name][i] = 0 if t is None else t.data_ptr()
            aligned = all(t is None or (t.data_ptr() & 15) == 0 for t in ptrs)
            h["flags"][i] = (int(h["flags"][i]) & ~N.OPT_FLAG_ALIGNED16) | (N.OPT_FLAG_ALIGNED16 if aligned else 0)

    def set_grad_pointers(self, grads: list[torch.Tensor | None]) -> None:
        """Refresh the gradient column (gradients are new tensors after zero_grad(set_to_none=True))."""
        ptrs = np.fromiter((0 if g is None else g.data_ptr() for g in grads), dtype=np.uint64, count=self.n)
        self.host["g"][:] = ptrs
        if np.any(ptrs & np.uint64(15)):
            for i in np.nonzero(ptrs & np.uint64(15))[0]:
                self.host["flags"][i] = int(self.host["flags"][i]) & ~N.OPT_FLAG_ALIGNED16
        for i in self._grad_checked_upto(grads):
            g = grads[i]
            if g is not None and (not g.is_contiguous() or g.numel() != self.numels[i]):
                raise ValueError("gradient must be contiguous and match its parameter")

    def _grad_checked_upto(self, grads):
        # full validation on the first call, then only a spot check per step (host time matters:
        # the reference trainer calls step() once per optimisation step on ~1.4 k tensors)
        if not getattr(self, "_grads_validated", False):
            self._grads_validated = True
            return range(len(grads))
        return range(0, len(grads), max(1, len(grads) // 8))

    def upload(self) -> None:
        """Host table -> device, asynchronously on the current stream.  NOT capturable: the host->pinned memcpy
        happens now, in python — a captured step uploads before the replay (harness/trainer.py)."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("OptimTable.upload() inside a CUDA-graph capture would freeze this step's scalars into every "
                               "replay; upload before graph.replay()")
        raw = self.host.view(np.uint8).reshape(-1)
        if self._pinned is not None:
            slot = self._ring.acquire()
            self._pinned[slot][: raw.size].numpy()[:] = raw
            self.dev[: raw.size].copy_(self._pinned[slot][: raw.size], non_blocking=True)
            self._ring.release(slot)
        else:
            self.dev[: raw.size].copy_(torch.from_numpy(raw.copy()))

    def launch_args(self, *, flavor=N.FLAVOR_TORCH_CUDA, seed=0, offset=0, ema_one_minus_decay=0.0, do_ema=False,
                    clip_coef: torch.Tensor | None = None, dyn: torch.Tensor | None = None) -> N.OptimLaunch:
        return N.OptimLaunch(
            tensors=self.dev.data_ptr(), chunk_tensor=self.chunk_tensor.data_ptr(), chunk_index=self.chunk_index.data_ptr(),
            n_tensors=self.n, n_chunks=self.n_chunks, chunk_elems=self.chunk_elems, flavor=int(flavor),
            seed=int(seed) & (2**64 - 1), offset=int(offset) & (2**64 - 1),
            ema_one_minus_decay=float(ema_one_minus_decay), do_ema=1 if do_ema else 0, clip_coef=N.ptr(clip_coef), dyn=N.ptr(dyn))


def adam_scalars(lr: float, beta1: float, beta2: float, eps: float, weight_decay: float, step: float,
                 decoupled: bool) -> dict:
    """Per-tensor float scalars exactly as the reference derives them in python doubles and torch
    then narrows to fp32 (adamw_extensions.py:83-132)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    f = np.float32
    return dict(
        decay_mul=f(1 - lr * weight_decay) if decoupled else f(1.0),
        l2=f(0.0) if decoupled else f(weight_decay),
        lerp_w=f(1 - beta1), beta2=f(beta2), one_minus_beta2=f(1 - beta2),
        bc2_sqrt=f(bc2_sqrt), inv_bc2_sqrt=f(1.0 / bc2_sqrt), eps=f(eps), neg_step_size=f(-step_size))


def adamw_sr_ema_multi(table: OptimTable, peer_store: "N.PeerStore | None" = None, **kw) -> None:
    """`peer_store` (K-z, dp.ShardedExchange.peer_store): every parameter store also goes to the other ranks' replicas of the
    symmetric parameter arena — the all-gather half of the sharded exchange folded into the optimizer launch."""
    a = table.launch_args(**kw)
    if peer_store is not None:
        N.check(N.lib().otb_adamw_sr_ema_multi_peers(C.byref(a), C.byref(peer_store), _stream()), "otb_adamw_sr_ema_multi_peers")
    else:
        N.check(N.lib().otb_adamw_sr_ema_multi(C.byref(a), _stream()), "otb_adamw_sr_ema_multi")
    _count()


def ema_multi(table: OptimTable, one_minus_decay: float) -> None:
    a = table.launch_args(ema_one_minus_decay=one_minus_decay, do_ema=True)
    N.check(N.lib().otb_ema_multi(C.byref(a), _stream()), "otb_ema_multi")
    _count()


def grad_norm_multi(table: OptimTable, max_norm: float, grad_prescale: float = 1.0) -> torch.Tensor:
    """Returns the device tensor [total_norm, clip_coef] (no host sync)."""
    a = table.launch_args()
    N.check(N.lib().otb_grad_norm_multi(C.byref(a), float(max_norm), float(grad_prescale), table.partials.data_ptr(),
                                        table.norm_coef.data_ptr(), _stream()), "otb_grad_norm_multi")
    _count(2)
    return table.norm_coef


def grad_norm_from_partials(table: OptimTable, partials: torch.Tensor, max_norm: float, grad_prescale: float = 1.0) -> torch.Tensor:
    """[total_norm, clip_coef] from squared-norm partials another kernel produced (the NVLS gradient exchange)."""
    N.require_cuda(partials)
    N.check(N.lib().otb_grad_norm_from_partials(partials.data_ptr(), partials.numel(), float(max_norm), float(grad_prescale),
                                                table.norm_coef.data_ptr(), _stream()), "otb_grad_norm_from_partials")
    _count()
    return table.norm_coef


def grad_scale_multi(table: OptimTable, coef: torch.Tensor) -> None:
    a = table.launch_args()
    N.check(N.lib().otb_grad_scale_multi(C.byref(a), coef.data_ptr(), _stream()), "otb_grad_scale_multi")
    _count()


def stochastic_round(mode: int, target: torch.Tensor, src: torch.Tensor, src2: torch.Tensor | None = None, *, alpha: float = 1.0,
                     seed: int = 0, offset: int = 0, rng_base: int = 0, flavor: int = N.FLAVOR_TORCH_CUDA) -> None:
    """One launch of otb_stochastic_round (include/otb200.h): N.SR_COPY target = SR(src); N.SR_ADD target = SR(fp32(src) +
    alpha * target); N.SR_ADDCDIV target = SR(target + alpha * src / src2).  target: contiguous bf16, written in place."""
    N.require_cuda(target, src, src2)
    if target.dtype != torch.bfloat16 or not target.is_contiguous():
        raise ValueError("stochastic_round: target must be a contiguous bf16 tensor")
    for t in (src, src2):
        if t is not None and (not t.is_contiguous() or t.numel() != target.numel() or t.dtype != src.dtype):
            raise ValueError("stochastic_round: operands must be contiguous, of the target's size and of one dtype")
    a = N.SrArgs(target=target.data_ptr(), src=src.data_ptr(), src2=N.ptr(src2), n=target.numel(), rng_base=int(rng_base),
                 seed=int(seed) & (2**64 - 1), offset=int(offset) & (2**64 - 1), mode=int(mode), src_dtype=N.dtype_code(src.dtype),
                 alpha=float(alpha), flavor=int(flavor))
    N.check(N.lib().otb_stochastic_round(C.byref(a), _stream()), "otb_stochastic_round")
    if target.numel():
        _count()


class ShadowTable:
    """bf16 (and transposed bf16) shadows of a list of 2-D master tensors, refreshed in one launch."""

    def __init__(self, masters: list[torch.Tensor], device, transposed: bool = True):
        self.device = torch.device(device)
        self.masters = masters
        total = sum(m.numel() for m in masters)
        self.flat = torch.empty((total,), dtype=torch.bfloat16, device=self.device)
        self.flat_t = torch.empty((total,), dtype=torch.bfloat16, device=self.device) if transposed else None
        host = np.zeros(len(masters), dtype=_SHADOW_DTYPE)
        self.views, self.views_t = [], []
        off, tile = 0, 0
        for i, m in enumerate(masters):
            if m.dim() != 2 or not m.is_contiguous():
                raise ValueError("ShadowTable masters must be contiguous 2-D tensors")
            # keep every shadow 16-byte aligned (TMA base alignment): pad offsets to 8 elements
            r, c = m.shape
            v = self.flat[off:off + r * c].view(r, c)
            vt = self.flat_t[off:off + r * c].view(c, r) if transposed else None
            self.views.append(v)
            self.views_t.append(vt)
            host[i]["dst"] = v.data_ptr()
            host[i]["dst_t"] = vt.data_ptr() if vt is not None else 0
            host[i]["rows"], host[i]["cols"] = r, c
            host[i]["src_dtype"] = N.dtype_code(m.dtype)
            host[i]["tile_base"] = tile
            tile += ((r + 31) // 32) * ((c + 31) // 32)
            off += r * c
            if (r * c) % 8:
                raise NotImplementedError("ShadowTable: tensor sizes must be multiples of 8 elements")
        self.n_tiles = tile
        self.host = host
        self.dev = torch.empty((max(len(masters), 1) * _SHADOW_DTYPE.itemsize,), dtype=torch.uint8, device=self.device)
        self.sync_pointers()

    def sync_pointers(self) -> None:
        for i, m in enumerate(self.masters):
            self.host[i]["src"] = m.data_ptr()
            self.host[i]["src_dtype"] = N.dtype_code(m.dtype)
        raw = self.host.view(np.uint8).reshape(-1)
        self.dev[: raw.size].copy_(torch.from_numpy(raw.copy()))
        self._ptrs = [m.data_ptr() for m in self.masters]

    def refresh(self) -> None:
        if any(m.data_ptr() != p for m, p in zip(self.masters, self._ptrs)):
            self.sync_pointers()
        N.check(N.lib().otb_shadow_refresh_multi(self.dev.data_ptr(), len(self.masters), self.n_tiles, _stream()),
                "otb_shadow_refresh_multi")
        _count()


def philox_dump(seed: int, offset: int, stream_id: int, first_group: int, n_groups: int, device) -> torch.Tensor:
    out = torch.empty((n_groups, 4), dtype=torch.int32, device=device)
    N.check(N.lib().otb_philox_dump(seed, offset, stream_id, first_group, n_groups, out.data_ptr(), _stream()),
            "otb_philox_dump")
    return out
