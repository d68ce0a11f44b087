"""Drop-in for the reference's modules/util/bf16_stochastic_rounding.py — same three functions, same argument meaning.

Reference: copy_stochastic_ (:5-30) draws a 16-bit integer per element with torch.randint_like, adds it to the int32 view of
the fp32 source, masks the low half off and copies the high half into the bf16 target — 4 launches and an int32 temporary;
add_stochastic_ (:33-45) and addcdiv_stochastic_ (:48-61) first build the fp32 result in another temporary.  The optimizers
fused in this repo (AdamW / Adam / Adafactor / CAME) do all of this inside their own launch; these stand-alone functions are
for every OTHER optimizer of the reference that imports them and calls them once per tensor.  Each call is ONE launch of
otb_stochastic_round (include/otb200.h, csrc/sr_round.cu): the fp32 result stays in registers.

Random bits: the reference draws from the default CUDA generator; so does this — the Philox key is that generator's
(seed, offset), and every call advances the offset, so `torch.manual_seed(s)` reproduces a run and successive calls differ.
Under data parallelism (net-new; the reference is single-process) the key is instead rank 0's seed and a call counter
(`dp.replicated_key`), so replicas of a parameter round identically whatever per-rank seeding the application uses.
torch's own bit stream is deliberately not reproduced; given the same 16-bit draws the result is bit-identical
to the reference (tests/golden/sr_helpers.npz, tests/test_gpu_kernels.py).

No fallback: CPU tensors, a target that is not bf16, or operands of another size raise.  A strided target is rounded in a dense
copy and written back (the reference's ATen ops accept any strides).
"""
from __future__ import annotations

import torch
from torch import Tensor

from ... import _native as N
from ... import dp, ops, rng


def _key(t: Tensor, *others: Tensor) -> tuple[int, int]:
    N.require_cuda(t, *others)                      # before anything touches a CUDA generator: CPU tensors get the clear error
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("bf16_stochastic_rounding: the default generator's offset cannot advance inside a CUDA-graph capture; "
                           "use the fused optimizers (their key is a device-side step counter)")
    if dp.world_size() > 1:
        # replicas must round identically or bf16 copies of one parameter drift apart: rank 0's seed and a call counter,
        # not this rank's generator (applications commonly seed with seed + rank)
        return dp.replicated_key()
    return rng.draw_key(None, t.device)


def _operand(t: Tensor, like: Tensor, what: str) -> Tensor:
    if t.numel() != like.numel():
        raise ValueError(f"{what} has {t.numel()} elements, the target has {like.numel()} (broadcasting is not part of this path)")
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.to(dtype=torch.float32)               # the reference's own `.to(dtype=torch.float32)`: exact for fp16
    return t.contiguous()


def _launch(mode: int, target: Tensor, src: Tensor, src2: Tensor | None = None, **kw) -> None:
    if target.dtype != torch.bfloat16:
        raise ValueError(f"stochastic rounding targets bf16 tensors; got {target.dtype}")
    if target.is_contiguous():
        ops.stochastic_round(mode, target, src, src2, **kw)
        return
    # a strided target (a transposed or sliced parameter): the kernel addresses dense memory, so round into a dense copy in the
    # target's logical element order — the order `_operand` gives the other operands — and write it back
    dense = target.contiguous()
    ops.stochastic_round(mode, dense, src, src2, **kw)
    target.copy_(dense)


def copy_stochastic_(target: Tensor, source: Tensor):
    """copies source (fp32) into target (bf16) using stochastic rounding"""
    if source.dtype != torch.float32:
        raise ValueError("copy_stochastic_: source must be float32 (the reference reinterprets it as int32)")
    seed, offset = _key(target, source)
    _launch(N.SR_COPY, target, _operand(source, target, "source"), seed=seed, offset=offset)


def add_stochastic_(input: Tensor, other: Tensor, alpha: float = 1.0):
    """input <- SR(fp32(other) + alpha * input): the reference's arithmetic (`result = other.float(); result.add_(input,
    alpha=alpha)`), in which alpha multiplies `input`."""
    seed, offset = _key(input, other)
    _launch(N.SR_ADD, input, _operand(other, input, "other"), alpha=alpha, seed=seed, offset=offset)


def addcdiv_stochastic_(input: Tensor, tensor1: Tensor, tensor2: Tensor, value: float = 1.0):
    """input <- SR(input + value * tensor1 / tensor2), rounded the way torch's CUDA addcdiv_ rounds (fma(value, t1 / t2, input))."""
    seed, offset = _key(input, tensor1, tensor2)
    t1, t2 = _operand(tensor1, input, "tensor1"), _operand(tensor2, input, "tensor2")
    if t1.dtype != t2.dtype:
        t1, t2 = t1.float(), t2.float()
    _launch(N.SR_ADDCDIV, input, t1, t2, alpha=value, seed=seed, offset=offset)
