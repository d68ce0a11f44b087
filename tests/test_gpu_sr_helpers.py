"""-m gpu parity of otb_stochastic_round (csrc/sr_round.cu) and of the drop-in module
onetrainer_b200/modules/util/bf16_stochastic_rounding.py, which replaces the reference's
modules/util/bf16_stochastic_rounding.py:5-61 (copy_stochastic_ / add_stochastic_ / addcdiv_stochastic_).

  (1) the golden fixture recorded from the unmodified reference (tests/golden/sr_helpers.npz, TORCH_CPU flavour): bit-exact,
  (2) the oracle on seeded inputs, ragged sizes and misaligned views (scalar path), both flavours: bit-exact,
  (3) the reference's own op sequence run by torch on this GPU with the same 16-bit draws (TORCH_CUDA flavour) — restated,
      and the unmodified reference file itself (baseline/_ref) with torch.randint_like answering from the same stream: bit-exact,
  (4) properties at the roofline size: unbiased within 5 sigma, every result one of the two bf16 neighbours.
Everything is integer / bit-level work once the draws are shared, so the bar is equality of bits."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ["copy", "add_update", "add_decay", "add_half", "add_bf16", "add_third", "add_third_bf16", "addcdiv_bf16", "addcdiv_f32"]


@pytest.fixture(scope="module")
def K(dev):
    from onetrainer_b200 import _native as N
    from onetrainer_b200 import ops
    N.lib()

    class Ns:
        pass

    ns = Ns()
    ns.N, ns.ops = N, ops
    return ns


def bits16(t):
    return t.detach().contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


def to_dev(a, dev):
    """fixture array -> device tensor: uint16 carries bf16 bits, float32 is itself"""
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).to(dev)
    return torch.from_numpy(a.copy()).to(dev)


@pytest.mark.parametrize("name", CASES)
def test_sr_helpers_bit_exact_vs_reference_golden(K, dev, golden, name):
    """Every case of the fixture through ONE launch each: the corners of the integer trick (zeros, exact bf16 values, a
    mantissa carrying into the exponent, denormals), `other` aliasing `input`, bf16 and fp32 operands, n = 4099 (ragged tail)."""
    g = golden("sr_helpers")
    seed, base = (int(x) for x in g["meta"])
    if name == "copy":
        src = to_dev(g["copy_src"], dev)
        tgt = torch.zeros(src.numel(), device=dev, dtype=torch.bfloat16)
        K.ops.stochastic_round(K.N.SR_COPY, tgt, src, seed=seed, offset=1, rng_base=base, flavor=K.N.FLAVOR_TORCH_CPU)
    else:
        alpha, off = g[f"{name}_meta"]
        tgt = to_dev(g[f"{name}_in"], dev)
        if name.startswith("addcdiv"):
            K.ops.stochastic_round(K.N.SR_ADDCDIV, tgt, to_dev(g[f"{name}_t1"], dev), to_dev(g[f"{name}_t2"], dev), alpha=float(alpha),
                                   seed=seed, offset=int(off), rng_base=base, flavor=K.N.FLAVOR_TORCH_CPU)
        else:
            other = tgt if name == "add_decay" else to_dev(g[f"{name}_other"], dev)       # add_decay: add_stochastic_(p, p, alpha)
            K.ops.stochastic_round(K.N.SR_ADD, tgt, other, alpha=float(alpha), seed=seed, offset=int(off), rng_base=base,
                                   flavor=K.N.FLAVOR_TORCH_CPU)
    torch.cuda.synchronize()
    got, want = bits16(tgt), g[f"{name}_out"]
    assert np.array_equal(got, want), f"{name}: {(got != want).sum()} of {want.size} differ"


@pytest.mark.parametrize("flavor_name", ["FLAVOR_TORCH_CPU", "FLAVOR_TORCH_CUDA"])
@pytest.mark.parametrize("n,skew", [(1, 0), (7, 0), (8, 0), (1000003, 0), (70001, 1), (4096, 3)])
def test_sr_helpers_bit_exact_vs_oracle(K, dev, oracle, flavor_name, n, skew):
    """Seeded inputs, ragged sizes, and views that start `skew` elements into an allocation (not 16-byte aligned: the scalar
    path) — the three modes in both arithmetic flavours equal the oracle bit for bit."""
    flavor = getattr(K.N, flavor_name)
    oflavor = getattr(oracle, flavor_name)
    gen = torch.Generator().manual_seed(n + skew)
    seed, off, base = 0x5EED + n, 11, 64

    def dv(t):      # device copy whose first element sits `skew` elements into its buffer
        buf = torch.empty(t.numel() + skew, dtype=t.dtype, device=dev)
        buf[skew:].copy_(t)
        return buf[skew:]

    draws = oracle.sr_bits(seed, off, base, n)
    inp = (torch.randn(n, generator=gen) * 0.02).bfloat16()
    for src_dtype in (torch.float32, torch.bfloat16):
        other = (torch.randn(n, generator=gen) * 1e-3).to(src_dtype)
        t2 = (torch.rand(n, generator=gen) * 1e-3 + 1e-8).to(src_dtype)
        as_np = (lambda t: bits16(t)) if src_dtype == torch.bfloat16 else (lambda t: t.numpy())
        if src_dtype == torch.float32:
            tgt = dv(torch.zeros(n, dtype=torch.bfloat16))
            K.ops.stochastic_round(K.N.SR_COPY, tgt, dv(other), seed=seed, offset=off, rng_base=base, flavor=flavor)
            want = oracle.sr_helper(oracle.SR_COPY, np.zeros(n, np.uint16), other.numpy(), bits=draws, flavor=oflavor)
            assert np.array_equal(bits16(tgt), want)
        tgt = dv(inp)
        K.ops.stochastic_round(K.N.SR_ADD, tgt, dv(other), alpha=-0.37, seed=seed, offset=off, rng_base=base, flavor=flavor)
        want = oracle.sr_helper(oracle.SR_ADD, bits16(inp), as_np(other), alpha=-0.37, bits=draws, flavor=oflavor)
        assert np.array_equal(bits16(tgt), want), f"add {src_dtype}"
        tgt = dv(inp)
        K.ops.stochastic_round(K.N.SR_ADDCDIV, tgt, dv(other), dv(t2), alpha=-3e-4, seed=seed, offset=off, rng_base=base, flavor=flavor)
        want = oracle.sr_helper(oracle.SR_ADDCDIV, bits16(inp), as_np(other), as_np(t2), alpha=-3e-4, bits=draws, flavor=oflavor)
        assert np.array_equal(bits16(tgt), want), f"addcdiv {src_dtype}"


def _ref_copy(target, source, draws):
    """modules/util/bf16_stochastic_rounding.py:5-30 with the draw handed in instead of torch.randint_like"""
    result = draws.clone()
    result.add_(source.view(dtype=torch.int32))
    result.bitwise_and_(-65536)
    target.copy_(result.view(dtype=torch.float32))


@pytest.mark.parametrize("src_dtype", [torch.float32, torch.bfloat16])
def test_sr_helpers_vs_reference_op_sequence_on_this_gpu(K, dev, oracle, src_dtype):
    """What the reference does on a B200 — its op sequence as torch CUDA eager (fp32 temporaries, add_ / addcdiv_ kernels of
    torch 2.11) — fed the same draws: the one-launch kernel in the TORCH_CUDA flavour gives the same bf16 bits."""
    n, seed, off, base = 300007, 99, 5, 0
    draws = torch.from_numpy(oracle.sr_bits(seed, off, base, n).copy()).to(dev)
    torch.manual_seed(3)
    inp = (torch.randn(n, device=dev) * 0.02).bfloat16()
    other = (torch.randn(n, device=dev) * 1e-2).to(src_dtype)
    t2 = (torch.rand(n, device=dev) * 1e-3 + 1e-8).to(src_dtype)
    # add_stochastic_ (:33-45)
    ref = inp.clone()
    result = other.clone() if other.dtype == torch.float32 else other.to(dtype=torch.float32)
    result.add_(ref, alpha=0.3)
    _ref_copy(ref, result, draws)
    tgt = inp.clone()
    K.ops.stochastic_round(K.N.SR_ADD, tgt, other, alpha=0.3, seed=seed, offset=off, rng_base=base, flavor=K.N.FLAVOR_TORCH_CUDA)
    assert np.array_equal(bits16(tgt), bits16(ref)), f"add: {(bits16(tgt) != bits16(ref)).sum()} of {n} differ"
    # addcdiv_stochastic_ (:48-61)
    ref = inp.clone()
    result = ref.to(dtype=torch.float32)
    result.addcdiv_(other, t2, value=-1.7e-3)
    _ref_copy(ref, result, draws)
    tgt = inp.clone()
    K.ops.stochastic_round(K.N.SR_ADDCDIV, tgt, other, t2, alpha=-1.7e-3, seed=seed, offset=off, rng_base=base, flavor=K.N.FLAVOR_TORCH_CUDA)
    assert np.array_equal(bits16(tgt), bits16(ref)), f"addcdiv: {(bits16(tgt) != bits16(ref)).sum()} of {n} differ"


def test_sr_helpers_vs_the_unmodified_reference_module_on_this_gpu(K, dev, oracle, monkeypatch):
    """The reference's own file (baseline/_ref/modules/util/bf16_stochastic_rounding.py, byte for byte) run on this GPU with
    torch.randint_like answering from the library's Philox stream: its three functions and the one-launch kernels in the
    TORCH_CUDA flavour leave the same bf16 bits — fp32 and bf16 operands, `other` aliasing `input` (CAME's decay call)."""
    from baseline import ref_arm
    if not ref_arm.available():
        pytest.skip("baseline/_ref did not travel")
    ref_arm.load()
    import modules.util.bf16_stochastic_rounding as ref
    n, seed, base = 200003, 4242, 8
    state = {"offset": 0}

    def fake_randint_like(src, dtype=None, low=0, high=0):
        assert low == 0 and high == (1 << 16) and dtype == torch.int32 and src.numel() == n
        return torch.from_numpy(oracle.sr_bits(seed, state["offset"], base, n).copy()).to(src.device).view(src.shape)

    monkeypatch.setattr(torch, "randint_like", fake_randint_like)
    torch.manual_seed(9)
    inp = (torch.randn(n, device=dev) * 0.02).bfloat16()
    src = torch.randn(n, device=dev) * 0.02
    kw = dict(seed=seed, rng_base=base, flavor=K.N.FLAVOR_TORCH_CUDA)

    state["offset"] = 1
    want, got = torch.empty_like(inp), torch.empty_like(inp)
    ref.copy_stochastic_(want, src)
    K.ops.stochastic_round(K.N.SR_COPY, got, src, offset=1, **kw)
    assert np.array_equal(bits16(got), bits16(want))
    for off, other, alpha in ((2, torch.randn(n, device=dev) * 1e-4, 1.0), (3, None, -3e-6),
                              (4, (torch.randn(n, device=dev) * 1e-2).bfloat16(), -0.7)):
        state["offset"] = off
        want, got = inp.clone(), inp.clone()
        ref.add_stochastic_(want, want if other is None else other, alpha=alpha)
        K.ops.stochastic_round(K.N.SR_ADD, got, got if other is None else other, alpha=alpha, offset=off, **kw)
        assert np.array_equal(bits16(got), bits16(want)), f"add_stochastic_ case {off}: {(bits16(got) != bits16(want)).sum()} of {n} differ"
    for off, dt, value in ((5, torch.bfloat16, -3e-4), (6, torch.float32, -1.7e-3)):
        state["offset"] = off
        t1 = (torch.randn(n, device=dev) * 1e-3).to(dt)
        t2 = (torch.rand(n, device=dev) * 1e-3 + 1e-8).to(dt)
        want, got = inp.clone(), inp.clone()
        ref.addcdiv_stochastic_(want, t1, t2, value=value)
        K.ops.stochastic_round(K.N.SR_ADDCDIV, got, t1, t2, alpha=value, offset=off, **kw)
        assert np.array_equal(bits16(got), bits16(want)), f"addcdiv_stochastic_ {dt}: {(bits16(got) != bits16(want)).sum()} of {n} differ"


def test_module_functions_are_seeded_unbiased_and_refuse_fallbacks(dev):
    """The drop-in module (same three names as the reference's): torch.manual_seed reproduces a call, successive calls draw
    different bits, the mean of 2^24 roundings of 1 + k/2^10 is the fp32 value within 5 sigma, every result is one of the
    two bf16 neighbours (SURVEY.md §8c's known-answer test at the scale of the north_star), and CPU tensors raise."""
    from onetrainer_b200.modules.util.bf16_stochastic_rounding import add_stochastic_, addcdiv_stochastic_, copy_stochastic_
    n = 1 << 24
    for k in (1, 3, 5):
        value = 1 + k * 2.0 ** -10
        src = torch.full((n,), value, device=dev)
        tgt = torch.empty(n, device=dev, dtype=torch.bfloat16)
        copy_stochastic_(tgt, src)
        vals = tgt.float()
        assert set(torch.unique(vals).tolist()) <= {1.0, 1.0078125}
        frac = k / 8.0
        assert abs(vals.double().mean().item() - value) < 5 * 2.0 ** -7 * math.sqrt(frac * (1 - frac) / n)
    src = torch.randn(100003, device=dev) * 0.02
    a, b, c = (torch.empty(100003, device=dev, dtype=torch.bfloat16) for _ in range(3))
    torch.manual_seed(17)
    copy_stochastic_(a, src)
    copy_stochastic_(b, src)
    torch.manual_seed(17)
    copy_stochastic_(c, src)
    assert torch.equal(a, c) and not torch.equal(a, b)
    # the other two: every element lands on one of the two bf16 neighbours of the fp32 result
    p = (torch.randn(100003, device=dev) * 0.02).bfloat16()
    upd = torch.randn(100003, device=dev) * 1e-4
    exact = upd + 1.0 * p.float()
    q = p.clone()
    add_stochastic_(q, upd)
    lo = (exact.view(torch.int32) & -65536).view(torch.float32)
    hi = ((exact.view(torch.int32) & -65536) + 65536).view(torch.float32)
    assert bool(((q.float() == lo) | (q.float() == hi)).all())
    t1 = (torch.randn(100003, device=dev) * 1e-3).bfloat16()
    t2 = (torch.rand(100003, device=dev) * 1e-3 + 1e-8).bfloat16()
    exact = p.float().addcdiv(t1.float(), t2.float(), value=-3e-4)
    q = p.clone()
    addcdiv_stochastic_(q, t1, t2, value=-3e-4)
    lo = (exact.view(torch.int32) & -65536).view(torch.float32)
    hi = ((exact.view(torch.int32) & -65536) + 65536).view(torch.float32)
    assert bool(((q.float() == lo) | (q.float() == hi)).all())
    with pytest.raises((RuntimeError, ValueError)):
        copy_stochastic_(torch.empty(8, dtype=torch.bfloat16), torch.zeros(8))


def test_module_functions_accept_strided_targets(dev):
    """The reference's ATen ops take any strides; the kernel addresses dense memory.  A transposed (non-contiguous) bf16 target
    is rounded in a dense copy in its logical element order and written back: same bits as the same call on a contiguous
    tensor holding the same values, under the same generator state."""
    from onetrainer_b200.modules.util.bf16_stochastic_rounding import add_stochastic_, copy_stochastic_
    torch.manual_seed(1)
    base = (torch.randn(48, 80, device=dev) * 0.02).bfloat16()
    upd = torch.randn(80, 48, device=dev) * 1e-4
    strided, dense = base.clone().t(), base.t().contiguous()
    assert not strided.is_contiguous() and torch.equal(strided, dense)
    torch.manual_seed(21)
    add_stochastic_(strided, upd, alpha=1.0)
    torch.manual_seed(21)
    add_stochastic_(dense, upd, alpha=1.0)
    assert torch.equal(strided, dense) and not torch.equal(dense, base.t())
    torch.manual_seed(22)
    copy_stochastic_(strided, upd * 100)
    torch.manual_seed(22)
    copy_stochastic_(dense, upd * 100)
    assert torch.equal(strided, dense)
