"""Adafactor and CAME on the fused multi-tensor kernels (-m gpu) against outputs of the UNMODIFIED reference
(tests/golden/adafactor.npz, came.npz — recorded by tools/make_golden.py from transformers' Adafactor + the reference's
patch_adafactor, and from the reference's CAME class), through the reference-facing API: patch_adafactor(optimizer, sr) /
CAME(params, ...), `step()` and `step_parameter(p, group, i)`."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def bf16_from_bits(a, dev):
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).to(dev)


def dec(g, key, bf16, dev):
    return bf16_from_bits(g[key], dev) if bf16 else torch.from_numpy(g[key]).to(dev)


ADAFACTOR_CASES = ["matrix_f32_relative", "vector_f32_relative", "matrix_f32_fixed_lr_beta1_wd", "conv_f32_warmup", "matrix_bf16_sr"]


@pytest.mark.parametrize("name", ADAFACTOR_CASES)
@pytest.mark.parametrize("entry", ["step", "step_parameter"])
def test_adafactor_vs_reference_golden(dev, golden, name, entry):
    """step_adafactor_parameter (adafactor_extensions.py:17-93), three steps.  fp32 elementwise arithmetic with the
    kernel's fixed-order reductions against torch's: 1e-5 relative on the parameter, the factored / unfactored second
    moments, the first moment and RMS(p); the bf16 case (whose "stochastic" rounding the reference overwrites with
    round-to-nearest) may differ by one bf16 ulp where the fp32 value sits on a rounding boundary."""
    from transformers.optimization import Adafactor
    from onetrainer_b200.modules.util.optimizer.adafactor_extensions import patch_adafactor
    g = golden("adafactor")
    lr, eps0, eps1, clip, decay, beta1, wd, scale_p, rel, warm = g[f"{name}_meta"]
    kw = dict(lr=None if np.isnan(lr) else float(lr), eps=(float(eps0), float(eps1)), clip_threshold=float(clip), decay_rate=float(decay),
              beta1=None if np.isnan(beta1) else float(beta1), weight_decay=float(wd), scale_parameter=bool(scale_p),
              relative_step=bool(rel), warmup_init=bool(warm))
    bf16 = name.endswith("bf16_sr")
    p = torch.nn.Parameter(dec(g, f"{name}_p0", bf16, dev))
    opt = Adafactor([p], **kw)
    patch_adafactor(opt, bf16)
    for step in (1, 2, 3):
        p.grad = dec(g, f"{name}_g{step}", bf16, dev)
        if entry == "step":
            opt.step()
        else:
            opt.step_parameter(p, opt.param_groups[0], 0)
        ref = dec(g, f"{name}_p{step}", bf16, dev)
        st = opt.state[p]
        if bf16:
            diff = (p.detach().float() - ref.float()).abs()
            assert (diff <= ref.float().abs() * 2.0 ** -7).all() and (diff != 0).float().mean().item() < 0.02
            with torch.no_grad():
                p.copy_(ref)                               # continue from the reference's values: boundary cases do not compound
        else:
            torch.testing.assert_close(p.detach(), ref, rtol=1e-5, atol=1e-8)
        assert st["step"] == step and st["RMS"].item() == pytest.approx(float(g[f"{name}_rms{step}"][0]), rel=1e-6)
        for key in ("exp_avg", "exp_avg_sq", "exp_avg_sq_row", "exp_avg_sq_col"):
            if f"{name}_{key}{step}" in g:
                r = torch.from_numpy(g[f"{name}_{key}{step}"]).to(dev)
                assert st[key].dtype == torch.float32 and st[key].shape == r.shape
                torch.testing.assert_close(st[key], r, rtol=2e-5, atol=1e-6 * r.abs().max().item())
        assert set(st) == {"step", "RMS"} | {k for k in ("exp_avg", "exp_avg_sq", "exp_avg_sq_row", "exp_avg_sq_col") if f"{name}_{k}1" in g}


CAME_CASES = ["matrix_f32", "vector_f32_wd", "conv_f32", "matrix_bf16_sr_wd"]


@pytest.mark.parametrize("name", CAME_CASES)
def test_came_vs_reference_golden(dev, golden, name):
    """CAME.step_parameter (CAME.py:80-176), three steps.  The bf16 fixture was recorded with the reference's
    randint_like draws replaced by this repo's Philox stochastic-rounding bits (seed, call counter, element index), which
    is what the kernel draws itself: identical except where the fp32 sum lands within an ulp of a draw's carry boundary
    (< 2 % of elements, one bf16 ulp).  fp32: 1e-5 relative."""
    from onetrainer_b200.modules.util.optimizer.CAME import CAME
    g = golden("came")
    lr, eps0, eps1, clip, b0, b1, b2, wd, seed, rng_base = g[f"{name}_meta"]
    bf16 = "bf16" in name
    p = torch.nn.Parameter(dec(g, f"{name}_p0", bf16, dev))
    pad = torch.nn.Parameter(torch.zeros(int(rng_base), device=dev, dtype=p.dtype))     # occupies elements [0, rng_base) of the SR stream
    opt = CAME([pad, p], lr=float(lr), eps=(float(eps0), float(eps1)), clip_threshold=float(clip), betas=(float(b0), float(b1), float(b2)),
               weight_decay=float(wd), stochastic_rounding=bf16)
    torch.manual_seed(int(seed))
    for step in (1, 2, 3):
        p.grad = dec(g, f"{name}_g{step}", bf16, dev)
        pad.grad = torch.zeros_like(pad)
        opt.step()
        ref = dec(g, f"{name}_p{step}", bf16, dev)
        st = opt.state[p]
        if bf16:
            diff = (p.detach().float() - ref.float()).abs()
            assert (diff <= ref.float().abs() * 2.0 ** -7).all() and (diff != 0).float().mean().item() < 0.02, (diff != 0).float().mean().item()
            with torch.no_grad():
                p.copy_(ref)
        else:
            torch.testing.assert_close(p.detach(), ref, rtol=1e-5, atol=1e-8)
        for key in ("exp_avg", "exp_avg_sq", "exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_res_row", "exp_avg_res_col"):
            if f"{name}_{key}{step}" in g:
                r = torch.from_numpy(g[f"{name}_{key}{step}"]).to(dev)
                torch.testing.assert_close(st[key], r, rtol=5e-5, atol=1e-6 * r.abs().max().item())
    assert st["step"] == 3


def test_factored_optimizers_multi_tensor_equals_per_tensor_and_is_reproducible(dev):
    """One launch over a mixed list (matrices, vectors, a conv weight, fp32 and bf16) equals per-tensor step_parameter
    calls bit for bit (one CTA per tensor, fixed-order reductions), twice in a row; create_optimizer builds both."""
    from transformers.optimization import Adafactor
    from onetrainer_b200.modules.util.optimizer.CAME import CAME
    from onetrainer_b200.modules.util.optimizer.adafactor_extensions import patch_adafactor
    shapes = [((32, 1280), torch.float32), ((1280, 32), torch.float32), ((77,), torch.float32), ((16, 24, 3, 3), torch.float32),
              ((640, 16), torch.bfloat16), ((2048, 1536), torch.bfloat16)]

    def make(kind):
        torch.manual_seed(3)
        ps = [torch.nn.Parameter((torch.randn(*s, device=dev) * 0.05).to(d)) for s, d in shapes]
        if kind == "adafactor":
            o = Adafactor(ps, beta1=0.9, weight_decay=0.01)
            patch_adafactor(o, False)
        else:
            o = CAME(ps, lr=1e-4, weight_decay=0.01, stochastic_rounding=False)
        return ps, o

    for kind in ("adafactor", "came"):
        runs = []
        for mode in ("multi", "multi", "single"):
            ps, o = make(kind)
            gen = torch.Generator(device=dev).manual_seed(9)
            for _ in range(3):
                for p in ps:
                    p.grad = (torch.randn(p.shape, device=dev, generator=gen) * 1e-2).to(p.dtype)
                if mode == "multi":
                    o.step()
                else:
                    for i, p in enumerate(ps):
                        o.step_parameter(p, o.param_groups[0], i)
            runs.append([p.detach().clone() for p in ps])
            assert all(torch.isfinite(p.float()).all() for p in ps)
        for a, b, c in zip(*runs):
            assert torch.equal(a, b) and torch.equal(a, c), kind
    # the trainer-facing constructor (create.py:914-951)
    from onetrainer_b200.modules.util import create
    from onetrainer_b200.modules.util.NamedParameterGroup import NamedParameterGroup, NamedParameterGroupCollection
    from onetrainer_b200.modules.util.config.TrainConfig import TrainConfig, TrainOptimizerConfig
    from onetrainer_b200.modules.util.enum import Optimizer
    for which, cls_name in ((Optimizer.ADAFACTOR, "Adafactor"), (Optimizer.CAME, "CAME")):
        col = NamedParameterGroupCollection()
        w = torch.nn.Parameter(torch.randn(8, 16, device=dev) * 0.05)
        col.add_group(NamedParameterGroup("unet_lora", [w], None))
        cfg = TrainConfig(train_device=str(dev), learning_rate=1e-4, optimizer=TrainOptimizerConfig(optimizer=which, relative_step=True))
        opt = create.create_optimizer(col, None, cfg)
        assert type(opt).__name__ == cls_name
        before = w.detach().clone()
        w.grad = torch.randn_like(w) * 1e-2
        opt.step()
        assert not torch.equal(w, before) and hasattr(opt, "step_parameter")


@pytest.mark.parametrize("kind", ["adafactor", "came"])
def test_factored_optimizers_large_and_odd_tensors_vs_reference_on_gpu(dev, kind):
    """The cluster path (8 CTAs per tensor, DSMEM sums) and both access widths against the UNMODIFIED reference optimizer
    (baseline/_ref: patch_adafactor over transformers' Adafactor; the reference's CAME) running its ATen op sequence on the same
    GPU: large matrices (> 128 k elements -> one cluster of 8 per tensor, 4-wide accesses), a matrix whose column count is not
    a multiple of 4 (scalar accesses), a batch of small matrices, a vector.  fp32, three steps: parameters to 2e-5 relative
    (reduction order differs), factored / first moments to 1e-4."""
    from baseline import ref_arm
    if not ref_arm.available():
        pytest.skip("baseline/_ref did not travel")
    ref_arm.load()
    from transformers.optimization import Adafactor
    shapes = [(1024, 640), (515, 333), (2560, 1280), (6, 96, 40), (3000,), (1280, 32)]

    def make(reference: bool):
        torch.manual_seed(11)
        ps = [torch.nn.Parameter(torch.randn(*s, device=dev) * 0.05) for s in shapes]
        if kind == "adafactor":
            o = Adafactor(ps, lr=None, beta1=0.9, weight_decay=0.01, relative_step=True, scale_parameter=True, warmup_init=False)
            if reference:
                from modules.util.optimizer.adafactor_extensions import patch_adafactor as ref_patch
                ref_patch(o, False)
            else:
                from onetrainer_b200.modules.util.optimizer.adafactor_extensions import patch_adafactor
                patch_adafactor(o, False)
        else:
            if reference:
                from modules.util.optimizer.CAME import CAME as RefCAME
                o = RefCAME(ps, lr=1e-4, weight_decay=0.01, stochastic_rounding=False)
            else:
                from onetrainer_b200.modules.util.optimizer.CAME import CAME
                o = CAME(ps, lr=1e-4, weight_decay=0.01, stochastic_rounding=False)
        return ps, o

    results = []
    for reference in (True, False):
        ps, o = make(reference)
        gen = torch.Generator(device=dev).manual_seed(5)
        for _ in range(3):
            for p in ps:
                p.grad = torch.randn(p.shape, device=dev, generator=gen) * 1e-2
            o.step()
        results.append((ps, o))
    (rp, ro), (mp, mo) = results
    for a, b, shape in zip(rp, mp, shapes):
        torch.testing.assert_close(b.detach(), a.detach(), rtol=2e-5, atol=1e-7, msg=lambda m: f"{kind} {shape}: {m}")
        sa, sb = ro.state[a], mo.state[b]
        for key in ("exp_avg", "exp_avg_sq", "exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_res_row", "exp_avg_res_col"):
            if key in sa:
                assert key in sb, (kind, shape, key)
                torch.testing.assert_close(sb[key], sa[key].to(sb[key].dtype), rtol=1e-4, atol=1e-6 * float(sa[key].abs().max()),
                                           msg=lambda m: f"{kind} {shape} {key}: {m}")
