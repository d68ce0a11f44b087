"""DoRA on the fused path (-m gpu): DoRAModule (reference modules/module/LoRAModule.py:333-419) against outputs of the
UNMODIFIED reference (tests/golden/dora.npz, CPU fp32) and against the reference's arithmetic restated in fp32 torch on
bf16-representable operands."""
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(dev):
    import onetrainer_b200.modules.module.LoRAModule as mod
    from onetrainer_b200 import _native as N
    N.lib()
    return mod


def _base(g, name, dev, dtype=torch.float32):
    w, bias = torch.from_numpy(g[f"{name}_w"]), g[f"{name}_bias"]
    if name.startswith("conv"):
        base = nn.Conv2d(w.shape[1], w.shape[0], w.shape[2], padding=w.shape[2] // 2, bias=bias.size > 0)
    else:
        base = nn.Linear(w.shape[1], w.shape[0], bias=bias.size > 0)
    with torch.no_grad():
        base.weight.copy_(w)
        if bias.size:
            base.bias.copy_(torch.from_numpy(bias))
    return base.to(dev).to(dtype).requires_grad_(False)


@pytest.mark.parametrize("name", ["linear", "linear_outaxis_eps", "conv3x3"])
def test_dora_module_vs_reference_golden(L, dev, golden, name):
    """The reference's CPU fp32 DoRA forward / backward vs the hooked fused module computing in bf16 (x, WP and dY are
    rounded to bf16 — 2^-9 relative each — and contracted over K <= 72 terms): 2 bf16 ulp of the value + 1 % of the
    tensor's scale, for y, dX, dA, dB and d(dora_scale); `dora_scale` is initialised to the reference's column norms and
    the state-dict keys are the reference's."""
    g = golden("dora")
    eps_flag, out_axis = (bool(v) for v in g[f"{name}_flags"])
    base = _base(g, name, dev)
    mod = L.DoRAModule("lora_unet.x", base, 8, 4.0, norm_epsilon=eps_flag, decompose_output_axis=out_axis, train_device=dev).to(dev)
    w = base.weight.detach()
    expect = w.reshape(w.shape[0], -1).norm(dim=1) if out_axis else w.transpose(0, 1).reshape(w.shape[1], -1).norm(dim=1)
    assert mod.dora_scale.shape == g[f"{name}_m"].shape and torch.allclose(mod.dora_scale.flatten(), expect, rtol=1e-6)
    with torch.no_grad():
        mod.lora_down.weight.copy_(torch.from_numpy(g[f"{name}_A"]))
        mod.lora_up.weight.copy_(torch.from_numpy(g[f"{name}_B"]))
        mod.dora_scale.copy_(torch.from_numpy(g[f"{name}_m"]))
    assert sorted(mod.state_dict(prefix=mod.prefix)) == list(g[f"{name}_keys"])
    mod.hook_to_module()
    x = torch.from_numpy(g[f"{name}_x"]).to(dev).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = base(x)
    assert y.dtype == torch.bfloat16 and y.shape == g[f"{name}_y"].shape
    dy = torch.from_numpy(g[f"{name}_dy"]).to(dev)
    dx, dA, dB, dm = torch.autograd.grad(y, [x, mod.lora_down.weight, mod.lora_up.weight, mod.dora_scale], dy.to(y.dtype))

    def close(a, ref, what):
        ref = torch.from_numpy(ref).to(dev)
        tol = ref.abs() * 2.0 ** -7 + 0.01 * ref.abs().max()
        assert a.shape == ref.shape
        assert ((a.float() - ref).abs() <= tol).all(), f"{what}: max err {(a.float() - ref).abs().max().item()} vs scale {ref.abs().max().item()}"

    close(y, g[f"{name}_y"], "y")
    close(dx, g[f"{name}_dx"], "dx")
    close(dA, g[f"{name}_dA"], "dA")
    close(dB, g[f"{name}_dB"], "dB")
    close(dm, g[f"{name}_dm"], "dm")
    assert dA.dtype == torch.float32 and dm.dtype == torch.float32


def test_dora_at_sdxl_leaf_shape_vs_fp32_restatement(L, dev):
    """A real leaf shape (M=4096, 1280 -> 1280, rank 32): the reference's formulae in fp32 torch on bf16-representable
    x / W / dY.  y and dX: 2 bf16 ulp (floor rms/8); the weight-space gradients go through G = dY^T x rounded to bf16 (as
    the reference's autocast GEMM leaves it): 1 % of the tensor's rms."""
    torch.manual_seed(0)
    M, K, N, r = 4096, 1280, 1280, 32
    r16 = lambda t: t.bfloat16().float()   # noqa: E731
    x, dy = r16(torch.randn(M, K, device=dev)), r16(torch.randn(M, N, device=dev))
    W = r16(torch.randn(N, K, device=dev) / K ** 0.5)
    base = nn.Linear(K, N, bias=False, device=dev, dtype=torch.bfloat16).requires_grad_(False)
    with torch.no_grad():
        base.weight.copy_(W)
    mod = L.DoRAModule("lora_unet.x", base, r, 16.0, norm_epsilon=True, decompose_output_axis=False, train_device=dev).to(dev)
    with torch.no_grad():
        mod.lora_up.weight.normal_(0, 0.02)
        mod.dora_scale.mul_(1 + 0.1 * torch.randn_like(mod.dora_scale))
    A, B, m = (t.detach().clone().requires_grad_(True) for t in (mod.lora_down.weight, mod.lora_up.weight, mod.dora_scale))
    V = W + (B @ A) * (16.0 / r)
    norm = V.detach().norm(dim=0, keepdim=True) + torch.finfo(torch.float32).eps
    WP = m * (V / norm)
    xr = x.clone().requires_grad_(True)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yr = xr @ WP.bfloat16().float().t()                 # the GEMM operand is bf16 on both sides (autocast casts WP)
        dxr, = torch.autograd.grad(yr, [xr], dy, retain_graph=True)
        (xr.detach() @ WP.t() * dy).sum().backward()        # weight-space gradients of the unrounded composition
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    mod.hook_to_module()
    xm = x.bfloat16().requires_grad_(True)
    ym = base(xm)
    dxm, dAm, dBm, dmm = torch.autograd.grad(ym, [xm, mod.lora_down.weight, mod.lora_up.weight, mod.dora_scale], dy.bfloat16())

    def ulps(mine, ref, what, floor):
        ref = ref.float()
        rms = ref.pow(2).mean().sqrt()
        tol = 2 * torch.exp2(torch.floor(torch.log2(torch.maximum(ref.abs(), rms * floor))) - 7)
        err = (mine.float() - ref).abs()
        assert (err <= tol).all(), (what, float((err / tol).max()))

    ulps(ym, yr.detach(), "y", 1 / 8)
    ulps(dxm, dxr, "dX", 1 / 8)
    for mine, ref, what in ((dAm, A.grad, "dA"), (dBm, B.grad, "dB"), (dmm, m.grad, "dm")):
        rms = ref.pow(2).mean().sqrt().item()
        assert (mine - ref).abs().max().item() <= 0.01 * rms * 4 + 1e-9, (what, (mine - ref).abs().max().item(), rms)


def test_wrapper_builds_dora_and_trains(dev):
    """config.lora_decompose selects DoRAModule in the wrapper (reference :448-457); the tiny stand-in trains through the
    harness step (fused noise / loss / AdamW) and the loss falls; state-dict round trip keeps dora_scale."""
    from onetrainer_b200.harness import trainer
    from onetrainer_b200.modules.module.LoRAModule import DoRAModule, LoRAModuleWrapper
    from onetrainer_b200.modules.util.config.TrainConfig import TrainConfig
    cfg = TrainConfig(train_device=str(dev), batch_size=4, lora_rank=16, lora_alpha=8.0, lora_decompose=True)
    model, setup = trainer.build_model("tiny", cfg, dev, stochastic_rounding=False, lora_up_std=0.05, lr=2e-3)
    mods = list(model.unet_lora.lora_modules.values())
    assert all(isinstance(m, DoRAModule) for m in mods) and not model.unet_lora.sibling_groups and not model.unet_lora.context_launches
    assert len(model.parameters) == 3 * len(mods)
    batch = trainer.batch_to_device(trainer.synthetic_batch("tiny", 4, dev, latent_hw=32), dev)
    losses = [trainer.train_step(setup, model, batch, cfg).item() for _ in range(20)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5])
    sd = model.unet_lora.state_dict()
    assert sum(k.endswith(".dora_scale") for k in sd) == len(mods)
    w2 = LoRAModuleWrapper(model.unet, "lora_unet", cfg, ["attentions"])
    w2.load_state_dict(sd)
    k0 = next(iter(model.unet_lora.lora_modules))
    assert torch.equal(w2.lora_modules[k0].dora_scale, model.unet_lora.lora_modules[k0].dora_scale)
