"""Generate tests/golden/*.npz by running the UNMODIFIED reference (sangoi-exe/OneTrainer) on CPU.

Runs only in the build container (needs /root/reference); the fixtures it writes are committed and
are what travels to the GPU box.  Two shims make the reference's hot-path modules importable here
(SURVEY.md F7): a `sys.modules["accelerate"]` stub and an alias for the renamed
`Optimizer._cuda_graph_capture_health_check`; a `diffusers` stub satisfies one type-hint import.
No reference arithmetic is rewritten.

Where the reference draws random numbers, the draws are REPLACED by this repo's Philox stream
(torch.rand / torch.normal / torch.randint_like patched for the duration of one call), so the
fixtures pin the deterministic transforms after the draw — the part that has to match bit for bit.

    python tools/make_golden.py            # writes tests/golden/
"""
import sys
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import transformers  # noqa: E402,F401  (must be imported before the accelerate stub)
import torch  # noqa: E402

acc = types.ModuleType("accelerate")


class _Accelerator:
    def __init__(self, *a, **k):
        self.device = torch.device("cpu")


acc.Accelerator = _Accelerator
sys.modules["accelerate"] = acc
dif = types.ModuleType("diffusers")
dif.DDIMScheduler = object
sys.modules.setdefault("diffusers", dif)
sys.path.insert(0, "/root/reference")
if not hasattr(torch.optim.Optimizer, "_cuda_graph_capture_health_check"):
    torch.optim.Optimizer._cuda_graph_capture_health_check = torch.optim.Optimizer._accelerator_graph_capture_health_check

from modules.modelSetup.mixin.ModelSetupDiffusionLossMixin import ModelSetupDiffusionLossMixin  # noqa: E402
from modules.modelSetup.mixin.ModelSetupDiffusionMixin import ModelSetupDiffusionMixin  # noqa: E402
from modules.modelSetup.mixin.ModelSetupFlowMatchingMixin import ModelSetupFlowMatchingMixin  # noqa: E402
from modules.modelSetup.mixin.ModelSetupNoiseMixin import ModelSetupNoiseMixin  # noqa: E402
from modules.module.EMAModule import EMAModuleWrapper  # noqa: E402
from modules.module.LoRAModule import DoRAModule, LoRAModule  # noqa: E402
from modules.util.bf16_stochastic_rounding import add_stochastic_, addcdiv_stochastic_, copy_stochastic_  # noqa: E402,F401
from modules.util.config.TrainConfig import TrainConfig  # noqa: E402
from modules.util.convert.rescale_noise_scheduler_to_zero_terminal_snr import rescale_noise_scheduler_to_zero_terminal_snr  # noqa: E402
from modules.util.DiffusionScheduleCoefficients import DiffusionScheduleCoefficients  # noqa: E402
from modules.util.enum.LossScaler import LossScaler  # noqa: E402
from modules.util.enum.LossWeight import LossWeight  # noqa: E402
from modules.util.enum.TimestepDistribution import TimestepDistribution  # noqa: E402
from modules.util.optimizer.adam_extensions import patch_adam  # noqa: E402
from modules.util.optimizer.adamw_extensions import patch_adamw  # noqa: E402

from oracle import otb_oracle as O  # noqa: E402

OUT = REPO / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)
torch.set_num_threads(1)


def bits16(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def scaled_linear_betas():
    return torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2


# ------------------------------------------------------------------------------------ schedule
def gen_schedule():
    betas = scaled_linear_betas()
    co = DiffusionScheduleCoefficients.from_betas(betas)

    class Sched:
        pass

    s = Sched()
    s.alphas_cumprod = torch.cumprod(1 - betas, 0)
    zt_betas = rescale_noise_scheduler_to_zero_terminal_snr(s)
    np.savez(OUT / "schedule.npz", betas=betas.numpy(), alphas_cumprod=co.alphas_cumprod.numpy(),
             sqrt_ac=co.sqrt_alphas_cumprod.numpy(), sqrt_1mac=co.sqrt_one_minus_alphas_cumprod.numpy(),
             zt_betas=zt_betas.numpy(), zt_alphas_cumprod=s.alphas_cumprod.numpy())


# ------------------------------------------------------------------------------------ optimizer
def gen_optimizer():
    out = {}
    n = 4099                      # not a multiple of 8: exercises the SR tail lanes
    lr, betas, eps, wd = 3e-4, (0.9, 0.999), 1e-8, 1e-2
    for name, dtype, sr, patch, cls in (("adamw_f32", torch.float32, False, patch_adamw, torch.optim.AdamW),
                                        ("adamw_bf16", torch.bfloat16, False, patch_adamw, torch.optim.AdamW),
                                        ("adamw_bf16_sr", torch.bfloat16, True, patch_adamw, torch.optim.AdamW),
                                        ("adam_f32", torch.float32, False, patch_adam, torch.optim.Adam),
                                        ("adam_bf16_sr", torch.bfloat16, True, patch_adam, torch.optim.Adam)):
        g = torch.Generator().manual_seed(11)
        p = torch.nn.Parameter((torch.randn(n, generator=g) * 0.02).to(dtype))
        opt = cls([p], lr=lr, betas=betas, eps=eps, weight_decay=wd)
        patch(opt, sr)
        seed, rng_base = 0xC0FFEE, 24
        out[f"{name}_p0"] = bits16(p) if dtype == torch.bfloat16 else p.detach().numpy().copy()
        for step in range(1, 4):
            grad = (torch.randn(n, generator=g) * 1e-3).to(dtype)
            p.grad = grad.clone()
            out[f"{name}_g{step}"] = bits16(grad) if dtype == torch.bfloat16 else grad.numpy().copy()
            orig = torch.randint_like
            if sr:
                def fake_randint_like(src, dtype=None, low=0, high=0, _step=step):
                    assert low == 0 and high == (1 << 16) and dtype == torch.int32
                    return torch.from_numpy(O.sr_bits(seed, _step, rng_base, src.numel()).copy()).view(src.shape)
                torch.randint_like = fake_randint_like
            try:
                opt.step_parameter(p, opt.param_groups[0], 0)
            finally:
                torch.randint_like = orig
            st = opt.state[p]
            for key, t in (("p", p), ("m", st["exp_avg"]), ("v", st["exp_avg_sq"])):
                out[f"{name}_{key}{step}"] = bits16(t) if dtype == torch.bfloat16 else t.detach().numpy().copy()
        out[f"{name}_meta"] = np.array([lr, betas[0], betas[1], eps, wd, seed, rng_base], dtype=np.float64)
    # the KAT of SURVEY.md §8c: 1e5 copies of 1 + 2^-10 through the reference's own copy_stochastic_
    torch.manual_seed(0)
    src = torch.full((100000,), 1 + 2 ** -10, dtype=torch.float32)
    tgt = torch.empty(100000, dtype=torch.bfloat16)
    copy_stochastic_(tgt, src)
    out["sr_kat_mean"] = np.array([tgt.float().mean().item()])
    np.savez(OUT / "optimizer.npz", **out)


# ------------------------------------------------------------------------------------ the SR helpers themselves
def gen_sr_helpers():
    """copy_stochastic_ / add_stochastic_ / addcdiv_stochastic_ (modules/util/bf16_stochastic_rounding.py) called directly,
    the way CAME.py:167-176 and other optimizers of the reference call them; the 16-bit draws come from this repo's Philox
    stream (torch.randint_like patched for the call)."""
    out = {}
    n = 4099
    seed, rng_base = 0xBADC0DE, 16
    g = torch.Generator().manual_seed(23)

    def with_draws(offset, fn, r32_key=None):
        orig = torch.randint_like

        def fake_randint_like(src, dtype=None, low=0, high=0):
            assert low == 0 and high == (1 << 16) and dtype == torch.int32
            if r32_key is not None:
                out[r32_key] = src.detach().numpy().copy()     # the reference's fp32 result BEFORE the rounding (pins the FMA flavour)
            return torch.from_numpy(O.sr_bits(seed, offset, rng_base, src.numel()).copy()).view(src.shape)
        torch.randint_like = fake_randint_like
        try:
            fn()
        finally:
            torch.randint_like = orig

    # copy: ordinary values plus the corners of the integer trick (zeros, exact bf16 values, a mantissa of all ones that
    # carries into the exponent, denormals, large magnitudes, both signs)
    src = torch.randn(n, generator=g) * 0.02
    special = torch.tensor([0.0, -0.0, 1.0, -1.0, 1.0 + 2 ** -10, 1.0078125, 1e-30, -1e-30, 1e-40, -3.5e20, 3.3895e38, 65504.0], dtype=torch.float32)
    carry = torch.tensor([0x3FFFFFFF, 0xBFFFFFFF, 0x3F7FFFFF, 0x007FFFFF, 0x00000001], dtype=torch.int64).to(torch.int32).view(torch.float32)
    src[:special.numel()] = special
    src[special.numel():special.numel() + carry.numel()] = carry
    tgt = torch.empty(n, dtype=torch.bfloat16)
    with_draws(1, lambda: copy_stochastic_(tgt, src))
    out["copy_src"], out["copy_out"] = src.numpy().copy(), bits16(tgt)
    # add: the two calls of CAME (p += -update; p += alpha * p) and a generic alpha
    for name, other_dtype, alpha, offset, alias in (("add_update", torch.float32, 1.0, 2, False), ("add_decay", torch.bfloat16, -3e-6, 3, True),
                                                    ("add_half", torch.float32, 0.5, 4, False), ("add_bf16", torch.bfloat16, 1.0, 5, False),
                                                    ("add_third", torch.float32, 0.3, 8, False), ("add_third_bf16", torch.bfloat16, -0.7, 9, False)):
        inp = (torch.randn(n, generator=g) * 0.02).bfloat16()
        other = inp if alias else (torch.randn(n, generator=g) * (1e-2 if "third" in name else 1e-4)).to(other_dtype)
        out[f"{name}_in"] = bits16(inp)
        out[f"{name}_other"] = bits16(other) if other_dtype == torch.bfloat16 else other.numpy().copy()
        with_draws(offset, lambda: add_stochastic_(inp, other, alpha=alpha), f"{name}_r32")            # noqa: B023
        out[f"{name}_out"] = bits16(inp)
        out[f"{name}_meta"] = np.array([alpha, offset], dtype=np.float64)
    # addcdiv: AdamW's parameter update (adamw_extensions.py:139-144) and an fp32-operand variant
    for name, dt, value, offset in (("addcdiv_bf16", torch.bfloat16, -3e-4, 6), ("addcdiv_f32", torch.float32, -1.7e-3, 7)):
        inp = (torch.randn(n, generator=g) * 0.02).bfloat16()
        t1 = (torch.randn(n, generator=g) * 1e-3).to(dt)
        t2 = (torch.rand(n, generator=g) * 1e-3 + 1e-8).to(dt)
        out[f"{name}_in"] = bits16(inp)
        for key, t in (("t1", t1), ("t2", t2)):
            out[f"{name}_{key}"] = bits16(t) if dt == torch.bfloat16 else t.numpy().copy()
        with_draws(offset, lambda: addcdiv_stochastic_(inp, t1, t2, value=value), f"{name}_r32")     # noqa: B023
        out[f"{name}_out"] = bits16(inp)
        out[f"{name}_meta"] = np.array([value, offset], dtype=np.float64)
    out["meta"] = np.array([seed, rng_base], dtype=np.int64)
    np.savez(OUT / "sr_helpers.npz", **out)


# ------------------------------------------------------------------------------------ Adafactor (oracle only so far)
def gen_adafactor():
    """step_adafactor_parameter (adafactor_extensions.py:17-93) on transformers' Adafactor, three steps per case.
    Pins the ORACLE's restatement; the fused optimizer kernel covers AdamW/Adam only.
    Note the reference's bf16 tail (:88-91): copy_stochastic_ is followed by an unconditional p.copy_(fp32), so the
    stochastic rounding is overwritten by round-to-nearest — the fixtures record exactly that."""
    from transformers import Adafactor
    from modules.util.optimizer.adafactor_extensions import patch_adafactor
    out = {}
    cases = (("matrix_f32_relative", (24, 40), torch.float32, False, dict()),
             ("vector_f32_relative", (301,), torch.float32, False, dict()),
             ("matrix_f32_fixed_lr_beta1_wd", (24, 40), torch.float32, False,
              dict(lr=1e-3, relative_step=False, scale_parameter=False, beta1=0.9, weight_decay=0.01)),
             ("conv_f32_warmup", (8, 6, 3, 3), torch.float32, False, dict(warmup_init=True)),
             ("matrix_bf16_sr", (24, 40), torch.bfloat16, True, dict()))
    for name, shape, dtype, sr, kw in cases:
        g = torch.Generator().manual_seed(13)
        p = torch.nn.Parameter((torch.randn(*shape, generator=g) * 0.05).to(dtype))
        opt = Adafactor([p], **kw)
        patch_adafactor(opt, sr)
        enc = bits16 if dtype == torch.bfloat16 else (lambda t: t.detach().numpy().copy())
        out[f"{name}_p0"] = enc(p)
        for step in range(1, 4):
            grad = (torch.randn(*shape, generator=g) * 1e-2).to(dtype)
            p.grad = grad.clone()
            out[f"{name}_g{step}"] = enc(grad)
            opt.step_parameter(p, opt.param_groups[0], 0)
            out[f"{name}_p{step}"] = enc(p)
            st = opt.state[p]
            for key in ("exp_avg", "exp_avg_sq", "exp_avg_sq_row", "exp_avg_sq_col"):
                if key in st:
                    out[f"{name}_{key}{step}"] = st[key].detach().float().numpy().copy()
            out[f"{name}_rms{step}"] = np.array([float(st["RMS"])])
        grp = opt.param_groups[0]
        out[f"{name}_meta"] = np.array([grp["lr"] if grp["lr"] is not None else np.nan, grp["eps"][0], grp["eps"][1], grp["clip_threshold"],
                                        grp["decay_rate"], grp["beta1"] if grp["beta1"] is not None else np.nan, grp["weight_decay"],
                                        float(grp["scale_parameter"]), float(grp["relative_step"]), float(grp["warmup_init"])], dtype=np.float64)
    np.savez(OUT / "adafactor.npz", **out)


# ------------------------------------------------------------------------------------ CAME (oracle only so far)
def gen_came():
    """CAME.step_parameter (modules/util/optimizer/CAME.py:80-176), three steps per case; in the bf16 + SR case the
    reference's randint_like draws are replaced by this repo's Philox SR bits (one draw per add_stochastic_ call)."""
    from modules.util.optimizer.CAME import CAME
    out = {}
    seed, rng_base = 0xCA3E, 8
    cases = (("matrix_f32", (24, 40), torch.float32, False, dict(lr=2e-4)),
             ("vector_f32_wd", (301,), torch.float32, False, dict(lr=2e-4, weight_decay=0.01)),
             ("conv_f32", (8, 6, 3, 3), torch.float32, False, dict(lr=1e-4, betas=(0.8, 0.99, 0.999), clip_threshold=0.5)),
             ("matrix_bf16_sr_wd", (24, 40), torch.bfloat16, True, dict(lr=2e-4, weight_decay=0.01)))
    for name, shape, dtype, sr, kw in cases:
        g = torch.Generator().manual_seed(17)
        p = torch.nn.Parameter((torch.randn(*shape, generator=g) * 0.05).to(dtype))
        opt = CAME([p], stochastic_rounding=sr, **kw)
        enc = bits16 if dtype == torch.bfloat16 else (lambda t: t.detach().numpy().copy())
        out[f"{name}_p0"] = enc(p)
        calls = [0]
        for step in range(1, 4):
            grad = (torch.randn(*shape, generator=g) * 1e-2).to(dtype)
            p.grad = grad.clone()
            out[f"{name}_g{step}"] = enc(grad)
            orig = torch.randint_like

            def fake_randint_like(src, dtype=None, low=0, high=0):
                assert low == 0 and high == (1 << 16) and dtype == torch.int32
                calls[0] += 1
                return torch.from_numpy(O.sr_bits(seed, calls[0], rng_base, src.numel()).copy()).view(src.shape)
            if sr:
                torch.randint_like = fake_randint_like
            try:
                opt.step_parameter(p, opt.param_groups[0], 0)
            finally:
                torch.randint_like = orig
            out[f"{name}_p{step}"] = enc(p)
            st = opt.state[p]
            for key in ("exp_avg", "exp_avg_sq", "exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_res_row", "exp_avg_res_col"):
                if key in st:
                    out[f"{name}_{key}{step}"] = st[key].detach().float().numpy().copy()
        grp = opt.param_groups[0]
        out[f"{name}_meta"] = np.array([grp["lr"], grp["eps"][0], grp["eps"][1], grp["clip_threshold"], *grp["betas"], grp["weight_decay"],
                                        seed, rng_base], dtype=np.float64)
    np.savez(OUT / "came.npz", **out)


# ------------------------------------------------------------------------------------ EMA
def gen_ema():
    out = {}
    for name, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        g = torch.Generator().manual_seed(5)
        p = torch.nn.Parameter((torch.randn(1003, generator=g) * 0.05).to(dtype))
        ema = EMAModuleWrapper([p], decay=0.999, update_step_interval=2, device=torch.device("cpu"))
        enc = bits16 if dtype == torch.bfloat16 else (lambda t: t.detach().numpy().copy())
        out[f"{name}_ema0"] = enc(ema.ema_parameters[0])
        for step in range(6):
            with torch.no_grad():
                p.add_((torch.randn(1003, generator=g) * 0.01).to(dtype))
            out[f"{name}_p{step}"] = enc(p)
            ema.step([p], step)
            out[f"{name}_ema{step + 1}"] = enc(ema.ema_parameters[0])
            out[f"{name}_decay{step}"] = np.array([ema.get_current_decay(step)])
    np.savez(OUT / "ema.npz", **out)


# ------------------------------------------------------------------------------------ timesteps
class _Noise(ModelSetupNoiseMixin):
    pass


def gen_timesteps():
    out = {}
    B = 64
    cases = [
        ("uniform", TimestepDistribution.UNIFORM, dict()),
        ("uniform_range_shift3", TimestepDistribution.UNIFORM, dict(min_noising_strength=0.1, max_noising_strength=0.9, timestep_shift=3.0)),
        ("logit_normal", TimestepDistribution.LOGIT_NORMAL, dict()),
        ("logit_normal_biased", TimestepDistribution.LOGIT_NORMAL, dict(noising_weight=0.5, noising_bias=-0.3, timestep_shift=2.0)),
        ("logit_normal_dynamic", TimestepDistribution.LOGIT_NORMAL, dict(dynamic_timestep_shifting=True)),
        ("heavy_tail", TimestepDistribution.HEAVY_TAIL, dict(noising_weight=1.2)),
    ]
    for name, dist, kw in cases:
        cfg = TrainConfig.default_values()
        cfg.timestep_distribution = dist
        for k, v in kw.items():
            setattr(cfg, k, v)
        seed, offset = 77, 4
        # the draws this repo's kernel would make for global samples 0..B-1
        u = np.zeros(B, dtype=np.float32)
        for b in range(B):
            r = O.philox(seed, offset, 3, b)
            if dist == TimestepDistribution.LOGIT_NORMAL:
                u1 = (float(r[0] >> 8) + 0.5) * 2.0 ** -24
                th = (float(np.int32(r[1]) >> 8) + 0.5) * 2.0 ** -23 * np.pi
                z = np.float32(np.sqrt(-2.0 * np.log(u1)) * np.cos(th))
                scale = np.float32(cfg.noising_weight) + np.float32(1.0)
                u[b] = np.float32(z * scale) + np.float32(cfg.noising_bias)
            else:
                u[b] = np.float32(r[0] >> 8) * np.float32(2.0 ** -24)
        draws = torch.from_numpy(u.copy())
        orig_rand, orig_normal = torch.rand, torch.normal
        torch.rand = lambda *a, **k: draws.clone()
        torch.normal = lambda *a, **k: draws.clone()
        try:
            gen = torch.Generator().manual_seed(0)
            ts = _Noise()._get_timestep_discrete(1000, False, gen, B, cfg, latent_width=128, latent_height=128)
        finally:
            torch.rand, torch.normal = orig_rand, orig_normal
        out[f"{name}_draws"] = u
        out[f"{name}_t"] = ts.numpy().astype(np.int32)
        out[f"{name}_cfg"] = np.array([cfg.min_noising_strength, cfg.max_noising_strength, cfg.noising_weight, cfg.noising_bias,
                                       cfg.timestep_shift, float(cfg.dynamic_timestep_shifting), seed, offset], dtype=np.float64)
    # COS_MAP / SIGMOID: a multinomial over a cached weight table (ModelSetupNoiseMixin.py:122-153).  torch's own CPU
    # generator draws (the repo keeps this distribution on torch ops too), so the samples themselves are pinned
    for name, dist, kw in (("cos_map", TimestepDistribution.COS_MAP, dict()),
                           ("cos_map_range_shift3", TimestepDistribution.COS_MAP, dict(min_noising_strength=0.1, max_noising_strength=0.9, timestep_shift=3.0)),
                           ("sigmoid", TimestepDistribution.SIGMOID, dict(noising_weight=2.0, noising_bias=0.1, timestep_shift=2.0))):
        cfg = TrainConfig.default_values()
        cfg.timestep_distribution = dist
        for k, v in kw.items():
            setattr(cfg, k, v)
        mix = _Noise()
        ts = mix._get_timestep_discrete(1000, False, torch.Generator().manual_seed(5), B, cfg)
        out[f"{name}_t"] = ts.numpy().astype(np.int32)
        out[f"{name}_weights"] = mix._ModelSetupNoiseMixin__weights.numpy()
        out[f"{name}_cfg"] = np.array([cfg.min_noising_strength, cfg.max_noising_strength, cfg.noising_weight, cfg.noising_bias,
                                       cfg.timestep_shift, 0.0, 5, 0], dtype=np.float64)
    det = _Noise()._get_timestep_discrete(1000, True, torch.Generator(), 4, TrainConfig.default_values())
    out["deterministic_t"] = det.numpy()
    np.savez(OUT / "timesteps.npz", **out)


# ------------------------------------------------------------------------------------ add noise
class _Diff(ModelSetupDiffusionMixin):
    pass


class _Flow(ModelSetupFlowMatchingMixin):
    pass


def gen_addnoise():
    g = torch.Generator().manual_seed(3)
    B, Cc, H, W = 4, 4, 16, 16
    latent = torch.randn(B, Cc, H, W, generator=g) * 0.8
    noise = torch.randn(B, Cc, H, W, generator=g)
    t = torch.tensor([0, 499, 731, 999], dtype=torch.int32)
    betas = scaled_linear_betas()
    scale = 0.13025
    x0 = latent * scale
    noisy = _Diff()._add_noise_discrete(x0, noise, t, betas)
    ac = torch.cumprod(1 - betas, 0)
    a = (ac[t.long()] ** 0.5).view(-1, 1, 1, 1)
    b = ((1 - ac[t.long()]) ** 0.5).view(-1, 1, 1, 1)
    velocity = a * noise - b * x0                       # diffusers DDIMScheduler.get_velocity @5873377 (not vendored)
    # flow: FLUX scaling (latent - shift) * scale, BaseFluxSetup.py:230
    fshift, fscale = 0.1159, 0.3611
    fx0 = (latent - fshift) * fscale
    fnoisy, sigmas = _Flow()._add_noise_discrete(fx0, noise, t, torch.zeros(1000))
    np.savez(OUT / "addnoise.npz", latent=latent.numpy(), noise=noise.numpy(), t=t.numpy(), scale=np.array([scale]),
             x0=x0.numpy(), noisy=noisy.numpy(), velocity=velocity.numpy(), flow_shift_scale=np.array([fshift, fscale]),
             flow_x0=fx0.numpy(), flow_noisy=fnoisy.numpy(), flow_sigmas=sigmas.flatten().numpy(),
             flow_target=(noise - fx0).numpy())


# ------------------------------------------------------------------------------------ losses
class _Loss(ModelSetupDiffusionLossMixin):
    pass


def gen_losses():
    out = {}
    g = torch.Generator().manual_seed(9)
    B, Cc, H, W = 4, 4, 16, 16
    pred32 = torch.randn(B, Cc, H, W, generator=g)
    target = torch.randn(B, Cc, H, W, generator=g)
    prior = torch.randn(B, Cc, H, W, generator=g)
    mask = (torch.rand(B, 1, H, W, generator=g) > 0.5).float()
    mask[0, 0, :2] = 0.37                                # non-binary values go through the clamp too
    lw = torch.tensor([1.0, 0.5, 2.0, 1.25])
    t = torch.tensor([3, 499, 640, 999], dtype=torch.int32)
    betas = scaled_linear_betas()
    out.update(pred=pred32.numpy(), target=target.numpy(), prior=prior.numpy(), mask=mask.numpy(), loss_weight=lw.numpy(), t=t.numpy())
    cases = {
        "mse": dict(),
        "mse_minsnr_eps": dict(loss_weight_fn=LossWeight.MIN_SNR_GAMMA),
        "mse_minsnr_v": dict(loss_weight_fn=LossWeight.MIN_SNR_GAMMA, _v=True),
        "mse_debiased_eps": dict(loss_weight_fn=LossWeight.DEBIASED_ESTIMATION),
        "mse_debiased_v": dict(loss_weight_fn=LossWeight.DEBIASED_ESTIMATION, _v=True),
        "mse_p2": dict(loss_weight_fn=LossWeight.P2, loss_weight_strength=1.0),
        "mse_mae": dict(mae_strength=0.5, mse_strength=0.7),
        "masked": dict(masked_training=True),
        "masked_norm": dict(masked_training=True, normalize_masked_area_loss=True, unmasked_weight=0.2),
        "masked_norm_prior_mae": dict(masked_training=True, normalize_masked_area_loss=True, masked_prior_preservation_weight=0.4,
                                      mae_strength=0.3, _prior=True),
        "masked_prior": dict(masked_training=True, masked_prior_preservation_weight=0.25, _prior=True),
        "scaler_both": dict(loss_scaler=LossScaler.BOTH, batch_size=4, gradient_accumulation_steps=2),
        "flow_sigma": dict(loss_weight_fn=LossWeight.SIGMA, _flow=True),
        "flow_const": dict(_flow=True),
    }
    for name, kw in cases.items():
        for ptag, pdtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
            cfg = TrainConfig.default_values()
            flags = {k: v for k, v in kw.items() if k.startswith("_")}
            for k, v in kw.items():
                if not k.startswith("_"):
                    setattr(cfg, k, v)
            pred = pred32.to(pdtype).requires_grad_(True)
            data = {"loss_type": "target", "predicted": pred, "target": target.clone(), "timestep": t,
                    "prediction_type": "v_prediction" if flags.get("_v") else "epsilon"}
            if flags.get("_prior"):
                data["prior_target"] = prior
            batch = {"latent_mask": mask, "loss_weight": lw}
            mix = _Loss()
            if flags.get("_flow"):
                losses = mix._flow_matching_losses(batch, data, cfg, torch.device("cpu"), sigmas=torch.zeros(1000))
            else:
                losses = mix._diffusion_losses(batch, data, cfg, torch.device("cpu"), betas=betas)
            total = losses.mean() / cfg.gradient_accumulation_steps
            (grad,) = torch.autograd.grad(total, pred)
            out[f"{name}_{ptag}_losses"] = losses.detach().numpy()
            out[f"{name}_{ptag}_grad"] = bits16(grad) if pdtype == torch.bfloat16 else grad.numpy()
    np.savez(OUT / "losses.npz", **out)


# ------------------------------------------------------------------------------------ LoRA
def gen_lora():
    out = {}

    def run(cases):
        for name, base, xshape in cases:
            mod = LoRAModule("lora_unet.x", base, 16, 8.0)
            torch.nn.init.normal_(mod.lora_up.weight, std=0.05)
            base.requires_grad_(False)
            mod.hook_to_module()
            x = torch.randn(*xshape, requires_grad=True)
            y = base(x)
            dy = torch.randn_like(y)
            dx, dA, dB = torch.autograd.grad(y, [x, mod.lora_down.weight, mod.lora_up.weight], dy)
            out.update({f"{name}_x": x.detach().numpy(), f"{name}_w": base.weight.detach().numpy(),
                        f"{name}_bias": base.bias.detach().numpy() if base.bias is not None else np.zeros(0, np.float32),
                        f"{name}_A": mod.lora_down.weight.detach().numpy(), f"{name}_B": mod.lora_up.weight.detach().numpy(),
                        f"{name}_y": y.detach().numpy(), f"{name}_dy": dy.numpy(), f"{name}_dx": dx.numpy(),
                        f"{name}_dA": dA.numpy(), f"{name}_dB": dB.numpy(),
                        f"{name}_scale": np.array([float(mod.alpha) / mod.rank])})
            out[f"{name}_keys"] = np.array(sorted(mod.state_dict(prefix=mod.prefix).keys()))

    torch.manual_seed(21)
    run((("linear", torch.nn.Linear(72, 40), (3, 10, 72)),
         ("linear_nobias", torch.nn.Linear(64, 48, bias=False), (5, 64)),
         ("conv1x1", torch.nn.Conv2d(24, 32, 1), (2, 24, 6, 6))))
    # k x k leaves of the `full` preset (LoRAModule.py:142-155): own seed, so the fixtures above keep their draws
    torch.manual_seed(22)
    run((("conv3x3", torch.nn.Conv2d(24, 32, 3, padding=1), (2, 24, 6, 6)),
         ("conv3x3_s2", torch.nn.Conv2d(16, 40, 3, stride=2, padding=1, bias=False), (2, 16, 9, 8))))
    np.savez(OUT / "lora.npz", **out)


def gen_lora_bf16in():
    """LoRAModule.forward / backward (LoRAModule.py:319-323) in the reference's fp32 on operands that are EXACTLY
    representable in bf16 (x, W, bias, A, B, dY rounded once, stored as fp32).  This is the north_star's comparison value
    for the bf16 kernel — "the value computed from bf16-rounded inputs in fp32" — so the GPU test can hold K-a to 2 bf16
    ulp instead of budgeting for the input rounding."""
    out = {}
    r16 = lambda t: t.detach().bfloat16().float()  # noqa: E731
    torch.manual_seed(23)
    for name, base, xshape in (("linear", torch.nn.Linear(264, 136), (4, 50, 264)),
                               ("linear_nobias", torch.nn.Linear(1280, 320, bias=False), (77, 1280)),
                               ("conv1x1", torch.nn.Conv2d(64, 96, 1), (2, 64, 12, 12)),
                               ("conv3x3", torch.nn.Conv2d(24, 32, 3, padding=1), (2, 24, 10, 10))):
        with torch.no_grad():
            base.weight.copy_(r16(base.weight))
            if base.bias is not None:
                base.bias.copy_(r16(base.bias))
        mod = LoRAModule("lora_unet.x", base, 16, 8.0)
        with torch.no_grad():
            mod.lora_down.weight.copy_(r16(mod.lora_down.weight))
            mod.lora_up.weight.copy_(r16(torch.randn_like(mod.lora_up.weight) * 0.05))
        base.requires_grad_(False)
        mod.hook_to_module()
        x = r16(torch.randn(*xshape)).requires_grad_(True)
        y = base(x)
        dy = r16(torch.randn_like(y))
        dx, dA, dB = torch.autograd.grad(y, [x, mod.lora_down.weight, mod.lora_up.weight], dy)
        out.update({f"{name}_x": x.detach().numpy(), f"{name}_w": base.weight.detach().numpy(),
                    f"{name}_bias": base.bias.detach().numpy() if base.bias is not None else np.zeros(0, np.float32),
                    f"{name}_A": mod.lora_down.weight.detach().numpy(), f"{name}_B": mod.lora_up.weight.detach().numpy(),
                    f"{name}_y": y.detach().numpy(), f"{name}_dy": dy.numpy(), f"{name}_dx": dx.numpy(),
                    f"{name}_dA": dA.numpy(), f"{name}_dB": dB.numpy()})
    np.savez(OUT / "lora_bf16in.npz", **out)


# ------------------------------------------------------------------------------------ DoRA (oracle only so far)
def gen_dora():
    """DoRAModule.forward (LoRAModule.py:385-419) and its autograd backward, CPU fp32.  Pins the ORACLE's restatement;
    the fused path still raises for DoRA."""
    out = {}
    torch.manual_seed(41)
    for name, base, xshape, kw in (("linear", torch.nn.Linear(48, 40), (3, 7, 48), dict()),
                                   ("linear_outaxis_eps", torch.nn.Linear(32, 24, bias=False), (5, 32), dict(norm_epsilon=True, decompose_output_axis=True)),
                                   ("conv3x3", torch.nn.Conv2d(8, 16, 3, padding=1), (2, 8, 5, 5), dict())):
        mod = DoRAModule("lora_unet.x", base, 8, 4.0, train_device=torch.device("cpu"), **kw)
        torch.nn.init.normal_(mod.lora_up.weight, std=0.05)
        with torch.no_grad():
            mod.dora_scale.mul_(1.0 + 0.1 * torch.randn_like(mod.dora_scale))
        base.requires_grad_(False)
        mod.hook_to_module()
        x = torch.randn(*xshape, requires_grad=True)
        y = base(x)
        dy = torch.randn_like(y)
        dx, dA, dB, dm = torch.autograd.grad(y, [x, mod.lora_down.weight, mod.lora_up.weight, mod.dora_scale], dy)
        out.update({f"{name}_x": x.detach().numpy(), f"{name}_w": base.weight.detach().numpy(),
                    f"{name}_bias": base.bias.detach().numpy() if base.bias is not None else np.zeros(0, np.float32),
                    f"{name}_A": mod.lora_down.weight.detach().numpy(), f"{name}_B": mod.lora_up.weight.detach().numpy(),
                    f"{name}_m": mod.dora_scale.detach().numpy(),
                    f"{name}_y": y.detach().numpy(), f"{name}_dy": dy.numpy(), f"{name}_dx": dx.numpy(),
                    f"{name}_dA": dA.numpy(), f"{name}_dB": dB.numpy(), f"{name}_dm": dm.numpy(),
                    f"{name}_scale": np.array([float(mod.alpha) / mod.rank]),
                    f"{name}_flags": np.array([int(kw.get("norm_epsilon", False)), int(kw.get("decompose_output_axis", False))])})
        out[f"{name}_keys"] = np.array(sorted(mod.state_dict(prefix=mod.prefix).keys()))
    np.savez(OUT / "dora.npz", **out)


# ------------------------------------------------------------------------------------ parameter groups
def gen_param_groups():
    """NamedParameterGroupCollection.parameters_for_optimizer (NamedParameterGroup.py:36-61) for every scaler."""
    from modules.util.enum.LearningRateScaler import LearningRateScaler
    from modules.util.NamedParameterGroup import NamedParameterGroup, NamedParameterGroupCollection
    out = {}
    for scaler in LearningRateScaler:
        cfg = TrainConfig.default_values()
        cfg.learning_rate, cfg.learning_rate_scaler, cfg.batch_size, cfg.gradient_accumulation_steps = 3e-4, scaler, 4, 2
        col = NamedParameterGroupCollection()
        col.add_group(NamedParameterGroup("unet_lora", [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2))], None, "UNet LoRA"))
        col.add_group(NamedParameterGroup("te_lora", [torch.nn.Parameter(torch.zeros(5))], 1e-5))
        groups = col.parameters_for_optimizer(cfg)
        out[f"{scaler}_lr"] = np.array([g["lr"] for g in groups], dtype=np.float64)
        out[f"{scaler}_initial_lr"] = np.array([g["initial_lr"] for g in groups], dtype=np.float64)
        out[f"{scaler}_counts"] = np.array([len(g["params"]) for g in groups])
        out[f"{scaler}_keys"] = np.array(sorted(groups[0].keys()))
        out["unique_names"], out["display_names"] = np.array(col.unique_name_mapping), np.array(col.display_name_mapping)
        out["n_parameters"] = np.array([len(col.parameters())])
    np.savez(OUT / "param_groups.npz", **out)


# ------------------------------------------------------------------------------------ clip
def gen_clip():
    g = torch.Generator().manual_seed(31)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (1000, 37, 4096)]
    out = {}
    for tag, scale in (("big", 1.0), ("small", 1e-3)):
        for i, p in enumerate(params):
            p.grad = torch.randn(p.numel(), generator=g) * scale
            out[f"{tag}_g{i}"] = p.grad.numpy().copy()
        norm = torch.nn.utils.clip_grad_norm_(params, 1.0)
        out[f"{tag}_norm"] = np.array([norm.item()], dtype=np.float32)
        for i, p in enumerate(params):
            out[f"{tag}_clipped{i}"] = p.grad.numpy().copy()
    np.savez(OUT / "clip.npz", **out)


if __name__ == "__main__":
    only = set(sys.argv[1:])              # e.g. `python tools/make_golden.py gen_lora` rewrites one fixture file
    for fn in (gen_schedule, gen_optimizer, gen_sr_helpers, gen_adafactor, gen_came, gen_ema, gen_timesteps, gen_addnoise, gen_losses, gen_lora, gen_lora_bf16in, gen_dora, gen_param_groups, gen_clip):
        if only and fn.__name__ not in only:
            continue
        fn()
        print("wrote", fn.__name__)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
