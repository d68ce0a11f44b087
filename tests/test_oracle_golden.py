"""The oracle (oracle/otb_oracle.c) against the reference's own outputs (tests/golden/, recorded by
tools/make_golden.py from the unmodified reference on CPU).  Integer / bit-level work must be
bit-exact; floating-point reductions carry their tolerance in the test."""
import numpy as np
import pytest


def _meta(g, name):
    lr, b1, b2, eps, wd, seed, rng_base = g[f"{name}_meta"]
    return lr, b1, b2, eps, wd, int(seed), int(rng_base)


@pytest.mark.parametrize("name,dt,sr,decoupled", [
    ("adamw_f32", 0, False, True), ("adamw_bf16", 1, False, True), ("adamw_bf16_sr", 1, True, True),
    ("adam_f32", 0, False, False), ("adam_bf16_sr", 1, True, False)])
def test_adam_step_vs_reference(oracle, golden, name, dt, sr, decoupled):
    """step_adamw_parameter / step_adam_parameter (adamw_extensions.py:17-150) over 3 steps, CPU flavour.

    exp_avg / exp_avg_sq: BIT-EXACT in every case.  bf16 p without stochastic rounding: BIT-EXACT.
    fp32 p, and bf16 p WITH stochastic rounding (whose outcome depends on all 32 bits of the fp32 update):
    bit-exact except where torch's CPU `sqrt` is not correctly rounded — its AVX512 kernel is off by one ulp
    on ~0.5 % of inputs (measured against fp64; torch CUDA and this oracle use IEEE sqrt), which moves that
    element's denominator.  Bound: <= 0.5 % of elements, <= 1 ulp of the parameter dtype, per step."""
    g = golden("optimizer")
    lr, b1, b2, eps, wd, seed, rng_base = _meta(g, name)
    p = g[f"{name}_p0"].copy()
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for step in (1, 2, 3):
        sc = oracle.adam_scalars(lr, b1, b2, eps, wd, float(step), decoupled)
        oracle.adam_step(p, g[f"{name}_g{step}"].copy(), m, v, dt, sc, flavor=oracle.FLAVOR_TORCH_CPU, sr=sr,
                         seed=seed, offset=step, rng_base=rng_base)
        for key, arr in (("p", p), ("m", m), ("v", v)):
            ref = g[f"{name}_{key}{step}"]
            ai = arr.view(np.int32 if dt == 0 else np.int16).astype(np.int64)
            ri = ref.view(np.int32 if dt == 0 else np.int16).astype(np.int64)
            exact = not (key == "p" and (dt == 0 or sr)) and not (name.startswith("adam_") and step > 1)
            if exact:
                assert np.array_equal(ai, ri), f"{name} step {step} {key}: {(ai != ri).sum()} of {ai.size} elements differ"
            else:
                ulp = np.abs(ai - ri)
                assert ulp.max() <= 1 and (ulp != 0).mean() <= 5e-3, f"{name} step {step} {key}: {(ulp != 0).sum()} differ, max {ulp.max()} ulp"
                arr[:] = ref          # resync so a one-ulp difference is not counted again next step


def test_sr_is_unbiased_like_reference_kat(oracle, golden):
    """1e5 copies of 1+2^-10 through copy_stochastic_ -> mean ~1.000977 (neighbours 1, 1.0078125)."""
    ref_mean = float(golden("optimizer")["sr_kat_mean"][0])
    n = 100000
    src = np.full(n, 1 + 2.0 ** -10, dtype=np.float32)
    bits = src.view(np.uint32) + oracle.sr_bits(123, 1, 0, n).astype(np.uint32)
    out = (bits & 0xFFFF0000).view(np.float32)
    assert set(np.unique(out)) <= {np.float32(1.0), np.float32(1.0078125)}
    # binomial std of the mean is 2^-7 * sqrt(p(1-p)/n) = 8.2e-6; tolerance 5 sigma, same for the reference's draw
    assert abs(out.mean() - (1 + 2.0 ** -10)) < 4.1e-5
    assert abs(ref_mean - (1 + 2.0 ** -10)) < 4.1e-5


SR_HELPER_CASES = {"copy": ("copy", 1), "add_update": ("add", None), "add_decay": ("add", None), "add_half": ("add", None),
                   "add_bf16": ("add", None), "add_third": ("add", None), "add_third_bf16": ("add", None),
                   "addcdiv_bf16": ("addcdiv", None), "addcdiv_f32": ("addcdiv", None)}


def run_sr_helper_case(oracle, g, name, flavor):
    """One case of tests/golden/sr_helpers.npz through the oracle: (bf16 bits, fp32 values before the rounding)."""
    seed, base = (int(x) for x in g["meta"])
    kind = SR_HELPER_CASES[name][0]
    if kind == "copy":
        n = g["copy_src"].size
        return oracle.sr_helper(oracle.SR_COPY, np.zeros(n, np.uint16), g["copy_src"], bits=oracle.sr_bits(seed, 1, base, n), want_r32=True)
    alpha, off = g[f"{name}_meta"]
    n = g[f"{name}_in"].size
    draws = oracle.sr_bits(seed, int(off), base, n)
    if kind == "add":
        return oracle.sr_helper(oracle.SR_ADD, g[f"{name}_in"], g[f"{name}_other"], alpha=float(alpha), bits=draws, flavor=flavor, want_r32=True)
    return oracle.sr_helper(oracle.SR_ADDCDIV, g[f"{name}_in"], g[f"{name}_t1"], g[f"{name}_t2"], alpha=float(alpha), bits=draws,
                            flavor=flavor, want_r32=True)


@pytest.mark.parametrize("name", list(SR_HELPER_CASES))
def test_sr_helpers_bit_exact_vs_reference(oracle, golden, name):
    """copy_stochastic_ / add_stochastic_ / addcdiv_stochastic_ (modules/util/bf16_stochastic_rounding.py) called directly:
    the oracle fed the reference run's own 16-bit draws reproduces the bf16 result AND the fp32 value the reference formed
    before rounding (captured at its randint_like call) bit for bit — including the corners of the integer trick (zeros,
    exact bf16 values, mantissas that carry into the exponent, denormals) and `other` aliasing `input`."""
    g = golden("sr_helpers")
    out, r32 = run_sr_helper_case(oracle, g, name, oracle.FLAVOR_TORCH_CPU)
    assert np.array_equal(out, g[f"{name}_out"])
    if name != "copy":
        assert np.array_equal(r32.view(np.int32), g[f"{name}_r32"].view(np.int32))


def test_sr_helper_fixture_pins_the_rounding_flavour(oracle, golden):
    """The fixture discriminates the arithmetic flavours: torch's CPU add_ with alpha is ONE fma, its addcdiv_ is
    p + (value * t1) / t2 — the other orders differ from the recorded fp32 values on hundreds of elements."""
    g = golden("sr_helpers")
    no_fma = oracle.FLAVOR_TORCH_CPU & ~oracle.FL_FMA_ADD_ALPHA
    for name in ("add_third", "add_third_bf16"):
        _, r32 = run_sr_helper_case(oracle, g, name, no_fma)
        assert (r32.view(np.int32) != g[f"{name}_r32"].view(np.int32)).sum() > 100
    for name in ("addcdiv_bf16", "addcdiv_f32"):
        _, r32 = run_sr_helper_case(oracle, g, name, oracle.FLAVOR_TORCH_CUDA)
        assert (r32.view(np.int32) != g[f"{name}_r32"].view(np.int32)).sum() > 100


@pytest.mark.parametrize("tag,dt", [("f32", 0), ("bf16", 1)])
def test_ema_bit_exact_vs_reference(oracle, golden, tag, dt):
    """EMAModuleWrapper.step (EMAModule.py:37-54): decay warm-up, update interval 2, three roundings."""
    g = golden("ema")
    ema = g[f"{tag}_ema0"].copy()
    for step in range(6):
        decay = min((1 + step) / (10 + step), 0.999)
        assert decay == float(g[f"{tag}_decay{step}"][0])
        if (step + 1) % 2 == 0:
            oracle.ema_step(ema, g[f"{tag}_p{step}"].copy(), dt, float(np.float32(1 - decay)))
        ref = g[f"{tag}_ema{step + 1}"]
        assert np.array_equal(ema.view(np.uint32 if dt == 0 else np.uint16), ref.view(np.uint32 if dt == 0 else np.uint16))


MODES = {"uniform": 0, "uniform_range_shift3": 0, "logit_normal": 1, "logit_normal_biased": 1, "logit_normal_dynamic": 1, "heavy_tail": 2}


@pytest.mark.parametrize("name", list(MODES))
def test_timesteps_vs_reference(oracle, golden, name):
    """_get_timestep_discrete (ModelSetupNoiseMixin.py:51-155) fed this repo's draws.  UNIFORM is pure
    IEEE arithmetic: exact.  LOGIT_NORMAL / HEAVY_TAIL pass through torch's sigmoid / cos, which are not
    correctly rounded (1 ulp): the integer may differ by one only where the pre-truncation value sits on
    an integer boundary — allowed for at most 1 sample in 64 and never by more than 1."""
    import math
    g = golden("timesteps")
    mn, mx, w, b, shift, dyn, seed, offset = g[f"{name}_cfg"]
    if dyn:
        m_ = (1.15 - 0.5) / (4096 - 256)
        shift = math.exp((128 // 2) * (128 // 2) * m_ + (0.5 - m_ * 256))
    ref = g[f"{name}_t"]
    mine = np.array([oracle.timestep(MODES[name], int(seed), int(offset), i, 1000, int(1000 * mn), int(1000 * mx), w, b, shift)
                     for i in range(ref.size)])
    from_draw = np.array([oracle.timestep_from_draw(MODES[name], float(d), 1000, int(1000 * mn), int(1000 * mx), w, b, shift)
                          for d in g[f"{name}_draws"]])
    assert np.array_equal(mine, from_draw)          # the Philox -> draw construction matches the fixture's draws
    diff = np.abs(mine - ref)
    if MODES[name] == 0:
        assert np.array_equal(mine, ref)
    else:
        assert diff.max() <= 1 and (diff != 0).sum() <= 1, f"{(diff != 0).sum()} timesteps differ, max {diff.max()}"
    assert ref.min() >= 0 and ref.max() < 1000


@pytest.mark.parametrize("name,dist,kw", [
    ("cos_map", "COS_MAP", dict()),
    ("cos_map_range_shift3", "COS_MAP", dict(min_noising_strength=0.1, max_noising_strength=0.9, timestep_shift=3.0)),
    ("sigmoid", "SIGMOID", dict(noising_weight=2.0, noising_bias=0.1, timestep_shift=2.0))])
def test_table_timesteps_vs_reference_weights(oracle, golden, name, dist, kw):
    """COS_MAP / SIGMOID (ModelSetupNoiseMixin.py:122-153): the reference draws torch.multinomial over a cached weight
    table.  (1) the host builds the SAME table, bit for bit; (2) the oracle's inverse-CDF draw (what the kernel does, per
    global sample) is distributed as weights / sum(weights); (3) the reference's own 64 recorded draws are consistent
    with that distribution (two-sample KS)."""
    from onetrainer_b200.modules.modelSetup.mixin.ModelSetupNoiseMixin import timestep_cdf, timestep_weights
    from onetrainer_b200.modules.util.config.TrainConfig import TrainConfig
    from onetrainer_b200.modules.util.enum import TimestepDistribution
    g = golden("timesteps")
    cfg = TrainConfig(timestep_distribution=getattr(TimestepDistribution, dist), **kw)
    w = timestep_weights(1000, cfg).numpy()
    assert np.array_equal(w.view(np.uint32), g[f"{name}_weights"].view(np.uint32))
    mn, mx = int(1000 * cfg.min_noising_strength), int(1000 * cfg.max_noising_strength)
    cdf = oracle.timestep_cdf(w)
    assert np.array_equal(cdf, timestep_cdf(__import__("torch").from_numpy(w)).numpy()) and cdf.size == mx - mn
    n = 20000
    draws = np.array([oracle.timestep_table(11, 3, i, cdf, mn) for i in range(n)])
    assert draws.min() >= mn and draws.max() < mx
    emp = np.cumsum(np.bincount(draws - mn, minlength=mx - mn)) / n
    assert np.abs(emp - cdf).max() < 1.63 / np.sqrt(n)            # KS, alpha = 1 %
    ref = g[f"{name}_t"]
    assert ref.min() >= mn and ref.max() < mx
    emp_ref = np.cumsum(np.bincount(ref - mn, minlength=mx - mn)) / ref.size
    assert np.abs(emp_ref - cdf).max() < 1.95 / np.sqrt(ref.size)  # the reference's own multinomial draws, alpha = 0.1 %
    # keyed by the GLOBAL sample index: a rank that owns samples [8, 16) draws what a single process draws for them
    assert [oracle.timestep_table(11, 3, 8 + i, cdf, mn) for i in range(8)] == draws[8:16].tolist()


def test_deterministic_timestep(oracle, golden):
    assert oracle.timestep(3, 0, 0, 0) == 499 == int(golden("timesteps")["deterministic_t"][0])


def test_add_noise_bit_exact_vs_reference(oracle, golden):
    """_add_noise_discrete (DDPM, ModelSetupDiffusionMixin.py:15-38), latent scaling and the v target."""
    g, s = golden("addnoise"), golden("schedule")
    x0, noisy, v = oracle.add_noise(g["latent"], g["noise"], g["t"], s["sqrt_ac"], s["sqrt_1mac"], latent_scale=float(np.float32(g["scale"][0])),
                                    target=oracle.TARGET_V)
    assert np.array_equal(x0, g["x0"])
    assert np.array_equal(noisy, g["noisy"])
    assert np.array_equal(v, g["velocity"])
    _, _, eps = oracle.add_noise(g["latent"], g["noise"], g["t"], s["sqrt_ac"], s["sqrt_1mac"], latent_scale=float(np.float32(g["scale"][0])),
                                 target=oracle.TARGET_EPSILON)
    assert np.array_equal(eps, g["noise"])


def test_flow_add_noise_bit_exact_vs_reference(oracle, golden):
    """Rectified flow (ModelSetupFlowMatchingMixin.py:14-39) with FLUX latent shift/scale and the eps-x0 target."""
    g = golden("addnoise")
    T = 1000
    sigma = (np.arange(1, T + 1, dtype=np.int32) / np.float32(T)).astype(np.float32)    # int32 tensor / python int -> fp32 division
    one_minus = (np.float32(1.0) - sigma).astype(np.float32)
    shift, scale = g["flow_shift_scale"]
    x0, noisy, tgt = oracle.add_noise(g["latent"], g["noise"], g["t"], one_minus, sigma, latent_scale=float(np.float32(scale)),
                                      latent_shift=float(np.float32(shift)), target=oracle.TARGET_FLOW)
    assert np.array_equal(sigma[g["t"]], g["flow_sigmas"])
    assert np.array_equal(x0, g["flow_x0"])
    assert np.array_equal(noisy, g["flow_noisy"])
    assert np.array_equal(tgt, g["flow_target"])


LOSS_CASES = {
    "mse": dict(),
    "mse_minsnr_eps": dict(lw_fn=1),
    "mse_minsnr_v": dict(lw_fn=1, v_pred=True),
    "mse_debiased_eps": dict(lw_fn=2),
    "mse_debiased_v": dict(lw_fn=2, v_pred=True),
    "mse_p2": dict(lw_fn=3, gamma=1.0),
    "mse_mae": dict(mae=0.5, mse=0.7),
    "masked": dict(masked=True),
    "masked_norm": dict(masked=True, normalize=True, unmasked_weight=0.2),
    "masked_norm_prior_mae": dict(masked=True, normalize=True, prior_w=0.4, mae=0.3, use_prior=True),
    "masked_prior": dict(masked=True, prior_w=0.25, use_prior=True),
    "scaler_both": dict(loss_scale=8.0, ga=2),
    "flow_sigma": dict(lw_fn=4),
    "flow_const": dict(),
}


@pytest.mark.parametrize("name", list(LOSS_CASES))
def test_losses_vs_reference(oracle, golden, name):
    """_diffusion_losses / _flow_matching_losses (ModelSetupDiffusionLossMixin.py:233-321) and their autograd
    gradient.  Tolerance (north_star): 1e-3 relative; the oracle sums in fp64 so it lands within 2e-6."""
    g, s = golden("losses"), golden("schedule")
    kw = dict(LOSS_CASES[name])
    use_prior = kw.pop("use_prior", False)
    ga = kw.pop("ga", 1)
    B = g["pred"].shape[0]
    sigmas = (np.arange(1, 1001, dtype=np.int32) / np.float32(1000)).astype(np.float32)
    losses, grad = oracle.losses(
        g["pred"], g["target"], prior=g["prior"] if use_prior else None, mask=g["mask"] if kw.get("masked") else None,
        loss_weight=g["loss_weight"], timestep_=g["t"], sqrt_ac=s["sqrt_ac"], sqrt_1mac=s["sqrt_1mac"], sigmas=sigmas, C_=4,
        upstream=1.0 / (B * ga), **kw)
    np.testing.assert_allclose(losses, g[f"{name}_f32_losses"], rtol=2e-6)
    ref_grad = g[f"{name}_f32_grad"]
    np.testing.assert_allclose(grad, ref_grad, rtol=2e-5, atol=1e-6 * np.abs(ref_grad).max())


CONV_GEOMETRY = {"conv3x3": dict(kernel=(3, 3), stride=(1, 1), padding=(1, 1)), "conv3x3_s2": dict(kernel=(3, 3), stride=(2, 2), padding=(1, 1))}


@pytest.mark.parametrize("name", ["linear", "linear_nobias", "conv1x1", "conv3x3", "conv3x3_s2"])
def test_lora_linear_vs_reference(oracle, golden, name):
    """LoRAModule.forward (LoRAModule.py:319-323) and its autograd backward on CPU fp32 (config C1's arithmetic).
    fp32 tolerance (north_star): 1e-3 relative; fp64-accumulated oracle lands within 1e-5."""
    g = golden("lora")
    x, w, A, B, dy = g[f"{name}_x"], g[f"{name}_w"], g[f"{name}_A"], g[f"{name}_B"], g[f"{name}_dy"]
    bias = g[f"{name}_bias"]
    s = float(g[f"{name}_scale"][0])
    to2d = lambda t: np.ascontiguousarray(t.transpose(0, 2, 3, 1).reshape(-1, t.shape[1]))  # noqa: E731
    geo = CONV_GEOMETRY.get(name)
    if geo is not None:        # k x k conv == Linear over the unfolded patches (create_layer, LoRAModule.py:142-155)
        x2, (ho, wo) = oracle.conv_rows(x, **geo)
        assert (ho, wo) == dy.shape[2:]
        dy2, y_ref, dx_ref = to2d(dy), to2d(g[f"{name}_y"]), None
    elif name == "conv1x1":    # NCHW 1x1 conv == Linear over pixels
        x2, dy2, y_ref, dx_ref = to2d(x), to2d(dy), to2d(g[f"{name}_y"]), to2d(g[f"{name}_dx"])
    else:
        x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, dy.shape[-1])
        y_ref, dx_ref = g[f"{name}_y"].reshape(dy2.shape), g[f"{name}_dx"].reshape(x2.shape)
    w2, A2, B2 = w.reshape(w.shape[0], -1), A.reshape(A.shape[0], -1), B.reshape(B.shape[0], -1)
    y, _ = oracle.lora_linear(x2, w2, A2, B2, bias if bias.size else None, s, round_t_bf16=False)
    np.testing.assert_allclose(y, y_ref, rtol=1e-5, atol=1e-5)
    dx, dA, dB = oracle.lora_backward(x2, dy2, w2, A2, B2, s, round_bf16=False)
    if geo is not None:
        dx, dx_ref = oracle.conv_rows_adjoint(dx, x.shape, **geo), g[f"{name}_dx"]
    np.testing.assert_allclose(dx, dx_ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dA, g[f"{name}_dA"].reshape(A2.shape), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dB, g[f"{name}_dB"].reshape(B2.shape), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["linear", "linear_outaxis_eps", "conv3x3"])
def test_dora_vs_reference(oracle, golden, name):
    """DoRAModule.forward (LoRAModule.py:385-419) + autograd backward, CPU fp32.  Oracle only: the fused path raises for
    DoRA until its kernel exists; this pins the restatement that kernel will be checked against."""
    g = golden("dora")
    x, dy = g[f"{name}_x"], g[f"{name}_dy"]
    eps_flag, out_axis = (bool(v) for v in g[f"{name}_flags"])
    to2d = lambda t: np.ascontiguousarray(t.transpose(0, 2, 3, 1).reshape(-1, t.shape[1]))  # noqa: E731
    geo = dict(kernel=(3, 3), stride=(1, 1), padding=(1, 1)) if name == "conv3x3" else None
    if geo is not None:
        x2, _ = oracle.conv_rows(x, **geo)
        dy2, y_ref = to2d(dy), to2d(g[f"{name}_y"])
    else:
        x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, dy.shape[-1])
        y_ref = g[f"{name}_y"].reshape(dy2.shape)
    y, dx, dA, dB, dm = oracle.dora(x2, dy2, g[f"{name}_w"], g[f"{name}_A"], g[f"{name}_B"], g[f"{name}_m"], g[f"{name}_bias"],
                                    float(g[f"{name}_scale"][0]), norm_epsilon=eps_flag, decompose_output_axis=out_axis)
    dx = oracle.conv_rows_adjoint(dx, x.shape, **geo) if geo is not None else dx.reshape(x.shape)
    np.testing.assert_allclose(y, y_ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dx, g[f"{name}_dx"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dA, g[f"{name}_dA"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dB, g[f"{name}_dB"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dm, g[f"{name}_dm"], rtol=1e-4, atol=1e-5)
    assert list(g[f"{name}_keys"]) == ["lora_unet.x.alpha", "lora_unet.x.dora_scale", "lora_unet.x.lora_down.weight", "lora_unet.x.lora_up.weight"]


@pytest.mark.parametrize("name", ["matrix_f32_relative", "vector_f32_relative", "matrix_f32_fixed_lr_beta1_wd", "conv_f32_warmup", "matrix_bf16_sr"])
def test_adafactor_vs_reference(oracle, golden, name):
    """step_adafactor_parameter (adafactor_extensions.py:17-93), three steps.  Oracle only: the fused optimizer kernel
    covers AdamW / Adam.  fp32 elementwise + fp64 reductions against torch's fp32 reductions:
    1e-5 relative on the parameter and the factored states; the bf16 case may differ by one bf16 ulp where the fp32
    value sits on a rounding boundary."""
    g = golden("adafactor")
    lr, eps0, eps1, clip, decay, beta1, wd, scale_p, rel, warm = g[f"{name}_meta"]
    kw = dict(lr=None if np.isnan(lr) else float(lr), eps=(float(eps0), float(eps1)), clip_threshold=float(clip), decay_rate=float(decay),
              beta1=None if np.isnan(beta1) else float(beta1), weight_decay=float(wd), scale_parameter=bool(scale_p),
              relative_step=bool(rel), warmup_init=bool(warm))
    bf16 = name.endswith("bf16_sr")
    dec = (lambda a: oracle.bf16_bits_to_f32(a)) if bf16 else (lambda a: a)
    p, state = dec(g[f"{name}_p0"]), {}
    for step in (1, 2, 3):
        p = oracle.adafactor_step(p, dec(g[f"{name}_g{step}"]), state, bf16=bf16, **kw)
        ref = dec(g[f"{name}_p{step}"])
        if bf16:
            ulp = np.abs(ref) * 2.0 ** -7
            assert (np.abs(p - ref) <= ulp).all() and (p != ref).mean() < 0.02
            p = ref                                    # continue from the reference's values: boundary cases do not compound
        else:
            np.testing.assert_allclose(p, ref, rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(float(state["RMS"]), float(g[f"{name}_rms{step}"][0]), rtol=1e-6)
        for key in ("exp_avg", "exp_avg_sq", "exp_avg_sq_row", "exp_avg_sq_col"):
            if f"{name}_{key}{step}" in g:
                ref_state = g[f"{name}_{key}{step}"]        # near-cancelling first-moment entries: tolerance on the tensor's scale
                np.testing.assert_allclose(state[key], ref_state, rtol=2e-5, atol=1e-6 * np.abs(ref_state).max())


@pytest.mark.parametrize("name", ["matrix_f32", "vector_f32_wd", "conv_f32", "matrix_bf16_sr_wd"])
def test_came_vs_reference(oracle, golden, name):
    """CAME.step_parameter (modules/util/optimizer/CAME.py:80-176), three steps; the bf16 case with stochastic rounding
    fed the same 16-bit draws as the reference run (one Philox call per add_stochastic_).  Oracle only (SURVEY.md §8f
    rank 4).  fp32: 1e-5 relative; bf16 + SR: identical except where the fp32 sum lands within an ulp of a draw's
    carry boundary (< 2 % of elements, one bf16 ulp)."""
    g = golden("came")
    lr, eps0, eps1, clip, b0, b1, b2, wd, seed, rng_base = g[f"{name}_meta"]
    kw = dict(lr=float(lr), eps=(float(eps0), float(eps1)), clip_threshold=float(clip), betas=(float(b0), float(b1), float(b2)), weight_decay=float(wd))
    bf16 = "bf16" in name
    dec = (lambda a: oracle.bf16_bits_to_f32(a)) if bf16 else (lambda a: a)
    p, state, calls = dec(g[f"{name}_p0"]), {}, [0]

    def draws(n):
        calls[0] += 1
        return oracle.sr_bits(int(seed), calls[0], int(rng_base), n)

    for step in (1, 2, 3):
        p = oracle.came_step(p, dec(g[f"{name}_g{step}"]), state, sr_draws=draws if bf16 else None, **kw)
        ref = dec(g[f"{name}_p{step}"])
        if bf16:
            assert (np.abs(p - ref) <= np.abs(ref) * 2.0 ** -7).all() and (p != ref).mean() < 0.02
            p = ref
        else:
            np.testing.assert_allclose(p, ref, rtol=1e-5, atol=1e-8)
        for key in ("exp_avg", "exp_avg_sq", "exp_avg_sq_row", "exp_avg_sq_col", "exp_avg_res_row", "exp_avg_res_col"):
            if f"{name}_{key}{step}" in g:
                ref_state = g[f"{name}_{key}{step}"]
                np.testing.assert_allclose(state[key], ref_state, rtol=5e-5, atol=1e-6 * np.abs(ref_state).max())
    assert calls[0] == (6 if bf16 else 0)


@pytest.mark.parametrize("tag", ["big", "small"])
def test_clip_grad_norm_vs_torch(oracle, golden, tag):
    """nn.utils.clip_grad_norm_(params, 1.0) as called at GenericTrainer.py:712-713."""
    g = golden("clip")
    grads = [g[f"{tag}_g{i}"] for i in range(3)]
    norm = oracle.grad_norm(grads)
    np.testing.assert_allclose(norm, float(g[f"{tag}_norm"][0]), rtol=1e-6)
    coef = oracle.clip_coef(float(g[f"{tag}_norm"][0]), 1.0)
    for i in range(3):
        np.testing.assert_allclose(grads[i] * np.float32(coef), g[f"{tag}_clipped{i}"], rtol=2e-7)


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for philox4x32-10."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, out in kat:
        assert list(oracle.philox_raw(ctr, key)) == out
