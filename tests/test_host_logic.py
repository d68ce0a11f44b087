"""Host-side mirror of the reference seams: API shape, state-dict keys, error behaviour, scalar derivation.
Runs on CPU (no kernel is launched)."""
import math

import numpy as np
import pytest
import torch
from torch import nn

from onetrainer_b200.modules.module.EMAModule import EMAModuleWrapper
from onetrainer_b200.modules.module.LoRAModule import DummyLoRAModule, LoRAModule, LoRAModuleWrapper
from onetrainer_b200.modules.util.config.TrainConfig import TrainConfig
from onetrainer_b200.modules.util.convert.rescale_noise_scheduler_to_zero_terminal_snr import rescale_noise_scheduler_to_zero_terminal_snr
from onetrainer_b200.modules.util.DiffusionScheduleCoefficients import DiffusionScheduleCoefficients, flow_sigmas, scaled_linear_betas
from onetrainer_b200.modules.util.enum import LossWeight, PeftType, TimestepDistribution, enum_name


def test_schedule_tables_match_reference_bit_for_bit(golden):
    """DiffusionScheduleCoefficients.from_betas (reference :36-60) and the diffusers scaled-linear betas."""
    g = golden("schedule")
    betas = scaled_linear_betas()
    assert np.array_equal(betas.numpy(), g["betas"])
    co = DiffusionScheduleCoefficients.from_betas(betas)
    assert np.array_equal(co.alphas_cumprod.numpy(), g["alphas_cumprod"])
    assert np.array_equal(co.sqrt_alphas_cumprod.numpy(), g["sqrt_ac"])
    assert np.array_equal(co.sqrt_one_minus_alphas_cumprod.numpy(), g["sqrt_1mac"])
    # SURVEY.md §8c known answers
    assert abs(g["sqrt_ac"][499] - 0.52694350) < 1e-7 and abs(g["sqrt_1mac"][999] - 0.99766725) < 1e-7


def test_zero_terminal_snr_rescale_matches_reference(golden):
    g = golden("schedule")

    class Sched:
        alphas_cumprod = torch.from_numpy(g["alphas_cumprod"].copy())

    s = Sched()
    betas = rescale_noise_scheduler_to_zero_terminal_snr(s)
    assert np.array_equal(betas.numpy(), g["zt_betas"])
    assert np.array_equal(s.alphas_cumprod.numpy(), g["zt_alphas_cumprod"])
    assert s.alphas_cumprod[-1].item() == 0.0            # zero terminal SNR


def test_flow_sigmas(golden):
    sigma, one_minus = flow_sigmas(1000, "cpu")
    t = golden("addnoise")["t"]
    assert np.array_equal(sigma.numpy()[t], golden("addnoise")["flow_sigmas"])
    assert sigma[0].item() == pytest.approx(0.001) and sigma[499].item() == 0.5 and one_minus[999].item() == 0.0


def test_enum_interop_with_reference_style_enums():
    from enum import Enum

    class RefStyle(Enum):       # same shape as /root/reference/modules/util/enum/LossWeight.py
        MIN_SNR_GAMMA = "MIN_SNR_GAMMA"

        def __str__(self):
            return self.value

    assert enum_name(RefStyle.MIN_SNR_GAMMA) == enum_name(LossWeight.MIN_SNR_GAMMA) == enum_name("MIN_SNR_GAMMA")
    assert str(TimestepDistribution.LOGIT_NORMAL) == "LOGIT_NORMAL"


def _net():
    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.to_q = nn.Linear(32, 48)
            self.proj = nn.Conv2d(8, 16, 1)
            self.conv3 = nn.Conv2d(8, 8, 3, padding=1)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.attentions = nn.ModuleList([Block()])
            self.other = nn.Linear(4, 4)

    return Net()


def test_wrapper_selects_leaves_by_substring_and_keeps_state_dict_keys(golden):
    cfg = TrainConfig(lora_rank=16, lora_alpha=8.0)
    net = _net()
    w = LoRAModuleWrapper(net, "lora_unet", cfg, ["attentions.0.to_q", " attentions.0.proj "])   # filters are stripped
    assert sorted(w.lora_modules) == ["attentions.0.proj", "attentions.0.to_q"]
    keys = sorted(w.state_dict())
    assert keys == ["lora_unet.attentions.0.proj.alpha", "lora_unet.attentions.0.proj.lora_down.weight",
                    "lora_unet.attentions.0.proj.lora_up.weight", "lora_unet.attentions.0.to_q.alpha",
                    "lora_unet.attentions.0.to_q.lora_down.weight", "lora_unet.attentions.0.to_q.lora_up.weight"]
    # same key shape the reference module produced in the fixture
    assert list(golden("lora")["linear_keys"]) == ["lora_unet.x.alpha", "lora_unet.x.lora_down.weight", "lora_unet.x.lora_up.weight"]
    m = w.lora_modules["attentions.0.to_q"]
    assert tuple(m.lora_down.weight.shape) == (16, 32) and tuple(m.lora_up.weight.shape) == (48, 16)
    assert torch.count_nonzero(m.lora_up.weight) == 0                      # up = 0 at init
    bound = 1 / math.sqrt(32)                                              # kaiming_uniform(a=sqrt(5)) bound
    assert m.lora_down.weight.abs().max() <= bound
    assert float(m.alpha) == 8.0 and not m.alpha.requires_grad
    c = w.lora_modules["attentions.0.proj"]
    assert tuple(c.lora_down.weight.shape) == (16, 8, 1, 1) and tuple(c.lora_up.weight.shape) == (16, 16, 1, 1)
    assert len(w.parameters()) == 4


def test_empty_filter_selects_everything_supported_and_unsupported_layers_raise():
    cfg = TrainConfig()
    w = LoRAModuleWrapper(_net(), "lora_unet", cfg, [""])                   # "" matches the 3x3 conv too (`full` preset)
    k3 = w.lora_modules["attentions.0.conv3"]
    assert tuple(k3.lora_down.weight.shape) == (cfg.lora_rank, 8, 3, 3) and tuple(k3.lora_up.weight.shape) == (8, cfg.lora_rank, 1, 1)
    assert k3.lora_down.padding == (1, 1) and k3.lora_down.stride == (1, 1)
    with pytest.raises(NotImplementedError, match="groups-1"):
        LoRAModule("p", nn.Conv2d(8, 8, 3, padding=1, groups=2), 8, 1.0)
    with pytest.raises(NotImplementedError, match="zero-padded"):
        LoRAModule("p", nn.Conv2d(8, 8, 3, padding=1, padding_mode="reflect"), 8, 1.0)
    # force_circular_padding (conv_util.py:6-9) flips EVERY conv of the UNet before the adapters are built: the 1x1 projections
    # pad nothing, so the mode changes nothing for them and they stay supported; a padded 3x3 in circular mode does not
    for conv in (nn.Conv2d(8, 8, 1), nn.Conv2d(8, 16, 1, stride=2)):
        conv.padding_mode = "circular"
        assert LoRAModule("p", conv, 8, 1.0).lora_down.weight.shape[1] == 8
    circ = nn.Conv2d(8, 8, 3, padding=1)
    circ.padding_mode = "circular"
    with pytest.raises(NotImplementedError, match="zero-padded"):
        LoRAModule("p", circ, 8, 1.0)
    from onetrainer_b200.modules.module.LoRAModule import DoRAModule
    dora = LoRAModuleWrapper(_net(), "lora_unet", TrainConfig(train_device="cpu", lora_decompose=True), ["to_q"])     # DoRA (reference :448-457)
    assert dora.lora_modules and all(isinstance(m, DoRAModule) and m.dora_scale.dtype == torch.float32 for m in dora.lora_modules.values())
    assert all(k.endswith(("lora_down.weight", "lora_up.weight", "alpha", "dora_scale")) for k in dora.state_dict())
    with pytest.raises(NotImplementedError):
        LoRAModuleWrapper(_net(), "lora_unet", TrainConfig(peft_type=PeftType.LOHA), ["to_q"])
    with pytest.raises(NotImplementedError, match="Only Linear and Conv2d"):
        LoRAModule("p", nn.LayerNorm(4), 8, 1.0)


def test_hook_swaps_and_restores_forward_train_eval():
    net = _net()
    lin = net.attentions[0].to_q
    w = LoRAModuleWrapper(net, "lora_unet", TrainConfig(), ["to_q"])
    orig_forward = lin.forward
    w.hook_to_module()
    m = w.lora_modules["attentions.0.to_q"]
    assert lin.forward == m.forward and m.is_applied
    lin.eval()
    assert not m.training and not lin.training
    lin.train()
    assert m.training
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lin(torch.randn(2, 32))                                            # product path refuses CPU tensors loudly
    w.remove_hook_from_module()
    assert lin.forward == orig_forward and not m.is_applied
    assert lin(torch.randn(2, 32)).shape == (2, 48)


def test_load_state_dict_prefix_filter_and_dummies():
    cfg = TrainConfig(lora_rank=8)
    w1 = LoRAModuleWrapper(_net(), "lora_unet", cfg, ["to_q"])
    sd = {k: torch.randn_like(v) for k, v in w1.state_dict().items()}
    sd["lora_unet.somewhere.else.alpha"] = torch.tensor(4.0)
    sd["lora_unet.somewhere.else.lora_down.weight"] = torch.randn(8, 4)
    sd["lora_te1.ignored.alpha"] = torch.tensor(1.0)
    w2 = LoRAModuleWrapper(_net(), "lora_unet", cfg, ["to_q"])
    w2.load_state_dict(sd)
    out = w2.state_dict()
    for k in w1.state_dict():
        assert torch.equal(out[k], sd[k])
    assert "lora_unet.somewhere.else.alpha" in out and "lora_te1.ignored.alpha" not in out     # kept by a dummy
    assert isinstance(w2.lora_modules["lora_unet.somewhere.else"], DummyLoRAModule)
    w2.prune()
    assert "lora_unet.somewhere.else.alpha" not in w2.state_dict()
    with pytest.raises(ValueError):
        w2.set_dropout(1.5)
    w2.set_dropout(0.25)
    assert all(m.dropout.p == 0.25 for m in w2.lora_modules.values())


def test_uninitialised_module_raises_like_the_reference():
    d = DummyLoRAModule("x", None, 8, 1.0)
    with pytest.raises(RuntimeError, match="state dict must be loaded"):
        d.state_dict()
    with pytest.raises(NotImplementedError):
        d.hook_to_module()


def test_ema_wrapper_api_and_state_dict():
    ps = [nn.Parameter(torch.randn(5)), nn.Parameter(torch.randn(3, 2))]
    ema = EMAModuleWrapper(ps, decay=0.999, update_step_interval=5, device=torch.device("cpu"))
    assert [e.shape for e in ema.ema_parameters] == [p.shape for p in ps]
    assert ema.get_current_decay(0) == 0.1 and ema.get_current_decay(10 ** 6) == 0.999
    sd = ema.state_dict()
    assert set(sd) == {"decay", "ema_parameters"} and sd["decay"] == 0.999
    with torch.no_grad():
        for p in ps:
            p.add_(1.0)
    ema.copy_ema_to(ps, store_temp=True)
    assert torch.equal(ps[0].data, ema.ema_parameters[0])
    ema.copy_temp_to(ps)
    assert ema.temp_stored_parameters is None        # (on CPU `.cpu()` aliases, exactly as in the reference)
    with pytest.raises(ValueError):
        ema.step(ps[:1], 4)                                                # strict zip, as the reference
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ema.step(ps, 4)                                                    # same-device EMA needs the CUDA kernel


def test_adam_scalars_narrowing_matches_oracle():
    from onetrainer_b200 import ops
    from oracle import otb_oracle as O
    for step in (1.0, 2.0, 1000.0):
        a = ops.adam_scalars(3e-4, 0.9, 0.999, 1e-8, 1e-2, step, True)
        b = O.adam_scalars(3e-4, 0.9, 0.999, 1e-8, 1e-2, step, True)
        for k in a:
            assert np.float32(a[k]) == np.float32(getattr(b, k)), k
    l2 = ops.adam_scalars(1e-3, 0.9, 0.99, 1e-8, 0.05, 3.0, False)
    assert l2["decay_mul"] == 1.0 and l2["l2"] == np.float32(0.05)


def test_patch_adamw_contract_and_unsupported_options():
    from onetrainer_b200.modules.util.optimizer.adam_extensions import patch_adam
    from onetrainer_b200.modules.util.optimizer.adamw_extensions import patch_adamw
    p = nn.Parameter(torch.zeros(8))
    opt = torch.optim.AdamW([p], lr=1e-3)
    patch_adamw(opt, True)
    assert opt.stochastic_rounding is True and callable(opt.step_parameter)
    assert opt.step() is None                                              # nothing has a gradient: no launch, no error
    assert opt.step(lambda: 1.5) == 1.5                                    # closure contract
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()
    opt2 = torch.optim.Adam([nn.Parameter(torch.zeros(8))], amsgrad=True)
    patch_adam(opt2, False)
    opt2.param_groups[0]["params"][0].grad = torch.ones(8)
    with pytest.raises(NotImplementedError, match="amsgrad"):
        opt2.step()


def test_timestep_kernel_args_follow_python_semantics():
    from onetrainer_b200.modules.modelSetup.mixin.ModelSetupNoiseMixin import timestep_kernel_args, timestep_shift
    cfg = TrainConfig(min_noising_strength=0.1234, max_noising_strength=0.9, timestep_shift=3.0)
    kw = timestep_kernel_args(1000, False, cfg)
    assert kw["min_timestep"] == int(1000 * 0.1234) == 123 and kw["max_timestep"] == 900 and kw["shift"] == 3.0
    dyn = TrainConfig(dynamic_timestep_shifting=True)
    with pytest.raises(NotImplementedError, match="Dynamic timestep shifting"):
        timestep_shift(dyn, None, None)
    mu = 4096 * ((1.15 - 0.5) / (4096 - 256)) + (0.5 - ((1.15 - 0.5) / (4096 - 256)) * 256)
    assert timestep_shift(dyn, 128, 128) == pytest.approx(math.exp(mu)) == pytest.approx(math.exp(1.15))
    assert timestep_kernel_args(1000, True, cfg)["ts_mode"] == 3


def test_loss_mixin_rejects_out_of_scope_terms():
    from onetrainer_b200.modules.modelSetup.mixin.ModelSetupDiffusionLossMixin import ModelSetupDiffusionLossMixin

    class S(ModelSetupDiffusionLossMixin):
        pass

    data = {"loss_type": "target", "predicted": torch.zeros(1, 4, 2, 2), "target": torch.zeros(1, 4, 2, 2)}
    batch = {"loss_weight": torch.ones(1)}
    with pytest.raises(NotImplementedError, match="log-cosh"):
        S()._diffusion_losses(batch, data, TrainConfig(log_cosh_strength=0.5), torch.device("cpu"))
    with pytest.raises(NotImplementedError, match="VB loss"):
        S()._diffusion_losses(batch, dict(data, predicted_var_values=torch.zeros(1)), TrainConfig(), torch.device("cpu"))
    with pytest.raises(NotImplementedError, match="alphas_cumprod_fun"):
        S()._diffusion_losses(batch, dict(data, timestep=torch.zeros(1, dtype=torch.int32)),
                              TrainConfig(loss_weight_fn=LossWeight.MIN_SNR_GAMMA), torch.device("cpu"), alphas_cumprod_fun=lambda t, d: t)


def test_flux_standin_has_the_reference_leaf_set():
    """`attn-mlp` on the FLUX stand-in selects 304 leaves / 35.5 M rank-16 parameters —
    ff_context and the AdaLN `linear`s do not match the reference's substring filter (FluxLoRASetup.py:12-16)."""
    from onetrainer_b200.harness import standin
    with torch.device("meta"):
        net = standin.flux_transformer()
    names = [n for n, m in net.named_modules() if isinstance(m, nn.Linear) and any(f in n for f in ["attn", "ff.net"])]
    assert len(names) == 304
    assert not any("ff_context" in n or "norm" in n for n in names)
    params = sum(16 * (m.in_features + m.out_features) for n, m in net.named_modules() if n in set(names))
    assert params == 35_487_744


@pytest.mark.parametrize("name", ["cos_map", "cos_map_range_shift3", "sigmoid"])
def test_table_timestep_distributions_host_side(golden, name):
    """COS_MAP / SIGMOID (ModelSetupNoiseMixin.py:122-153): the host reduces the config to the reference's weight table
    (bit for bit) and its running sum for the kernel's inverse-CDF draw; like every other stage there is no CPU path."""
    from onetrainer_b200 import _native as N
    from onetrainer_b200.modules.modelSetup.mixin.ModelSetupNoiseMixin import ModelSetupNoiseMixin, timestep_kernel_args, timestep_weights

    class S(ModelSetupNoiseMixin):
        pass

    g = golden("timesteps")
    lo, hi, weight, bias, shift, _, seed, _ = g[f"{name}_cfg"]
    cfg = TrainConfig(timestep_distribution=TimestepDistribution.SIGMOID if name == "sigmoid" else TimestepDistribution.COS_MAP,
                      min_noising_strength=lo, max_noising_strength=hi, noising_weight=weight, noising_bias=bias, timestep_shift=shift)
    assert np.array_equal(timestep_weights(1000, cfg).numpy(), g[f"{name}_weights"])
    kw = timestep_kernel_args(1000, False, cfg, device="cpu")
    cdf = kw["ts_cdf"]
    assert kw["ts_mode"] == N.TS_TABLE and cdf.numel() == kw["max_timestep"] - kw["min_timestep"] == int(1000 * hi) - int(1000 * lo)
    assert cdf.dtype == torch.float32 and cdf[-1].item() == 1.0 and bool((cdf[1:] >= cdf[:-1]).all()) and cdf[0].item() > 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S()._get_timestep_discrete(1000, False, torch.Generator().manual_seed(int(seed)), 64, cfg)


# ------------------------------------------------------------------------------------------ A15 trainer glue
def _collection():
    from onetrainer_b200.modules.util.NamedParameterGroup import NamedParameterGroup, NamedParameterGroupCollection
    col = NamedParameterGroupCollection()
    col.add_group(NamedParameterGroup("unet_lora", [nn.Parameter(torch.zeros(3)), nn.Parameter(torch.zeros(2))], None, "UNet LoRA"))
    col.add_group(NamedParameterGroup("te_lora", [nn.Parameter(torch.zeros(5))], 1e-5))
    return col


@pytest.mark.parametrize("scaler", ["NONE", "BATCH", "GRADIENT_ACCUMULATION", "BOTH"])
def test_parameter_groups_for_optimizer_vs_reference(golden, scaler):
    """NamedParameterGroupCollection.parameters_for_optimizer (NamedParameterGroup.py:36-61): the same dicts, bit for bit."""
    from onetrainer_b200.modules.util.enum import LearningRateScaler
    g = golden("param_groups")
    cfg = TrainConfig(learning_rate=3e-4, learning_rate_scaler=LearningRateScaler[scaler], batch_size=4, gradient_accumulation_steps=2)
    col = _collection()
    groups = col.parameters_for_optimizer(cfg)
    assert [grp["lr"] for grp in groups] == g[f"{scaler}_lr"].tolist()
    assert [grp["initial_lr"] for grp in groups] == g[f"{scaler}_initial_lr"].tolist()
    assert [len(grp["params"]) for grp in groups] == g[f"{scaler}_counts"].tolist() and sorted(groups[0]) == list(g[f"{scaler}_keys"])
    assert col.unique_name_mapping == list(g["unique_names"]) and col.display_name_mapping == list(g["display_names"])
    assert len(col.parameters()) == int(g["n_parameters"][0]) and col.parameters()[0] is groups[0]["params"][0]


def test_create_optimizer_and_ema_keep_the_reference_contract():
    """create_optimizer ADAM / ADAMW branches (create.py:481-534), create_ema (:1089-1111), init_model_parameters
    (optimizer_util.py:50-65): defaults, errors, and the attributes the trainer and the savers read afterwards."""
    from onetrainer_b200.modules.util import create
    from onetrainer_b200.modules.util.config.TrainConfig import TrainOptimizerConfig
    from onetrainer_b200.modules.util.enum import EMAMode, Optimizer
    from onetrainer_b200.modules.util.optimizer_util import init_model_parameters
    cfg = TrainConfig(learning_rate=3e-4, train_device="cpu", optimizer=TrainOptimizerConfig(stochastic_rounding=True))
    opt = create.create_optimizer(_collection(), None, cfg)
    assert type(opt) is torch.optim.AdamW and opt.stochastic_rounding and callable(opt.step_parameter)
    g0, g1 = opt.param_groups
    assert (g0["lr"], g1["lr"], g0["betas"], g0["eps"], g0["weight_decay"]) == (3e-4, 1e-5, (0.9, 0.999), 1e-8, 1e-2)
    adam = create.create_optimizer(_collection(), None, TrainConfig(optimizer=TrainOptimizerConfig(optimizer=Optimizer.ADAM, beta2=0.99)))
    assert type(adam) is torch.optim.Adam and adam.param_groups[0]["weight_decay"] == 0 and adam.param_groups[0]["betas"] == (0.9, 0.99)
    with pytest.raises(RuntimeError, match='"stochastic_rounding" is only allowed'):
        create.create_optimizer(_collection(), None, TrainConfig(optimizer=TrainOptimizerConfig(stochastic_rounding=True, foreach=True)))
    with pytest.raises(RuntimeError, match='"fused_back_pass" is only allowed'):
        create.create_optimizer(_collection(), None, TrainConfig(optimizer=TrainOptimizerConfig(fused_back_pass=True, fused=True)))
    with pytest.raises(NotImplementedError, match="LION"):
        create.create_optimizer(_collection(), None, TrainConfig(optimizer=TrainOptimizerConfig(optimizer="LION")))
    # ADAFACTOR (create.py:914-936) and CAME (:939-951): the reference's classes and defaults, step rebound to the fused launch
    af = create.create_optimizer(_collection(), None, TrainConfig(optimizer=TrainOptimizerConfig(optimizer=Optimizer.ADAFACTOR, relative_step=True)))
    assert type(af).__name__ == "Adafactor" and af.param_groups[0]["relative_step"] and af.param_groups[0]["lr"] is None
    assert af.param_groups[0]["eps"] == (1e-30, 1e-3) and af.param_groups[0]["decay_rate"] == -0.8 and hasattr(af, "step_parameter")
    af2 = create.create_optimizer(_collection(), None, TrainConfig(learning_rate=1e-4, optimizer=TrainOptimizerConfig(
        optimizer=Optimizer.ADAFACTOR, relative_step=False, scale_parameter=False, beta1=0.9)))
    assert af2.param_groups[0]["lr"] == 1e-4 and af2.param_groups[0]["beta1"] == 0.9 and af2.stochastic_rounding is False
    came = create.create_optimizer(_collection(), None, TrainConfig(learning_rate=2e-4, optimizer=TrainOptimizerConfig(
        optimizer=Optimizer.CAME, stochastic_rounding=True)))
    assert type(came).__name__ == "CAME" and came.defaults["betas"] == (0.9, 0.999, 0.9999) and came.defaults["eps"] == (1e-30, 1e-16)
    assert came.stochastic_rounding is True and came.param_groups[1]["lr"] == 1e-5
    assert create.create_ema(_collection().parameters(), None, cfg) is None
    with pytest.raises(NotImplementedError, match="CPU"):
        create.create_ema(_collection().parameters(), None, TrainConfig(ema=EMAMode.CPU))
    ema = create.create_ema(_collection().parameters(), None, TrainConfig(ema=EMAMode.GPU, train_device="cpu", ema_decay=0.99))
    assert isinstance(ema, EMAModuleWrapper) and ema.decay == 0.99 and len(ema.ema_parameters) == 3

    class Model:
        pass

    model = Model()
    model.train_config = TrainConfig(ema=EMAMode.GPU, train_device="cpu")
    model.optimizer_state_dict = model.ema_state_dict = None
    init_model_parameters(model, _collection(), torch.device("cpu"))
    assert model.param_group_mapping == ["unet_lora", "te_lora"] and model.optimizer_state_dict is None and model.ema_state_dict is None
    assert isinstance(model.optimizer, torch.optim.AdamW) and isinstance(model.ema, EMAModuleWrapper)
    assert len(model.parameters.parameters()) == 3


# ------------------------------------------------------------------------------------------ f4: on-disk LoRA key formats
def test_lora_saver_key_formats_round_trip(tmp_path):
    """LoRASaverMixin._save (modelSaver/mixin/LoRASaverMixin.py:38-106): legacy diffusers keys (adapter prefix + '_' + the
    module path with '.' -> '_'), dtype conversion, the INTERNAL layout, and loading a legacy file back into a wrapper."""
    from safetensors import safe_open
    from safetensors.torch import load_file
    from onetrainer_b200.modules.modelSaver.mixin.LoRASaverMixin import (ModelFormat, StandInLoRASaver, convert_legacy_diffusers_to_internal,
                                                                        convert_to_legacy_diffusers, sdxl_lora_key_sets)
    net = _net()
    wrap = LoRAModuleWrapper(net, "lora_unet", TrainConfig(lora_rank=8), ["to_q", "to_out"])
    for m in wrap.lora_modules.values():
        nn.init.normal_(m.lora_up.weight, std=0.02)

    class Model:
        unet_lora, sdxl = wrap, True
        lora_state_dict = {"bundle_emb.<tok>.clip_l": torch.ones(2, 4)}

    sd = wrap.state_dict()
    legacy = convert_to_legacy_diffusers(sd, sdxl_lora_key_sets())
    k_internal = next(k for k in sd if k.endswith("to_q.lora_down.weight"))
    k_legacy = "lora_unet_" + k_internal[len("lora_unet."):-len(".lora_down.weight")].replace(".", "_") + ".lora_down.weight"
    assert k_legacy in legacy and "." not in k_legacy[: -len(".lora_down.weight")] and len(legacy) == len(sd)
    saver = StandInLoRASaver()
    dst = tmp_path / "out" / "lora.safetensors"
    saver.save(Model, ModelFormat.SAFETENSORS, str(dst), torch.bfloat16)
    loaded = load_file(str(dst))
    assert set(loaded) == set(legacy) | {"bundle_emb.<tok>.clip_l"} and all(v.dtype == torch.bfloat16 for v in loaded.values())
    with safe_open(str(dst), "pt") as f:
        assert f.metadata()["ss_base_model_version"] == "sdxl_"
    saver.save(Model, ModelFormat.LEGACY_SAFETENSORS, str(tmp_path / "legacy.safetensors"), None)
    assert set(load_file(str(tmp_path / "legacy.safetensors"))) == set(loaded)
    saver.save(Model, ModelFormat.INTERNAL, str(tmp_path / "backup"), None)
    internal = load_file(str(tmp_path / "backup" / "lora" / "lora.safetensors"))
    assert set(internal) == set(sd) | {"bundle_emb.<tok>.clip_l"} and internal[k_internal].dtype == torch.float32
    with pytest.raises(NotImplementedError):
        saver.save(Model, ModelFormat.DIFFUSERS, str(tmp_path / "x"), None)
    with pytest.raises(NotImplementedError, match="omi_model_standards"):
        saver._save(Model, ModelFormat.SAFETENSORS, str(tmp_path / "omi.safetensors"), None, enable_omi_format=True)
    # load the legacy file into a fresh wrapper
    w2 = LoRAModuleWrapper(_net(), "lora_unet", TrainConfig(lora_rank=8), ["to_q", "to_out"])
    back = convert_legacy_diffusers_to_internal({k: v.float() for k, v in loaded.items()}, w2.state_dict().keys())
    w2.load_state_dict(back)
    assert torch.equal(w2.state_dict()[k_internal], sd[k_internal].detach().bfloat16().float())


def test_sharded_exchange_plan_is_contiguous_tensor_aligned_and_balanced():
    """dp.ShardedExchange.plan (K-z host logic): every tensor on an 8-element boundary (= the replicated optimizer table's
    rng_base), buckets contiguous / tensor-aligned / covering everything, owners round-robin, shares balanced."""
    import random
    from onetrainer_b200 import dp
    from onetrainer_b200 import ops
    rnd = random.Random(0)
    numels = [rnd.choice([320, 640, 1280, 77, 13, 1280 * 1280, 640 * 640, 10240 * 1280, 5120 * 640, 9]) for _ in range(400)]
    for world in (2, 4, 8):
        offsets, total, bounds, first = dp.ShardedExchange.plan(numels, world)
        assert all(o % 8 == 0 for o in offsets) and total % 8 == 0
        run = 0
        for n, o in zip(numels, offsets):           # same running offset as ops.OptimTable (rng_base += ceil8(n))
            assert o == run
            run += (n + 7) // 8 * 8
        assert bounds[0] == 0 and bounds[-1] == total and bounds == sorted(set(bounds))
        assert all(bounds[b] == offsets[first[b]] for b in range(len(bounds) - 1))
        assert first[0] == 0 and first[-1] == len(numels) and all(a < b for a, b in zip(first, first[1:]))
        nb = len(bounds) - 1
        assert nb >= world
        share = [0] * world
        for b in range(nb):
            share[b % world] += bounds[b + 1] - bounds[b]
        assert max(share) <= 1.35 * total / world, (world, share)
    # fewer tensors than buckets asked for: never an empty bucket
    offsets, total, bounds, first = dp.ShardedExchange.plan([100, 200, 300], 2, n_buckets=8)
    assert len(bounds) - 1 <= 3 and all(a < b for a, b in zip(bounds, bounds[1:]))


def test_grad_queue_never_carries_leaves_past_the_pass_that_queued_them(monkeypatch):
    """ops.GradQueue (K-a' host logic): every queued dA/dB contraction is launched by the end of the backward pass that queued
    it — also when an EARLIER pass raised after queueing (the engine drops a failed pass's callbacks), and when a re-entrant
    backward runs inside another.  The launches are recorded instead of executed (no GPU here)."""
    import contextlib
    from onetrainer_b200 import ops
    launched = []
    monkeypatch.setattr(ops, "lora_grad_batched", lambda leaves: launched.extend(leaves))
    monkeypatch.setattr(torch.cuda, "device", lambda idx: contextlib.nullcontext())
    q = ops.GradQueue()
    dev = torch.device("cuda", 0)

    class Leaf(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, tag, boom):
            ctx.tag, ctx.boom = tag, boom
            return x * 1.0

        @staticmethod
        def backward(ctx, g):
            q.push((ctx.tag,), dev)
            if ctx.boom:
                raise RuntimeError("boom")
            return g, None, None

    def chain(tags, boom_at=None):
        x = torch.ones(3, requires_grad=True)
        y = x
        for t in tags:
            y = Leaf.apply(y, t, t == boom_at)
        return x, y.sum()

    # outside a backward pass: launched at once
    q.push(("eager",), dev)
    assert launched == [("eager",)]
    launched.clear()
    # a pass that raises after queueing two leaves: nothing was flushed by it ...
    _, loss = chain(["a3", "a2", "a1"], boom_at="a2")
    with pytest.raises(RuntimeError, match="boom"):
        loss.backward()
    assert launched == [] and len(q._pending[0]) == 2
    # ... the next pass launches the stale leaves first and still ends with its own flush
    _, loss = chain(["b2", "b1"])
    loss.backward()
    assert [t[0] for t in launched] == ["a1", "a2", "b1", "b2"] and q._pending[0] == []
    launched.clear()
    # a full batch flushes by itself, the remainder at the end of the pass
    q.FLUSH_LEAVES = 2
    _, loss = chain(["c3", "c2", "c1"])
    loss.backward()
    assert [t[0] for t in launched] == ["c1", "c2", "c3"] and q._pending[0] == []
    launched.clear()
    q.FLUSH_LEAVES = 96

    # re-entrant backward inside a backward pass: the inner pass flushes the outer pass's queued leaf and its own
    class Nested(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x * 1.0

        @staticmethod
        def backward(ctx, g):
            with torch.enable_grad():
                _, inner = chain(["inner"])
            inner.backward()
            assert [t[0] for t in launched] == ["outer1", "inner"]
            return g

    x = torch.ones(3, requires_grad=True)
    y = Leaf.apply(Nested.apply(Leaf.apply(x, "outer2", False)), "outer1", False)
    y.sum().backward()
    assert [t[0] for t in launched] == ["outer1", "inner", "outer2"] and q._pending[0] == []


def test_sr_helper_module_keeps_the_reference_signatures_and_has_no_cpu_path():
    """modules/util/bf16_stochastic_rounding.py: the three functions of the reference's file under the same names with the same
    parameter names and defaults (compared with the unmodified file when baseline/_ref is present, else with the signatures
    written out here from bf16_stochastic_rounding.py:5,33,48), and CPU tensors get the 'no CPU fallback' error — not a
    silent torch path, and not an index error from a CUDA generator that does not exist."""
    import inspect
    from onetrainer_b200.modules.util import bf16_stochastic_rounding as mine
    want = {"copy_stochastic_": ["target", "source"], "add_stochastic_": ["input", "other", "alpha"],
            "addcdiv_stochastic_": ["input", "tensor1", "tensor2", "value"]}
    from baseline import ref_arm
    if ref_arm.available():
        ref_arm.load()
        import modules.util.bf16_stochastic_rounding as ref
        for name in want:
            ps = inspect.signature(getattr(ref, name)).parameters
            assert list(ps) == want[name], (name, list(ps))
            assert [p.default for p in ps.values()] == [p.default for p in inspect.signature(getattr(mine, name)).parameters.values()]
    for name, params in want.items():
        assert list(inspect.signature(getattr(mine, name)).parameters) == params
    p = torch.zeros(16, dtype=torch.bfloat16)
    for call in (lambda: mine.copy_stochastic_(p, torch.zeros(16)), lambda: mine.add_stochastic_(p, torch.zeros(16), alpha=0.5),
                 lambda: mine.addcdiv_stochastic_(p, torch.ones(16), torch.ones(16), value=-1.0)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    with pytest.raises(ValueError, match="float32"):
        mine.copy_stochastic_(p, torch.zeros(16, dtype=torch.bfloat16))


def test_patched_optimizers_survive_the_lr_scheduler_step_wrapper():
    """The reference builds its LambdaLR right after the optimizer (create.py: create_optimizer, then create_lr_scheduler);
    torch's scheduler wraps `optimizer.step` by re-binding `optimizer.step.__func__` to the optimizer
    (lr_scheduler.py: patch_track_step_called).  The reference's patches are bound to the optimizer
    (adamw_extensions.py:202-205), so must these be: a scheduler created AFTER the patch steps through the fused driver,
    records the call, rewrites `lr` — and `step_parameter` keeps its (p, group, i) signature."""
    import inspect
    from transformers.optimization import Adafactor
    from onetrainer_b200.modules.util.optimizer.adafactor_extensions import patch_adafactor
    from onetrainer_b200.modules.util.optimizer.adam_extensions import patch_adam
    from onetrainer_b200.modules.util.optimizer.adamw_extensions import patch_adamw
    from onetrainer_b200.modules.util.optimizer.CAME import CAME

    def params():
        return [nn.Parameter(torch.zeros(8, 4))]

    opts = []
    o = torch.optim.AdamW(params(), lr=1e-3); patch_adamw(o, True); opts.append(o)
    o = torch.optim.Adam(params(), lr=1e-3); patch_adam(o, False); opts.append(o)
    o = Adafactor(params(), lr=1e-3, relative_step=False, scale_parameter=False); patch_adafactor(o, False); opts.append(o)
    opts.append(CAME(params(), lr=1e-3))
    for opt in opts:
        assert opt.step.__self__ is opt and opt.step_parameter.__self__ is opt, type(opt).__name__
        assert list(inspect.signature(opt.step_parameter).parameters) == ["p", "group", "i"]
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda step: 0.5 ** step)
        assert getattr(opt.step, "_wrapped_by_lr_sched", False)
        assert opt.step() is None and opt.step(lambda: 2.5) == 2.5          # no gradient anywhere: no launch, closure contract
        assert opt._opt_called
        sched.step()
        assert opt.param_groups[0]["lr"] == pytest.approx(5e-4)
        p = opt.param_groups[0]["params"][0]
        p.grad = torch.ones_like(p)
        with pytest.raises(RuntimeError, match="no CPU fallback"):           # the wrapper reached the fused driver
            opt.step()
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            opt.step_parameter(p, opt.param_groups[0], 0)


def test_saved_optimizer_state_follows_the_group_names_on_resume():
    """create_optimizer's hand-over of a backup's optimizer state (create.py:1038-1084 with the saver's two mapping lists,
    InternalModelSaverMixin.py:19-26): groups are matched by unique name, so a resumed run may reorder, add or drop groups;
    a group trained by another optimizer type starts empty; learning rates always come from the new config; without the
    mapping (old backups) only the rates are refreshed.  The SR stream position travels in the same dict."""
    from onetrainer_b200.modules.util import create
    from onetrainer_b200.modules.util.config.TrainConfig import TrainOptimizerConfig
    from onetrainer_b200.modules.util.enum import Optimizer
    from onetrainer_b200.modules.util.NamedParameterGroup import NamedParameterGroup, NamedParameterGroupCollection

    def collection(order):
        sizes = {"unet_lora": (3, 2), "te_lora": (5,), "embeddings": (7,)}
        col = NamedParameterGroupCollection()
        for name in order:
            col.add_group(NamedParameterGroup(name, [nn.Parameter(torch.zeros(n)) for n in sizes[name]], None))
        return col

    cfg = TrainConfig(learning_rate=3e-4, train_device="cpu", optimizer=TrainOptimizerConfig(stochastic_rounding=True))
    first = collection(["unet_lora", "te_lora"])
    opt = create.create_optimizer(first, None, cfg)
    for k, p in enumerate(first.parameters()):          # what three steps would have left behind, recognisable per tensor
        opt.state[p] = {"step": torch.tensor(3.0), "exp_avg": torch.full_like(p, k + 1.0), "exp_avg_sq": torch.full_like(p, 10.0 * (k + 1))}
    opt.otb_driver.step_count = 3
    saved = opt.state_dict()
    assert saved["otb_sr_step_count"] == 3
    saved["param_group_mapping"] = first.unique_name_mapping
    saved["param_group_optimizer_mapping"] = [str(cfg.optimizer.optimizer) for _ in first.unique_name_mapping]
    import copy

    # resumed with the groups reordered, one added, a new learning rate
    cfg2 = TrainConfig(learning_rate=1e-4, train_device="cpu", optimizer=TrainOptimizerConfig(stochastic_rounding=True))
    second = collection(["te_lora", "embeddings", "unet_lora"])
    opt2 = create.create_optimizer(second, copy.deepcopy(saved), cfg2)
    te, emb, (u0, u1) = second._groups[0].parameters[0], second._groups[1].parameters[0], second._groups[2].parameters
    assert float(opt2.state[te]["exp_avg"][0]) == 3.0 and opt2.state[te]["exp_avg"].shape == (5,)      # was the third tensor saved
    assert float(opt2.state[u0]["exp_avg"][0]) == 1.0 and float(opt2.state[u1]["exp_avg_sq"][0]) == 20.0
    assert len(opt2.state[emb]) == 0                                                                   # new group: empty state
    assert [g["lr"] for g in opt2.param_groups] == [1e-4] * 3 and [g["initial_lr"] for g in opt2.param_groups] == [1e-4] * 3
    assert opt2.otb_driver.step_count == 3 and opt2.otb_driver.table is None

    # the same backup under another optimizer type: every group starts empty
    cfg3 = TrainConfig(learning_rate=1e-4, train_device="cpu", optimizer=TrainOptimizerConfig(optimizer=Optimizer.ADAM))
    third = collection(["unet_lora", "te_lora"])
    opt3 = create.create_optimizer(third, copy.deepcopy(saved), cfg3)
    assert all(len(opt3.state[p]) == 0 for p in third.parameters())

    # an old backup without the mapping: loaded positionally, rates refreshed
    old = copy.deepcopy(saved)
    del old["param_group_mapping"], old["param_group_optimizer_mapping"]
    fourth = collection(["unet_lora", "te_lora"])
    opt4 = create.create_optimizer(fourth, old, cfg2)
    assert float(opt4.state[fourth.parameters()[2]]["exp_avg"][0]) == 3.0 and opt4.param_groups[0]["lr"] == 1e-4
