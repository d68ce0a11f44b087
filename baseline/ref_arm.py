"""Drives the UNMODIFIED reference (sangoi-exe/OneTrainer) through the same step as the own arm.

BENCH / TEST INFRASTRUCTURE ONLY: nothing under onetrainer_b200/ imports this file.  The reference sources themselves
are not in this repository: `tools/make_baseline_ref.py` copies the import closure of the hot-path modules (SURVEY.md
§8a) from /root/reference into git-ignored `baseline/_ref/`, which travels to the GPU box with the snapshot.

What runs here is the reference's own code: `LoRAModuleWrapper` / `LoRAModule.forward` (modules/module/LoRAModule.py),
`ModelSetupNoiseMixin`, `ModelSetupDiffusionMixin`, `ModelSetupDiffusionLossMixin` (modules/modelSetup/mixin/),
`patch_adamw` (modules/util/optimizer/adamw_extensions.py), `EMAModuleWrapper` (modules/module/EMAModule.py) and
`TrainConfig` (modules/util/config/TrainConfig.py), on CPU fp32 or on the GPU under bf16 autocast.  What is restated
(the reference's versions need `diffusers` / `mgds`, which are not installed): the family `predict()` /
`calculate_loss()` bodies (BaseStableDiffusionXLSetup.py:188-291,357,360-373) and the trainer's step sequence
(GenericTrainer.py:683-747) — the same restatement the own arm's harness uses, with the reference's mixins underneath —
and the wrapped network is the same seeded stand-in (onetrainer_b200/harness/standin.py).

Two import shims: a stub `accelerate` module (modules/util/torch_util.py:8-13 builds an Accelerator only
to read `.device`) and an alias for torch's renamed `Optimizer._cuda_graph_capture_health_check`
(adamw_extensions.py:185 calls the torch-2.6 name).
"""
from __future__ import annotations

import sys
import time
import types
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REF_DIR = REPO / "baseline" / "_ref"

_ns = None


def install_shims() -> None:
    import transformers  # noqa: F401  (must be imported before the accelerate stub: it probes the real package's version)
    import torch
    if "accelerate" not in sys.modules or not hasattr(sys.modules["accelerate"], "_otb_stub"):
        acc = types.ModuleType("accelerate")

        class Accelerator:
            def __init__(self, *a, **k):
                self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

        acc.Accelerator = Accelerator
        acc._otb_stub = True
        sys.modules["accelerate"] = acc
    try:
        import diffusers  # noqa: F401
    except Exception:
        dif = types.ModuleType("diffusers")
        dif.DDIMScheduler = object         # a type annotation in rescale_noise_scheduler_to_zero_terminal_snr.py:3
        sys.modules["diffusers"] = dif
    if not hasattr(torch.optim.Optimizer, "_cuda_graph_capture_health_check"):
        torch.optim.Optimizer._cuda_graph_capture_health_check = torch.optim.Optimizer._accelerator_graph_capture_health_check


def available() -> bool:
    return (REF_DIR / "modules" / "module" / "LoRAModule.py").exists()


def load():
    """Import the reference's hot-path modules from baseline/_ref (once).  Raises if the copy is missing."""
    global _ns
    if _ns is not None:
        return _ns
    if not available():
        raise RuntimeError(f"{REF_DIR} is missing: run `python tools/make_baseline_ref.py` in the build container "
                           "(needs /root/reference); the directory travels to the GPU box with the snapshot")
    install_shims()
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    ns = types.SimpleNamespace()
    from modules.module.EMAModule import EMAModuleWrapper
    from modules.module.LoRAModule import LoRAModule, LoRAModuleWrapper
    from modules.modelSetup.mixin.ModelSetupDiffusionLossMixin import ModelSetupDiffusionLossMixin
    from modules.modelSetup.mixin.ModelSetupDiffusionMixin import ModelSetupDiffusionMixin
    from modules.modelSetup.mixin.ModelSetupFlowMatchingMixin import ModelSetupFlowMatchingMixin
    from modules.modelSetup.mixin.ModelSetupNoiseMixin import ModelSetupNoiseMixin
    from modules.util.config.TrainConfig import TrainConfig
    from modules.util.enum.LossWeight import LossWeight
    from modules.util.enum.TimestepDistribution import TimestepDistribution
    from modules.util.optimizer.adamw_extensions import patch_adamw
    import modules.module.LoRAModule as lora_file
    assert Path(lora_file.__file__).resolve().is_relative_to(REF_DIR.resolve()), "reference modules resolved outside baseline/_ref"
    ns.__dict__.update(locals())
    _ns = ns
    return ns


def reference_config(**overrides):
    """`TrainConfig.default_values()` of the reference with the given fields set."""
    R = load()
    cfg = R.TrainConfig.default_values()
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"reference TrainConfig has no field {k!r}")
        setattr(cfg, k, v)
    return cfg


def _setup_classes():
    import torch
    R = load()

    class RefDiffusionSetup(R.ModelSetupNoiseMixin, R.ModelSetupDiffusionMixin, R.ModelSetupDiffusionLossMixin):
        """BaseStableDiffusionXLSetup.predict (:188-291,357) / calculate_loss (:360-373) with cached text-encoder
        outputs, on the reference's own mixins — call by call, as the reference runs it."""

        def __init__(self, train_device):
            super().__init__()
            self.train_device = train_device

        def predict(self, model, batch, config, train_progress, *, deterministic=False, draws=None):
            with model.autocast_context:
                generator = torch.Generator(device=config.train_device)
                generator.manual_seed(0 if deterministic else train_progress.global_step)
                scaled_latent_image = batch["latent_image"] * model.vae_scaling_factor
                T = model.noise_scheduler.config["num_train_timesteps"]
                if draws is None:
                    latent_noise = self._create_noise(scaled_latent_image, config, generator)
                    timestep = self._get_timestep_discrete(T, deterministic, generator, scaled_latent_image.shape[0], config)
                else:           # parity runs feed both arms the same draws (the two RNG streams differ by design)
                    latent_noise, timestep = draws["noise"].to(scaled_latent_image.dtype), draws["timestep"]
                noisy = self._add_noise_discrete(scaled_latent_image, latent_noise, timestep, model.noise_scheduler.betas)
                latent_input = noisy.to(dtype=model.train_dtype)
                ptype = model.noise_scheduler.config["prediction_type"]
                target = latent_noise if ptype == "epsilon" else model.noise_scheduler.get_velocity(scaled_latent_image, latent_noise, timestep)
                kwargs = {}
                if model.sdxl:
                    add_time_ids = torch.stack([
                        batch["original_resolution"][0], batch["original_resolution"][1], batch["crop_offset"][0], batch["crop_offset"][1],
                        batch["crop_resolution"][0], batch["crop_resolution"][1]], dim=1).to(dtype=latent_input.dtype, device=latent_input.device)
                    kwargs["added_cond_kwargs"] = {"text_embeds": batch["text_encoder_2_pooled_state"], "time_ids": add_time_ids}
                predicted = model.unet(latent_input, timestep, batch["text_encoder_hidden_state"].to(dtype=model.train_dtype), **kwargs).sample
                data = {"loss_type": "target", "timestep": timestep, "predicted": predicted, "target": target}
            data["prediction_type"] = ptype
            return data

        def calculate_loss(self, model, batch, data, config):
            return self._diffusion_losses(batch=batch, data=data, config=config, train_device=self.train_device,
                                          betas=model.noise_scheduler.betas.to(device=self.train_device)).mean()

    class RefFlowSetup(R.ModelSetupNoiseMixin, R.ModelSetupFlowMatchingMixin, R.ModelSetupDiffusionLossMixin):
        """BaseFluxSetup.predict (:202-313) / calculate_loss (:377-390), cached text-encoder outputs."""

        def __init__(self, train_device, guidance_scale=1.0):
            super().__init__()
            self.train_device = train_device
            self.guidance_scale = guidance_scale

        def predict(self, model, batch, config, train_progress, *, deterministic=False, draws=None):
            with model.autocast_context:
                generator = torch.Generator(device=config.train_device)
                generator.manual_seed(0 if deterministic else train_progress.global_step)
                latent_image = batch["latent_image"]
                T = model.noise_scheduler.config["num_train_timesteps"]
                B, C, H, W = latent_image.shape
                scaled_latent_image = (latent_image - model.vae_shift_factor) * model.vae_scaling_factor
                if draws is None:
                    latent_noise = self._create_noise(scaled_latent_image, config, generator)
                    timestep = self._get_timestep_discrete(T, deterministic, generator, B, config, latent_height=H, latent_width=W)
                else:
                    latent_noise, timestep = draws["noise"].to(scaled_latent_image.dtype), draws["timestep"]
                noisy, _sigma = self._add_noise_discrete(scaled_latent_image, latent_noise, timestep, model.noise_scheduler.timesteps)
                latent_input = noisy.to(dtype=model.train_dtype)
                guidance = None
                if model.transformer.config["guidance_embeds"]:
                    guidance = torch.full((B,), float(self.guidance_scale), device=latent_input.device, dtype=model.train_dtype)
                packed = model.pack_latents(latent_input, B, C, H, W)
                packed_predicted_flow = model.transformer(
                    hidden_states=packed.to(dtype=model.train_dtype), timestep=timestep / 1000, guidance=guidance,
                    pooled_projections=batch["text_encoder_1_pooled_state"].to(dtype=model.train_dtype),
                    encoder_hidden_states=batch["text_encoder_2_hidden_state"].to(dtype=model.train_dtype),
                    txt_ids=None, img_ids=None, joint_attention_kwargs=None, return_dict=True).sample
                predicted_flow = model.unpack_latents(packed_predicted_flow, H, W)
                flow = latent_noise - scaled_latent_image
                data = {"loss_type": "target", "timestep": timestep, "predicted": predicted_flow, "target": flow}
            return data

        def calculate_loss(self, model, batch, data, config):
            return self._flow_matching_losses(batch=batch, data=data, config=config, train_device=self.train_device,
                                              sigmas=model.noise_scheduler.sigmas.to(device=self.train_device)).mean()

    return RefDiffusionSetup, RefFlowSetup


def build_reference_model(kind: str, config, device, *, train_dtype, lora_weight_dtype=None, prediction_type="epsilon", ema=False,
                          stochastic_rounding=True, seed=1234, lr=3e-4, lora_up_std=0.0, module_filter=None, network=None):
    """The own arm's `trainer.build_model` / `build_flux_model` with the REFERENCE's adapter, optimizer patch and EMA:
    same seeded stand-in (or `network`, an already-built frozen stand-in to share its weights), reference
    LoRAModuleWrapper hooked on it (StableDiffusionXLLoRASetup.setup_model :103-158), torch AdamW + the reference's
    patch_adamw (create.py:509-534), reference EMAModuleWrapper.  `config` is a reference TrainConfig."""
    import torch
    from onetrainer_b200.harness import standin, trainer
    R = load()
    RefDiffusionSetup, RefFlowSetup = _setup_classes()
    device = torch.device(device)
    flux = kind in ("flux", "tiny_flux")
    if network is None:
        torch.manual_seed(seed)
        with torch.device(device):
            if flux:
                network = {"flux": standin.flux_transformer, "tiny_flux": standin.tiny_flux}[kind]()
            else:
                network = {"sdxl": standin.sdxl_unet, "sd15": standin.sd15_unet, "tiny": standin.tiny_unet}[kind]()
        network = network.to(dtype=train_dtype)
        network.requires_grad_(False)
        network.train()
    if flux:
        model = trainer.StandInFluxModel(network, standin.StandInFlowScheduler(1000, device), train_dtype)
    else:
        sched = standin.StandInScheduler(trainer.standin_betas().to(device), prediction_type)
        model = trainer.StandInModel(network, sched, 0.13025 if kind == "sdxl" else 0.18215, train_dtype, sdxl=(kind == "sdxl"))
    if device.type == "cpu":
        import contextlib
        model.autocast_context = contextlib.nullcontext()          # config C1: fp32 on the CPU, no autocast
    if module_filter is None:
        module_filter = ["attn", "ff.net"] if flux else ["attentions"]
    wrap = R.LoRAModuleWrapper(network, "lora_transformer" if flux else "lora_unet", config, module_filter)
    wrap.set_dropout(config.dropout_probability)
    wrap.to(dtype=lora_weight_dtype or torch.float32)
    if lora_up_std > 0:
        g = torch.Generator(device=device).manual_seed(seed + 1)
        for m in wrap.lora_modules.values():
            m.lora_up.weight.data.normal_(0.0, lora_up_std, generator=g)
    wrap.hook_to_module()
    wrap.requires_grad_(True)
    model.unet_lora = wrap
    if flux:
        model.transformer_lora = wrap
    model.parameters = wrap.parameters()
    model.optimizer = torch.optim.AdamW(model.parameters, lr=lr, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    R.patch_adamw(model.optimizer, stochastic_rounding)
    if ema:
        model.ema = R.EMAModuleWrapper(model.parameters, decay=config.ema_decay, update_step_interval=config.ema_update_step_interval,
                                       device=device)
    setup = RefFlowSetup(device) if flux else RefDiffusionSetup(device)
    return model, setup


def reference_train_step(setup, model, batch, config, *, draws=None, timings: dict | None = None):
    """The step body of GenericTrainer.train (:683-747), gradient accumulation 1: predict -> loss -> backward ->
    loss.item() (:699) -> nn.utils.clip_grad_norm_ (:712-713) -> optimizer.step() -> zero_grad(set_to_none=True) ->
    ema.step.  Returns the loss as a python float (the reference syncs on it every micro-step)."""
    import torch
    t0 = time.perf_counter()
    data = setup.predict(model, batch, config, model.train_progress, draws=draws)
    loss = setup.calculate_loss(model, batch, data, config)
    loss = loss / config.gradient_accumulation_steps
    loss.backward()
    value = loss.item()
    t1 = time.perf_counter()
    if config.clip_grad_norm is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters, config.clip_grad_norm)
    model.optimizer.step()
    model.optimizer.zero_grad(set_to_none=True)
    if model.ema is not None:
        model.ema.step(model.parameters, model.train_progress.global_step // config.gradient_accumulation_steps)
    model.train_progress.next_step(config.batch_size)
    if timings is not None:
        if torch.cuda.is_available() and str(config.train_device).startswith("cuda"):
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        timings["forward_backward"] = timings.get("forward_backward", 0.0) + (t1 - t0)
        timings["clip_optimizer_ema"] = timings.get("clip_optimizer_ema", 0.0) + (t2 - t1)
    return value
