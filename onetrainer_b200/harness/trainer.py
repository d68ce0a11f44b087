"""Training-step harness: the reference's step body on the B200 hot path (HARNESS, not product).

`predict` / `calculate_loss` restate the family setup bodies the reference keeps unchanged
(BaseStableDiffusionXLSetup.predict :188-291,357 and calculate_loss :360-373; SD 1.5 is the same
minus added_cond_kwargs, BaseStableDiffusionSetup.py:144-224,312-325) with the mixins of this package
underneath; `train_step` follows GenericTrainer.train :683-747 (predict -> loss -> /GA -> backward ->
[allreduce] -> clip -> optimizer.step -> zero_grad -> ema.step).  Conditioning comes from the batch
(cached text-encoder states), as with latent/text caching in the reference.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import dp, step_state
from ..rng import PhiloxKey
from ..modules.module.EMAModule import EMAModuleWrapper
from ..modules.module.LoRAModule import GradArena, LoRAModuleWrapper
from ..modules.modelSetup.mixin.ModelSetupDiffusionLossMixin import ModelSetupDiffusionLossMixin
from ..modules.modelSetup.mixin.ModelSetupDiffusionMixin import ModelSetupDiffusionMixin
from ..modules.modelSetup.mixin.ModelSetupFlowMatchingMixin import ModelSetupFlowMatchingMixin
from ..modules.modelSetup.mixin.ModelSetupFusedNoiseMixin import ModelSetupFusedNoiseMixin
from ..modules.modelSetup.mixin.ModelSetupNoiseMixin import ModelSetupNoiseMixin
from ..modules.util.optimizer.adamw_extensions import patch_adamw
from . import standin


class TrainProgress:
    def __init__(self):
        self.global_step = 0

    def next_step(self, batch_size: int):
        self.global_step += 1


class StandInModel:
    """The fields of the reference's BaseModel / StableDiffusionXLModel the step touches."""

    def __init__(self, unet: nn.Module, scheduler: standin.StandInScheduler, vae_scaling_factor: float,
                 train_dtype: torch.dtype, sdxl: bool):
        self.unet = unet
        self.noise_scheduler = scheduler
        self.vae_scaling_factor = vae_scaling_factor
        self.train_dtype = train_dtype
        self.sdxl = sdxl
        self.unet_lora: LoRAModuleWrapper | None = None
        self.full_grad_arena: dp.FlatGradArena | None = None      # full fine-tuning: persistent flat gradients
        self.grad_reducer: dp.OverlappedArenaReducer | None = None  # ... and their allreduce, overlapped with backward
        self.optimizer = None
        self.ema: EMAModuleWrapper | None = None
        self.parameters: list[nn.Parameter] = []
        self.train_progress = TrainProgress()
        self.autocast_context = torch.autocast(device_type="cuda", dtype=train_dtype)


class DiffusionSetup(ModelSetupFusedNoiseMixin, ModelSetupNoiseMixin, ModelSetupDiffusionMixin, ModelSetupDiffusionLossMixin):
    """SD 1.5 / SDXL flavoured setup (epsilon or v-prediction, DDPM schedule)."""

    def __init__(self, train_device: torch.device, fused_noise: bool = True):
        super().__init__()
        self.train_device = train_device
        self.fused_noise = fused_noise

    def predict(self, model: StandInModel, batch: dict, config, train_progress: TrainProgress, *, deterministic: bool = False) -> dict:
        with model.autocast_context:
            batch_seed = 0 if deterministic else train_progress.global_step
            if step_state.active() is not None:      # inside a CUDA-graph capture torch cannot build a CUDA generator
                generator = PhiloxKey(batch_seed, config.train_device)
            else:
                generator = torch.Generator(device=config.train_device)
                generator.manual_seed(batch_seed)
            latent_image = batch["latent_image"]
            T = model.noise_scheduler.config["num_train_timesteps"]
            ptype = model.noise_scheduler.config["prediction_type"]
            if self.fused_noise:
                st = self._noise_timestep_addnoise(
                    latent_image, config, generator, num_train_timesteps=T, deterministic=deterministic,
                    prediction_type=ptype, betas=model.noise_scheduler.betas, latent_scale=model.vae_scaling_factor,
                    train_dtype=model.train_dtype)
                timestep, latent_input, target = st["timestep"], st["noisy"], st["target"]
            else:   # the reference's call-by-call sequence, each call one launch
                scaled_latent_image = latent_image * model.vae_scaling_factor
                latent_noise = self._create_noise(scaled_latent_image, config, generator)
                timestep = self._get_timestep_discrete(T, deterministic, generator, scaled_latent_image.shape[0], config)
                noisy = self._add_noise_discrete(scaled_latent_image, latent_noise, timestep, model.noise_scheduler.betas)
                latent_input = noisy.to(dtype=model.train_dtype)
                target = latent_noise if ptype == "epsilon" else model.noise_scheduler.get_velocity(
                    scaled_latent_image, latent_noise, timestep)
            kwargs = {}
            if model.sdxl:
                add_time_ids = torch.stack([
                    batch["original_resolution"][0], batch["original_resolution"][1],
                    batch["crop_offset"][0], batch["crop_offset"][1],
                    batch["crop_resolution"][0], batch["crop_resolution"][1]], dim=1).to(dtype=latent_input.dtype, device=latent_input.device)
                kwargs["added_cond_kwargs"] = {"text_embeds": batch["text_encoder_2_pooled_state"], "time_ids": add_time_ids}
            predicted = model.unet(latent_input, timestep, batch["text_encoder_hidden_state"].to(dtype=model.train_dtype), **kwargs).sample
            data = {"loss_type": "target", "timestep": timestep, "predicted": predicted, "target": target}
        data["prediction_type"] = ptype
        return data

    def calculate_loss(self, model: StandInModel, batch: dict, data: dict, config) -> torch.Tensor:
        return self._diffusion_losses(batch=batch, data=data, config=config, train_device=self.train_device,
                                      betas=model.noise_scheduler.betas.to(device=self.train_device)).mean()


class StandInFluxModel:
    """The fields of the reference's FluxModel the step touches (modules/model/FluxModel.py)."""

    def __init__(self, transformer: nn.Module, scheduler, train_dtype: torch.dtype):
        self.transformer = transformer
        self.unet = transformer                       # the harness' generic name for "the wrapped network"
        self.noise_scheduler = scheduler
        self.vae_scaling_factor, self.vae_shift_factor = 0.3611, 0.1159      # FLUX.1 VAE config
        self.train_dtype = train_dtype
        self.sdxl = False
        self.unet_lora: LoRAModuleWrapper | None = None
        self.transformer_lora = None
        self.optimizer = None
        self.ema: EMAModuleWrapper | None = None
        self.parameters: list[nn.Parameter] = []
        self.train_progress = TrainProgress()
        self.autocast_context = torch.autocast(device_type="cuda", dtype=train_dtype)

    # FluxModel.pack_latents / unpack_latents, modules/model/FluxModel.py:319-344 (pure views + one copy each)
    @staticmethod
    def pack_latents(latents, batch_size, num_channels_latents, height, width):
        latents = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
        latents = latents.permute(0, 2, 4, 1, 3, 5)
        return latents.reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)

    @staticmethod
    def unpack_latents(latents, height, width):
        batch_size, num_patches, channels = latents.shape
        height, width = height // 2, width // 2
        latents = latents.view(batch_size, height, width, channels // 4, 2, 2)
        latents = latents.permute(0, 3, 1, 4, 2, 5)
        return latents.reshape(batch_size, channels // 4, height * 2, width * 2)


class FlowSetup(ModelSetupFusedNoiseMixin, ModelSetupNoiseMixin, ModelSetupFlowMatchingMixin, ModelSetupDiffusionLossMixin):
    """FLUX-flavoured setup (rectified flow): restates BaseFluxSetup.predict :202-313 and calculate_loss :377-390
    with cached text-encoder outputs in the batch."""

    def __init__(self, train_device: torch.device, fused_noise: bool = True, guidance_scale: float = 1.0, fold_pack: bool = True):
        super().__init__()
        self.train_device = train_device
        self.fused_noise = fused_noise
        self.guidance_scale = guidance_scale
        # K-b1 writes the network input in packed layout and K-b2 reads the packed prediction, so
        # pack_latents / unpack_latents (and the unpack's backward) disappear from the step
        self.fold_pack = fold_pack and fused_noise

    def predict(self, model: StandInFluxModel, batch: dict, config, train_progress: TrainProgress, *, deterministic: bool = False) -> dict:
        with model.autocast_context:
            batch_seed = 0 if deterministic else train_progress.global_step
            if step_state.active() is not None:
                generator = PhiloxKey(batch_seed, config.train_device)
            else:
                generator = torch.Generator(device=config.train_device)
                generator.manual_seed(batch_seed)
            latent_image = batch["latent_image"]
            T = model.noise_scheduler.config["num_train_timesteps"]
            H, W = latent_image.shape[-2], latent_image.shape[-1]
            if self.fused_noise:
                st = self._noise_timestep_addnoise(
                    latent_image, config, generator, num_train_timesteps=T, deterministic=deterministic, prediction_type="flow",
                    flow=True, latent_scale=model.vae_scaling_factor, latent_shift=model.vae_shift_factor,
                    train_dtype=model.train_dtype, latent_height=H, latent_width=W, pack_latents=self.fold_pack)
                timestep, latent_input, flow = st["timestep"], st["noisy"], st["target"]
            else:   # the reference's call-by-call sequence
                scaled_latent_image = (latent_image - model.vae_shift_factor) * model.vae_scaling_factor
                latent_noise = self._create_noise(scaled_latent_image, config, generator)
                timestep = self._get_timestep_discrete(T, deterministic, generator, scaled_latent_image.shape[0], config,
                                                       latent_height=H, latent_width=W)
                noisy, _sigma = self._add_noise_discrete(scaled_latent_image, latent_noise, timestep, model.noise_scheduler.timesteps)
                latent_input = noisy.to(dtype=model.train_dtype)
                flow = latent_noise - scaled_latent_image
            B, C = latent_image.shape[0], latent_image.shape[1]
            guidance = None
            if model.transformer.config["guidance_embeds"]:
                guidance = torch.full((B,), float(self.guidance_scale), device=latent_input.device, dtype=model.train_dtype)
            packed = latent_input if self.fold_pack else model.pack_latents(latent_input, B, C, H, W)
            packed_predicted_flow = model.transformer(
                hidden_states=packed.to(dtype=model.train_dtype), timestep=timestep / 1000, guidance=guidance,
                pooled_projections=batch["text_encoder_1_pooled_state"].to(dtype=model.train_dtype),
                encoder_hidden_states=batch["text_encoder_2_hidden_state"].to(dtype=model.train_dtype),
                txt_ids=None, img_ids=None, joint_attention_kwargs=None, return_dict=True).sample
            if self.fold_pack:
                data = {"loss_type": "target", "timestep": timestep, "predicted": packed_predicted_flow, "target": flow,
                        "predicted_packed_width": W}
            else:
                predicted_flow = model.unpack_latents(packed_predicted_flow, H, W)
                data = {"loss_type": "target", "timestep": timestep, "predicted": predicted_flow, "target": flow}
        return data

    def calculate_loss(self, model: StandInFluxModel, batch: dict, data: dict, config) -> torch.Tensor:
        return self._flow_matching_losses(batch=batch, data=data, config=config, train_device=self.train_device,
                                          sigmas=model.noise_scheduler.sigmas.to(device=self.train_device)).mean()


def build_flux_model(kind: str, config, device, *, train_dtype=torch.bfloat16, lora_weight_dtype=torch.float32,
                     stochastic_rounding: bool = True, fused_clip: bool = True, seed: int = 1234, lr: float = 3e-4,
                     lora_up_std: float = 0.0) -> tuple[StandInFluxModel, FlowSetup]:
    """Random-init FLUX stand-in + LoRA wrapper (`attn-mlp` preset) + patched AdamW, as FluxLoRASetup.setup_model :105-161."""
    torch.manual_seed(seed)
    with torch.device(device):
        net = {"flux": standin.flux_transformer, "tiny_flux": standin.tiny_flux}[kind]()
    net = net.to(dtype=train_dtype)
    net.requires_grad_(False)
    net.train()
    model = StandInFluxModel(net, standin.StandInFlowScheduler(1000, device), train_dtype)
    GradArena.allocator = dp.symmetric_zeros if dp.peer_exchange_wanted() else None    # K-r needs a peer-mapped arena
    model.unet_lora = model.transformer_lora = LoRAModuleWrapper(net, "lora_transformer", config, ["attn", "ff.net"])
    model.unet_lora.set_dropout(config.dropout_probability)
    model.unet_lora.to(dtype=lora_weight_dtype)
    if lora_up_std > 0:
        g = torch.Generator(device=device).manual_seed(seed + 1)
        for m in model.unet_lora.lora_modules.values():
            m.lora_up.weight.data.normal_(0.0, lora_up_std, generator=g)
    model.unet_lora.hook_to_module()
    model.unet_lora.requires_grad_(True)
    model.parameters = model.unet_lora.parameters()
    model.optimizer = torch.optim.AdamW(model.parameters, lr=lr, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    patch_adamw(model.optimizer, stochastic_rounding, clip_grad_norm=config.clip_grad_norm if fused_clip else None)
    return model, FlowSetup(torch.device(device))


def build_model(kind: str, config, device, *, train_dtype=torch.bfloat16, lora_weight_dtype=torch.float32,
                prediction_type: str = "epsilon", ema: bool = False, stochastic_rounding: bool = True,
                fused_clip: bool = True, fused_ema: bool = True, seed: int = 1234, lr: float = 3e-4,
                lora_up_std: float = 0.0, full_finetune: bool = False) -> tuple[StandInModel, DiffusionSetup]:
    """Random-init stand-in + LoRA wrapper + patched AdamW (+EMA), as StableDiffusionXLLoRASetup.setup_model
    (:103-158) and init_model_parameters (optimizer_util.py:50-65) wire them.  `full_finetune=True` is config C4
    (StableDiffusionXLFineTuneSetup.setup_model :90-113): no adapters, every UNet tensor trainable in the train
    dtype, AdamW state and EMA in that dtype too — the path is then K-b1, K-b2 and the multi-tensor SR-AdamW + EMA."""
    torch.manual_seed(seed)
    with torch.device(device):
        unet = {"sdxl": standin.sdxl_unet, "sd15": standin.sd15_unet, "tiny": standin.tiny_unet}[kind]()
    unet = unet.to(dtype=train_dtype)
    unet.requires_grad_(False)
    unet.train()
    sched = standin.StandInScheduler(standin_betas().to(device), prediction_type)   # device-resident: `.to()` is then a no-op under capture
    model = StandInModel(unet, sched, 0.13025 if kind == "sdxl" else 0.18215, train_dtype, sdxl=(kind == "sdxl"))
    if full_finetune:
        unet.requires_grad_(True)
        model.parameters = list(unet.parameters())
        opt_params = model.parameters
        if dp.sharded_optimizer_wanted():
            # K-z: parameters and gradients in symmetric arenas, every bucket owned by one rank, optimizer state and EMA
            # only for the owner's tensors (reduce-scatter -> sharded K-c -> all-gather)
            model.full_grad_arena = model.grad_reducer = dp.ShardedExchange(model.parameters)
            opt_params = model.grad_reducer.owned_parameters()
        else:
            model.full_grad_arena = dp.FlatGradArena(model.parameters)
            model.grad_reducer = dp.OverlappedArenaReducer(model.full_grad_arena) if dp.world_size() > 1 else None
        model.optimizer_parameters = opt_params
        model.optimizer = torch.optim.AdamW(opt_params, lr=lr, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
        if ema:
            model.ema = EMAModuleWrapper(opt_params, decay=config.ema_decay,
                                         update_step_interval=config.ema_update_step_interval, device=torch.device(device))
        progress = model.train_progress
        patch_adamw(model.optimizer, stochastic_rounding,
                    ema=model.ema if (ema and fused_ema) else None,
                    ema_step_fn=(lambda: progress.global_step // config.gradient_accumulation_steps) if (ema and fused_ema) else None,
                    clip_grad_norm=config.clip_grad_norm if fused_clip else None)
        if isinstance(model.grad_reducer, dp.ShardedExchange):
            model.optimizer.otb_driver.rng_base_of = model.grad_reducer.rng_base_of
            if model.grad_reducer.peer_store is not None:      # the all-gather rides on K-c's parameter stores
                model.optimizer.otb_driver.peer_store = model.grad_reducer.peer_store
                model.grad_reducer.gather_is_fused = True
        return model, DiffusionSetup(torch.device(device))
    GradArena.allocator = dp.symmetric_zeros if dp.peer_exchange_wanted() else None    # K-r needs a peer-mapped arena
    model.unet_lora = LoRAModuleWrapper(unet, "lora_unet", config, ["attentions"])
    model.unet_lora.set_dropout(config.dropout_probability)
    model.unet_lora.to(dtype=lora_weight_dtype)
    if lora_up_std > 0:
        g = torch.Generator(device=device).manual_seed(seed + 1)
        for m in model.unet_lora.lora_modules.values():
            m.lora_up.weight.data.normal_(0.0, lora_up_std, generator=g)
    model.unet_lora.hook_to_module()
    model.unet_lora.requires_grad_(True)
    model.parameters = model.unet_lora.parameters()
    model.optimizer = torch.optim.AdamW(model.parameters, lr=lr, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    if ema:
        model.ema = EMAModuleWrapper(model.parameters, decay=config.ema_decay,
                                     update_step_interval=config.ema_update_step_interval, device=torch.device(device))
    progress = model.train_progress
    patch_adamw(model.optimizer, stochastic_rounding,
                ema=model.ema if (ema and fused_ema) else None,
                ema_step_fn=(lambda: progress.global_step // config.gradient_accumulation_steps) if (ema and fused_ema) else None,
                clip_grad_norm=config.clip_grad_norm if fused_clip else None)
    return model, DiffusionSetup(torch.device(device))


def standin_betas():
    from ..modules.util.DiffusionScheduleCoefficients import scaled_linear_betas
    return scaled_linear_betas()


def synthetic_batch(kind: str, batch_size: int, device, seed: int = 1234, latent_hw: int | None = None, masked: bool = False,
                    rank: int | None = None) -> dict:
    """Seeded synthetic batch with the dataloader's keys.  `rank` selects another rank's batch
    (the data-parallel self-check rebuilds the global batch on one GPU)."""
    g = torch.Generator(device="cpu").manual_seed(seed + (dp.rank() if rank is None else rank))
    if kind in ("flux", "tiny_flux"):          # FLUX.1: 16-channel latents, T5 states [B,512,4096], CLIP pooled [B,768]
        hw = latent_hw or (128 if kind == "flux" else 16)
        ctx, pooled, n_txt = (4096, 768, 512) if kind == "flux" else (64, 32, 8)
        return {
            "latent_image": torch.randn(batch_size, 16, hw, hw, generator=g) * 0.8,
            "loss_weight": torch.ones(batch_size),
            "text_encoder_2_hidden_state": torch.randn(batch_size, n_txt, ctx, generator=g),
            "text_encoder_1_pooled_state": torch.randn(batch_size, pooled, generator=g),
        }
    hw = latent_hw or (128 if kind == "sdxl" else 64)
    ctx = {"sdxl": 2048, "sd15": 768, "tiny": 64}[kind]
    b = {
        "latent_image": torch.randn(batch_size, 4, hw, hw, generator=g) * 0.8,
        "loss_weight": torch.ones(batch_size),
        "text_encoder_hidden_state": torch.randn(batch_size, 77, ctx, generator=g),
    }
    if masked:
        blocks = (torch.rand(batch_size, 1, hw // 8, hw // 8, generator=g) > 0.5).float()
        b["latent_mask"] = blocks.repeat_interleave(8, dim=2).repeat_interleave(8, dim=3)
    if kind == "sdxl":
        b["text_encoder_2_pooled_state"] = torch.randn(batch_size, 1280, generator=g)
        px = float(hw * 8)
        b["original_resolution"] = [torch.full((batch_size,), px), torch.full((batch_size,), px)]
        b["crop_offset"] = [torch.zeros(batch_size), torch.zeros(batch_size)]
        b["crop_resolution"] = [torch.full((batch_size,), px), torch.full((batch_size,), px)]
    return b


def batch_to_device(batch: dict, device, non_blocking: bool = True) -> dict:
    out = {}
    for k, v in batch.items():
        if isinstance(v, list):
            out[k] = [t.to(device, non_blocking=non_blocking) for t in v]
        else:
            out[k] = v.to(device, non_blocking=non_blocking)
    return out


def train_step(setup: DiffusionSetup, model: StandInModel, batch: dict, config, bucket: dp.GradientBucket | None = None) -> torch.Tensor:
    """One optimisation step (gradient accumulation 1), GenericTrainer.py:683-747.  Returns the loss as a
    DEVICE scalar — the reference's `loss.item()` every micro-step (:699) is left to the caller."""
    data = setup.predict(model, batch, config, model.train_progress)
    loss = setup.calculate_loss(model, batch, data, config)
    loss = loss / config.gradient_accumulation_steps
    reducer = getattr(model, "grad_reducer", None)
    if reducer is not None:
        reducer.arm()
    loss.backward()
    driver = model.optimizer.otb_driver
    if reducer is not None:
        driver.grad_prescale = reducer.finish()
        driver.norm_partials = getattr(reducer, "norm_partials", None)     # K-z: the reduce already accumulated the global norm
    elif dp.world_size() > 1:
        arena = model.unet_lora.grad_arena if model.unet_lora is not None else getattr(model, "full_grad_arena", None)
        bucket = bucket or dp.GradientBucket(arena.flat if isinstance(arena, dp.FlatGradArena) else arena)
        driver.grad_prescale = bucket.allreduce([p.grad for p in model.parameters if p.grad is not None])
        driver.norm_partials = bucket.norm_partials
    if config.clip_grad_norm is not None and driver.clip_max_norm is None:
        from ..modules.util.grad_clip import clip_grad_norm_
        clip_grad_norm_(model.parameters, config.clip_grad_norm, driver.grad_prescale)
    model.optimizer.step()
    if hasattr(reducer, "gather"):
        reducer.gather()                                                    # K-z: the owners' updated ranges -> every replica
    _zero_grads(model)
    if model.ema is not None:
        model.ema.step(getattr(model, "optimizer_parameters", None) or model.parameters,
                       model.train_progress.global_step // config.gradient_accumulation_steps)
    model.train_progress.next_step(config.batch_size)
    return loss.detach()


def _zero_grads(model) -> None:
    """`optimizer.zero_grad(set_to_none=True)` (GenericTrainer.py:717), or one memset when the gradients are the
    persistent views of a FlatGradArena."""
    arena = getattr(model, "full_grad_arena", None)
    if arena is not None:
        arena.zero()
    else:
        model.optimizer.zero_grad(set_to_none=True)


class GraphedTrainStep:
    """The whole step as ONE CUDA graph (streams and graphs instead of a tracing compiler).

    The step launches ~11 k kernels (722 leaves x 3 fused launches, plus the stand-in's attention /
    norm / GEGLU glue); eager python dispatch of those costs more than the GPU work.  Capturing the
    step once removes the host from the loop: per step the host only (1) copies the next batch into
    pinned staging buffers, (2) prepares a few words — Philox keys, EMA schedule, per-tensor AdamW scalars
    (onetrainer_b200/step_state.py, FusedAdamDriver.host_prepare) — and uploads them with two small async copies
    from ring-buffered pinned slots, and (3) replays.  The H2D batch copy, the noise/loss kernels, every leaf,
    the NCCL allreduce, grad-norm, AdamW and the shadow refresh are all inside the graph, on one stream.
    """

    def __init__(self, setup: DiffusionSetup, model: StandInModel, config, example_batch: dict, warmup: int = 3, *,
                 pool=None, state=None):
        """`pool` / `state`: a graph memory pool and a DeviceStepState shared with other captures of the same model
        (BucketedGraphedTrainStep: one capture per batch shape, never replayed concurrently).  `warmup=0` captures straight
        away — only after an eager step on this batch shape has initialised the shape-dependent library plans."""
        from .. import step_state
        self.setup, self.model, self.config = setup, model, config
        dev = setup.train_device
        self.state = state if state is not None else step_state.DeviceStepState(dev)
        self.host_batch = {k: ([torch.empty_like(t).pin_memory() for t in v] if isinstance(v, list) else torch.empty_like(v).pin_memory())
                           for k, v in example_batch.items()}
        self.dev_batch = {k: ([torch.empty_like(t, device=dev) for t in v] if isinstance(v, list) else torch.empty_like(v, device=dev))
                          for k, v in example_batch.items()}
        self.bucket = dp.GradientBucket(None)
        self.driver = model.optimizer.otb_driver
        self.loss = None
        self.graph = None
        self._replayed = None
        self.set_batch(example_batch)
        step_state.activate(self.state)
        try:
            self._warm_up_and_capture(warmup, pool)
        except BaseException:
            step_state.activate(None)            # a failed capture must not leave its device step state active
            raise

    def _warm_up_and_capture(self, warmup: int, pool) -> None:
        from .. import step_state
        dev = self.setup.train_device
        model = self.model
        # warm-up on a side stream (allocator, lazy tables, cuBLAS/SDPA workspaces) as torch's capture rules require
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._device_part(first=True)
                self._after_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.driver.freeze_for_graph()
        self._frozen_params = [p for p, _ in self.driver._frozen_active]
        self._frozen_grad_ptrs = [int(v) for v in self.driver.table.host["g"]]
        self._host_part()
        from .. import ops
        before = ops.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self._device_part(first=False)
        self.launches_per_step = ops.LAUNCHES - before      # libotb200 kernels inside one replay
        # capture records but does not execute; the host side of this step (Philox keys, AdamW step counters and
        # bias corrections) has been prepared, so run it now — otherwise the step would be counted but never taken
        self.graph.replay()
        self._after_step()

    # host-side work that must precede every replay
    def _host_part(self) -> None:
        step = self.model.train_progress.global_step
        self.state.set_noise_key(step, 0)           # predict(): generator.manual_seed(global_step), first draw
        self.driver.host_prepare()
        # this step's words and per-tensor AdamW scalars go to the device now, on the stream the replay follows on
        # (ring-buffered pinned sources: the host may be a step ahead of the GPU)
        self.state.upload()
        self.driver.upload_table()

    def _device_part(self, first: bool) -> None:
        if first:                                   # eager warm-up: the optimizer's host part runs after backward (below)
            self.state.set_noise_key(self.model.train_progress.global_step, 0)
            self.state.upload()
        for k, v in self.host_batch.items():
            if isinstance(v, list):
                for d, h in zip(self.dev_batch[k], v):
                    d.copy_(h, non_blocking=True)
            else:
                self.dev_batch[k].copy_(v, non_blocking=True)
        model, cfg = self.model, self.config
        data = self.setup.predict(model, self.dev_batch, cfg, model.train_progress)
        loss = self.setup.calculate_loss(model, self.dev_batch, data, cfg) / cfg.gradient_accumulation_steps
        reducer = getattr(model, "grad_reducer", None)
        if reducer is not None:
            reducer.arm()
        loss.backward()
        if reducer is not None:
            self.driver.grad_prescale = reducer.finish()
            self.driver.norm_partials = getattr(reducer, "norm_partials", None)
        elif dp.world_size() > 1:
            if self.bucket.arena is None:
                self.bucket.arena = model.unet_lora.grad_arena if model.unet_lora is not None else model.full_grad_arena.flat
            self.driver.grad_prescale = self.bucket.allreduce([p.grad for p in model.parameters if p.grad is not None])
            self.driver.norm_partials = self.bucket.norm_partials
        if first:
            self.driver.host_prepare()              # discovers the gradients, builds the table, fills the scalars
            self.state.upload()                     # the optimizer's dynamic words were written just now
        if not first:
            # the captured optimizer launch reads gradients through a pointer table frozen after the warm-up steps: refuse
            # to capture (nothing has executed yet) if autograd handed out gradient tensors at other addresses
            moved = [p for p, g in zip(self._frozen_params, self._frozen_grad_ptrs) if p.grad is None or p.grad.data_ptr() != g]
            if moved:
                raise RuntimeError(f"GraphedTrainStep: {len(moved)} gradient tensors do not live at stable addresses; give the "
                                   "parameters persistent gradients (LoRA: the wrapper's arena; full fine-tuning: dp.FlatGradArena)")
        self.driver.device_step(upload=first)       # captured: the table was uploaded by _host_part()
        if hasattr(reducer, "gather"):
            reducer.gather()
        _zero_grads(model)
        self.loss = loss.detach()

    def _after_step(self) -> None:
        self.model.train_progress.next_step(self.config.batch_size)

    def set_batch(self, batch: dict) -> None:
        """Stage the next batch in pinned memory (host memcpy); the H2D copy happens inside the graph.  The previous
        replay must have consumed the staging buffers first."""
        if self._replayed is not None:
            self._replayed.synchronize()
        for k, v in batch.items():
            if isinstance(v, list):
                for h, t in zip(self.host_batch[k], v):
                    h.copy_(t)
            else:
                self.host_batch[k].copy_(v)

    def step(self, batch: dict | None = None) -> torch.Tensor:
        """One optimisation step; returns the loss as a device scalar (read it with .item() when needed)."""
        if batch is not None:
            self.set_batch(batch)
        from .. import step_state
        step_state.activate(self.state)
        self._host_part()
        self.graph.replay()
        if self._replayed is None:
            self._replayed = torch.cuda.Event()
        self._replayed.record()
        self._after_step()
        return self.loss

    def close(self) -> None:
        from .. import step_state
        step_state.activate(None)


class BucketedGraphedTrainStep:
    """One captured step per BATCH SHAPE.  The reference trains aspect-ratio buckets — the latent height x width changes
    from step to step (modules/dataLoader/mixin/DataLoaderText2ImageMixin.py: aspect bucketing) — and a CUDA graph is
    bound to one shape.  The first batch of a shape takes the eager step (which also initialises the shape-dependent
    library plans); the second one captures the step for that shape and is stepped by the capture's first replay; later
    ones replay.  Every batch is stepped exactly once.  All captures share one graph memory pool (they never replay
    concurrently, so the activations of all buckets occupy the memory of the largest one) and one DeviceStepState.
    `max_graphs` bounds the cache (least recently used capture dropped)."""

    def __init__(self, setup: DiffusionSetup, model: StandInModel, config, max_graphs: int = 16):
        from .. import step_state
        self.setup, self.model, self.config = setup, model, config
        self.state = step_state.DeviceStepState(setup.train_device)
        self.max_graphs = max_graphs
        self._steps: dict[tuple, GraphedTrainStep] = {}
        self._seen: set = set()
        self._pool = None
        self._bucket = dp.GradientBucket(None)
        self._side = None
        self.captures = 0
        self.eager_steps = 0

    @staticmethod
    def shape_key(batch: dict) -> tuple:
        return tuple(sorted((k, tuple(tuple(t.shape) for t in v) if isinstance(v, list) else tuple(v.shape)) for k, v in batch.items()))

    def step(self, batch: dict) -> torch.Tensor:
        from .. import step_state
        key = self.shape_key(batch)
        st = self._steps.pop(key, None)
        if st is not None:
            self._steps[key] = st                       # most recently used last
            return st.step(batch)
        if key not in self._seen:                       # first batch of this shape: the eager step
            self._seen.add(key)
            step_state.activate(None)
            self.eager_steps += 1
            # on a side stream, like the warm-up steps of a capture: autograd remembers the stream a leaf's AccumulateGrad
            # node first ran on, and a later capture must not be made to synchronise with the legacy default stream
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.setup.train_device)
            cur = torch.cuda.current_stream(self.setup.train_device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                loss = train_step(self.setup, self.model, batch_to_device(batch, self.setup.train_device), self.config, self._bucket)
            cur.wait_stream(self._side)
            return loss
        while len(self._steps) >= self.max_graphs:
            old = self._steps.pop(next(iter(self._steps)))
            old.graph = None
        # second batch of this shape: capture; the capture's first replay is this batch's step
        st = GraphedTrainStep(self.setup, self.model, self.config, batch, warmup=0, pool=self._pool, state=self.state)
        if self._pool is None:
            self._pool = st.graph.pool()
        self._steps[key] = st
        self.captures += 1
        return st.loss

    def close(self) -> None:
        from .. import step_state
        self._steps.clear()
        step_state.activate(None)
