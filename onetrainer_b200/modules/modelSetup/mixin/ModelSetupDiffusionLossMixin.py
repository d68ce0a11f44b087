"""Diffusion / flow-matching training losses on the B200 kernel — drop-in for
/root/reference/modules/modelSetup/mixin/ModelSetupDiffusionLossMixin.py:233-321.

`_diffusion_losses` / `_flow_matching_losses` keep the reference's signatures and return the same
fp32 [B] tensor with autograd history to `data['predicted']`.  Underneath, the reference's ~20
forward and ~20 backward launches (casts, mse_loss, mask clamp/multiply/mean, per-sample mean,
scaling, SNR weights, and their autograd mirror images — with >= 4 full-size fp32 temporaries) are
ONE launch of otb_loss_fwd_bwd (K-b2): it produces the per-sample losses and d(loss)/d(predicted)
in the same pass, assuming the upstream gradient the trainer's `.mean()` -> `/GA` -> `.backward()`
chain produces (GenericTrainer.py:690-696).  `backward` only launches a fix-up kernel that exits
without touching memory when that assumption holds.

Not implemented on this path (raise, never emulate): log-cosh loss, the learned-variance VB loss
(PixArt), continuous-time SNR (`alphas_cumprod_fun`, Würstchen) — SURVEY.md §2.1 marks them out of scope.
"""
from __future__ import annotations

from abc import ABCMeta
from collections.abc import Callable

import torch
from torch import Tensor

from .... import ops
from .... import _native as N
from ...util.DiffusionScheduleCoefficients import DiffusionScheduleCoefficients, flow_sigmas
from ...util.enum import enum_name

_LW = {"CONSTANT": N.LW_CONSTANT, "MIN_SNR_GAMMA": N.LW_MIN_SNR, "DEBIASED_ESTIMATION": N.LW_DEBIASED,
       "P2": N.LW_P2, "SIGMA": N.LW_SIGMA}


class _FusedLoss(torch.autograd.Function):
    """losses[B] = f(predicted); the gradient is computed in the forward launch."""

    @staticmethod
    def forward(ctx, predicted: Tensor, kernel_kwargs: dict):
        want_grad = ctx.needs_input_grad[0]      # grad mode is off inside Function.forward: ask autograd instead
        losses, dpred = ops.loss_fwd_bwd(predicted.detach(), want_grad=want_grad, **kernel_kwargs)
        ctx.assumed = kernel_kwargs["upstream"]
        ctx.dpred = dpred
        return losses

    @staticmethod
    def backward(ctx, grad_losses: Tensor):
        dpred = ctx.dpred
        if dpred is None:
            raise RuntimeError("loss backward called but no gradient was requested in forward")
        ops.loss_bwd_rescale(dpred, grad_losses, ctx.assumed)
        ctx.dpred = None
        return dpred, None


class ModelSetupDiffusionLossMixin(metaclass=ABCMeta):
    __coefficients: DiffusionScheduleCoefficients | None
    __sigmas: Tensor | None

    def __init__(self):
        super().__init__()
        self.__coefficients = None
        self.__alphas_cumprod_fun = None
        self.__sigmas = None

    # ---- shared body ---------------------------------------------------------------------
    def __losses(self, batch: dict, data: dict, config, lw_fn: int, tables: dict) -> Tensor:
        if data.get("loss_type") != "target":
            raise NotImplementedError(f"loss_type {data.get('loss_type')!r} is not supported")
        if getattr(config, "log_cosh_strength", 0) != 0:
            raise NotImplementedError("log-cosh loss is outside the B200 hot path (SURVEY.md §2.1)")
        if getattr(config, "vb_loss_strength", 0) != 0 and "predicted_var_values" in data:
            raise NotImplementedError("learned-variance VB loss (PixArt) is outside the B200 hot path (SURVEY.md §2.1)")
        predicted, target = data["predicted"], data["target"]
        B = predicted.shape[0]
        # LossScaler (ModelSetupDiffusionLossMixin.py:243-248)
        scaler = enum_name(config.loss_scaler)
        # "batch" is the GLOBAL batch under data parallelism (as in NamedParameterGroup's LR scaler), so that an N-rank run
        # at per-rank batch b reproduces the 1-rank run at batch N*b
        from .... import dp
        batch_scale = 1 if scaler in ("NONE", "GRADIENT_ACCUMULATION") else config.batch_size * dp.world_size()
        ga_scale = 1 if scaler in ("NONE", "BATCH") else config.gradient_accumulation_steps
        masked = bool(config.masked_training) and not config.model_type.has_conditioning_image_input()
        use_mask_mean_only = (not masked) and bool(config.masked_training) and bool(config.normalize_masked_area_loss)
        mask = batch["latent_mask"] if (masked or use_mask_mean_only) else None
        kw = dict(
            target=target, prior_target=data.get("prior_target") if masked else None, mask=mask,
            loss_weight=batch["loss_weight"], timestep=data.get("timestep"),
            masked=masked, normalize_masked_area_loss=bool(config.normalize_masked_area_loss) and mask is not None,
            loss_weight_fn=lw_fn if "timestep" in data else N.LW_CONSTANT,
            v_prediction=data.get("prediction_type", "") == "v_prediction",
            mse_strength=config.mse_strength, mae_strength=config.mae_strength,
            unmasked_weight=config.unmasked_weight,
            masked_prior_preservation_weight=config.masked_prior_preservation_weight,
            loss_scale=float(batch_scale * ga_scale), loss_weight_strength=config.loss_weight_strength,
            # what `.mean()` over B then `/ gradient_accumulation_steps` hands back to us
            upstream=1.0 / (B * config.gradient_accumulation_steps),
            # opt-in: data['predicted'] is the transformer's PACKED output (FluxModel.unpack_latents skipped)
            pred_pack_w=int(data.get("predicted_packed_width", 0)), **tables)
        return _FusedLoss.apply(predicted, kw)

    # ---- reference API -------------------------------------------------------------------
    def _diffusion_losses(self, batch: dict, data: dict, config, train_device: torch.device, betas: Tensor | None = None,
                          alphas_cumprod_fun: Callable[[Tensor, int], Tensor] | None = None) -> Tensor:
        if self.__coefficients is None and betas is not None:
            self.__coefficients = DiffusionScheduleCoefficients.from_betas(betas.to(device=train_device))
        self.__alphas_cumprod_fun = alphas_cumprod_fun
        name = enum_name(config.loss_weight_fn)
        lw_fn = _LW[name] if name in ("MIN_SNR_GAMMA", "DEBIASED_ESTIMATION", "P2") else N.LW_CONSTANT
        tables = {}
        if lw_fn != N.LW_CONSTANT:
            if self.__coefficients is None:
                raise NotImplementedError(
                    "SNR loss weighting from a continuous alphas_cumprod_fun (Würstchen) is outside the B200 hot path")
            tables = dict(sqrt_ac=self.__coefficients.sqrt_alphas_cumprod,
                          sqrt_1mac=self.__coefficients.sqrt_one_minus_alphas_cumprod,
                          num_train_timesteps=self.__coefficients.num_timesteps)
        return self.__losses(batch, data, config, lw_fn, tables)

    def _flow_matching_losses(self, batch: dict, data: dict, config, train_device: torch.device,
                              sigmas: Tensor | None = None) -> Tensor:
        if self.__sigmas is None and sigmas is not None:
            self.__sigmas, _ = flow_sigmas(sigmas.shape[0], sigmas.device)
        lw_fn = N.LW_SIGMA if enum_name(config.loss_weight_fn) == "SIGMA" else N.LW_CONSTANT
        tables = {}
        if lw_fn == N.LW_SIGMA:
            if self.__sigmas is None:
                raise RuntimeError("SIGMA loss weighting needs the scheduler's sigmas")
            tables = dict(sigmas=self.__sigmas.to(device=train_device), num_train_timesteps=self.__sigmas.shape[0])
        return self.__losses(batch, data, config, lw_fn, tables)
