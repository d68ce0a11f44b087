"""One-launch pre-network step (K-b1) — an ADDITION to the reference's mixin API.

The reference's family `predict()` bodies call, in sequence: latent scaling, `_create_noise`,
`_get_timestep_discrete`, `_add_noise_discrete`, a `.to(train_dtype)` cast and (v-prediction)
`scheduler.get_velocity` (BaseStableDiffusionXLSetup.py:214-236,269,285; BaseFluxSetup.py:230-313).
Those stay available one-to-one on the other mixins.  A `predict()` that wants the whole stage in
ONE memory-bound launch calls `_noise_timestep_addnoise` instead; results are bit-identical to the
call-by-call sequence on the same generator state (same Philox key, same fp32 op order).
"""
from __future__ import annotations

from abc import ABCMeta

import torch
from torch import Generator, Tensor

from .... import dp, ops, step_state
from .... import _native as N
from ....rng import draw_key
from ....util_shapes import flatten_chw
from ...util.DiffusionScheduleCoefficients import DiffusionScheduleCoefficients, flow_sigmas
from .ModelSetupNoiseMixin import timestep_kernel_args

_TARGETS = {"epsilon": N.TARGET_EPSILON, "v_prediction": N.TARGET_V, "flow": N.TARGET_FLOW}


class ModelSetupFusedNoiseMixin(metaclass=ABCMeta):

    def __init__(self):
        super().__init__()
        self.__tables = None

    def _noise_timestep_addnoise(self, latent_image: Tensor, config, generator: Generator, *,
                                 num_train_timesteps: int, deterministic: bool, prediction_type: str,
                                 betas: Tensor | None = None, flow: bool = False, latent_scale: float = 1.0,
                                 latent_shift: float = 0.0, train_dtype: torch.dtype | None = None,
                                 latent_width: int | None = None, latent_height: int | None = None,
                                 want_noise: bool = False, want_scaled_latent: bool = False,
                                 pack_latents: bool = False) -> dict:
        """Returns {'timestep' int32[B], 'noisy' (train dtype), 'target', ['sigma'], ['noise'], ['scaled_latent']}.
        `pack_latents=True` writes 'noisy' directly in FLUX's packed layout [B, (H/2)(W/2), 4C]
        (FluxModel.pack_latents, modules/model/FluxModel.py:319-331) — no permute copy before the transformer."""
        dev = latent_image.device
        if self.__tables is None:
            if flow:
                sigma, one_minus = flow_sigmas(num_train_timesteps, dev)
                self.__tables = (one_minus, sigma)
            else:
                co = DiffusionScheduleCoefficients.from_betas(betas.to(device=dev))
                self.__tables = (co.sqrt_alphas_cumprod, co.sqrt_one_minus_alphas_cumprod)
        coef_x0, coef_noise = self.__tables
        B, C, hw = flatten_chw(latent_image.shape)
        seed, offset = draw_key(generator, dev)
        state = step_state.active()
        rng_dev = None
        if state is not None:                      # CUDA-graph mode: the key lives in device memory
            state.set_noise_key(seed, offset)
            rng_dev = state.noise_ptr
        kw = timestep_kernel_args(num_train_timesteps, deterministic, config, latent_width, latent_height, device=dev)
        out = ops.noise_addnoise(
            latent=latent_image.contiguous(), coef_x0=coef_x0, coef_noise=coef_noise, B=B, C_=C, hw=hw,
            latent_dtype=latent_image.dtype, device=dev, want_x0=want_scaled_latent, want_noise=want_noise,
            cast_dtype=train_dtype or latent_image.dtype, target=_TARGETS["flow" if flow else prediction_type],
            latent_scale=latent_scale, latent_shift=latent_shift,
            offset_noise_weight=config.offset_noise_weight, perturbation_noise_weight=config.perturbation_noise_weight,
            seed=seed, offset=offset, sample_base=dp.global_sample_base(B), want_sigma=flow, shape=latent_image.shape,
            rng_dev=rng_dev, pack_w=latent_image.shape[-1] if pack_latents else 0, **kw)
        res = {"timestep": out["timestep"], "noisy": out["noisy_cast"], "target": out["target"]}
        if flow:
            res["sigma"] = out["sigma"].view(B, *([1] * (latent_image.dim() - 1)))
        if want_noise:
            res["noise"] = out["noise"]
        if want_scaled_latent:
            res["scaled_latent"] = out["x0"]
        return res
