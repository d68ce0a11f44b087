"""DDPM forward process on the B200 kernel — drop-in for
/root/reference/modules/modelSetup/mixin/ModelSetupDiffusionMixin.py:15-38.

x_t = x0 * sqrt(abar_t) + eps * sqrt(1 - abar_t): the reference's 2 gathers + unsqueeze loop +
2 casts + 2 mul + add + cast (9 launches, 4 full-size temporaries) are one launch.  The products
and the sum are rounded separately (no FMA), so fp32 results are bit-identical to the reference's.
"""
from __future__ import annotations

from abc import ABCMeta

from torch import Tensor

from .... import ops
from .... import _native as N
from ....util_shapes import flatten_chw
from ...util.DiffusionScheduleCoefficients import DiffusionScheduleCoefficients


class ModelSetupDiffusionMixin(metaclass=ABCMeta):

    def __init__(self):
        super().__init__()
        self.__coefficients = None

    def _schedule_tables(self, betas: Tensor, device) -> DiffusionScheduleCoefficients:
        if self.__coefficients is None:
            self.__coefficients = DiffusionScheduleCoefficients.from_betas(betas.to(device=device))
        return self.__coefficients

    def _add_noise_discrete(self, scaled_latent_image: Tensor, latent_noise: Tensor, timestep: Tensor, betas: Tensor) -> Tensor:
        co = self._schedule_tables(betas, scaled_latent_image.device)
        B, C, hw = flatten_chw(scaled_latent_image.shape)
        if timestep.numel() == 1 and B > 1:       # deterministic mode hands over a [1] tensor: broadcast like indexing would
            timestep = timestep.expand(B)
        out = ops.noise_addnoise(
            latent=scaled_latent_image.contiguous(), noise_in=latent_noise.to(dtype=scaled_latent_image.dtype).contiguous(),
            coef_x0=co.sqrt_alphas_cumprod, coef_noise=co.sqrt_one_minus_alphas_cumprod, timestep=timestep,
            B=B, C_=C, hw=hw, latent_dtype=scaled_latent_image.dtype, device=scaled_latent_image.device,
            want_noisy=True, ts_mode=N.TS_GIVEN, generate_noise=False, num_train_timesteps=co.num_timesteps,
            shape=scaled_latent_image.shape)
        return out["noisy"]

    def _add_noise_continuous(self, scaled_latent_image, latent_noise, timestep, alphas_cumprod_fun):
        raise NotImplementedError(
            "continuous-time add-noise (Würstchen only, ModelSetupDiffusionMixin.py:40-57) is outside the "
            "B200 hot path (SURVEY.md §8a row A6); it is not silently emulated")
