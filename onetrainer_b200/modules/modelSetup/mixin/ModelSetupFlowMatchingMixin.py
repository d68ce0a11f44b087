"""Rectified-flow interpolation on the B200 kernel — drop-in for
/root/reference/modules/modelSetup/mixin/ModelSetupFlowMatchingMixin.py:14-39.

sigma_t = (t+1)/T;  x_t = eps * sigma + x0 * (1 - sigma);  returns (x_t, sigma[B,1,...,1]).
"""
from __future__ import annotations

from abc import ABCMeta

from torch import Tensor

from .... import ops
from .... import _native as N
from ....util_shapes import flatten_chw
from ...util.DiffusionScheduleCoefficients import flow_sigmas


class ModelSetupFlowMatchingMixin(metaclass=ABCMeta):

    def __init__(self):
        super().__init__()
        self.__sigma = None
        self.__one_minus_sigma = None

    def _flow_tables(self, num_timesteps: int, device):
        if self.__sigma is None:
            self.__sigma, self.__one_minus_sigma = flow_sigmas(num_timesteps, device)
        return self.__sigma, self.__one_minus_sigma

    def _add_noise_discrete(self, scaled_latent_image: Tensor, latent_noise: Tensor, timestep: Tensor,
                            timesteps: Tensor) -> tuple[Tensor, Tensor]:
        sigma, one_minus = self._flow_tables(timesteps.shape[-1], scaled_latent_image.device)
        B, C, hw = flatten_chw(scaled_latent_image.shape)
        if timestep.numel() == 1 and B > 1:
            timestep = timestep.expand(B)
        out = ops.noise_addnoise(
            latent=scaled_latent_image.contiguous(), noise_in=latent_noise.to(dtype=scaled_latent_image.dtype).contiguous(),
            coef_x0=one_minus, coef_noise=sigma, timestep=timestep,
            B=B, C_=C, hw=hw, latent_dtype=scaled_latent_image.dtype, device=scaled_latent_image.device,
            want_noisy=True, want_sigma=True, ts_mode=N.TS_GIVEN, generate_noise=False,
            num_train_timesteps=sigma.shape[0], shape=scaled_latent_image.shape)
        sigmas = out["sigma"].view(B, *([1] * (scaled_latent_image.dim() - 1)))
        return out["noisy"], sigmas
