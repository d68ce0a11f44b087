"""Schedule tables consumed by the add-noise and loss kernels.

Mirrors the two tables of the reference's DiffusionScheduleCoefficients that the hot path reads
(/root/reference/modules/util/DiffusionScheduleCoefficients.py:36-60: sqrt_alphas_cumprod,
sqrt_one_minus_alphas_cumprod) plus alphas_cumprod itself.  The remaining posterior/VB tables
exist only for the learned-variance loss (PixArt), which is outside SURVEY.md §8.

The tables are built once, with torch ops on the training device — a 1000-element setup step,
not hot-path work — so they carry exactly the bits the reference's tables carry on that device.
"""
from __future__ import annotations

import torch
from torch import Tensor


class DiffusionScheduleCoefficients:
    def __init__(self, betas: Tensor, alphas_cumprod: Tensor):
        self.num_timesteps = betas.shape[0]
        self.betas = betas
        self.alphas_cumprod = alphas_cumprod
        self.sqrt_alphas_cumprod = torch.sqrt(alphas_cumprod).contiguous()
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1 - alphas_cumprod).contiguous()

    @staticmethod
    def from_tables(betas: Tensor, alphas_cumprod: Tensor, sqrt_alphas_cumprod: Tensor,
                    sqrt_one_minus_alphas_cumprod: Tensor) -> "DiffusionScheduleCoefficients":
        """Adopt tables computed elsewhere (e.g. loaded from a backup, or a fixture) without recomputing them."""
        co = DiffusionScheduleCoefficients.__new__(DiffusionScheduleCoefficients)
        co.num_timesteps, co.betas, co.alphas_cumprod = betas.shape[0], betas, alphas_cumprod
        co.sqrt_alphas_cumprod = sqrt_alphas_cumprod.contiguous()
        co.sqrt_one_minus_alphas_cumprod = sqrt_one_minus_alphas_cumprod.contiguous()
        return co

    @staticmethod
    def from_betas(betas: Tensor) -> "DiffusionScheduleCoefficients":
        betas = betas.to(dtype=torch.float32)
        return DiffusionScheduleCoefficients(betas, torch.cumprod(1 - betas, dim=0))


def flow_sigmas(num_timesteps: int, device) -> tuple[Tensor, Tensor]:
    """sigma_t = (t+1)/T and 1 - sigma_t, as the reference builds them
    (ModelSetupFlowMatchingMixin.py:22-26, ModelSetupDiffusionLossMixin.py:297-300)."""
    ts = torch.arange(start=1, end=num_timesteps + 1, step=1, dtype=torch.int32, device=device)
    sigma = ts / num_timesteps
    return sigma.contiguous(), (1.0 - sigma).contiguous()


def scaled_linear_betas(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> Tensor:
    """diffusers' `scaled_linear` schedule (DDIMScheduler @ git 5873377, not vendored in the reference;
    same constants as /root/reference/resources/model_config/stable_diffusion/v1-inference.yaml:5-9)."""
    return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
