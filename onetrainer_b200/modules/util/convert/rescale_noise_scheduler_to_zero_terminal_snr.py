"""Zero-terminal-SNR rescale of a DDPM schedule (arXiv 2305.08891, Algorithm 1).

Same contract as the reference's helper
(/root/reference/modules/util/convert/rescale_noise_scheduler_to_zero_terminal_snr.py:6-41):
the scheduler object (anything with `alphas_cumprod`, e.g. a diffusers DDIMScheduler) gets new
`betas`, `alphas`, `alphas_cumprod`, and the new betas are returned.  One-time 1000-element setup.
"""
from __future__ import annotations

import torch
from torch import Tensor


def rescale_betas_to_zero_terminal_snr(alphas_cumprod: Tensor) -> tuple[Tensor, Tensor, Tensor]:
    """Returns (betas, alphas, alphas_cumprod) of the rescaled schedule."""
    root = alphas_cumprod ** 0.5
    first, last = root[0].clone(), root[-1].clone()
    root = root - last                       # terminal sqrt(alpha_bar) -> 0
    root = root * (first / (first - last))   # restore the initial value
    ac = root ** 2
    alphas = torch.cat([ac[0:1], ac[1:] / ac[:-1]])
    return 1 - alphas, alphas, ac


def rescale_noise_scheduler_to_zero_terminal_snr(noise_scheduler) -> Tensor:
    betas, alphas, ac = rescale_betas_to_zero_terminal_snr(noise_scheduler.alphas_cumprod)
    noise_scheduler.betas = betas
    noise_scheduler.alphas = alphas
    noise_scheduler.alphas_cumprod = ac
    return betas
