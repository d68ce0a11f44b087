"""Noise and timestep sampling on the B200 kernels — drop-in for the reference mixin
(/root/reference/modules/modelSetup/mixin/ModelSetupNoiseMixin.py): same method names, argument
meaning, return dtypes/shapes and errors.

What changes underneath: the reference issues 1-3 `randn` launches plus ~6-10 tiny elementwise
launches; here `_create_noise` and `_get_timestep_discrete` are one launch each of
otb_noise_addnoise (include/otb200.h), and `_noise_timestep_addnoise` does noise + timestep +
add-noise + target + the train-dtype cast in ONE launch (the pre-network half of the fused step
kernel, K-b1).  Draws come from a counter-based Philox stream keyed by the generator's seed and
the GLOBAL sample index (onetrainer_b200/rng.py, dp.py), not from torch's stream.
"""
from __future__ import annotations

import math
from abc import ABCMeta

import torch
from torch import Generator, Tensor

from .... import dp, ops
from .... import _native as N
from ....rng import draw_key
from ....util_shapes import flatten_chw
from ...util.enum import enum_name

_TS_MODES = {"UNIFORM": N.TS_UNIFORM, "LOGIT_NORMAL": N.TS_LOGIT_NORMAL, "HEAVY_TAIL": N.TS_HEAVY_TAIL}


def timestep_shift(config, latent_width: int | None, latent_height: int | None) -> float:
    """Static shift, or FLUX's resolution-dependent exp(mu) (ModelSetupNoiseMixin.py:73-89)."""
    shift = config.timestep_shift
    if getattr(config, "dynamic_timestep_shifting", False):
        if not latent_width or not latent_height:
            raise NotImplementedError("Dynamic timestep shifting not support by this model")
        base_seq_len, max_seq_len, base_shift, max_shift, patch_size = 256, 4096, 0.5, 1.15, 2
        image_seq_len = (latent_width // patch_size) * (latent_height // patch_size)
        m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
        b = base_shift - m * base_seq_len
        shift = math.exp(image_seq_len * m + b)
    return shift


def timestep_kernel_args(num_train_timesteps: int, deterministic: bool, config, latent_width=None, latent_height=None) -> dict:
    """Host-side reduction of the config to the scalars the kernel takes — python ints/doubles
    exactly where the reference computes in python (ModelSetupNoiseMixin.py:69-71,91-121)."""
    if deterministic:
        return dict(ts_mode=N.TS_DETERMINISTIC, num_train_timesteps=num_train_timesteps)
    dist_name = enum_name(config.timestep_distribution)
    if dist_name not in _TS_MODES:
        raise NotImplementedError(f"timestep distribution {dist_name} is sampled on the host path")
    return dict(
        ts_mode=_TS_MODES[dist_name], num_train_timesteps=num_train_timesteps,
        min_timestep=int(num_train_timesteps * config.min_noising_strength),
        max_timestep=int(num_train_timesteps * config.max_noising_strength),
        noising_weight=config.noising_weight, noising_bias=config.noising_bias,
        shift=timestep_shift(config, latent_width, latent_height))


class ModelSetupNoiseMixin(metaclass=ABCMeta):

    def __init__(self):
        super().__init__()
        self.__weights = None

    def _create_noise(self, source_tensor: Tensor, config, generator: Generator) -> Tensor:
        B, C, hw = flatten_chw(source_tensor.shape)
        seed, offset = draw_key(generator, config.train_device)
        out = ops.noise_addnoise(
            B=B, C_=C, hw=hw, latent_dtype=source_tensor.dtype, device=torch.device(config.train_device),
            want_noise=True, do_add_noise=False, ts_mode=N.TS_DETERMINISTIC,
            offset_noise_weight=config.offset_noise_weight, perturbation_noise_weight=config.perturbation_noise_weight,
            seed=seed, offset=offset, sample_base=dp.global_sample_base(B), shape=source_tensor.shape)
        return out["noise"]

    def _get_timestep_discrete(self, num_train_timesteps: int, deterministic: bool, generator: Generator,
                               batch_size: int, config, latent_width: int | None = None,
                               latent_height: int | None = None) -> Tensor:
        if deterministic:
            # -1 is for zero-based indexing; shape [1], int64 — as the reference returns it
            return torch.tensor(int(num_train_timesteps * 0.5) - 1, dtype=torch.long, device=generator.device).unsqueeze(0)
        dist_name = enum_name(config.timestep_distribution)
        if dist_name in ("COS_MAP", "SIGMOID"):
            return self.__multinomial_timesteps(num_train_timesteps, generator, batch_size, config, latent_width, latent_height)
        kw = timestep_kernel_args(num_train_timesteps, False, config, latent_width, latent_height)
        seed, offset = draw_key(generator, generator.device)
        out = ops.noise_addnoise(B=batch_size, C_=1, hw=4, latent_dtype=torch.float32, device=generator.device,
                                 generate_noise=False, do_add_noise=False, seed=seed, offset=offset,
                                 sample_base=dp.global_sample_base(batch_size), **kw)
        return out["timestep"]

    def __multinomial_timesteps(self, num_train_timesteps, generator, batch_size, config, latent_width, latent_height):
        """COS_MAP / SIGMOID: a B-element multinomial over a cached 1000-entry weight table
        (ModelSetupNoiseMixin.py:122-153).  Off the kernel path: torch ops on B elements."""
        min_timestep = int(num_train_timesteps * config.min_noising_strength)
        max_timestep = int(num_train_timesteps * config.max_noising_strength)
        num_timestep = max_timestep - min_timestep
        shift = timestep_shift(config, latent_width, latent_height)
        if self.__weights is None:
            lin = torch.linspace(0, 1, num_timestep)
            derivative = shift / (shift + lin - (lin * shift)).pow(2)
            lin = lin / (shift - shift * lin + lin)
            if enum_name(config.timestep_distribution) == "COS_MAP":
                weights = 2.0 / (math.pi - 2.0 * math.pi * lin + 2.0 * math.pi * lin ** 2.0)
            else:
                bias, weight = config.noising_bias + 0.5, config.noising_weight
                weights = lin / (shift - shift * lin + lin)
                weights = 1 / (1 + torch.exp(-weight * (weights - bias)))
            self.__weights = (weights * derivative).to(device=generator.device)
        samples = torch.multinomial(self.__weights, num_samples=batch_size, replacement=True, generator=generator) + min_timestep
        return samples.to(dtype=torch.long, device=generator.device).int()

    def _get_timestep_continuous(self, deterministic: bool, generator: Generator, batch_size: int, config) -> Tensor:
        if deterministic:
            return torch.full(size=(batch_size,), fill_value=0.5, device=generator.device)
        discrete_timesteps = 10000
        discrete = self._get_timestep_discrete(discrete_timesteps, False, generator, batch_size, config) + 1
        return discrete.float() / discrete_timesteps
