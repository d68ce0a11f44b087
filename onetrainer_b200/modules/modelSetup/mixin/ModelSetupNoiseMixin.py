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


def timestep_weights(num_train_timesteps: int, config, latent_width=None, latent_height=None) -> Tensor:
    """The reference's multinomial weight table for COS_MAP / SIGMOID (ModelSetupNoiseMixin.py:122-153), op for op, on
    the CPU as the reference builds it (its `linspace` has no device argument)."""
    min_timestep = int(num_train_timesteps * config.min_noising_strength)
    max_timestep = int(num_train_timesteps * config.max_noising_strength)
    num_timestep = max_timestep - min_timestep
    shift = timestep_shift(config, latent_width, latent_height)
    lin = torch.linspace(0, 1, num_timestep)
    lin = lin / (shift - shift * lin + lin)
    derivative = torch.linspace(0, 1, num_timestep)
    derivative = shift / (shift + derivative - (derivative * shift)).pow(2)
    if enum_name(config.timestep_distribution) == "COS_MAP":
        weights = 2.0 / (math.pi - 2.0 * math.pi * lin + 2.0 * math.pi * lin ** 2.0)
    else:
        bias, weight = config.noising_bias + 0.5, config.noising_weight
        weights = lin / (shift - shift * lin + lin)
        weights = 1 / (1 + torch.exp(-weight * (weights - bias)))
    weights *= derivative
    return weights


def timestep_cdf(weights: Tensor) -> Tensor:
    """Normalised running sum of the weight table (fp64 accumulate, one rounding to fp32, last entry exactly 1): what
    the kernel's inverse-CDF draw (OTB_TS_TABLE) searches.  (The CPU checker restates the same construction.)"""
    c = torch.cumsum(weights.detach().to("cpu", torch.float64), 0)
    c = (c / c[-1]).to(torch.float32)
    c[-1] = 1.0
    return c


_cdf_cache: dict = {}


def timestep_kernel_args(num_train_timesteps: int, deterministic: bool, config, latent_width=None, latent_height=None,
                         device=None) -> dict:
    """Host-side reduction of the config to the scalars the kernel takes — python ints/doubles
    exactly where the reference computes in python (ModelSetupNoiseMixin.py:69-71,91-121)."""
    if deterministic:
        return dict(ts_mode=N.TS_DETERMINISTIC, num_train_timesteps=num_train_timesteps)
    dist_name = enum_name(config.timestep_distribution)
    min_t, max_t = int(num_train_timesteps * config.min_noising_strength), int(num_train_timesteps * config.max_noising_strength)
    if dist_name in ("COS_MAP", "SIGMOID"):
        shift = timestep_shift(config, latent_width, latent_height)
        key = (dist_name, num_train_timesteps, min_t, max_t, float(shift), float(config.noising_weight), float(config.noising_bias), str(device))
        cdf = _cdf_cache.get(key)
        if cdf is None:
            if len(_cdf_cache) > 16:
                _cdf_cache.clear()
            cdf = _cdf_cache[key] = timestep_cdf(timestep_weights(num_train_timesteps, config, latent_width, latent_height)).to(device)
        return dict(ts_mode=N.TS_TABLE, num_train_timesteps=num_train_timesteps, min_timestep=min_t, max_timestep=max_t, ts_cdf=cdf)
    if dist_name not in _TS_MODES:
        raise NotImplementedError(f"timestep distribution {dist_name} is not implemented on the B200 path")
    return dict(
        ts_mode=_TS_MODES[dist_name], num_train_timesteps=num_train_timesteps, min_timestep=min_t, max_timestep=max_t,
        noising_weight=config.noising_weight, noising_bias=config.noising_bias,
        shift=timestep_shift(config, latent_width, latent_height))


class ModelSetupNoiseMixin(metaclass=ABCMeta):

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
        kw = timestep_kernel_args(num_train_timesteps, False, config, latent_width, latent_height, device=generator.device)
        seed, offset = draw_key(generator, generator.device)
        out = ops.noise_addnoise(B=batch_size, C_=1, hw=4, latent_dtype=torch.float32, device=generator.device,
                                 generate_noise=False, do_add_noise=False, seed=seed, offset=offset,
                                 sample_base=dp.global_sample_base(batch_size), **kw)
        return out["timestep"]

    def _get_timestep_continuous(self, deterministic: bool, generator: Generator, batch_size: int, config) -> Tensor:
        if deterministic:
            return torch.full(size=(batch_size,), fill_value=0.5, device=generator.device)
        discrete_timesteps = 10000
        discrete = self._get_timestep_discrete(discrete_timesteps, False, generator, batch_size, config) + 1
        return discrete.float() / discrete_timesteps
