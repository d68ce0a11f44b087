"""CPU baseline of the hot path: the oracle port timed on the host cores.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's `cpu_baseline` leg and `--impl reference` arm).
The reference's own CPU path cannot travel to the GPU box (it is Python under /root/reference, which
does not exist there), so the timed code is this repo's port of it: the C restatement
(otb_oracle.c, OpenMP over all host threads) for the element-wise stages, and the reference's LoRA
leaf arithmetic (base GEMM, down GEMM, up GEMM, scale, add; autograd backward) in torch CPU fp32
(MKL) for the contractions — which IS what the reference executes on CPU (config C1).

A full SDXL step is ~55 TFLOP of GEMMs (minutes on a CPU), so one "step" here is a BOUNDED SAMPLE:
  - noise + timestep + add-noise, loss fwd+bwd, grad-norm, AdamW over ALL 85.1 M LoRA parameters — full size
  - the 10 hooked leaves of ONE C=1280 transformer block (M=4096) and ONE C=640 block (M=16384), fwd+bwd
and the step time is extrapolated by leaf count (60 + 10 blocks; the 22 proj_in/out leaves by FLOPs).
"""
from __future__ import annotations

import os
import time

import numpy as np


def _block_leaves(C: int, ctx: int = 2048):
    """(K, N, cross) of the 10 hooked leaves of one BasicTransformerBlock at width C."""
    return [(C, C, False)] * 4 + [(C, C, False), (ctx, C, True), (ctx, C, True), (C, C, False), (C, 8 * C, False), (4 * C, C, False)]


def _time_block(C: int, M: int, B: int, rank: int, torch) -> float:
    scale = 1.0 / rank
    t = 0.0
    for (K, N, cross) in _block_leaves(C):
        rows = B * 77 if cross else M
        x = torch.randn(rows, K, requires_grad=True)
        W = torch.randn(N, K) / K ** 0.5
        A = (torch.randn(rank, K) / K ** 0.5).requires_grad_(True)
        Bm = (torch.randn(N, rank) * 0.02).requires_grad_(True)
        dy = torch.randn(rows, N)
        t0 = time.perf_counter()
        y = torch.nn.functional.linear(x, W) + torch.nn.functional.linear(torch.nn.functional.linear(x, A), Bm) * scale
        y.backward(dy)
        t += time.perf_counter() - t0
    return t


def run_sample(threads: int | None = None, batch: int = 4, rank: int = 32) -> dict:
    import torch
    from . import otb_oracle as O
    threads = threads or os.cpu_count() or 1
    O.set_threads(threads)
    torch.set_num_threads(threads)
    O.lib()
    T = 1000
    betas = (np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, T, dtype=np.float32) ** 2).astype(np.float32)
    ac = np.cumprod(1 - betas).astype(np.float32)
    sa, s1 = np.sqrt(ac), np.sqrt(1 - ac)
    B, Cc, hw = batch, 4, 128 * 128
    rng = np.random.default_rng(0)
    lat = rng.standard_normal((B, Cc * hw), dtype=np.float32)
    pred = rng.standard_normal((B, Cc * hw), dtype=np.float32)
    n_params = 85_114_880
    p = rng.standard_normal(n_params, dtype=np.float32) * 0.02
    g = rng.standard_normal(n_params, dtype=np.float32) * 1e-3
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    t0 = time.perf_counter()
    noise = O.create_noise(B, Cc, hw, seed=1, offset=0)
    ts = np.array([O.timestep(O.TS_UNIFORM, 1, 0, i) for i in range(B)], dtype=np.int32)
    _, noisy, tgt = O.add_noise(lat, noise, ts, sa, s1, latent_scale=0.13025, target=O.TARGET_EPSILON)
    t_noise = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.losses(pred, tgt, timestep_=ts, sqrt_ac=sa, sqrt_1mac=s1, C_=Cc, lw_fn=O.LW_MIN_SNR, upstream=1.0 / B)
    t_loss = time.perf_counter() - t0
    t0 = time.perf_counter()
    norm = O.grad_norm([g])
    coef = O.clip_coef(norm, 1.0)
    O.adam_step(p, g, m, v, O.DT_F32, O.adam_scalars(3e-4, 0.9, 0.999, 1e-8, 1e-2, 1.0), clip_coef=coef)
    t_opt = time.perf_counter() - t0
    t1280 = _time_block(1280, B * 1024, B, rank, torch)
    t640 = _time_block(640, B * 4096, B, rank, torch)
    # proj_in/proj_out leaves: 12 at C=1280 (M=4096) + 10 at C=640 (M=16384); cost ~ one to_q-sized leaf each
    frac1280 = (1280 * 1280) / sum(K * N for K, N, _ in _block_leaves(1280))
    frac640 = (640 * 640) / sum(K * N for K, N, _ in _block_leaves(640))
    t_gemm = 60 * t1280 + 10 * t640 + 12 * t1280 * frac1280 + 10 * t640 * frac640
    step = t_noise + t_loss + t_opt + t_gemm
    return dict(step_seconds=step, sample_seconds=t_noise + t_loss + t_opt + t1280 + t640, threads=threads,
                parts=dict(noise_addnoise=t_noise, loss=t_loss, clip_adamw=t_opt, block1280=t1280, block640=t640, gemm_extrapolated=t_gemm),
                sample="full-size noise/add-noise + loss fwd+bwd + clip + AdamW over 85.1M fp32 LoRA params (C oracle, OpenMP); "
                       "LoRA leaves of 1 of 60 C=1280 blocks and 1 of 10 C=640 blocks fwd+bwd (torch CPU fp32), "
                       "extrapolated to 722 leaves by count/FLOPs")
