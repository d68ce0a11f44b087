"""numpy front-end of oracle/otb_oracle.c.

TEST INFRASTRUCTURE ONLY — see the header of otb_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.

Besides the ctypes wrappers of the C restatement, this file holds the numpy restatements (fp64 accumulation) of:
  lora_backward        LoRAModule.forward's autograd backward           modules/module/LoRAModule.py:319-323
  conv_rows / _adjoint k x k conv leaves as row matrices                 modules/module/LoRAModule.py:142-155
  dora                 DoRAModule.forward + backward                     modules/module/LoRAModule.py:385-419
  adafactor_step       step_adafactor_parameter                          modules/util/optimizer/adafactor_extensions.py:17-93
  came_step            CAME.step_parameter                               modules/util/optimizer/CAME.py:80-176
  stochastic_round_bf16  copy_stochastic_                                modules/util/bf16_stochastic_rounding.py:5-30
  sr_helper            copy_ / add_ / addcdiv_stochastic_ (C restatement) modules/util/bf16_stochastic_rounding.py:5-61
all pinned on outputs of the unmodified reference (tests/golden/, tools/make_golden.py).  The last four have no kernel
yet: the oracle is in place first (task order: oracle, boundary, kernel).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "otb_oracle.c"
LIB = HERE / "_build" / "libotb_oracle.so"

FL_DIV_RECIPROCAL, FL_ADDCDIV_MUL_LAST, FL_FMA_LERP, FL_FMA_ADDCMUL, FL_FMA_ADDCDIV, FL_FMA_ADD_ALPHA, FL_ADDCMUL_SQUARE_FIRST = 1, 2, 4, 8, 16, 32, 64
FL_ALPHA_IN_DTYPE = 128
FLAVOR_TORCH_CPU = FL_FMA_LERP | FL_FMA_ADDCMUL | FL_FMA_ADD_ALPHA | FL_ALPHA_IN_DTYPE
FLAVOR_TORCH_CUDA = (FL_DIV_RECIPROCAL | FL_ADDCDIV_MUL_LAST | FL_FMA_LERP | FL_FMA_ADDCMUL | FL_FMA_ADDCDIV
                     | FL_FMA_ADD_ALPHA | FL_ADDCMUL_SQUARE_FIRST)
DT_F32, DT_BF16 = 0, 1
TS_UNIFORM, TS_LOGIT_NORMAL, TS_HEAVY_TAIL, TS_DETERMINISTIC = 0, 1, 2, 3
TARGET_EPSILON, TARGET_V, TARGET_FLOW = 1, 2, 3
LW_CONSTANT, LW_MIN_SNR, LW_DEBIASED, LW_P2, LW_SIGMA = range(5)


def build(force: bool = False) -> Path:
    LIB.parent.mkdir(exist_ok=True)
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                        "-o", str(LIB), str(SRC), "-lm"], check=True)
    return LIB


class AdamScalars(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("decay_mul", "l2", "lerp_w", "beta2", "one_minus_beta2", "bc2_sqrt",
                                         "inv_bc2_sqrt", "eps", "neg_step_size")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.otb_o_sqsum.restype = C.c_double
        _lib.otb_o_timestep.restype = C.c_int32
        _lib.otb_o_timestep_from_draw.restype = C.c_int32
        _lib.otb_o_timestep_table.restype = C.c_int32
        _lib.otb_o_timestep_weight.restype = C.c_float
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def set_threads(n: int) -> None:
    os.environ["OMP_NUM_THREADS"] = str(n)


# ---- Philox -------------------------------------------------------------------------------------
def philox(seed: int, offset: int, stream: int, group: int) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint32)
    lib().otb_o_philox(C.c_uint64(seed), C.c_uint64(offset), C.c_uint32(stream), C.c_uint64(group), _p(out))
    return out


def philox_raw(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().otb_o_philox_raw(_p(c), _p(k), _p(out))
    return out


def sr_bits(seed: int, offset: int, first_elem: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.int32)
    lib().otb_o_sr_bits(C.c_uint64(seed), C.c_uint64(offset), C.c_int64(first_elem), C.c_int64(n), _p(out))
    return out


# ---- bf16 helpers (numpy has no bf16: carried as uint16) -------------------------------------
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + (0x7FFF + ((u >> 16) & 1))) >> 16).astype(np.uint16)
    r[nan] = 0x7FC0
    return r


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


# ---- optimizer -------------------------------------------------------------------------------------
def adam_scalars(lr, beta1, beta2, eps, weight_decay, step, decoupled=True) -> AdamScalars:
    """python-double -> fp32 narrowing exactly as adamw_extensions.py:83-132 + torch's scalar casts."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    f = np.float32
    return AdamScalars(float(f(1 - lr * weight_decay)) if decoupled else 1.0, 0.0 if decoupled else float(f(weight_decay)),
                       float(f(1 - beta1)), float(f(beta2)), float(f(1 - beta2)), float(f(bc2_sqrt)),
                       float(f(1.0 / bc2_sqrt)), float(f(eps)), float(f(-step_size)))


def adam_step(p, g, m, v, dt, scalars: AdamScalars, flavor=FLAVOR_TORCH_CPU, sr=False, maximize=False, seed=0, offset=0,
              rng_base=0, clip_coef=None) -> None:
    """In place on contiguous numpy arrays (float32, or uint16 carrying bf16 bits)."""
    n = p.size
    tg = np.empty(n, dtype=np.float32)
    td = np.empty(n, dtype=np.float32)
    lib().otb_o_adam_step(_p(p), _p(g), _p(m), _p(v), C.c_int64(n), C.c_int(dt), C.byref(scalars), C.c_uint32(flavor),
                          C.c_int(int(sr)), C.c_int(int(maximize)), C.c_uint64(seed), C.c_uint64(offset), C.c_int64(rng_base),
                          C.c_int(0 if clip_coef is None else 1), C.c_float(0.0 if clip_coef is None else clip_coef),
                          _p(tg), _p(td))


def ema_step(ema, p, dt, one_minus_decay: float) -> None:
    lib().otb_o_ema_step(_p(ema), _p(p), C.c_int64(p.size), C.c_int(dt), C.c_float(one_minus_decay))


def grad_norm(grads: list[np.ndarray], dt=DT_F32) -> float:
    return math.sqrt(sum(lib().otb_o_sqsum(_p(g), C.c_int64(g.size), C.c_int(dt)) for g in grads))


def clip_coef(total_norm: float, max_norm: float) -> float:
    c = np.float32(max_norm) / (np.float32(total_norm) + np.float32(1e-6))
    return float(min(c, np.float32(1.0)))


# ---- timesteps / noise -----------------------------------------------------------------------------
def timestep(mode, seed, offset, global_sample, T=1000, min_t=0, max_t=1000, noising_weight=0.0, noising_bias=0.0, shift=1.0) -> int:
    return int(lib().otb_o_timestep(C.c_int(mode), C.c_uint64(seed), C.c_uint64(offset), C.c_int64(global_sample), C.c_int(T),
                                    C.c_int(min_t), C.c_int(max_t), C.c_float(noising_weight), C.c_float(noising_bias),
                                    C.c_float(float(T * shift)), C.c_float(float(shift - 1))))


def timestep_cdf(weights) -> np.ndarray:
    """Normalised running sum of the reference's multinomial weight table (fp64 accumulate, one rounding, last entry 1)."""
    c = np.cumsum(np.asarray(weights, dtype=np.float64))
    c = (c / c[-1]).astype(np.float32)
    c[-1] = 1.0
    return c


def timestep_table(seed, offset, global_sample, cdf, min_t=0) -> int:
    """COS_MAP / SIGMOID draw for one global sample (ModelSetupNoiseMixin.py:122-153 restated as an inverse CDF)."""
    cdf = np.ascontiguousarray(cdf, dtype=np.float32)
    return int(lib().otb_o_timestep_table(C.c_uint64(seed), C.c_uint64(offset), C.c_int64(global_sample),
                                          cdf.ctypes.data_as(C.c_void_p), C.c_int(cdf.size), C.c_int(min_t)))


def timestep_from_draw(mode, draw, T=1000, min_t=0, max_t=1000, noising_weight=0.0, noising_bias=0.0, shift=1.0) -> int:
    return int(lib().otb_o_timestep_from_draw(C.c_int(mode), C.c_float(draw), C.c_int(T), C.c_int(min_t), C.c_int(max_t),
                                              C.c_float(noising_weight), C.c_float(noising_bias),
                                              C.c_float(float(T * shift)), C.c_float(float(shift - 1))))


def create_noise(B, Cc, hw, offset_w=0.0, perturb_w=0.0, seed=0, offset=0, sample_base=0) -> np.ndarray:
    out = np.empty((B, Cc * hw), dtype=np.float32)
    lib().otb_o_create_noise(_p(out), C.c_int(B), C.c_int(Cc), C.c_int(hw), C.c_float(offset_w), C.c_float(perturb_w),
                             C.c_uint64(seed), C.c_uint64(offset), C.c_int64(sample_base))
    return out


def add_noise(latent, noise, timestep_, coef_x0, coef_noise, latent_scale=1.0, latent_shift=0.0, target=TARGET_EPSILON):
    latent = np.ascontiguousarray(latent, dtype=np.float32)
    noise = np.ascontiguousarray(noise, dtype=np.float32)
    ts = np.ascontiguousarray(timestep_, dtype=np.int32)
    B = latent.shape[0]
    chw = latent.size // B
    x0, noisy, tgt = np.empty_like(latent), np.empty_like(latent), np.empty_like(latent)
    lib().otb_o_add_noise(_p(latent), _p(noise), _p(ts), _p(np.ascontiguousarray(coef_x0, dtype=np.float32)),
                          _p(np.ascontiguousarray(coef_noise, dtype=np.float32)), C.c_int(B), C.c_int64(chw),
                          C.c_float(latent_scale), C.c_float(latent_shift), C.c_int(target), _p(x0), _p(noisy), _p(tgt))
    return x0, noisy, tgt


# ---- losses --------------------------------------------------------------------------------------
def losses(pred, target, *, prior=None, mask=None, loss_weight=None, timestep_=None, sqrt_ac=None, sqrt_1mac=None, sigmas=None,
           C_=1, masked=False, normalize=False, lw_fn=LW_CONSTANT, v_pred=False, mse=1.0, mae=0.0, unmasked_weight=0.1,
           prior_w=0.0, loss_scale=1.0, gamma=5.0, upstream=1.0, want_grad=True):
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    target = np.ascontiguousarray(target, dtype=np.float32)
    B = pred.shape[0]
    hw = pred.size // B // C_
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    prior, mask, loss_weight, sqrt_ac, sqrt_1mac, sigmas = map(f32, (prior, mask, loss_weight, sqrt_ac, sqrt_1mac, sigmas))
    ts = None if timestep_ is None else np.ascontiguousarray(timestep_, dtype=np.int32)
    out = np.empty(B, dtype=np.float32)
    dp = np.empty_like(pred) if want_grad else None
    lib().otb_o_losses(_p(pred), _p(target), _p(prior), _p(mask), _p(loss_weight), _p(ts), _p(sqrt_ac), _p(sqrt_1mac), _p(sigmas),
                       C.c_int(B), C.c_int(C_), C.c_int(hw), C.c_int(int(masked)), C.c_int(int(normalize)), C.c_int(lw_fn),
                       C.c_int(int(v_pred)), C.c_float(mse), C.c_float(mae), C.c_float(unmasked_weight), C.c_float(prior_w),
                       C.c_float(loss_scale), C.c_float(gamma), C.c_float(upstream), _p(out), _p(dp))
    return out, dp


# ---- LoRA linear ---------------------------------------------------------------------------------
def lora_linear(x, w, down, up, bias, scale, round_t_bf16=True):
    x, w = (np.ascontiguousarray(a, dtype=np.float32) for a in (x, w))
    M, K = x.shape
    Nn = w.shape[0]
    R = 0 if down is None else down.shape[0]
    down = None if down is None else np.ascontiguousarray(down, dtype=np.float32)
    up = None if up is None else np.ascontiguousarray(up, dtype=np.float32)
    bias = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    y = np.empty((M, Nn), dtype=np.float32)
    t = np.empty((M, max(R, 1)), dtype=np.float32)
    lib().otb_o_lora_linear(_p(x), _p(w), _p(down), _p(up), _p(bias), C.c_int(M), C.c_int(Nn), C.c_int(K), C.c_int(R),
                            C.c_float(scale), C.c_int(int(round_t_bf16)), _p(y), _p(t))
    return y, t[:, :R]


def lora_backward(x, dy, w, down, up, scale, round_bf16=True):
    """dx, dA, dB of y = x w^T + (s x down^T) up^T  (SURVEY.md §3.4), fp64 accumulate via numpy."""
    x64, dy64, w64, a64, b64 = (np.asarray(a, dtype=np.float64) for a in (x, dy, w, down, up))
    t = scale * (x64 @ a64.T)
    g = scale * (dy64 @ b64)
    if round_bf16:
        t = bf16_bits_to_f32(f32_to_bf16_bits(t.astype(np.float32))).astype(np.float64)
        g = bf16_bits_to_f32(f32_to_bf16_bits(g.astype(np.float32))).astype(np.float64)
    dx = dy64 @ w64 + g @ a64
    dA = g.T @ x64
    dB = dy64.T @ t
    return dx.astype(np.float32), dA.astype(np.float32), dB.astype(np.float32)


# ---- k x k convolution leaves as row/column matrices ---------------------------------------------
def conv_rows(x, kernel, stride=(1, 1), padding=(0, 0), dilation=(1, 1)):
    """NCHW -> one row per output pixel, columns ordered (c, ki, kj) like `weight.reshape(Cout, -1)`, zero padded:
    the k x k leaves of LoRAModule.create_layer (LoRAModule.py:142-155) then read as the Linear arithmetic above."""
    x = np.asarray(x)
    Bn, Cc, H, W = x.shape
    (kh, kw), (sh, sw), (ph, pw), (dh, dw) = kernel, stride, padding, dilation
    Ho = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    Wo = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    xp = np.zeros((Bn, Cc, H + 2 * ph, W + 2 * pw), dtype=x.dtype)
    xp[:, :, ph:ph + H, pw:pw + W] = x
    rows = np.empty((Bn, Ho, Wo, Cc, kh, kw), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            rows[:, :, :, :, i, j] = xp[:, :, i * dh:i * dh + sh * (Ho - 1) + 1:sh, j * dw:j * dw + sw * (Wo - 1) + 1:sw].transpose(0, 2, 3, 1)
    return rows.reshape(Bn * Ho * Wo, Cc * kh * kw), (Ho, Wo)


def conv_rows_adjoint(drows, xshape, kernel, stride=(1, 1), padding=(0, 0), dilation=(1, 1)):
    """Transpose of conv_rows: scatter-add the row gradients back onto the NCHW input (fp64 accumulate)."""
    Bn, Cc, H, W = xshape
    (kh, kw), (sh, sw), (ph, pw), (dh, dw) = kernel, stride, padding, dilation
    Ho = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    Wo = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    d = np.asarray(drows, dtype=np.float64).reshape(Bn, Ho, Wo, Cc, kh, kw)
    dxp = np.zeros((Bn, Cc, H + 2 * ph, W + 2 * pw), dtype=np.float64)
    for i in range(kh):
        for j in range(kw):
            dxp[:, :, i * dh:i * dh + sh * (Ho - 1) + 1:sh, j * dw:j * dw + sw * (Wo - 1) + 1:sw] += d[:, :, :, :, i, j].transpose(0, 3, 1, 2)
    return dxp[:, :, ph:ph + H, pw:pw + W].astype(np.float32)


# ---- DoRA (oracle ahead of the kernel: the fused path still raises for it) -----------------------
def dora(x2, dy2, w, down, up, m, bias, scale, norm_epsilon=False, decompose_output_axis=False):
    """DoRAModule.forward (LoRAModule.py:385-419) and its backward with the norm detached (:396-400), fp64 accumulate.
    x2 / dy2 are row matrices ([M, Cin*kh*kw] via conv_rows for a conv leaf, [M, Cout]); w, down, up, m keep the module's
    own shapes.  Returns y2, dx2, d_down, d_up, d_m."""
    w64 = np.asarray(w, dtype=np.float64)
    a2 = np.asarray(down, dtype=np.float64).reshape(down.shape[0], -1)
    b2 = np.asarray(up, dtype=np.float64).reshape(up.shape[0], -1)
    x64, dy64, m64 = (np.asarray(t, dtype=np.float64) for t in (x2, dy2, m))
    v = w64 + scale * (b2 @ a2).reshape(w64.shape)                       # make_weight(A, B) * alpha / rank  (:76-91, :389)
    axes = tuple(range(1, v.ndim)) if decompose_output_axis else (0,) + tuple(range(2, v.ndim))
    norm = np.sqrt((v * v).sum(axis=axes, keepdims=True)) + (float(np.finfo(np.float32).eps) if norm_epsilon else 0.0)
    wp = m64 * (v / norm)
    wp2 = wp.reshape(w64.shape[0], -1)
    y = x64 @ wp2.T
    if bias is not None and np.size(bias):
        y = y + np.asarray(bias, dtype=np.float64)
    g = (dy64.T @ x64).reshape(w64.shape)                                # dL/dWP
    d_m = (g * v / norm).sum(axis=axes, keepdims=True)
    dv2 = (g * m64 / norm).reshape(w64.shape[0], -1)                     # the norm is a constant for backpropagation
    d_up = scale * (dv2 @ a2.T)
    d_down = scale * (b2.T @ dv2)
    dx = dy64 @ wp2
    f32 = lambda t: t.astype(np.float32)  # noqa: E731
    return f32(y), f32(dx), f32(d_down).reshape(down.shape), f32(d_up).reshape(up.shape), f32(d_m)


# ---- Adafactor (oracle ahead of the kernel: the fused optimizer covers AdamW / Adam) -------------
def adafactor_step(p, grad, state, *, lr=None, eps=(1e-30, 1e-3), clip_threshold=1.0, decay_rate=-0.8, beta1=None,
                   weight_decay=0.0, scale_parameter=True, relative_step=True, warmup_init=False, bf16=False):
    """One step_adafactor_parameter (adafactor_extensions.py:17-93) over transformers' Adafactor helpers (_get_lr,
    _rms, _approx_sq_grad; transformers pinned by requirements-global.txt).  `p` / `grad` are fp32 arrays holding the
    tensor's values (bf16 tensors: the widened values, `bf16=True`); `state` is the per-tensor dict, {} before the first
    step.  Elementwise work in fp32 like the reference, reductions accumulated in fp64.  Returns the new parameter.
    The reference's bf16 tail overwrites copy_stochastic_'s result with p.copy_(fp32) (:88-91), i.e. round-to-nearest
    whether or not stochastic rounding is on; that is what this does."""
    f32 = np.float32
    p32 = np.asarray(p, dtype=f32).copy()
    g = np.asarray(grad, dtype=f32)
    factored, use_first_moment = g.ndim >= 2, beta1 is not None                       # _get_options
    if not state:
        state["step"] = 0
        if use_first_moment:
            state["exp_avg"] = np.zeros_like(g)
        if factored:
            state["exp_avg_sq_row"] = np.zeros(g.shape[:-1], dtype=f32)
            state["exp_avg_sq_col"] = np.zeros(g.shape[:-2] + g.shape[-1:], dtype=f32)
        else:
            state["exp_avg_sq"] = np.zeros_like(g)
    rms = lambda t: f32(np.sqrt((t.astype(np.float64) ** 2).sum()) / (t.size ** 0.5))  # noqa: E731  (_rms)
    mean = lambda t, axis, keep=False: (t.astype(np.float64).mean(axis=axis, keepdims=keep)).astype(f32)  # noqa: E731
    state["step"] += 1
    state["RMS"] = rms(p32)
    step_size = lr                                                                     # _get_lr
    if relative_step:
        min_step = 1e-6 * state["step"] if warmup_init else 1e-2
        step_size = min(min_step, 1.0 / math.sqrt(state["step"]))
    param_scale = max(eps[1], float(state["RMS"])) if scale_parameter else 1.0
    lr_t = f32(param_scale * step_size)
    beta2t = 1.0 - math.pow(state["step"], decay_rate)
    update = g * g + f32(eps[0])
    if factored:
        row, col = state["exp_avg_sq_row"], state["exp_avg_sq_col"]
        row *= f32(beta2t)
        row += f32(1.0 - beta2t) * mean(update, -1)
        col *= f32(beta2t)
        col += f32(1.0 - beta2t) * mean(update, -2)
        r_factor = (f32(1.0) / np.sqrt(row / mean(row, -1, keep=True)))[..., None]     # _approx_sq_grad
        c_factor = (f32(1.0) / np.sqrt(col))[..., None, :]
        update = (r_factor * c_factor) * g
    else:
        v = state["exp_avg_sq"]
        v *= f32(beta2t)
        v += f32(1.0 - beta2t) * update
        update = (f32(1.0) / np.sqrt(v)) * g
    update = update / max(f32(rms(update) / f32(clip_threshold)), f32(1.0))
    update = update * lr_t
    if use_first_moment:
        m = state["exp_avg"]
        m *= f32(beta1)
        m += f32(1.0 - beta1) * update
        update = m
    if weight_decay != 0:
        p32 = p32 + p32 * f32(-weight_decay * float(lr_t))
    p32 = p32 - update
    return bf16_bits_to_f32(f32_to_bf16_bits(p32)) if bf16 else p32


# ---- CAME (oracle ahead of the kernel) -------------------------------------------------------------
def _factored_rsqrt(row, col):
    """_approx_sq_grad of Adafactor / CAME: rsqrt(row / mean(row)) x rsqrt(col)."""
    f32 = np.float32
    r = (f32(1.0) / np.sqrt(row / row.astype(np.float64).mean(axis=-1, keepdims=True).astype(f32)))[..., None]
    c = (f32(1.0) / np.sqrt(col))[..., None, :]
    return r * c


def stochastic_round_bf16(x32, bits):
    """copy_stochastic_ (bf16_stochastic_rounding.py:5-30): int32 view + 16 random bits, low half masked off."""
    u = np.ascontiguousarray(x32, dtype=np.float32).view(np.int32) + np.asarray(bits, dtype=np.int32).reshape(np.shape(x32))
    return (u & np.int32(-65536)).view(np.float32)


SR_COPY, SR_ADD, SR_ADDCDIV = 0, 1, 2


def sr_helper(mode: int, target_bits, a, b=None, *, alpha: float = 1.0, bits, flavor=FLAVOR_TORCH_CPU, want_r32: bool = False):
    """copy_stochastic_ / add_stochastic_ / addcdiv_stochastic_ (modules/util/bf16_stochastic_rounding.py:5-61) as their
    callers see them.  `target_bits`: uint16 bf16 bits of `target` / `input` (read by ADD and ADDCDIV); `a`, `b`: the other
    operands, float32 arrays or uint16 arrays carrying bf16 bits; `bits`: the 16-bit draw of every element.  Returns the
    new bf16 bits (the input array is not modified); with `want_r32` also the fp32 values before the rounding."""
    out = np.ascontiguousarray(target_bits, dtype=np.uint16).copy()
    a = np.ascontiguousarray(a)
    src_dt = DT_BF16 if a.dtype == np.uint16 else DT_F32
    assert a.dtype in (np.uint16, np.float32) and a.size == out.size and (mode != SR_COPY or src_dt == DT_F32)
    if b is not None:
        b = np.ascontiguousarray(b)
        assert b.dtype == a.dtype and b.size == out.size
    bits = np.ascontiguousarray(bits, dtype=np.int32)
    assert bits.size == out.size and (mode == SR_ADDCDIV) == (b is not None)
    r32 = np.empty(out.size, dtype=np.float32) if want_r32 else None
    lib().otb_o_sr_helper(_p(out), _p(a), _p(b), C.c_int64(out.size), C.c_int(mode), C.c_int(src_dt), C.c_float(alpha),
                          C.c_uint32(flavor), _p(bits), _p(r32))
    return (out, r32) if want_r32 else out


def came_step(p, grad, state, *, lr, eps=(1e-30, 1e-16), clip_threshold=1.0, betas=(0.9, 0.999, 0.9999), weight_decay=0.0,
              sr_draws=None):
    """One CAME.step_parameter (modules/util/optimizer/CAME.py:80-176).  `p` / `grad`: fp32 arrays with the tensor's
    values; `state`: per-tensor dict, {} before the first step.  `sr_draws` (bf16 parameter with stochastic rounding):
    a callable n -> int32[n] of 16-bit draws, called once per add_stochastic_ in the reference's order (weight decay
    first, then the update); None = fp32 parameter, plain adds.  Returns the new parameter values."""
    f32 = np.float32
    p32 = np.asarray(p, dtype=f32).copy()
    g = np.asarray(grad, dtype=f32)
    factored = g.ndim >= 2
    if not state:
        state["step"] = 0
        state["exp_avg"] = np.zeros_like(g)
        if factored:
            for key in ("exp_avg_sq_row", "exp_avg_res_row"):
                state[key] = np.zeros(g.shape[:-1], dtype=f32)
            for key in ("exp_avg_sq_col", "exp_avg_res_col"):
                state[key] = np.zeros(g.shape[:-2] + g.shape[-1:], dtype=f32)
        else:
            state["exp_avg_sq"] = np.zeros_like(g)
    mean = lambda t, axis: t.astype(np.float64).mean(axis=axis).astype(f32)  # noqa: E731
    rms = lambda t: f32(np.sqrt((t.astype(np.float64) ** 2).sum()) / (t.size ** 0.5))  # noqa: E731
    state["step"] += 1
    state["RMS"] = rms(p32)

    def ema_(buf, value, beta):
        buf *= f32(beta)
        buf += f32(1.0 - beta) * value

    update = g * g + f32(eps[0])
    if factored:
        ema_(state["exp_avg_sq_row"], mean(update, -1), betas[1])
        ema_(state["exp_avg_sq_col"], mean(update, -2), betas[1])
        update = _factored_rsqrt(state["exp_avg_sq_row"], state["exp_avg_sq_col"]) * g
    else:
        ema_(state["exp_avg_sq"], update, betas[1])
        update = (f32(1.0) / np.sqrt(state["exp_avg_sq"])) * g
    update = update / max(f32(rms(update) / f32(clip_threshold)), f32(1.0))
    exp_avg = state["exp_avg"]
    ema_(exp_avg, update, betas[0])
    res = (update - exp_avg) ** 2 + f32(eps[1])                         # instability (:137-139)
    if factored:
        ema_(state["exp_avg_res_row"], mean(res, -1), betas[2])
        ema_(state["exp_avg_res_col"], mean(res, -2), betas[2])
        update = _factored_rsqrt(state["exp_avg_res_row"], state["exp_avg_res_col"]) * exp_avg
    else:
        update = exp_avg.copy()
    if weight_decay != 0:
        decayed = p32 + f32(-weight_decay * lr) * p32
        p32 = stochastic_round_bf16(decayed, sr_draws(p32.size)) if sr_draws is not None else decayed
    update = update * f32(lr)
    stepped = -update + p32                                            # add_stochastic_: result = (-update).clone(); result.add_(p)
    return stochastic_round_bf16(stepped, sr_draws(p32.size)) if sr_draws is not None else stepped
