"""ctypes binding of libotb200.so (include/otb200.h).

PyTorch tensors stop here: everything below this file sees raw device pointers, sizes and a
stream handle.  There is deliberately no fallback — if the library is missing or a call fails
the caller gets a RuntimeError / NotImplementedError, never a silent PyTorch path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libotb200.so"

# ---- enums (include/otb200.h) ---------------------------------------------------------
ABI_VERSION = 6          # include/otb200.h: OTB_ABI_VERSION
OTB_OK, OTB_EINVAL, OTB_EUNSUPPORTED, OTB_ECUDA, OTB_EARCH = 0, -1, -2, -3, -4
F32, BF16, F16 = 0, 1, 2

FLAVOR_DIV_RECIPROCAL = 1
FLAVOR_ADDCDIV_MUL_LAST = 2
FLAVOR_FMA_LERP = 4
FLAVOR_FMA_ADDCMUL = 8
FLAVOR_FMA_ADDCDIV = 16
FLAVOR_FMA_ADD_ALPHA = 32
FLAVOR_ADDCMUL_SQUARE_FIRST = 64
FLAVOR_ALPHA_IN_DTYPE = 128
FLAVOR_TORCH_CPU = FLAVOR_FMA_LERP | FLAVOR_FMA_ADDCMUL | FLAVOR_FMA_ADD_ALPHA | FLAVOR_ALPHA_IN_DTYPE
FLAVOR_TORCH_CUDA = (FLAVOR_DIV_RECIPROCAL | FLAVOR_ADDCDIV_MUL_LAST | FLAVOR_FMA_LERP | FLAVOR_FMA_ADDCMUL
                     | FLAVOR_FMA_ADDCDIV | FLAVOR_FMA_ADD_ALPHA | FLAVOR_ADDCMUL_SQUARE_FIRST)

STREAM_NOISE, STREAM_OFFSET, STREAM_PERTURB, STREAM_TIMESTEP, STREAM_SR, STREAM_DROPOUT = range(6)

TS_UNIFORM, TS_LOGIT_NORMAL, TS_HEAVY_TAIL, TS_DETERMINISTIC, TS_GIVEN, TS_TABLE = range(6)
TARGET_NONE, TARGET_EPSILON, TARGET_V, TARGET_FLOW = range(4)
LW_CONSTANT, LW_MIN_SNR, LW_DEBIASED, LW_P2, LW_SIGMA = range(5)

OPT_FLAG_SR, OPT_FLAG_MAXIMIZE, OPT_FLAG_ALIGNED16 = 1, 2, 4

_vp, _i32, _i64, _u32, _u64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float


class OptimTensor(C.Structure):
    _fields_ = [
        ("p", _vp), ("g", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("ema", _vp), ("shadow_bf16", _vp),
        ("numel", _i64), ("rng_base", _i64),
        ("decay_mul", _f32), ("l2", _f32), ("lerp_w", _f32), ("beta2", _f32), ("one_minus_beta2", _f32),
        ("bc2_sqrt", _f32), ("inv_bc2_sqrt", _f32), ("eps", _f32), ("neg_step_size", _f32),
        ("p_dtype", _i32), ("flags", _i32),
    ]


class OptimLaunch(C.Structure):
    _fields_ = [
        ("tensors", _vp), ("chunk_tensor", _vp), ("chunk_index", _vp),
        ("n_tensors", _i32), ("n_chunks", _i32), ("chunk_elems", _i32), ("flavor", _u32),
        ("seed", _u64), ("offset", _u64),
        ("ema_one_minus_decay", _f32), ("do_ema", _i32),
        ("clip_coef", _vp), ("dyn", _vp),
    ]


class OptimDyn(C.Structure):
    _fields_ = [("seed", _u64), ("offset", _u64), ("ema_one_minus_decay", _f32), ("do_ema", _i32)]


class ShadowTensor(C.Structure):
    _fields_ = [
        ("src", _vp), ("dst", _vp), ("dst_t", _vp),
        ("rows", _i32), ("cols", _i32), ("src_dtype", _i32), ("tile_base", _i32),
    ]


class NoiseArgs(C.Structure):
    _fields_ = [
        ("latent", _vp), ("noise_in", _vp), ("x0_out", _vp), ("noise_out", _vp), ("noisy_out", _vp),
        ("noisy_cast_out", _vp), ("target_out", _vp), ("timestep", _vp), ("sigma_out", _vp),
        ("coef_x0", _vp), ("coef_noise", _vp),
        ("B", _i32), ("C", _i32), ("hw", _i32),
        ("latent_dtype", _i32), ("cast_dtype", _i32),
        ("target", _i32), ("ts_mode", _i32),
        ("num_train_timesteps", _i32), ("generate_noise", _i32), ("do_add_noise", _i32),
        ("latent_scale", _f32), ("latent_shift", _f32),
        ("offset_noise_weight", _f32), ("perturbation_noise_weight", _f32),
        ("min_timestep", _i32), ("max_timestep", _i32),
        ("noising_weight", _f32), ("noising_bias", _f32),
        ("shift_c1", _f32), ("shift_c2", _f32),
        ("seed", _u64), ("offset", _u64), ("sample_base", _i64), ("rng_dev", _vp), ("pack_w", _i32),
        ("ts_cdf", _vp), ("ts_cdf_len", _i32),
    ]


class LossArgs(C.Structure):
    _fields_ = [
        ("predicted", _vp), ("target", _vp), ("prior_target", _vp), ("mask", _vp), ("loss_weight", _vp),
        ("timestep", _vp), ("sqrt_ac", _vp), ("sqrt_1mac", _vp), ("sigmas", _vp),
        ("losses", _vp), ("dpred", _vp), ("partials", _vp), ("counters", _vp), ("mask_stats", _vp),
        ("B", _i32), ("C", _i32), ("hw", _i32),
        ("pred_dtype", _i32), ("target_dtype", _i32), ("mask_dtype", _i32),
        ("masked", _i32), ("normalize_masked_area_loss", _i32), ("loss_weight_fn", _i32), ("v_prediction", _i32),
        ("num_train_timesteps", _i32), ("blocks_per_sample", _i32),
        ("mse_strength", _f32), ("mae_strength", _f32),
        ("unmasked_weight", _f32), ("masked_prior_preservation_weight", _f32),
        ("loss_scale", _f32), ("loss_weight_strength", _f32), ("upstream", _f32), ("pred_pack_w", _i32),
    ]


class LoraLinearArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("w", _vp), ("down", _vp), ("up", _vp), ("bias", _vp), ("y", _vp), ("t_out", _vp),
        ("M", _i32), ("N", _i32), ("K", _i32), ("R", _i32),
        ("bias_dtype", _i32), ("scale", _f32), ("dropout_p", _f32),
        ("seed", _u64), ("offset", _u64),
        ("tile_n", _i32), ("stages", _i32), ("flags", _i32),
    ]


class LoraGradArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("dy", _vp), ("gs", _vp), ("ts", _vp), ("dA", _vp), ("dB", _vp),
        ("M", _i32), ("N", _i32), ("K", _i32), ("R", _i32),
    ]


FACTORED_FACTORED, FACTORED_FIRST_MOMENT, FACTORED_SCALE_PARAMETER, FACTORED_SR = 1, 2, 4, 8


class FactoredTensor(C.Structure):
    _fields_ = [
        ("p", _vp), ("g", _vp), ("exp_avg", _vp), ("sq_row", _vp), ("sq_col", _vp), ("sq", _vp), ("res_row", _vp), ("res_col", _vp),
        ("rms_out", _vp),
        ("numel", _i64), ("rng_base", _i64), ("scratch_offset", _i64),
        ("batch", _i32), ("rows", _i32), ("cols", _i32), ("p_dtype", _i32), ("flags", _i32),
        ("beta1", _f32), ("beta2", _f32), ("beta3", _f32), ("one_minus_beta1", _f32), ("one_minus_beta2", _f32), ("one_minus_beta3", _f32),
        ("eps0", _f32), ("eps1", _f32), ("clip_threshold", _f32),
        ("lr", _f32), ("weight_decay", _f32),
    ]


class DoraArgs(C.Structure):
    _fields_ = [
        ("w", _vp), ("ba", _vp), ("dora_scale", _vp), ("norm", _vp), ("scratch", _vp), ("wp", _vp), ("wpt", _vp), ("G", _vp), ("dV", _vp),
        ("d_scale", _vp),
        ("N", _i32), ("J", _i32), ("group_size", _i32), ("out_axis", _i32), ("w_dtype", _i32), ("scale", _f32), ("eps", _f32),
    ]


class FactoredLaunch(C.Structure):
    _fields_ = [("tensors", _vp), ("scratch", _vp), ("n_tensors", _i32), ("cluster", _i32), ("seed", _u64), ("offset", _u64)]


MAX_RANKS = 16           # OTB_MAX_RANKS
GRAD_BATCH_MAX = 32      # OTB_GRAD_BATCH_MAX
LORA_GROUP_MAX = 32      # OTB_LORA_GROUP_MAX
LINEAR_STATIC_WEIGHTS = 1


class GradExchangeArgs(C.Structure):
    _fields_ = [
        ("peer_grads", _vp * MAX_RANKS), ("peer_partials", _vp * MAX_RANKS), ("peer_flags", _vp * MAX_RANKS),
        ("mc_grads", _vp), ("mc_partials", _vp), ("n_elems", _i64),
        ("rank", _i32), ("world", _i32), ("blocks", _i32), ("mode", _i32), ("grad_prescale", _f32),
    ]


EXCHANGE_P2P, EXCHANGE_NVLS = 0, 1
SR_COPY, SR_ADD, SR_ADDCDIV = 0, 1, 2        # OTB_SR_*
PEER_PHASES = 4          # include/otb200.h: OTB_PEER_PHASES


class PeerStore(C.Structure):
    _fields_ = [("delta", _i64 * MAX_RANKS), ("mc_delta", _i64), ("rank", _i32), ("world", _i32), ("mode", _i32), ("reserved", _i32)]


class SrArgs(C.Structure):
    _fields_ = [("target", _vp), ("src", _vp), ("src2", _vp), ("n", _i64), ("rng_base", _i64), ("seed", _u64), ("offset", _u64),
                ("mode", _i32), ("src_dtype", _i32), ("alpha", _f32), ("flavor", _u32)]


class ShardExchangeArgs(C.Structure):
    _fields_ = [
        ("peer_base", _vp * MAX_RANKS), ("peer_partials", _vp * MAX_RANKS), ("mc_base", _vp), ("mc_partials", _vp),
        ("begin", _i64), ("end", _i64), ("dtype", _i32), ("rank", _i32), ("world", _i32), ("mode", _i32), ("blocks", _i32),
        ("partial_slot", _i32), ("grad_prescale", _f32),
    ]


# every symbol include/otb200.h declares: (name, restype, argtypes)
SYMBOLS = {
    "otb_last_error": (C.c_char_p, []),
    "otb_abi_version": (C.c_int, []),
    "otb_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "otb_adamw_sr_ema_multi": (C.c_int, [C.POINTER(OptimLaunch), _vp]),
    "otb_ema_multi": (C.c_int, [C.POINTER(OptimLaunch), _vp]),
    "otb_grad_norm_multi": (C.c_int, [C.POINTER(OptimLaunch), _f32, _f32, _vp, _vp, _vp]),
    "otb_grad_norm_from_partials": (C.c_int, [_vp, _i32, _f32, _f32, _vp, _vp]),
    "otb_grad_scale_multi": (C.c_int, [C.POINTER(OptimLaunch), _vp, _vp]),
    "otb_grad_exchange_blocks": (C.c_int, [_i64, _i32]),
    "otb_grad_exchange": (C.c_int, [C.POINTER(GradExchangeArgs), _vp]),
    "otb_adamw_sr_ema_multi_peers": (C.c_int, [C.POINTER(OptimLaunch), C.POINTER(PeerStore), _vp]),
    "otb_shard_reduce": (C.c_int, [C.POINTER(ShardExchangeArgs), _vp]),
    "otb_shard_gather": (C.c_int, [C.POINTER(ShardExchangeArgs), _vp]),
    "otb_peer_barrier": (C.c_int, [C.POINTER(_vp), _i32, _i32, _i32, _vp]),
    "otb_peer_signal": (C.c_int, [_vp, _vp]),
    "otb_peer_wait": (C.c_int, [_vp, _i32, _i32, _vp]),
    "otb_shadow_refresh_multi": (C.c_int, [_vp, _i32, _i32, _vp]),
    "otb_noise_addnoise": (C.c_int, [C.POINTER(NoiseArgs), _vp]),
    "otb_loss_fwd_bwd": (C.c_int, [C.POINTER(LossArgs), _vp]),
    "otb_loss_blocks_per_sample": (C.c_int, [_i32, _i32, _i32]),
    "otb_loss_bwd_rescale": (C.c_int, [_vp, _i32, _vp, _f32, _i32, _i64, _vp]),
    "otb_lora_linear": (C.c_int, [C.POINTER(LoraLinearArgs), _vp]),
    "otb_lora_linear_grouped": (C.c_int, [C.POINTER(LoraLinearArgs), C.c_int32, _vp]),
    "otb_lora_grad_ab": (C.c_int, [C.POINTER(LoraGradArgs), _vp]),
    "otb_lora_grad_batched": (C.c_int, [C.POINTER(LoraGradArgs), C.c_int32, _vp, _i64, C.c_int32, _vp]),
    "otb_lora_grad_workspace_bytes": (_i64, [C.POINTER(LoraGradArgs), C.c_int32]),
    "otb_dora_compose": (C.c_int, [C.POINTER(DoraArgs), _vp]),
    "otb_dora_backward": (C.c_int, [C.POINTER(DoraArgs), _vp]),
    "otb_adafactor_multi": (C.c_int, [C.POINTER(FactoredLaunch), _vp]),
    "otb_came_multi": (C.c_int, [C.POINTER(FactoredLaunch), _vp]),
    "otb_stochastic_round": (C.c_int, [C.POINTER(SrArgs), _vp]),
    "otb_philox_dump": (C.c_int, [_u64, _u64, _u32, _i64, _i64, _vp, _vp]),
    "otb_struct_sizes": (C.c_int, [C.POINTER(C.c_int32)]),
}
_STRUCTS = (OptimTensor, OptimLaunch, ShadowTensor, NoiseArgs, LossArgs, LoraLinearArgs, LoraGradArgs, GradExchangeArgs,
            FactoredTensor, FactoredLaunch, DoraArgs, ShardExchangeArgs, PeerStore, SrArgs)

_lib = None


def lib() -> C.CDLL:
    """Load libotb200.so (once).  Raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing. Build it with `python -m onetrainer_b200.build` "
                "(or __graft_entry__.build()). There is no non-CUDA fallback for this path."
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.otb_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH}: ABI version {handle.otb_abi_version()} != {ABI_VERSION}; rebuild the library")
        sizes = (C.c_int32 * len(_STRUCTS))()
        handle.otb_struct_sizes(sizes)
        for got, st in zip(sizes, _STRUCTS):
            if got != C.sizeof(st):
                raise RuntimeError(f"ABI drift: sizeof({st.__name__}) is {got} in libotb200.so but "
                                   f"{C.sizeof(st)} in the ctypes binding; rebuild the library")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    """Translate a C-ABI return code into the exceptions the reference's Python API raises."""
    if rc == OTB_OK:
        return
    msg = lib().otb_last_error().decode("utf-8", "replace")
    if rc == OTB_EUNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == OTB_EINVAL:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def dtype_code(dtype) -> int:
    import torch
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    if dtype == torch.float16:
        return F16
    raise NotImplementedError(f"dtype {dtype} is not supported by libotb200")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "onetrainer_b200 runs on sm_100a CUDA tensors only; got a tensor on "
                f"{t.device}. There is no CPU fallback for this path."
            )
