// noise_loss.cu — K-b1 (noise + timestep + add-noise + target) and K-b2 (loss fwd+bwd).
//
// Reference behaviour restated:
//   _create_noise            modules/modelSetup/mixin/ModelSetupNoiseMixin.py:18-49
//   _get_timestep_discrete   modules/modelSetup/mixin/ModelSetupNoiseMixin.py:51-155
//   _add_noise_discrete      modules/modelSetup/mixin/ModelSetupDiffusionMixin.py:15-38
//                            modules/modelSetup/mixin/ModelSetupFlowMatchingMixin.py:14-39
//   targets                  modules/modelSetup/BaseStableDiffusionXLSetup.py:277-291, BaseFluxSetup.py:307
//   _diffusion_losses / _flow_matching_losses  mixin/ModelSetupDiffusionLossMixin.py:233-321
//   masked_losses_with_prior modules/util/loss/masked_loss.py:21-45
//
// Both kernels are HBM-bound with O(1) flop/byte: 128-bit accesses, warp-shuffle reductions,
// one deterministic last-block reduction per sample.  Separate torch ops in the reference are
// separate roundings here (explicit _rn intrinsics, no contraction).
#include "common.cuh"
#include <stdlib.h>

#define NL_THREADS 256
#define NB1_UNROLL 2   /* groups of 4 elements per thread per iteration, K-b1 */
#define NB2_UNROLL 2   /* K-b2: measured 4.94 TB/s at 2 vs 4.61 (1) / 3.08 (4) on B200.  Groups of 8 elements (128-bit bf16
                          accesses) were tried at the end of round 2 and are SLOWER: 0.394 ms (one group per thread) / 0.431 ms
                          (two) against 0.379 ms for this layout at N = 2^28 (profiles/r03b_logs/kb2_ab.log) */

namespace {

// ------------------------------------------------------------------------------------
// timestep for global sample `gs` — fp32 op order of ModelSetupNoiseMixin.py:97-121
// transcendental pieces are evaluated in fp64 and rounded once to fp32 so the CPU oracle
// (libm) and the GPU agree bit-for-bit on the resulting integer.
// ------------------------------------------------------------------------------------
__device__ int32_t sample_timestep(const otb_noise_args& a, int64_t gs) {
    const int T = a.num_train_timesteps;
    if (a.ts_mode == OTB_TS_DETERMINISTIC) return (int)((double)T * 0.5) - 1;
    Philox4 r = otb_rng(a.seed, a.offset, OTB_STREAM_TIMESTEP, (uint64_t)gs);
    if (a.ts_mode == OTB_TS_TABLE) {
        // inverse CDF over the reference's weight table (ModelSetupNoiseMixin.py:122-153: multinomial, with replacement):
        // first index whose running sum exceeds u; the shift is already inside the weights
        const float u = otb_u01(r.x);
        int lo = 0, hi = a.ts_cdf_len - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (a.ts_cdf[mid] > u) hi = mid; else lo = mid + 1;
        }
        return a.min_timestep + lo;
    }
    const float num_t = (float)(a.max_timestep - a.min_timestep);
    const float min_t = (float)a.min_timestep;
    float ts;
    if (a.ts_mode == OTB_TS_UNIFORM) {
        float u = otb_u01(r.x);
        ts = __fadd_rn(min_t, __fmul_rn(num_t, u));
    } else if (a.ts_mode == OTB_TS_LOGIT_NORMAL) {
        double u1 = ((double)(r.x >> 8) + 0.5) * 5.9604644775390625e-08;
        double th = ((double)((int32_t)r.y >> 8) + 0.5) * 1.1920928955078125e-07 * 3.14159265358979323846;
        float z = (float)(sqrt(-2.0 * log(u1)) * cos(th));
        float scale = __fadd_rn(a.noising_weight, 1.0f);
        float nrm = __fadd_rn(__fmul_rn(z, scale), a.noising_bias);
        float sg = (float)(1.0 / (1.0 + exp(-(double)nrm)));
        ts = __fadd_rn(__fmul_rn(sg, num_t), min_t);
    } else {  // HEAVY_TAIL
        float u = otb_u01(r.x);
        float ang = __fmul_rn(1.57079632679489661923f, u);
        float c = (float)cos((double)ang);
        float c2 = __fmul_rn(c, c);
        float e = __fadd_rn(__fsub_rn(c2, 1.0f), u);
        float f = __fmul_rn(a.noising_weight, e);
        float h = __fsub_rn(__fsub_rn(1.0f, u), f);
        ts = __fadd_rn(__fmul_rn(h, num_t), min_t);
    }
    // timestep = T*shift*ts / ((shift-1)*ts + T)
    float num = __fmul_rn(a.shift_c1, ts);
    float den = __fadd_rn(__fmul_rn(a.shift_c2, ts), (float)T);
    ts = __fdiv_rn(num, den);
    return (int32_t)ts;   // .int(): truncation toward zero
}

__device__ __forceinline__ void load4(const void* p, int64_t i, int dtype, float (&f)[4]) {
    if (dtype == OTB_F32) {
        uint4 r = ld_stream_u4((const float*)p + i);
        f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
    } else if (dtype == OTB_BF16) {
        uint2 r = ld_stream_u2((const uint16_t*)p + i);
        f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xFFFF0000u);
        f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xFFFF0000u);
    } else {
        uint2 r = ld_stream_u2((const __half*)p + i);
        __half2 a = *reinterpret_cast<__half2*>(&r.x), b = *reinterpret_cast<__half2*>(&r.y);
        f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
    }
}
__device__ __forceinline__ void store4(void* p, int64_t i, int dtype, const float (&f)[4]) {
    if (dtype == OTB_F32) {
        st_stream_u4((float*)p + i, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
    } else if (dtype == OTB_BF16) {
        st_stream_u2((uint16_t*)p + i, make_uint2(pack2_bf16_rn(f[0], f[1]), pack2_bf16_rn(f[2], f[3])));
    } else {
        __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
        st_stream_u2((__half*)p + i, make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b)));
    }
}

__device__ __forceinline__ float rnd_dt(float x, int dtype) {
    if (dtype == OTB_BF16) return round_bf16(x);
    if (dtype == OTB_F16) return __half2float(__float2half_rn(x));
    return x;
}

// FLUX packed layout (FluxModel.pack_latents, modules/model/FluxModel.py:319-331): element (c, h, w) of a sample lives
// at token (h/2)(W/2) + w/2, channel 4c + 2(h&1) + (w&1) of [L, 4C].  A group of 4 consecutive w (w % 4 == 0) is two
// adjacent channel PAIRS of two adjacent tokens.  Returns the element offset of the first pair inside the sample; the
// second pair is 4C elements further.
__device__ __forceinline__ int64_t packed_pair_offset(uint32_t i, int C, int hw, int W) {
    const uint32_t c = i / (uint32_t)hw, rem = i - c * (uint32_t)hw;
    const uint32_t h = rem / (uint32_t)W, w = rem - h * (uint32_t)W;
    const uint32_t token = (h >> 1) * ((uint32_t)W >> 1) + (w >> 1);
    return (int64_t)token * (4 * C) + (int64_t)(4 * c + 2 * (h & 1));
}
__device__ __forceinline__ void load_packed4(const void* p, int64_t off, int C, int dtype, float (&f)[4]) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int64_t o = off + (int64_t)half * 4 * C;
        if (dtype == OTB_BF16) {
            const uint32_t r = *reinterpret_cast<const uint32_t*>((const uint16_t*)p + o);
            f[2 * half] = __uint_as_float(r << 16); f[2 * half + 1] = __uint_as_float(r & 0xFFFF0000u);
        } else if (dtype == OTB_F32) {
            const float2 r = *reinterpret_cast<const float2*>((const float*)p + o);
            f[2 * half] = r.x; f[2 * half + 1] = r.y;
        } else {
            const __half2 r = *reinterpret_cast<const __half2*>((const __half*)p + o);
            f[2 * half] = __low2float(r); f[2 * half + 1] = __high2float(r);
        }
    }
}
__device__ __forceinline__ void store_packed4(void* p, int64_t off, int C, int dtype, const float (&f)[4]) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int64_t o = off + (int64_t)half * 4 * C;
        if (dtype == OTB_BF16) *reinterpret_cast<uint32_t*>((uint16_t*)p + o) = pack2_bf16_rn(f[2 * half], f[2 * half + 1]);
        else if (dtype == OTB_F32) *reinterpret_cast<float2*>((float*)p + o) = make_float2(f[2 * half], f[2 * half + 1]);
        else *reinterpret_cast<__half2*>((__half*)p + o) = __floats2half2_rn(f[2 * half], f[2 * half + 1]);
    }
}

// four N(0,1) draws for global element group `grp` of stream `stream`
__device__ __forceinline__ void normal4(const otb_noise_args& a, uint32_t stream, uint64_t grp, float (&z)[4]) {
    Philox4 r = otb_rng(a.seed, a.offset, stream, grp);
    otb_box_muller_fast(r.x, r.y, z[0], z[1]);
    otb_box_muller_fast(r.z, r.w, z[2], z[3]);
}
__device__ __forceinline__ float normal1(const otb_noise_args& a, uint32_t stream, uint64_t elem) {
    float z[4];
    normal4(a, stream, elem >> 2, z);
    return z[elem & 3];
}

// grid = (blocks_per_sample, B); each thread owns groups of 4 consecutive elements.
// LD: compile-time latent dtype (-1 = read a.latent_dtype)
// FT: fast-path target kind (OTB_TARGET_*), or -1 for the generic kernel that reads every option at run time.
//     FT >= 0 is the training configuration — draw the noise, no offset / perturbation noise, add-noise on,
//     latent given — with every option a compile-time constant, so the loop body is the
//     Philox rounds, four MUFU ops per pair of draws and the affine maps (profiles/: 57 -> ~27 thread
//     instructions per element; the generic kernel was issue-bound at 58 % of the HBM roofline).
// PACK: noisy_cast_out is written in FLUX's packed layout (a.pack_w = latent width); compile-time so that the plain
// kernels carry none of the index arithmetic (as a run-time branch it cost K-b2 20 % of its bandwidth)
template <int U, int LD, int FT, bool PACK>
__global__ void __launch_bounds__(NL_THREADS) noise_addnoise_kernel(otb_noise_args a) {
    constexpr bool FAST = FT >= 0;
    if (a.rng_dev != nullptr) { a.seed = a.rng_dev[0]; a.offset = a.rng_dev[1]; }
    const int b = blockIdx.y;
    const int64_t chw = (int64_t)a.C * a.hw;
    const int64_t gs = a.sample_base + b;
    __shared__ int32_t s_t;
    if (threadIdx.x == 0) {
        int32_t t = (a.ts_mode == OTB_TS_GIVEN) ? a.timestep[b] : sample_timestep(a, gs);
        s_t = t;
        if (blockIdx.x == 0) {
            if (a.ts_mode != OTB_TS_GIVEN && a.timestep) a.timestep[b] = t;
            if (a.sigma_out && a.coef_noise) a.sigma_out[b] = a.coef_noise[t];
        }
    }
    __syncthreads();
    const int32_t t = s_t;
    const bool add = FAST ? true : a.do_add_noise != 0;
    const float cx = add ? a.coef_x0[t] : 0.0f;
    const float cn = add ? a.coef_noise[t] : 0.0f;
    const int ldt = LD >= 0 ? LD : a.latent_dtype;
    const bool need_noise = FAST ? true : (a.generate_noise || a.noise_in);
    const bool generate = FAST ? true : a.generate_noise != 0;
    const int target = FAST ? FT : a.target;
    const int64_t base = (int64_t)b * chw;

    if (FAST || (chw & 3) == 0) {
        const int64_t ngroups = chw >> 2;
        for (int64_t g0 = (int64_t)blockIdx.x * (NL_THREADS * U) + threadIdx.x; g0 < ngroups; g0 += (int64_t)gridDim.x * (NL_THREADS * U)) {
            float x[U][4], n[U][4];
            bool live[U];
            // all global loads of the iteration are issued before any math (memory-level parallelism)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = (g0 + (int64_t)u * NL_THREADS) << 2;
                live[u] = i < chw;
#pragma unroll
                for (int k = 0; k < 4; ++k) { x[u][k] = 0.0f; n[u][k] = 0.0f; }
                if (live[u]) {
                    if (FAST || a.latent) load4(a.latent, base + i, ldt, x[u]);
                    if (!FAST && !a.generate_noise && a.noise_in) load4(a.noise_in, base + i, ldt, n[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!live[u]) continue;
                const int64_t i = (g0 + (int64_t)u * NL_THREADS) << 2;
                if (generate) {
                    const uint64_t ge = (uint64_t)(gs * chw + i);
                    normal4(a, OTB_STREAM_NOISE, ge >> 2, n[u]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) n[u][k] = rnd_dt(n[u][k], ldt);
                    if (!FAST && a.offset_noise_weight > 0.0f) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            int c = (int)((uint32_t)(i + k) / (uint32_t)a.hw);
                            float on = rnd_dt(normal1(a, OTB_STREAM_OFFSET, (uint64_t)(gs * a.C + c)), ldt);
                            n[u][k] = rnd_dt(__fadd_rn(n[u][k], rnd_dt(__fmul_rn(a.offset_noise_weight, on), ldt)), ldt);
                        }
                    }
                    if (!FAST && a.perturbation_noise_weight > 0.0f) {
                        float pn[4];
                        normal4(a, OTB_STREAM_PERTURB, ge >> 2, pn);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            n[u][k] = rnd_dt(__fadd_rn(n[u][k], rnd_dt(__fmul_rn(a.perturbation_noise_weight, rnd_dt(pn[k], ldt)), ldt)), ldt);
                    }
                }
                if (a.noise_out && need_noise) store4(a.noise_out, base + i, ldt, n[u]);
                if (!FAST && !a.latent) continue;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    x[u][k] = rnd_dt(__fmul_rn((FAST && FT != OTB_TARGET_FLOW) ? x[u][k]      // eps / v fast paths: shift == 0, x - 0 == x
                                                                              : rnd_dt(__fsub_rn(x[u][k], a.latent_shift), ldt), a.latent_scale), ldt);
                if (a.x0_out) store4(a.x0_out, base + i, ldt, x[u]);
                if (add) {
                    float y[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) y[k] = __fadd_rn(__fmul_rn(x[u][k], cx), __fmul_rn(n[u][k], cn));
                    if (a.noisy_out) store4(a.noisy_out, base + i, ldt, y);
                    if (a.noisy_cast_out) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) y[k] = rnd_dt(y[k], ldt);
                        if constexpr (PACK) store_packed4(a.noisy_cast_out, base + packed_pair_offset((uint32_t)i, a.C, a.hw, a.pack_w), a.C, a.cast_dtype, y);
                        else store4(a.noisy_cast_out, base + i, a.cast_dtype, y);
                    }
                }
                if (a.target_out && target != OTB_TARGET_NONE) {
                    float tg[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (target == OTB_TARGET_EPSILON) tg[k] = n[u][k];
                        else if (target == OTB_TARGET_V) tg[k] = __fsub_rn(__fmul_rn(cx, n[u][k]), __fmul_rn(cn, x[u][k]));
                        else tg[k] = __fsub_rn(n[u][k], x[u][k]);
                    }
                    store4(a.target_out, base + i, ldt, tg);
                }
            }
        }
    } else if constexpr (!FAST) {
        // ragged tail path (chw not a multiple of 4): one element per thread, same RNG indexing
        for (int64_t i = (int64_t)blockIdx.x * NL_THREADS + threadIdx.x; i < chw; i += (int64_t)gridDim.x * NL_THREADS) {
            float n = 0.0f;
            if (a.generate_noise) {
                const uint64_t ge = (uint64_t)(gs * chw + i);
                n = rnd_dt(normal1(a, OTB_STREAM_NOISE, ge), ldt);
                if (a.offset_noise_weight > 0.0f) {
                    int c = (int)(i / a.hw);
                    float on = rnd_dt(normal1(a, OTB_STREAM_OFFSET, (uint64_t)(gs * a.C + c)), ldt);
                    n = rnd_dt(__fadd_rn(n, rnd_dt(__fmul_rn(a.offset_noise_weight, on), ldt)), ldt);
                }
                if (a.perturbation_noise_weight > 0.0f) {
                    float pn = rnd_dt(normal1(a, OTB_STREAM_PERTURB, ge), ldt);
                    n = rnd_dt(__fadd_rn(n, rnd_dt(__fmul_rn(a.perturbation_noise_weight, pn), ldt)), ldt);
                }
            } else if (a.noise_in) {
                n = load_as_float(a.noise_in, base + i, ldt);
            }
            if (a.noise_out && need_noise) store_from_float(a.noise_out, base + i, ldt, n);
            if (!a.latent) continue;
            float x = load_as_float(a.latent, base + i, ldt);
            x = rnd_dt(__fmul_rn(rnd_dt(__fsub_rn(x, a.latent_shift), ldt), a.latent_scale), ldt);
            if (a.x0_out) store_from_float(a.x0_out, base + i, ldt, x);
            if (add) {
                float y = __fadd_rn(__fmul_rn(x, cx), __fmul_rn(n, cn));
                if (a.noisy_out) store_from_float(a.noisy_out, base + i, ldt, y);
                if (a.noisy_cast_out) store_from_float(a.noisy_cast_out, base + i, a.cast_dtype, rnd_dt(y, ldt));
            }
            if (a.target_out && a.target != OTB_TARGET_NONE) {
                float tg = a.target == OTB_TARGET_EPSILON ? n
                         : a.target == OTB_TARGET_V ? __fsub_rn(__fmul_rn(cx, n), __fmul_rn(cn, x))
                                                    : __fsub_rn(n, x);
                store_from_float(a.target_out, base + i, ldt, tg);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// K-b2
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* sm) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = threadIdx.x < NL_THREADS / 32 ? sm[threadIdx.x] : 0.0f;
    if (threadIdx.x < 32) r = warp_sum(r);
    __syncthreads();
    return r;  // valid in warp 0
}

// mean(clamp(mask, uw, 1)) and mean(1 - clamp(mask, uw, 1)) per sample; grid = B
__global__ void __launch_bounds__(NL_THREADS) mask_stats_kernel(const void* mask, int mask_dtype, int hw, float uw, float* stats) {
    __shared__ float sm[NL_THREADS / 32];
    const int b = blockIdx.x;
    float s0 = 0.0f, s1 = 0.0f;
    for (int i = threadIdx.x; i < hw; i += NL_THREADS) {
        float m = load_as_float(mask, (int64_t)b * hw + i, mask_dtype);
        m = fminf(fmaxf(m, uw), 1.0f);
        s0 += m;
        s1 += __fsub_rn(1.0f, m);
    }
    s0 = block_sum(s0, sm);
    s1 = block_sum(s1, sm);
    if (threadIdx.x == 0) {
        stats[2 * b] = s0 / (float)hw;
        stats[2 * b + 1] = s1 / (float)hw;
    }
}

// timestep weight of sample b — ModelSetupDiffusionLossMixin.py:170-231
__device__ float timestep_weight(const otb_loss_args& a, int b) {
    if (a.loss_weight_fn == OTB_LW_CONSTANT || a.timestep == nullptr) return 1.0f;
    const int t = a.timestep[b];
    if (a.loss_weight_fn == OTB_LW_SIGMA) return a.sigmas ? a.sigmas[t] : 1.0f;
    if (a.sqrt_ac == nullptr) return 1.0f;
    float q = __fdiv_rn(a.sqrt_ac[t], a.sqrt_1mac[t]);
    float snr = __fmul_rn(q, q);
    if (a.loss_weight_fn == OTB_LW_MIN_SNR) {
        float mn = fminf(snr, a.loss_weight_strength);
        if (a.v_prediction) snr = __fadd_rn(snr, 1.0f);
        return __fdiv_rn(mn, snr);
    }
    if (a.loss_weight_fn == OTB_LW_DEBIASED) {
        float w = fminf(snr, 1.0e3f);
        if (a.v_prediction) w = __fadd_rn(w, 1.0f);
        return __fdiv_rn(1.0f, __fsqrt_rn(w));
    }
    // P2
    if (a.v_prediction) snr = __fadd_rn(snr, 1.0f);
    return powf(__fadd_rn(1.0f, snr), -a.loss_weight_strength);
}

// grid = (blocks_per_sample, B).  PD / TD: compile-time predicted / target dtype (-1 = read a.*_dtype)
// FEAT: -1 = every option read at run time; otherwise a compile-time feature set — bit 0 masked, bit 1 prior
// preservation term, bit 2 MAE term — so the training configurations (FEAT 0: plain MSE, FEAT 1: masked MSE,
// FEAT 3: masked MSE + prior) carry no dead branches or dead registers.
// PACK: predicted / dpred are in FLUX's packed layout (a.pred_pack_w = latent width), compile-time like the features
template <int U, int PD, int TD, int FEAT, bool PACK>
__global__ void __launch_bounds__(NL_THREADS, (FEAT >= 0 && U <= 2 && !PACK) ? 6 : 1) loss_fwd_bwd_kernel(otb_loss_args a) {
    const int pred_dt = PD >= 0 ? PD : a.pred_dtype;
    const int tgt_dt = TD >= 0 ? TD : a.target_dtype;
    constexpr bool FIXED = FEAT >= 0;
    __shared__ float sm[NL_THREADS / 32];
    __shared__ bool s_last;
    const int b = blockIdx.y;
    const int64_t chw = (int64_t)a.C * a.hw;
    const int64_t base = (int64_t)b * chw;
    const bool masked = FIXED ? (FEAT & 1) != 0 : (a.masked != 0 && a.mask != nullptr);
    const bool prior = FIXED ? (FEAT & 2) != 0 : (masked && a.prior_target != nullptr && a.masked_prior_preservation_weight != 0.0f);
    const bool mae = FIXED ? (FEAT & 4) != 0 : a.mae_strength != 0.0f;
    // per-sample constants
    float inv_mm = 1.0f, inv_mp = 1.0f;
    if (a.normalize_masked_area_loss && a.mask != nullptr) {
        inv_mm = 1.0f / a.mask_stats[2 * b];
        inv_mp = 1.0f / a.mask_stats[2 * b + 1];
    }
    const float wt = timestep_weight(a, b);
    const float lw = a.loss_weight ? a.loss_weight[b] : 1.0f;
    const float inv_n = 1.0f / (float)chw;
    // post factor applied to the per-sample mean: scale * loss_weight * timestep weight
    const float post = a.loss_scale * lw * wt;
    // coefficients of the element terms
    const float k_main = masked ? inv_mm : ((a.normalize_masked_area_loss && a.mask) ? inv_mm : 1.0f);
    const float k_prior = prior ? a.masked_prior_preservation_weight * inv_mp : 0.0f;
    const float gfac = post * inv_n * a.upstream;        // multiplies d(elem term)/d(pred)
    const bool want_grad = a.dpred != nullptr;

    float acc = 0.0f;
    if ((chw & 3) == 0 && (a.hw & 3) == 0) {
        const int64_t ngroups = chw >> 2;
                for (int64_t g0 = (int64_t)blockIdx.x * (NL_THREADS * U) + threadIdx.x; g0 < ngroups; g0 += (int64_t)gridDim.x * (NL_THREADS * U)) {
            float p[U][4], t[U][4], pt[U][4], m[U][4];
            bool live[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = (g0 + (int64_t)u * NL_THREADS) << 2;
                live[u] = i < chw;
#pragma unroll
                for (int k = 0; k < 4; ++k) { p[u][k] = 0.0f; t[u][k] = 0.0f; pt[u][k] = 0.0f; m[u][k] = 1.0f; }
                if (!live[u]) continue;
                if constexpr (PACK) load_packed4(a.predicted, base + packed_pair_offset((uint32_t)i, a.C, a.hw, a.pred_pack_w), a.C, pred_dt, p[u]);
                else load4(a.predicted, base + i, pred_dt, p[u]);
                load4(a.target, base + i, tgt_dt, t[u]);
                if (prior) load4(a.prior_target, base + i, tgt_dt, pt[u]);
                if (masked) {
                    const int64_t mi = (int64_t)b * a.hw + (int64_t)((uint32_t)i % (uint32_t)a.hw);
                    if (a.mask_dtype == OTB_F32) {   // mask is re-read C times: keep it cacheable
                        float4 mv = *reinterpret_cast<const float4*>((const float*)a.mask + mi);
                        m[u][0] = mv.x; m[u][1] = mv.y; m[u][2] = mv.z; m[u][3] = mv.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) m[u][k] = load_as_float(a.mask, mi + k, a.mask_dtype);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!live[u]) continue;
                const int64_t i = (g0 + (int64_t)u * NL_THREADS) << 2;
                float g[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float mk = masked ? fminf(fmaxf(m[u][k], a.unmasked_weight), 1.0f) : 1.0f;
                    const float d = __fsub_rn(p[u][k], t[u][k]);
                    const float wm = mk * k_main;
                    float term = a.mse_strength * (d * d) * wm;
                    float gr = a.mse_strength * 2.0f * d * wm;
                    if (mae) {
                        term += a.mae_strength * fabsf(d) * wm;
                        gr += a.mae_strength * (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * wm;
                    }
                    if (prior) {
                        const float dp = __fsub_rn(p[u][k], pt[u][k]);
                        const float wp = __fsub_rn(1.0f, mk) * k_prior;
                        term += a.mse_strength * (dp * dp) * wp;
                        gr += a.mse_strength * 2.0f * dp * wp;
                        if (mae) {
                            term += a.mae_strength * fabsf(dp) * wp;
                            gr += a.mae_strength * (dp > 0.0f ? 1.0f : (dp < 0.0f ? -1.0f : 0.0f)) * wp;
                        }
                    }
                    acc += term;
                    g[k] = gr * gfac;
                }
                if (want_grad) {
                    if constexpr (PACK) store_packed4(a.dpred, base + packed_pair_offset((uint32_t)i, a.C, a.hw, a.pred_pack_w), a.C, pred_dt, g);
                    else store4(a.dpred, base + i, pred_dt, g);
                }
            }
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * NL_THREADS + threadIdx.x; i < chw; i += (int64_t)gridDim.x * NL_THREADS) {
            const float p = load_as_float(a.predicted, base + i, pred_dt);
            const float t = load_as_float(a.target, base + i, tgt_dt);
            float m = 1.0f;
            if (masked) m = fminf(fmaxf(load_as_float(a.mask, (int64_t)b * a.hw + (i % a.hw), a.mask_dtype), a.unmasked_weight), 1.0f);
            const float d = __fsub_rn(p, t);
            const float wm = m * k_main;
            float term = a.mse_strength * (d * d) * wm;
            float gr = a.mse_strength * 2.0f * d * wm;
            if (mae) {
                term += a.mae_strength * fabsf(d) * wm;
                gr += a.mae_strength * (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * wm;
            }
            if (prior) {
                const float dp = __fsub_rn(p, load_as_float(a.prior_target, base + i, tgt_dt));
                const float wp = __fsub_rn(1.0f, m) * k_prior;
                term += a.mse_strength * (dp * dp) * wp;
                gr += a.mse_strength * 2.0f * dp * wp;
                if (mae) {
                    term += a.mae_strength * fabsf(dp) * wp;
                    gr += a.mae_strength * (dp > 0.0f ? 1.0f : (dp < 0.0f ? -1.0f : 0.0f)) * wp;
                }
            }
            acc += term;
            if (want_grad) store_from_float(a.dpred, base + i, pred_dt, gr * gfac);
        }
    }
    acc = block_sum(acc, sm);
    // deterministic cross-block reduction: the last block to arrive sums the partials in order
    if (threadIdx.x == 0) {
        a.partials[(int64_t)b * gridDim.x + blockIdx.x] = acc;
        __threadfence();
        unsigned prev = atomicAdd(&a.counters[b], 1u);
        s_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float s = 0.0f;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += NL_THREADS)
            s += __ldcg(&a.partials[(int64_t)b * gridDim.x + k]);
        s = block_sum(s, sm);
        if (threadIdx.x == 0) {
            a.losses[b] = s * inv_n * post;
            a.counters[b] = 0;
        }
    }
}

__global__ void __launch_bounds__(NL_THREADS) loss_bwd_rescale_kernel(void* dpred, int dtype, const float* grad_out,
                                                                      float assumed, int64_t chw) {
    const int b = blockIdx.y;
    const float go = grad_out[b];
    if (go == assumed) return;              // the normal case: nothing to do, nothing touched
    const float ratio = go / assumed;
    const int64_t base = (int64_t)b * chw;
    for (int64_t i = (int64_t)blockIdx.x * NL_THREADS + threadIdx.x; i < chw; i += (int64_t)gridDim.x * NL_THREADS)
        store_from_float(dpred, base + i, dtype, load_as_float(dpred, base + i, dtype) * ratio);
}

// tuning knob for A/B runs on the GPU box (unset in production: the compiled default is used)
int tuning_unroll(const char* env, int dflt) {
    const char* v = getenv(env);
    if (!v) return dflt;
    int u = atoi(v);
    return (u == 1 || u == 2 || u == 4) ? u : dflt;
}

int blocks_per_sample_for(int64_t chw, int B, int unroll) {
    // One block-iteration (a "slab") is 4*unroll elements per thread.  A block that loops over several slabs
    // serialises their memory latency (load -> math -> store -> next load), which is what bounded the first
    // version at ~60 % of the HBM roofline (profiles/r01_ncu_kb2.md: 21 long-scoreboard stalls per issue), so
    // every slab gets its own block and the block scheduler provides the pipelining; only beyond ~64 waves of
    // resident blocks do blocks start to loop.
    int64_t slabs = (chw / 4 + NL_THREADS * unroll - 1) / (NL_THREADS * unroll);
    if (slabs < 1) slabs = 1;
    // A/B knob.  Measured at N = 2^28 (profiles/r03c_logs/ab.log): lifting the cap so that every block runs exactly one slab
    // is SLOWER — K-b2 0.378 -> 0.483 ms, K-b1 0.470 -> 0.517 ms — the per-block prologue (timestep weight, barriers, the
    // partial-sum hand-off) outweighs the serialised second iteration of the looping blocks.
    int waves = 64;
    if (const char* v = getenv("OTB_NB_WAVES")) { const int w = atoi(v); if (w >= 1 && w <= 4096) waves = w; }
    int64_t want = (int64_t)otb_sm_count() * 8 * waves / (B > 0 ? B : 1);
    if (want < 1) want = 1;
    int64_t bps = slabs < want ? slabs : want;
    if (bps > 65535) bps = 65535;
    return (int)bps;
}

}  // namespace

extern "C" int otb_noise_addnoise(const otb_noise_args* a, void* stream) {
    OTB_CHECK_ARG(a != nullptr, "null args");
    OTB_CHECK_ARG(a->B > 0 && a->C > 0 && a->hw > 0 && (int64_t)a->C * a->hw < (1ll << 31), "bad shape B=%d C=%d hw=%d", a->B, a->C, a->hw);
    OTB_CHECK_ARG(a->B <= 65535, "batch %d too large for one launch", a->B);
    OTB_CHECK_ARG(a->latent_dtype >= OTB_F32 && a->latent_dtype <= OTB_F16, "bad latent dtype");
    OTB_CHECK_ARG(a->ts_mode >= OTB_TS_UNIFORM && a->ts_mode <= OTB_TS_TABLE, "bad timestep mode %d", a->ts_mode);
    OTB_CHECK_ARG(a->ts_mode != OTB_TS_TABLE || (a->ts_cdf != nullptr && a->ts_cdf_len > 0), "OTB_TS_TABLE needs ts_cdf / ts_cdf_len");
    OTB_CHECK_ARG(a->timestep != nullptr || (!a->do_add_noise && a->ts_mode != OTB_TS_GIVEN && a->target != OTB_TARGET_V),
                  "timestep buffer required");
    if (a->do_add_noise || a->target == OTB_TARGET_V) {
        OTB_CHECK_ARG(a->coef_x0 && a->coef_noise, "schedule tables required for add-noise / v-target");
        OTB_CHECK_ARG(a->latent != nullptr, "latent required for add-noise");
    }
    OTB_CHECK_ARG(!(a->generate_noise && a->noise_in), "generate_noise and noise_in are mutually exclusive");
    OTB_CHECK_ARG(a->num_train_timesteps > 0, "num_train_timesteps must be positive");
    if (a->ts_mode != OTB_TS_GIVEN && a->ts_mode != OTB_TS_DETERMINISTIC)
        OTB_CHECK_ARG(a->max_timestep >= a->min_timestep && a->min_timestep >= 0 && a->max_timestep <= a->num_train_timesteps,
                      "bad timestep range [%d,%d)", a->min_timestep, a->max_timestep);
    const int64_t chw = (int64_t)a->C * a->hw;
    if (a->pack_w != 0) {
        OTB_CHECK_ARG(a->pack_w > 0 && a->hw % a->pack_w == 0 && a->pack_w % 4 == 0 && (a->hw / a->pack_w) % 2 == 0,
                      "pack_w %d: the latent must be H x W with W %% 4 == 0 and H even (hw = %d)", a->pack_w, a->hw);
        OTB_CHECK_ARG(a->noisy_cast_out != nullptr, "pack_w applies to noisy_cast_out, which is null");
    }
    OTB_BIND(a->timestep ? (const void*)a->timestep : (a->latent ? a->latent : a->noise_out));
    const int u = tuning_unroll("OTB_NB1_UNROLL", NB1_UNROLL);
    dim3 grid(blocks_per_sample_for(chw, a->B, u), a->B);
    cudaStream_t st = (cudaStream_t)stream;
    // the training configuration takes the specialised kernel (see the template comment)
    const bool fast = a->latent_dtype == OTB_F32 && a->generate_noise && a->latent && a->do_add_noise && !a->noise_in &&
                      a->offset_noise_weight <= 0.0f && a->perturbation_noise_weight <= 0.0f &&
                      (a->latent_shift == 0.0f || a->target == OTB_TARGET_FLOW) &&
                      (chw & 3) == 0 && a->target != OTB_TARGET_NONE && getenv("OTB_NB1_GENERIC") == nullptr;
#define OTB_NB1_LAUNCH(UU, FT) noise_addnoise_kernel<UU, OTB_F32, FT, false><<<grid, NL_THREADS, 0, st>>>(*a)
#define OTB_NB1_BY_TARGET(UU)                                                   \
    do {                                                                        \
        if (!fast) OTB_NB1_LAUNCH(UU, -1);                                      \
        else if (a->target == OTB_TARGET_EPSILON) OTB_NB1_LAUNCH(UU, OTB_TARGET_EPSILON); \
        else if (a->target == OTB_TARGET_V) OTB_NB1_LAUNCH(UU, OTB_TARGET_V);   \
        else OTB_NB1_LAUNCH(UU, OTB_TARGET_FLOW);                               \
    } while (0)
    if (a->pack_w > 0) {
        // packed network input: the FLUX training configuration (fp32 latents, flow target) or the generic kernel
        OTB_UNSUPPORTED((chw & 3) != 0, "pack_w needs C*H*W to be a multiple of 4");
        if (fast && a->target == OTB_TARGET_FLOW) noise_addnoise_kernel<NB1_UNROLL, OTB_F32, OTB_TARGET_FLOW, true><<<grid, NL_THREADS, 0, st>>>(*a);
        else noise_addnoise_kernel<NB1_UNROLL, -1, -1, true><<<grid, NL_THREADS, 0, st>>>(*a);
    } else if (a->latent_dtype == OTB_F32) {
        if (u == 1) OTB_NB1_BY_TARGET(1);
        else if (u == 4) OTB_NB1_BY_TARGET(4);
        else OTB_NB1_BY_TARGET(2);
    } else {
        noise_addnoise_kernel<NB1_UNROLL, -1, -1, false><<<grid, NL_THREADS, 0, st>>>(*a);
    }
#undef OTB_NB1_BY_TARGET
#undef OTB_NB1_LAUNCH
    OTB_LAUNCH_CHECK("noise_addnoise_kernel");
    return OTB_OK;
}

extern "C" int otb_loss_fwd_bwd(const otb_loss_args* a, void* stream) {
    OTB_CHECK_ARG(a != nullptr, "null args");
    OTB_CHECK_ARG(a->B > 0 && a->C > 0 && a->hw > 0 && a->B <= 65535 && (int64_t)a->C * a->hw < (1ll << 31), "bad shape B=%d C=%d hw=%d", a->B, a->C, a->hw);
    OTB_CHECK_ARG(a->predicted && a->target && a->losses && a->partials && a->counters, "null tensor");
    OTB_CHECK_ARG(a->blocks_per_sample > 0 && a->blocks_per_sample <= 65535, "bad blocks_per_sample");
    OTB_CHECK_ARG(!(a->masked && !a->mask), "masked loss needs a mask");
    if (a->pred_pack_w != 0)
        OTB_CHECK_ARG(a->pred_pack_w > 0 && a->hw % a->pred_pack_w == 0 && a->pred_pack_w % 4 == 0 && (a->hw / a->pred_pack_w) % 2 == 0,
                      "pred_pack_w %d: the latent must be H x W with W %% 4 == 0 and H even (hw = %d)", a->pred_pack_w, a->hw);
    if (a->loss_weight_fn == OTB_LW_MIN_SNR || a->loss_weight_fn == OTB_LW_DEBIASED || a->loss_weight_fn == OTB_LW_P2)
        OTB_CHECK_ARG(a->timestep == nullptr || (a->sqrt_ac && a->sqrt_1mac), "snr weighting needs the schedule tables");
    OTB_BIND(a->predicted);
    if (a->normalize_masked_area_loss && a->mask) {
        OTB_CHECK_ARG(a->mask_stats != nullptr, "mask_stats workspace required");
        mask_stats_kernel<<<a->B, NL_THREADS, 0, (cudaStream_t)stream>>>(a->mask, a->mask_dtype, a->hw, a->unmasked_weight, a->mask_stats);
        OTB_LAUNCH_CHECK("mask_stats_kernel");
    }
    dim3 grid(a->blocks_per_sample, a->B);
    cudaStream_t st = (cudaStream_t)stream;
    const bool masked = a->masked != 0 && a->mask != nullptr;
    const bool prior = masked && a->prior_target != nullptr && a->masked_prior_preservation_weight != 0.0f;
    int feat = (masked ? 1 : 0) | (prior ? 2 : 0) | (a->mae_strength != 0.0f ? 4 : 0);
    if ((feat != 0 && feat != 1 && feat != 3) || getenv("OTB_NB2_GENERIC") != nullptr) feat = -1;
    if (a->pred_pack_w > 0) {                  // packed prediction (FLUX): plain MSE specialisation or the generic kernel
        if (feat == 0 && a->pred_dtype == OTB_BF16 && a->target_dtype == OTB_F32)
            loss_fwd_bwd_kernel<NB2_UNROLL, OTB_BF16, OTB_F32, 0, true><<<grid, NL_THREADS, 0, st>>>(*a);
        else
            loss_fwd_bwd_kernel<NB2_UNROLL, -1, -1, -1, true><<<grid, NL_THREADS, 0, st>>>(*a);
    }
    else if (a->pred_dtype == OTB_BF16 && a->target_dtype == OTB_F32) {     // the training configuration
        const int u = tuning_unroll("OTB_NB2_UNROLL", NB2_UNROLL);
#define OTB_NB2_LAUNCH(UU, FF) loss_fwd_bwd_kernel<UU, OTB_BF16, OTB_F32, FF, false><<<grid, NL_THREADS, 0, st>>>(*a)
#define OTB_NB2_BY_FEAT(UU)                          \
    do {                                             \
        if (feat == 0) OTB_NB2_LAUNCH(UU, 0);        \
        else if (feat == 1) OTB_NB2_LAUNCH(UU, 1);   \
        else if (feat == 3) OTB_NB2_LAUNCH(UU, 3);   \
        else OTB_NB2_LAUNCH(UU, -1);                 \
    } while (0)
        if (u == 1) OTB_NB2_BY_FEAT(1);
        else if (u == 2) OTB_NB2_BY_FEAT(2);
        else OTB_NB2_BY_FEAT(4);
#undef OTB_NB2_BY_FEAT
#undef OTB_NB2_LAUNCH
    }
    else if (a->pred_dtype == OTB_F32 && a->target_dtype == OTB_F32)
        loss_fwd_bwd_kernel<NB2_UNROLL, OTB_F32, OTB_F32, -1, false><<<grid, NL_THREADS, 0, st>>>(*a);
    else
        loss_fwd_bwd_kernel<NB2_UNROLL, -1, -1, -1, false><<<grid, NL_THREADS, 0, st>>>(*a);
    OTB_LAUNCH_CHECK("loss_fwd_bwd_kernel");
    return OTB_OK;
}

extern "C" int otb_loss_blocks_per_sample(int32_t B, int32_t C, int32_t hw) {
    return blocks_per_sample_for((int64_t)C * hw, B, tuning_unroll("OTB_NB2_UNROLL", NB2_UNROLL));
}

extern "C" int otb_loss_bwd_rescale(void* dpred, int32_t dtype, const float* grad_out, float assumed,
                                    int32_t B, int64_t chw, void* stream) {
    OTB_CHECK_ARG(dpred && grad_out && B > 0 && chw > 0 && assumed != 0.0f, "bad rescale args");
    OTB_BIND(dpred);
    dim3 grid(blocks_per_sample_for(chw, B, 1), B);
    loss_bwd_rescale_kernel<<<grid, NL_THREADS, 0, (cudaStream_t)stream>>>(dpred, dtype, grad_out, assumed, chw);
    OTB_LAUNCH_CHECK("loss_bwd_rescale_kernel");
    return OTB_OK;
}
