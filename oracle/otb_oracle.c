/*
 * otb_oracle.c — CPU restatement of the reference's algorithm for the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under onetrainer_b200/ may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker or as the timed CPU baseline — never as the product.
 *
 * It follows the reference (sangoi-exe/OneTrainer, 100 % Python/torch) operation by operation:
 * every torch op of the reference is one loop here, with one rounding to the tensor dtype at
 * the end of each loop, exactly like a chain of eager kernels.  Compile with
 *   gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -fno-fast-math
 * so the compiler never fuses a multiply-add: FMA appears only where written (fmaf), selected
 * by the same `flavor` bits the CUDA library takes (include/otb200.h).
 *
 * Pinned against the reference itself: tests/golden/*.npz are produced by
 * tools/make_golden.py, which imports the unmodified reference modules from /root/reference
 * and records their outputs; tests/test_oracle_golden.py replays them through this file.
 *
 *   adam_step            modules/util/optimizer/adamw_extensions.py:79-146,
 *                        modules/util/optimizer/adam_extensions.py:71-90,
 *                        modules/util/bf16_stochastic_rounding.py:5-30,48-61
 *   sr_helper            modules/util/bf16_stochastic_rounding.py:5-61 (copy_ / add_ / addcdiv_stochastic_)
 *   ema_step             modules/module/EMAModule.py:41-47
 *   clip (scale part)    torch.nn.utils.clip_grad_norm_ as called at GenericTrainer.py:712-713
 *   timesteps            modules/modelSetup/mixin/ModelSetupNoiseMixin.py:51-155
 *   add_noise / targets  mixin/ModelSetupDiffusionMixin.py:15-38, ModelSetupFlowMatchingMixin.py:14-39,
 *                        BaseStableDiffusionXLSetup.py:215,277-291, BaseFluxSetup.py:230,307
 *   losses               mixin/ModelSetupDiffusionLossMixin.py:36-321, util/loss/masked_loss.py:5-45
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FL_DIV_RECIPROCAL   1u
#define FL_ADDCDIV_MUL_LAST 2u
#define FL_FMA_LERP         4u
#define FL_FMA_ADDCMUL      8u
#define FL_FMA_ADDCDIV      16u
#define FL_FMA_ADD_ALPHA    32u
#define FL_ADDCMUL_SQUARE_FIRST 64u
#define FL_ALPHA_IN_DTYPE   128u

enum { DT_F32 = 0, DT_BF16 = 1 };

/* ---------------- bf16 <-> fp32 (round to nearest even, torch's c10::BFloat16) ---------------- */
static inline float bits_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bf16_to_f(uint16_t h) { return bits_f((uint32_t)h << 16); }
static inline uint16_t f_to_bf16(float f) {
    uint32_t u = f_bits(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;               /* NaN */
    uint32_t bias = 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)((u + bias) >> 16);
}
static inline float ld(const void* p, int64_t i, int dt) {
    return dt == DT_F32 ? ((const float*)p)[i] : bf16_to_f(((const uint16_t*)p)[i]);
}
static inline void st(void* p, int64_t i, int dt, float v) {
    if (dt == DT_F32) ((float*)p)[i] = v; else ((uint16_t*)p)[i] = f_to_bf16(v);
}
static inline float rnd(float v, int dt) { return dt == DT_F32 ? v : bf16_to_f(f_to_bf16(v)); }

/* ---------------- Philox4x32-10 ---------------- */
void otb_o_philox(uint64_t seed, uint64_t offset, uint32_t stream, uint64_t group, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)group, c1 = (uint32_t)(group >> 32), c2 = (uint32_t)offset;
    uint32_t c3 = (stream << 24) | ((uint32_t)(offset >> 32) & 0x00ffffffu);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* raw counter/key interface for the published Random123 known-answer vectors */
void otb_o_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* the 16 random bits element `elem` of the stochastic-rounding stream receives */
static inline uint32_t sr_rand16(uint64_t seed, uint64_t offset, uint64_t elem) {
    uint32_t r[4];
    otb_o_philox(seed, offset, 4u, elem >> 3, r);
    uint32_t w = r[(elem & 7) >> 1];
    return (elem & 1) ? (w >> 16) : (w & 0xffffu);
}
void otb_o_sr_bits(uint64_t seed, uint64_t offset, int64_t first_elem, int64_t n, int32_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)sr_rand16(seed, offset, (uint64_t)(first_elem + i));
}

/* ---------------- optimizer ---------------- */
typedef struct {
    float decay_mul, l2, lerp_w, beta2, one_minus_beta2, bc2_sqrt, inv_bc2_sqrt, eps, neg_step_size;
} adam_scalars;

/* One reference step over one tensor, one torch op per loop.
 * tmp_g / tmp_d are caller-provided fp32 scratch of n elements. */
void otb_o_adam_step(void* p, const void* g_in, void* m, void* v, int64_t n, int dt, const adam_scalars* s,
                     uint32_t flavor, int sr, int maximize, uint64_t seed, uint64_t offset, int64_t rng_base,
                     int has_clip, float clip_coef, float* tmp_g, float* tmp_d) {
    int64_t i;
    /* grad (optionally negated), clip_grad_norm_'s in-place g.mul_(clip_coef) */
#pragma omp parallel for
    for (i = 0; i < n; ++i) {
        float g = ld(g_in, i, dt);
        if (maximize) g = -g;
        if (has_clip) g = rnd(g * clip_coef, dt);
        tmp_g[i] = g;
    }
    if (s->l2 != 0.0f) {
        /* Adam: grad = grad.add(p, alpha=weight_decay); torch's CPU kernel narrows alpha to the tensor dtype */
        const float l2 = (flavor & FL_ALPHA_IN_DTYPE) ? rnd(s->l2, dt) : s->l2;
#pragma omp parallel for
        for (i = 0; i < n; ++i) {
            float pv = ld(p, i, dt);
            tmp_g[i] = rnd((flavor & FL_FMA_ADD_ALPHA) ? fmaf(l2, pv, tmp_g[i]) : tmp_g[i] + l2 * pv, dt);
        }
    } else {
        /* AdamW: p.mul_(1 - lr * weight_decay) */
#pragma omp parallel for
        for (i = 0; i < n; ++i) st(p, i, dt, ld(p, i, dt) * s->decay_mul);
    }
    /* exp_avg.lerp_(grad, 1 - beta1)   (ATen lerp: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)) */
#pragma omp parallel for
    for (i = 0; i < n; ++i) {
        float a = ld(m, i, dt), b = tmp_g[i], d = b - a, r;
        if (s->lerp_w < 0.5f) r = (flavor & FL_FMA_LERP) ? fmaf(s->lerp_w, d, a) : a + s->lerp_w * d;
        else { float w1 = 1.0f - s->lerp_w; r = (flavor & FL_FMA_LERP) ? fmaf(-w1, d, b) : b - d * w1; }
        st(m, i, dt, r);
    }
    /* exp_avg_sq.mul_(beta2) */
#pragma omp parallel for
    for (i = 0; i < n; ++i) st(v, i, dt, ld(v, i, dt) * s->beta2);
    /* .addcmul_(grad, grad, value=1 - beta2) */
#pragma omp parallel for
    for (i = 0; i < n; ++i) {
        float gv = tmp_g[i], a = ld(v, i, dt), x, y;
        if (flavor & FL_ADDCMUL_SQUARE_FIRST) { x = s->one_minus_beta2; y = gv * gv; }   /* a + alpha * (b * c) */
        else { x = s->one_minus_beta2 * gv; y = gv; }                                     /* a + alpha * b * c   */
        st(v, i, dt, (flavor & FL_FMA_ADDCMUL) ? fmaf(x, y, a) : a + x * y);
    }
    /* denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps) — three ops, three roundings */
#pragma omp parallel for
    for (i = 0; i < n; ++i) tmp_d[i] = rnd(sqrtf(ld(v, i, dt)), dt);
#pragma omp parallel for
    for (i = 0; i < n; ++i)
        tmp_d[i] = rnd((flavor & FL_DIV_RECIPROCAL) ? tmp_d[i] * s->inv_bc2_sqrt : tmp_d[i] / s->bc2_sqrt, dt);
#pragma omp parallel for
    for (i = 0; i < n; ++i) tmp_d[i] = rnd(tmp_d[i] + s->eps, dt);
    /* p.addcdiv_(exp_avg, denom, value=-step_size), or the stochastic-rounding variant:
     * result = p.float(); result.addcdiv_(...); copy_stochastic_(p, result) */
#pragma omp parallel for
    for (i = 0; i < n; ++i) {
        float pv = ld(p, i, dt), mv = ld(m, i, dt), r;
        if (flavor & FL_ADDCDIV_MUL_LAST) {
            float q = mv / tmp_d[i];
            r = (flavor & FL_FMA_ADDCDIV) ? fmaf(s->neg_step_size, q, pv) : pv + s->neg_step_size * q;
        } else {
            r = pv + (s->neg_step_size * mv) / tmp_d[i];
        }
        if (dt == DT_BF16 && sr) {
            uint32_t bits = f_bits(r) + sr_rand16(seed, offset, (uint64_t)(rng_base + i));
            ((uint16_t*)p)[i] = (uint16_t)((bits & 0xffff0000u) >> 16);
        } else {
            st(p, i, dt, r);
        }
    }
}

/* The three helpers of modules/util/bf16_stochastic_rounding.py as their callers see them (CAME.py:167-176,
 * adamw_extensions.py:139-144 use them; other optimizers of the reference import them the same way).  `bits16` = the 16-bit
 * draw of every element (torch.randint_like(.., 0, 1 << 16) in the reference), `target` = bf16 bits, updated in place.
 *   mode 0  copy_stochastic_(target, source):                 r = source                                   (:5-30)
 *   mode 1  add_stochastic_(input, other, alpha):             r = fp32(other); r.add_(input, alpha=alpha)  (:33-45)
 *           — alpha multiplies INPUT, not `other` (the docstring says otherwise; the code is what callers get)
 *   mode 2  addcdiv_stochastic_(input, tensor1, tensor2, v):  r = fp32(input); r.addcdiv_(t1, t2, value=v) (:48-61)
 * then bits = int32(r) + draw; bits &= 0xFFFF0000; target = high half. */
void otb_o_sr_helper(uint16_t* target, const void* a, const void* b, int64_t n, int mode, int src_dt, float alpha,
                     uint32_t flavor, const int32_t* bits16, float* r_out /* the fp32 value before rounding, or NULL */) {
    int64_t i;
#pragma omp parallel for
    for (i = 0; i < n; ++i) {
        float r;
        if (mode == 0) {
            r = ((const float*)a)[i];
        } else if (mode == 1) {
            const float o = ld(a, i, src_dt), in = bf16_to_f(target[i]);
            r = (flavor & FL_FMA_ADD_ALPHA) ? fmaf(alpha, in, o) : o + alpha * in;
        } else {
            const float in = bf16_to_f(target[i]), t1 = ld(a, i, src_dt), t2 = ld(b, i, src_dt);
            if (flavor & FL_ADDCDIV_MUL_LAST) {
                const float q = t1 / t2;
                r = (flavor & FL_FMA_ADDCDIV) ? fmaf(alpha, q, in) : in + alpha * q;
            } else {
                r = in + (alpha * t1) / t2;
            }
        }
        if (r_out) r_out[i] = r;
        const uint32_t u = f_bits(r) + (uint32_t)bits16[i];
        target[i] = (uint16_t)((u & 0xffff0000u) >> 16);
    }
}

/* ema.add_(one_minus_decay * (p - ema)): sub, mul, add — three roundings in the EMA dtype */
void otb_o_ema_step(void* ema, const void* p, int64_t n, int dt, float one_minus_decay) {
    int64_t i;
#pragma omp parallel for
    for (i = 0; i < n; ++i) {
        float e = ld(ema, i, dt);
        float t1 = rnd(ld(p, i, dt) - e, dt);
        float t2 = rnd(t1 * one_minus_decay, dt);
        st(ema, i, dt, e + t2);
    }
}

double otb_o_sqsum(const void* g, int64_t n, int dt) {
    double acc = 0.0;
    int64_t i;
#pragma omp parallel for reduction(+ : acc)
    for (i = 0; i < n; ++i) { double x = ld(g, i, dt); acc += x * x; }
    return acc;
}

/* ---------------- timesteps ---------------- */
enum { TS_UNIFORM = 0, TS_LOGIT_NORMAL = 1, TS_HEAVY_TAIL = 2, TS_DETERMINISTIC = 3 };

/* the continuous part of _get_timestep_discrete after the random draw, fp32 op by op;
 * exposed so the golden test can feed the reference's own draws */
int32_t otb_o_timestep_from_draw(int mode, float draw, int T, int min_t, int max_t, float noising_weight,
                                 float noising_bias, float shift_c1, float shift_c2) {
    float num_t = (float)(max_t - min_t), fmin = (float)min_t, ts;
    if (mode == TS_UNIFORM) {
        ts = fmin + num_t * draw;                        /* min + (max-min) * rand */
    } else if (mode == TS_LOGIT_NORMAL) {
        float sg = (float)(1.0 / (1.0 + exp(-(double)draw)));   /* draw = normal(bias, scale) */
        float t1 = sg * num_t;
        ts = t1 + fmin;
    } else {
        float ang = 1.57079632679489661923f * draw;      /* draw = rand */
        float c = (float)cos((double)ang);
        float c2 = c * c;
        float e = (c2 - 1.0f) + draw;
        float f = noising_weight * e;
        float h = (1.0f - draw) - f;
        float t1 = h * num_t;
        ts = t1 + fmin;
    }
    float num = shift_c1 * ts;
    float den = shift_c2 * ts;
    den = den + (float)T;
    ts = num / den;
    (void)noising_bias;
    return (int32_t)ts;
}

int32_t otb_o_timestep(int mode, uint64_t seed, uint64_t offset, int64_t global_sample, int T, int min_t, int max_t,
                       float noising_weight, float noising_bias, float shift_c1, float shift_c2) {
    if (mode == TS_DETERMINISTIC) return (int32_t)((double)T * 0.5) - 1;
    uint32_t r[4];
    otb_o_philox(seed, offset, 3u, (uint64_t)global_sample, r);
    float draw;
    if (mode == TS_LOGIT_NORMAL) {
        double u1 = ((double)(r[0] >> 8) + 0.5) * 5.9604644775390625e-08;
        double th = ((double)((int32_t)r[1] >> 8) + 0.5) * 1.1920928955078125e-07 * 3.14159265358979323846;
        float z = (float)(sqrt(-2.0 * log(u1)) * cos(th));
        float scale = noising_weight + 1.0f;
        float t1 = z * scale;
        draw = t1 + noising_bias;
    } else {
        draw = (float)(r[0] >> 8) * 5.9604644775390625e-08f;
    }
    return otb_o_timestep_from_draw(mode, draw, T, min_t, max_t, noising_weight, noising_bias, shift_c1, shift_c2);
}

/* COS_MAP / SIGMOID (ModelSetupNoiseMixin.py:122-153): the reference draws torch.multinomial(weights, B, replacement=True)
 * + min_timestep.  multinomial-with-replacement IS an i.i.d. draw from weights / sum(weights); restated as the inverse
 * CDF of one U[0,1) draw per GLOBAL sample from the timestep stream: first i with cdf[i] > u. */
int32_t otb_o_timestep_table(uint64_t seed, uint64_t offset, int64_t global_sample, const float* cdf, int len, int min_t) {
    uint32_t r[4];
    otb_o_philox(seed, offset, 3u, (uint64_t)global_sample, r);
    float u = (float)(r[0] >> 8) * 5.9604644775390625e-08f;
    int i = 0;
    while (i < len - 1 && !(cdf[i] > u)) ++i;         /* linear scan: the definition, not the kernel's binary search */
    return min_t + i;
}

/* ---------------- noise ---------------- */
/* N(0,1) draw for global element `elem` of `stream` (ideal evaluation of the kernel's definition) */
static inline float normal_at(uint64_t seed, uint64_t offset, uint32_t stream, uint64_t elem) {
    uint32_t r[4];
    otb_o_philox(seed, offset, stream, elem >> 2, r);
    int pair = (int)((elem & 3) >> 1);
    uint32_t a = r[2 * pair], b = r[2 * pair + 1];
    double u1 = ((double)(a >> 8) + 0.5) * 5.9604644775390625e-08;
    /* the kernel forms theta in fp32: ((float)s + 0.5f) * 2^-23 * pi_f */
    float thf = ((float)((int32_t)b >> 8) + 0.5f) * 1.1920928955078125e-07f * 3.14159265358979323846f;
    double rr = sqrt(-2.0 * log(u1));
    return (float)((elem & 1) ? rr * sin((double)thf) : rr * cos((double)thf));
}

/* _create_noise: noise (+ offset noise broadcast over H,W) (+ perturbation noise), fp32 latents */
void otb_o_create_noise(float* noise, int B, int C, int hw, float offset_w, float perturb_w, uint64_t seed,
                        uint64_t offset, int64_t sample_base) {
    int64_t chw = (int64_t)C * hw;
    int64_t total = (int64_t)B * chw, idx;
#pragma omp parallel for
    for (idx = 0; idx < total; ++idx) {
        int b = (int)(idx / chw);
        int64_t i = idx % chw;
        int64_t gs = sample_base + b;
        uint64_t ge = (uint64_t)(gs * chw + i);
        float n = normal_at(seed, offset, 0u, ge);
        if (offset_w > 0.0f) {
            int c = (int)(i / hw);
            float on = normal_at(seed, offset, 1u, (uint64_t)(gs * C + c));
            float t = offset_w * on;
            n = n + t;
        }
        if (perturb_w > 0.0f) {
            float pn = normal_at(seed, offset, 2u, ge);
            float t = perturb_w * pn;
            n = n + t;
        }
        noise[idx] = n;
    }
}

/* x0 = (latent - shift) * scale ; noisy = x0*cx[t] + noise*cn[t] ; target per prediction type */
void otb_o_add_noise(const float* latent, const float* noise, const int32_t* timestep, const float* coef_x0,
                     const float* coef_noise, int B, int64_t chw, float latent_scale, float latent_shift, int target,
                     float* x0_out, float* noisy_out, float* target_out) {
    int64_t total = (int64_t)B * chw, idx;
#pragma omp parallel for
    for (idx = 0; idx < total; ++idx) {
        int b = (int)(idx / chw);
        int t = timestep[b];
        float cx = coef_x0[t], cn = coef_noise[t];
        float x = latent[idx] - latent_shift;
        x = x * latent_scale;
        float n = noise[idx];
        if (x0_out) x0_out[idx] = x;
        if (noisy_out) { float a = x * cx, c = n * cn; noisy_out[idx] = a + c; }
        if (target_out) {
            if (target == 1) target_out[idx] = n;
            else if (target == 2) { float a = cx * n, c = cn * x; target_out[idx] = a - c; }
            else target_out[idx] = n - x;
        }
    }
}

/* ---------------- losses ---------------- */
enum { LW_CONSTANT = 0, LW_MIN_SNR = 1, LW_DEBIASED = 2, LW_P2 = 3, LW_SIGMA = 4 };

float otb_o_timestep_weight(int fn, int t, const float* sqrt_ac, const float* sqrt_1mac, const float* sigmas,
                            float gamma, int v_pred) {
    if (fn == LW_CONSTANT) return 1.0f;
    if (fn == LW_SIGMA) return sigmas[t];
    float q = sqrt_ac[t] / sqrt_1mac[t];
    float snr = q * q;
    if (fn == LW_MIN_SNR) {
        float mn = snr < gamma ? snr : gamma;
        if (v_pred) snr = snr + 1.0f;
        return mn / snr;
    }
    if (fn == LW_DEBIASED) {
        float w = snr < 1.0e3f ? snr : 1.0e3f;
        if (v_pred) w = w + 1.0f;
        return 1.0f / sqrtf(w);
    }
    if (v_pred) snr = snr + 1.0f;
    return powf(1.0f + snr, -gamma);
}

/* per-sample losses [B] and, if dpred != NULL, d(sum_b upstream*losses[b])/d pred (fp32).
 * pred/target/prior fp32 (callers upcast, as the reference does with .to(float32)). */
void otb_o_losses(const float* pred, const float* target, const float* prior, const float* mask, const float* loss_weight,
                  const int32_t* timestep, const float* sqrt_ac, const float* sqrt_1mac, const float* sigmas, int B, int C,
                  int hw, int masked, int normalize, int lw_fn, int v_pred, float mse_s, float mae_s, float unmasked_weight,
                  float prior_w, float loss_scale, float gamma, float upstream, float* losses, float* dpred) {
    int64_t chw = (int64_t)C * hw;
    for (int b = 0; b < B; ++b) {
        double mm = 1.0, mp = 1.0;
        if (mask && normalize) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = 0; i < hw; ++i) {
                float m = mask[(int64_t)b * hw + i];
                m = m < unmasked_weight ? unmasked_weight : (m > 1.0f ? 1.0f : m);
                s0 += m; s1 += 1.0f - m;
            }
            mm = s0 / hw; mp = s1 / hw;
        }
        int use_prior = masked && prior != NULL && prior_w != 0.0f;
        double acc = 0.0;
        for (int64_t i = 0; i < chw; ++i) {
            float p = pred[b * chw + i], d = p - target[b * chw + i];
            double m = 1.0;
            if (masked) {
                float mv = mask[(int64_t)b * hw + (i % hw)];
                mv = mv < unmasked_weight ? unmasked_weight : (mv > 1.0f ? 1.0f : mv);
                m = mv;
            }
            double km = (masked || (mask && normalize)) ? m / mm : m;
            acc += mse_s * (double)d * d * km + mae_s * fabs((double)d) * km;
            if (use_prior) {
                float dp = p - prior[b * chw + i];
                double kp = (1.0 - m) * prior_w / (normalize ? mp : 1.0);
                acc += mse_s * (double)dp * dp * kp + mae_s * fabs((double)dp) * kp;
            }
        }
        float wt = timestep ? otb_o_timestep_weight(lw_fn, timestep[b], sqrt_ac, sqrt_1mac, sigmas, gamma, v_pred) : 1.0f;
        double post = (double)loss_scale * (loss_weight ? loss_weight[b] : 1.0f) * wt;
        losses[b] = (float)(acc / (double)chw * post);
        if (dpred) {
            double gf = post / (double)chw * upstream;
            for (int64_t i = 0; i < chw; ++i) {
                float p = pred[b * chw + i], d = p - target[b * chw + i];
                double m = 1.0;
                if (masked) {
                    float mv = mask[(int64_t)b * hw + (i % hw)];
                    mv = mv < unmasked_weight ? unmasked_weight : (mv > 1.0f ? 1.0f : mv);
                    m = mv;
                }
                double km = (masked || (mask && normalize)) ? m / mm : m;
                double g = mse_s * 2.0 * d * km + mae_s * ((d > 0) - (d < 0)) * km;
                if (use_prior) {
                    float dp = p - prior[b * chw + i];
                    double kp = (1.0 - m) * prior_w / (normalize ? mp : 1.0);
                    g += mse_s * 2.0 * dp * kp + mae_s * ((dp > 0) - (dp < 0)) * kp;
                }
                dpred[b * chw + i] = (float)(g * gf);
            }
        }
    }
}

/* ---------------- LoRA linear (fp32 reference of the fused contract) ---------------- */
/* y = x w^T + bias + bf16(s * x down^T) up^T, fp32 accumulate in double; T optionally rounded to bf16 */
void otb_o_lora_linear(const float* x, const float* w, const float* down, const float* up, const float* bias, int M, int N,
                       int K, int R, float scale, int round_t_bf16, float* y, float* t_out) {
    int m;
#pragma omp parallel for
    for (m = 0; m < M; ++m) {
        float t[64];
        for (int r = 0; r < R; ++r) {
            double a = 0.0;
            for (int k = 0; k < K; ++k) a += (double)x[(int64_t)m * K + k] * down[(int64_t)r * K + k];
            float tv = (float)a * scale;
            t[r] = round_t_bf16 ? bf16_to_f(f_to_bf16(tv)) : tv;
            if (t_out) t_out[(int64_t)m * R + r] = t[r];
        }
        for (int n = 0; n < N; ++n) {
            double a = bias ? bias[n] : 0.0;
            for (int k = 0; k < K; ++k) a += (double)x[(int64_t)m * K + k] * w[(int64_t)n * K + k];
            for (int r = 0; r < R; ++r) a += (double)t[r] * up[(int64_t)n * R + r];
            y[(int64_t)m * N + n] = (float)a;
        }
    }
}
