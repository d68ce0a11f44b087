// optim.cu — K-c: multi-tensor AdamW/Adam + bf16 stochastic rounding + EMA (+ clip scaling,
// + bf16 shadow emission), the global grad-norm reduction, and the LoRA bf16 shadow refresh.
//
// Reference behaviour restated (rounding point by rounding point):
//   step_adamw_parameter  /root/reference/modules/util/optimizer/adamw_extensions.py:79-146
//   step_adam_parameter   /root/reference/modules/util/optimizer/adam_extensions.py:71-90
//   addcdiv_stochastic_ / copy_stochastic_  modules/util/bf16_stochastic_rounding.py:5-30,48-61
//   EMAModuleWrapper.step modules/module/EMAModule.py:41-47
//   clip_grad_norm_       modules/trainer/GenericTrainer.py:712-713
//
// HBM-bound: 14 B/elem (bf16 params) / 28 B/elem (fp32 params), +4/+8 on EMA steps.
// Every thread owns 8 consecutive elements per iteration (one 128-bit access per bf16
// stream, two per fp32 stream); the 8 x 16 random bits of the SR step are one Philox call.
#include "common.cuh"

#define OPT_THREADS 256
#define OPT_FLAG_SR        1
#define OPT_FLAG_MAXIMIZE  2
#define OPT_FLAG_ALIGNED16 4   /* set by the host when every pointer of the tensor is 16 B aligned */

namespace {

struct OptScalars {
    float decay_mul, l2, lerp_w, beta2, omb2, bc2, inv_bc2, eps, nss, clip, ema_omd;
    uint32_t flavor;
    bool sr, maximize, bf16, do_ema, has_clip;
};

// in-place read (not .nc: the same addresses are written below)
__device__ __forceinline__ uint4 ld_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <bool BF16>
__device__ __forceinline__ float rnd(float x) {
    if (BF16) return round_bf16(x);
    return x;
}
// round two values at once: one F2FP.BF16.PACK_AB instead of two scalar conversions
template <bool BF16>
__device__ __forceinline__ void rnd2(float& a, float& b) {
    if (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        uint32_t u = *reinterpret_cast<uint32_t*>(&h);
        a = __uint_as_float(u << 16);
        b = __uint_as_float(u & 0xFFFF0000u);
    }
}

// Two elements of the reference update.  All arithmetic through explicit _rn intrinsics so
// nvcc cannot contract; FMA is used only where the flavour says torch's kernel does.
// FL is the compile-time flavour when >= 0 (the hot instantiation is OTB_FLAVOR_TORCH_CUDA),
// or -1 to read the run-time bits.
template <bool BF16, int FL>
__device__ __forceinline__ void adam_pair(const OptScalars& s, float (&p)[2], float (&g)[2], float (&m)[2],
                                          float (&v)[2], float (&e)[2], const uint32_t (&rnd16)[2], bool has_ema) {
    const uint32_t fl = FL >= 0 ? (uint32_t)FL : s.flavor;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (s.maximize) g[k] = -g[k];
        if (s.has_clip) g[k] = __fmul_rn(g[k], s.clip);                 // grads.mul_(clip_coef)
    }
    if (s.has_clip) rnd2<BF16>(g[0], g[1]);
    if (s.l2 != 0.0f) {                                                 // Adam: grad.add(p, alpha=wd)
        const float l2 = (fl & OTB_FLAVOR_ALPHA_IN_DTYPE) ? rnd<BF16>(s.l2) : s.l2;
#pragma unroll
        for (int k = 0; k < 2; ++k)
            g[k] = (fl & OTB_FLAVOR_FMA_ADD_ALPHA) ? __fmaf_rn(l2, p[k], g[k]) : __fadd_rn(g[k], __fmul_rn(l2, p[k]));
        rnd2<BF16>(g[0], g[1]);
    } else {                                                            // AdamW: p.mul_(1 - lr*wd)
#pragma unroll
        for (int k = 0; k < 2; ++k) p[k] = __fmul_rn(p[k], s.decay_mul);
        rnd2<BF16>(p[0], p[1]);
    }
    // exp_avg.lerp_(grad, 1-beta1)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float d = __fsub_rn(g[k], m[k]);
        if (s.lerp_w < 0.5f) {
            m[k] = (fl & OTB_FLAVOR_FMA_LERP) ? __fmaf_rn(s.lerp_w, d, m[k]) : __fadd_rn(m[k], __fmul_rn(s.lerp_w, d));
        } else {
            float w1 = __fsub_rn(1.0f, s.lerp_w);
            m[k] = (fl & OTB_FLAVOR_FMA_LERP) ? __fmaf_rn(-w1, d, g[k]) : __fsub_rn(g[k], __fmul_rn(d, w1));
        }
    }
    rnd2<BF16>(m[0], m[1]);
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
#pragma unroll
    for (int k = 0; k < 2; ++k) v[k] = __fmul_rn(v[k], s.beta2);
    rnd2<BF16>(v[0], v[1]);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float x, y;
        if (fl & OTB_FLAVOR_ADDCMUL_SQUARE_FIRST) { x = s.omb2; y = __fmul_rn(g[k], g[k]); }
        else { x = __fmul_rn(s.omb2, g[k]); y = g[k]; }
        v[k] = (fl & OTB_FLAVOR_FMA_ADDCMUL) ? __fmaf_rn(x, y, v[k]) : __fadd_rn(v[k], __fmul_rn(x, y));
    }
    rnd2<BF16>(v[0], v[1]);
    // denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    float d[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) d[k] = __fsqrt_rn(v[k]);
    rnd2<BF16>(d[0], d[1]);
#pragma unroll
    for (int k = 0; k < 2; ++k) d[k] = (fl & OTB_FLAVOR_DIV_RECIPROCAL) ? __fmul_rn(d[k], s.inv_bc2) : __fdiv_rn(d[k], s.bc2);
    rnd2<BF16>(d[0], d[1]);
#pragma unroll
    for (int k = 0; k < 2; ++k) d[k] = __fadd_rn(d[k], s.eps);
    rnd2<BF16>(d[0], d[1]);
    // p.addcdiv_(exp_avg, denom, value=-step_size)   [fp32 math; SR or RN write-back]
    float r[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (fl & OTB_FLAVOR_ADDCDIV_MUL_LAST) {
            float q = __fdiv_rn(m[k], d[k]);
            r[k] = (fl & OTB_FLAVOR_FMA_ADDCDIV) ? __fmaf_rn(s.nss, q, p[k]) : __fadd_rn(p[k], __fmul_rn(s.nss, q));
        } else {
            r[k] = __fadd_rn(p[k], __fdiv_rn(__fmul_rn(s.nss, m[k]), d[k]));
        }
    }
    if (BF16 && s.sr) {
#pragma unroll
        for (int k = 0; k < 2; ++k)                                      // copy_stochastic_
            p[k] = __uint_as_float((__float_as_uint(r[k]) + rnd16[k]) & 0xFFFF0000u);
    } else {
        p[0] = r[0]; p[1] = r[1];
        rnd2<BF16>(p[0], p[1]);
    }
    if (has_ema) {                                                      // ema.add_(omd * (p - ema))
        float t[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) t[k] = __fsub_rn(p[k], e[k]);
        rnd2<BF16>(t[0], t[1]);
#pragma unroll
        for (int k = 0; k < 2; ++k) t[k] = __fmul_rn(t[k], s.ema_omd);
        rnd2<BF16>(t[0], t[1]);
#pragma unroll
        for (int k = 0; k < 2; ++k) e[k] = __fadd_rn(e[k], t[k]);
        rnd2<BF16>(e[0], e[1]);
    }
}

template <bool BF16> struct Pack8;
template <> struct Pack8<true> {
    __device__ static __forceinline__ void load(const void* base, int64_t i, float (&f)[8], bool nc) {
        const char* p = (const char*)base + i * 2;
        uint4 r = nc ? ld_stream_u4(p) : ld_u4(p);
        uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[2 * k] = __uint_as_float(w[k] << 16);
            f[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
        }
    }
    __device__ static __forceinline__ void store(void* base, int64_t i, const float (&f)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = (__float_as_uint(f[2 * k]) >> 16) | (__float_as_uint(f[2 * k + 1]) & 0xFFFF0000u);
        st_stream_u4((char*)base + i * 2, make_uint4(w[0], w[1], w[2], w[3]));
    }
};
template <> struct Pack8<false> {
    __device__ static __forceinline__ void load(const void* base, int64_t i, float (&f)[8], bool nc) {
        const char* p = (const char*)base + i * 4;
        uint4 a = nc ? ld_stream_u4(p) : ld_u4(p);
        uint4 b = nc ? ld_stream_u4(p + 16) : ld_u4(p + 16);
        f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
        f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
    }
    __device__ static __forceinline__ void store(void* base, int64_t i, const float (&f)[8]) {
        char* p = (char*)base + i * 4;
        st_stream_u4(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
        st_stream_u4(p + 16, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
    }
};

__device__ __forceinline__ uint32_t sr_bits(const Philox4& r, int lane) {
    uint32_t w = (lane >> 1) == 0 ? r.x : (lane >> 1) == 1 ? r.y : (lane >> 1) == 2 ? r.z : r.w;
    return (lane & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// ---- K-z: the updated parameters leave for the other replicas from the optimizer kernel itself (otb_peer_store): the
// all-gather half of the sharded exchange rides on K-c's stores instead of being a second pass over the owner's range
__device__ __forceinline__ void st_sys_u4(void* p, uint4 v) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_u4(void* mc, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}
// 16 bytes at byte offset `off` of my replica -> every other replica
__device__ __forceinline__ void peer_store16(const otb_peer_store& ps, void* mine, int64_t off, uint4 v) {
    if (ps.mode == OTB_EXCHANGE_NVLS) {
        multimem_st_u4((char*)mine + off + ps.mc_delta, v);
    } else {
        for (int r = 0; r < ps.world; ++r)
            if (r != ps.rank) st_sys_u4((char*)mine + off + ps.delta[r], v);
    }
}
template <bool BF16>
__device__ __forceinline__ void peer_store8(const otb_peer_store& ps, void* base, int64_t i, const float (&f)[8]) {
    if (BF16) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = (__float_as_uint(f[2 * k]) >> 16) | (__float_as_uint(f[2 * k + 1]) & 0xFFFF0000u);
        peer_store16(ps, base, i * 2, make_uint4(w[0], w[1], w[2], w[3]));
    } else {
        peer_store16(ps, base, i * 4, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
        peer_store16(ps, base, i * 4 + 16, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
    }
}
// scalar tail: plain peer stores in both modes (a multicast store needs 4-byte granularity at least)
template <bool BF16>
__device__ __forceinline__ void peer_store1(const otb_peer_store& ps, void* base, int64_t i, float v) {
    for (int r = 0; r < ps.world; ++r) {
        if (r == ps.rank) continue;
        char* dst = (char*)base + ps.delta[r];
        if (BF16) {
            const uint16_t b = float_to_bf16_bits_rn(v);
            asm volatile("st.relaxed.sys.global.u16 [%0], %1;" :: "l"((uint16_t*)dst + i), "h"(b) : "memory");
        } else {
            asm volatile("st.relaxed.sys.global.f32 [%0], %1;" :: "l"((float*)dst + i), "f"(v) : "memory");
        }
    }
}

template <bool BF16, int FL, bool PEERS = false>
__device__ __forceinline__ void adam_chunk(const otb_optim_tensor& t, const OptScalars& s, int64_t start, int64_t end,
                                           uint64_t seed, uint64_t offset, const otb_peer_store* ps = nullptr) {
    const bool has_ema = s.do_ema && t.ema != nullptr;
    const bool aligned = (t.flags & OPT_FLAG_ALIGNED16) != 0;
    int64_t vec_end = aligned ? start + ((end - start) & ~int64_t(7)) : start;
    // vector body
    for (int64_t i = start + (int64_t)threadIdx.x * 8; i < vec_end; i += (int64_t)OPT_THREADS * 8) {
        float p[8], g[8], m[8], v[8], e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        Pack8<BF16>::load(t.p, i, p, false);
        Pack8<BF16>::load(t.g, i, g, true);
        Pack8<BF16>::load(t.exp_avg, i, m, false);
        Pack8<BF16>::load(t.exp_avg_sq, i, v, false);
        if (has_ema) Pack8<BF16>::load(t.ema, i, e, false);
        Philox4 r{0, 0, 0, 0};
        if (BF16 && s.sr) r = otb_rng(seed, offset, OTB_STREAM_SR, (uint64_t)(t.rng_base + i) >> 3);
        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float pp[2] = {p[2 * k], p[2 * k + 1]}, gg[2] = {g[2 * k], g[2 * k + 1]}, mm[2] = {m[2 * k], m[2 * k + 1]};
            float vv[2] = {v[2 * k], v[2 * k + 1]}, ee[2] = {e[2 * k], e[2 * k + 1]};
            const uint32_t r16[2] = {rw[k] & 0xFFFFu, rw[k] >> 16};
            adam_pair<BF16, FL>(s, pp, gg, mm, vv, ee, r16, has_ema);
            p[2 * k] = pp[0]; p[2 * k + 1] = pp[1]; m[2 * k] = mm[0]; m[2 * k + 1] = mm[1];
            v[2 * k] = vv[0]; v[2 * k + 1] = vv[1]; e[2 * k] = ee[0]; e[2 * k + 1] = ee[1];
        }
        Pack8<BF16>::store(t.p, i, p);
        if (PEERS) peer_store8<BF16>(*ps, t.p, i, p);
        Pack8<BF16>::store(t.exp_avg, i, m);
        Pack8<BF16>::store(t.exp_avg_sq, i, v);
        if (has_ema) Pack8<BF16>::store(t.ema, i, e);
        if (t.shadow_bf16) {
            float q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = round_bf16(p[k]);     // RN for fp32 masters, identity for bf16
            Pack8<true>::store(t.shadow_bf16, i, q);
        }
    }
    // scalar tail (and the whole chunk for unaligned tensors)
    const int dt = BF16 ? OTB_BF16 : OTB_F32;
    for (int64_t i = vec_end + threadIdx.x; i < end; i += OPT_THREADS) {
        float p[2] = {load_as_float(t.p, i, dt), 0.0f}, g[2] = {load_as_float(t.g, i, dt), 0.0f};
        float m[2] = {load_as_float(t.exp_avg, i, dt), 0.0f}, v[2] = {load_as_float(t.exp_avg_sq, i, dt), 1.0f};
        float e[2] = {has_ema ? load_as_float(t.ema, i, dt) : 0.0f, 0.0f};
        uint32_t r16[2] = {0, 0};
        if (BF16 && s.sr) {
            uint64_t ge = (uint64_t)(t.rng_base + i);
            r16[0] = sr_bits(otb_rng(seed, offset, OTB_STREAM_SR, ge >> 3), (int)(ge & 7));
        }
        adam_pair<BF16, FL>(s, p, g, m, v, e, r16, has_ema);
        store_from_float(t.p, i, dt, p[0]);
        if (PEERS) peer_store1<BF16>(*ps, t.p, i, p[0]);
        store_from_float(t.exp_avg, i, dt, m[0]);
        store_from_float(t.exp_avg_sq, i, dt, v[0]);
        if (has_ema) store_from_float(t.ema, i, dt, e[0]);
        if (t.shadow_bf16) ((uint16_t*)t.shadow_bf16)[i] = float_to_bf16_bits_rn(p[0]);
    }
}

__global__ void __launch_bounds__(OPT_THREADS, 4) adamw_sr_ema_kernel(otb_optim_launch a) {
    if (a.dyn != nullptr) {
        a.seed = a.dyn->seed; a.offset = a.dyn->offset;
        a.ema_one_minus_decay = a.dyn->ema_one_minus_decay; a.do_ema = a.dyn->do_ema;
    }
    const int chunk = blockIdx.x;
    const otb_optim_tensor t = a.tensors[a.chunk_tensor[chunk]];
    const int64_t start = (int64_t)a.chunk_index[chunk] * a.chunk_elems;
    int64_t end = start + a.chunk_elems;
    if (end > t.numel) end = t.numel;
    OptScalars s;
    s.decay_mul = t.decay_mul; s.l2 = t.l2; s.lerp_w = t.lerp_w; s.beta2 = t.beta2; s.omb2 = t.one_minus_beta2;
    s.bc2 = t.bc2_sqrt; s.inv_bc2 = t.inv_bc2_sqrt; s.eps = t.eps; s.nss = t.neg_step_size;
    s.has_clip = a.clip_coef != nullptr;
    s.clip = s.has_clip ? *a.clip_coef : 1.0f;
    s.ema_omd = a.ema_one_minus_decay;
    s.flavor = a.flavor;
    s.sr = (t.flags & OPT_FLAG_SR) != 0;
    s.maximize = (t.flags & OPT_FLAG_MAXIMIZE) != 0;
    s.do_ema = a.do_ema != 0;
    const bool hot = a.flavor == OTB_FLAVOR_TORCH_CUDA;      // block-uniform
    if (t.p_dtype == OTB_BF16) {
        if (hot) adam_chunk<true, (int)OTB_FLAVOR_TORCH_CUDA>(t, s, start, end, a.seed, a.offset);
        else adam_chunk<true, -1>(t, s, start, end, a.seed, a.offset);
    } else {
        if (hot) adam_chunk<false, (int)OTB_FLAVOR_TORCH_CUDA>(t, s, start, end, a.seed, a.offset);
        else adam_chunk<false, -1>(t, s, start, end, a.seed, a.offset);
    }
}

// the same step with the parameter stores replicated into the peers' arenas (hot flavour only)
__global__ void __launch_bounds__(OPT_THREADS, 4) adamw_sr_ema_peers_kernel(otb_optim_launch a, const otb_peer_store ps) {
    if (a.dyn != nullptr) {
        a.seed = a.dyn->seed; a.offset = a.dyn->offset;
        a.ema_one_minus_decay = a.dyn->ema_one_minus_decay; a.do_ema = a.dyn->do_ema;
    }
    const int chunk = blockIdx.x;
    const otb_optim_tensor t = a.tensors[a.chunk_tensor[chunk]];
    const int64_t start = (int64_t)a.chunk_index[chunk] * a.chunk_elems;
    int64_t end = start + a.chunk_elems;
    if (end > t.numel) end = t.numel;
    OptScalars s;
    s.decay_mul = t.decay_mul; s.l2 = t.l2; s.lerp_w = t.lerp_w; s.beta2 = t.beta2; s.omb2 = t.one_minus_beta2;
    s.bc2 = t.bc2_sqrt; s.inv_bc2 = t.inv_bc2_sqrt; s.eps = t.eps; s.nss = t.neg_step_size;
    s.has_clip = a.clip_coef != nullptr;
    s.clip = s.has_clip ? *a.clip_coef : 1.0f;
    s.ema_omd = a.ema_one_minus_decay;
    s.flavor = a.flavor;
    s.sr = (t.flags & OPT_FLAG_SR) != 0;
    s.maximize = (t.flags & OPT_FLAG_MAXIMIZE) != 0;
    s.do_ema = a.do_ema != 0;
    if (t.p_dtype == OTB_BF16) adam_chunk<true, (int)OTB_FLAVOR_TORCH_CUDA, true>(t, s, start, end, a.seed, a.offset, &ps);
    else adam_chunk<false, (int)OTB_FLAVOR_TORCH_CUDA, true>(t, s, start, end, a.seed, a.offset, &ps);
    __threadfence_system();                       // the peer stores are performed before the barrier that follows on the stream
}

// ------------------------------- EMA only -------------------------------------------
template <bool BF16>
__device__ void ema_chunk(const otb_optim_tensor& t, float omd, int64_t start, int64_t end) {
    const bool aligned = (t.flags & OPT_FLAG_ALIGNED16) != 0;
    int64_t vec_end = aligned ? start + ((end - start) & ~int64_t(7)) : start;
    for (int64_t i = start + (int64_t)threadIdx.x * 8; i < vec_end; i += (int64_t)OPT_THREADS * 8) {
        float p[8], e[8];
        Pack8<BF16>::load(t.p, i, p, true);
        Pack8<BF16>::load(t.ema, i, e, false);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float t1 = rnd<BF16>(__fsub_rn(p[k], e[k]));
            float t2 = rnd<BF16>(__fmul_rn(t1, omd));
            e[k] = rnd<BF16>(__fadd_rn(e[k], t2));
        }
        Pack8<BF16>::store(t.ema, i, e);
    }
    const int dt = BF16 ? OTB_BF16 : OTB_F32;
    for (int64_t i = vec_end + threadIdx.x; i < end; i += OPT_THREADS) {
        float p = load_as_float(t.p, i, dt), e = load_as_float(t.ema, i, dt);
        float t1 = rnd<BF16>(__fsub_rn(p, e));
        float t2 = rnd<BF16>(__fmul_rn(t1, omd));
        store_from_float(t.ema, i, dt, rnd<BF16>(__fadd_rn(e, t2)));
    }
}

__global__ void __launch_bounds__(OPT_THREADS) ema_kernel(otb_optim_launch a) {
    const otb_optim_tensor t = a.tensors[a.chunk_tensor[blockIdx.x]];
    if (t.ema == nullptr) return;
    const int64_t start = (int64_t)a.chunk_index[blockIdx.x] * a.chunk_elems;
    int64_t end = start + a.chunk_elems;
    if (end > t.numel) end = t.numel;
    if (t.p_dtype == OTB_BF16) ema_chunk<true>(t, a.ema_one_minus_decay, start, end);
    else ema_chunk<false>(t, a.ema_one_minus_decay, start, end);
}

// ------------------------------- grad norm ------------------------------------------
template <bool BF16>
__device__ float sqsum_chunk(const otb_optim_tensor& t, float prescale, int64_t start, int64_t end) {
    const bool aligned = (t.flags & OPT_FLAG_ALIGNED16) != 0;
    int64_t vec_end = aligned ? start + ((end - start) & ~int64_t(7)) : start;
    float acc = 0.0f;
    for (int64_t i = start + (int64_t)threadIdx.x * 8; i < vec_end; i += (int64_t)OPT_THREADS * 8) {
        float g[8];
        Pack8<BF16>::load(t.g, i, g, true);
#pragma unroll
        for (int k = 0; k < 8; ++k) { float x = g[k] * prescale; acc = fmaf(x, x, acc); }
    }
    const int dt = BF16 ? OTB_BF16 : OTB_F32;
    for (int64_t i = vec_end + threadIdx.x; i < end; i += OPT_THREADS) {
        float x = load_as_float(t.g, i, dt) * prescale;
        acc = fmaf(x, x, acc);
    }
    return acc;
}

__device__ __forceinline__ float block_sum_256(float v) {
    __shared__ float sm[OPT_THREADS / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.0f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < OPT_THREADS / 32 ? sm[threadIdx.x] : 0.0f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;   // valid on warp 0
}

__global__ void __launch_bounds__(OPT_THREADS) grad_sqnorm_kernel(otb_optim_launch a, float prescale, float* partials) {
    const otb_optim_tensor t = a.tensors[a.chunk_tensor[blockIdx.x]];
    const int64_t start = (int64_t)a.chunk_index[blockIdx.x] * a.chunk_elems;
    int64_t end = start + a.chunk_elems;
    if (end > t.numel) end = t.numel;
    float acc = (t.g == nullptr) ? 0.0f
              : (t.p_dtype == OTB_BF16) ? sqsum_chunk<true>(t, prescale, start, end)
                                        : sqsum_chunk<false>(t, prescale, start, end);
    acc = block_sum_256(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// fixed-order final reduction: thread i sums partials[i], [i+1024], ... then a tree
__global__ void __launch_bounds__(1024) grad_norm_finish_kernel(const float* partials, int n, float max_norm, float prescale, float* out) {
    __shared__ double sm[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += (double)partials[i];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float total = sqrtf((float)sm[0]);
        float coef = max_norm / (total + 1e-6f);          // torch: clip_coef = max_norm / (total_norm + 1e-6)
        out[0] = total;
        out[1] = (coef < 1.0f ? coef : 1.0f) * prescale;  // clamp(max=1.0); the 1/world_size average rides along
    }
}

template <bool BF16>
__device__ void scale_chunk(const otb_optim_tensor& t, float c, int64_t start, int64_t end) {
    const int dt = BF16 ? OTB_BF16 : OTB_F32;
    const bool aligned = (t.flags & OPT_FLAG_ALIGNED16) != 0;
    int64_t vec_end = aligned ? start + ((end - start) & ~int64_t(7)) : start;
    for (int64_t i = start + (int64_t)threadIdx.x * 8; i < vec_end; i += (int64_t)OPT_THREADS * 8) {
        float g[8];
        Pack8<BF16>::load(t.g, i, g, false);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = rnd<BF16>(__fmul_rn(g[k], c));
        Pack8<BF16>::store((void*)t.g, i, g);
    }
    for (int64_t i = vec_end + threadIdx.x; i < end; i += OPT_THREADS)
        store_from_float((void*)t.g, i, dt, rnd<BF16>(__fmul_rn(load_as_float(t.g, i, dt), c)));
}

__global__ void __launch_bounds__(OPT_THREADS) grad_scale_kernel(otb_optim_launch a, const float* coef) {
    const otb_optim_tensor t = a.tensors[a.chunk_tensor[blockIdx.x]];
    if (t.g == nullptr) return;
    const int64_t start = (int64_t)a.chunk_index[blockIdx.x] * a.chunk_elems;
    int64_t end = start + a.chunk_elems;
    if (end > t.numel) end = t.numel;
    const float c = *coef;
    if (t.p_dtype == OTB_BF16) scale_chunk<true>(t, c, start, end);
    else scale_chunk<false>(t, c, start, end);
}

// ------------------------------- bf16 shadow refresh --------------------------------
// one 32x32 tile per block (256 threads = 32 x 8), padded smem transpose
__global__ void __launch_bounds__(256) shadow_refresh_kernel(const otb_shadow_tensor* tensors, int n_tensors) {
    __shared__ float tile[32][33];
    // binary search the tensor owning this tile
    int lo = 0, hi = n_tensors - 1;
    const int tile_id = blockIdx.x;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (tensors[mid].tile_base <= tile_id) lo = mid; else hi = mid - 1;
    }
    const otb_shadow_tensor t = tensors[lo];
    const int local = tile_id - t.tile_base;
    const int tiles_c = (t.cols + 31) >> 5;
    const int tr = local / tiles_c, tc = local % tiles_c;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int r = tr * 32 + ty + j * 8, c = tc * 32 + tx;
        float v = 0.0f;
        if (r < t.rows && c < t.cols) {
            v = load_as_float(t.src, (int64_t)r * t.cols + c, t.src_dtype);
            if (t.dst) ((uint16_t*)t.dst)[(int64_t)r * t.cols + c] = float_to_bf16_bits_rn(v);
        }
        tile[ty + j * 8][tx] = v;
    }
    __syncthreads();
    if (t.dst_t) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = tc * 32 + ty + j * 8, r = tr * 32 + tx;   // output row = source col
            if (r < t.rows && c < t.cols)
                ((uint16_t*)t.dst_t)[(int64_t)c * t.rows + r] = float_to_bf16_bits_rn(tile[tx][ty + j * 8]);
        }
    }
}

int check_launch(const otb_optim_launch* a) {
    OTB_CHECK_ARG(a != nullptr, "null launch args");
    OTB_CHECK_ARG(a->n_chunks >= 0 && a->n_tensors >= 0, "negative sizes");
    if (a->n_chunks == 0) return 1;
    OTB_CHECK_ARG(a->tensors && a->chunk_tensor && a->chunk_index, "null table pointer");
    OTB_CHECK_ARG(a->chunk_elems > 0 && a->chunk_elems % 2048 == 0, "chunk_elems must be a positive multiple of 2048");
    OTB_BIND(a->tensors);
    return 0;
}

}  // namespace

extern "C" int otb_adamw_sr_ema_multi(const otb_optim_launch* a, void* stream) {
    int c = check_launch(a);
    if (c) return c < 0 ? c : OTB_OK;
    adamw_sr_ema_kernel<<<a->n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(*a);
    OTB_LAUNCH_CHECK("adamw_sr_ema_kernel");
    return OTB_OK;
}

extern "C" int otb_adamw_sr_ema_multi_peers(const otb_optim_launch* a, const otb_peer_store* ps, void* stream) {
    OTB_CHECK_ARG(ps != nullptr, "null peer-store description");
    OTB_CHECK_ARG(ps->world >= 2 && ps->world <= OTB_MAX_RANKS && ps->rank >= 0 && ps->rank < ps->world, "bad rank %d of %d", ps->rank, ps->world);
    OTB_CHECK_ARG(ps->mode == OTB_EXCHANGE_P2P || ps->mode == OTB_EXCHANGE_NVLS, "unknown peer-store mode %d", ps->mode);
    OTB_CHECK_ARG(ps->delta[ps->rank] == 0, "delta of the own rank must be 0");
    for (int r = 0; r < ps->world; ++r) OTB_CHECK_ARG((ps->delta[r] & 15) == 0, "arena of rank %d is not 16-byte congruent with mine", r);
    OTB_CHECK_ARG(ps->mode != OTB_EXCHANGE_NVLS || (ps->mc_delta != 0 && (ps->mc_delta & 15) == 0), "NVLS mode needs the multicast offset");
    int c = check_launch(a);
    if (c) return c < 0 ? c : OTB_OK;
    OTB_UNSUPPORTED(a->flavor != OTB_FLAVOR_TORCH_CUDA, "the peer-store form of the optimizer step exists for OTB_FLAVOR_TORCH_CUDA only");
    adamw_sr_ema_peers_kernel<<<a->n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(*a, *ps);
    OTB_LAUNCH_CHECK("adamw_sr_ema_peers_kernel");
    return OTB_OK;
}

extern "C" int otb_ema_multi(const otb_optim_launch* a, void* stream) {
    int c = check_launch(a);
    if (c) return c < 0 ? c : OTB_OK;
    ema_kernel<<<a->n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(*a);
    OTB_LAUNCH_CHECK("ema_kernel");
    return OTB_OK;
}

extern "C" int otb_grad_norm_multi(const otb_optim_launch* a, float max_norm, float grad_prescale,
                                   float* partials, float* out, void* stream) {
    OTB_CHECK_ARG(out != nullptr, "null output");
    OTB_BIND(out);
    int c = check_launch(a);
    if (c < 0) return c;
    if (c == 0) {
        OTB_CHECK_ARG(partials != nullptr, "null partials workspace");
        grad_sqnorm_kernel<<<a->n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(*a, grad_prescale, partials);
        OTB_LAUNCH_CHECK("grad_sqnorm_kernel");
    }
    grad_norm_finish_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, c == 0 ? a->n_chunks : 0, max_norm, grad_prescale, out);
    OTB_LAUNCH_CHECK("grad_norm_finish_kernel");
    return OTB_OK;
}

extern "C" int otb_grad_norm_from_partials(const float* partials, int32_t n_partials, float max_norm, float grad_prescale,
                                           float* out, void* stream) {
    OTB_CHECK_ARG(out != nullptr && partials != nullptr && n_partials > 0, "null output / partials");
    OTB_BIND(out);
    grad_norm_finish_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, n_partials, max_norm, grad_prescale, out);
    OTB_LAUNCH_CHECK("grad_norm_finish_kernel");
    return OTB_OK;
}

extern "C" int otb_grad_scale_multi(const otb_optim_launch* a, const float* coef, void* stream) {
    OTB_CHECK_ARG(coef != nullptr, "null coef");
    int c = check_launch(a);
    if (c) return c < 0 ? c : OTB_OK;
    grad_scale_kernel<<<a->n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(*a, coef);
    OTB_LAUNCH_CHECK("grad_scale_kernel");
    return OTB_OK;
}

extern "C" int otb_shadow_refresh_multi(const otb_shadow_tensor* tensors, int32_t n_tensors, int32_t n_tiles, void* stream) {
    if (n_tensors == 0 || n_tiles == 0) return OTB_OK;
    OTB_CHECK_ARG(tensors != nullptr && n_tensors > 0 && n_tiles > 0, "bad shadow table");
    OTB_BIND(tensors);
    shadow_refresh_kernel<<<n_tiles, 256, 0, (cudaStream_t)stream>>>(tensors, n_tensors);
    OTB_LAUNCH_CHECK("shadow_refresh_kernel");
    return OTB_OK;
}
