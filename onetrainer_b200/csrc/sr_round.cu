// sr_round.cu — the reference's bf16 stochastic-rounding helpers as ONE launch each (HBM bound).
//
// Replaces modules/util/bf16_stochastic_rounding.py of the reference:
//   copy_stochastic_(target, source)                    :5-30   randint_like + add_ + bitwise_and_ + copy_  (4 launches, an
//                                                                int32 temporary of 4 B per element)
//   add_stochastic_(input, other, alpha)                :33-45  clone / to(fp32) + add_ + the four above
//   addcdiv_stochastic_(input, tensor1, tensor2, value) :48-61  clone / to(fp32) + addcdiv_ + the four above
// The fp32 intermediate never leaves registers and the 16 random bits per element come from the library's Philox stream
// (OTB_STREAM_SR, one Philox4x32-10 call per 8 elements — the same assignment K-c uses, restated in oracle/otb_oracle.c:
// sr_rand16).  Algorithmic bytes per element: copy 4 + 2, add 2 + {4|2} + 2, addcdiv 2 + 2·{4|2} + 2.
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int SR_THREADS = 256;
// groups of 8 elements per thread, all loaded before any is stored: 2 where the arithmetic is light (copy, add); addcdiv's 16
// IEEE divisions per thread cost registers (74 -> 3 blocks per SM), so it runs one group per thread at twice the occupancy.
// Measured at n = 2^28 (profiles/r03b_logs/sr_ab.log): copy 0.235 ms = 6.84 TB/s, add 0.353 ms = 6.09 TB/s, addcdiv 0.449 ms =
// 4.79 TB/s with one group (0.603 ms with two; 0.681 ms before the flavour became a compile-time constant) — the reference's
// op sequence on the same B200: 1.57 / 2.74 / 4.14 ms.  addcdiv is issue-bound: 8 correctly rounded divisions per group.
template <int MODE> constexpr int sr_unroll() { return MODE == OTB_SR_ADDCDIV ? 1 : 2; }

__device__ __forceinline__ uint4 ld_u4_plain(const void* p) { return *reinterpret_cast<const uint4*>(p); }

// 8 consecutive values of a bf16 or fp32 operand as floats
template <bool BF16>
__device__ __forceinline__ void load8(const void* base, int64_t i, bool vec, int cnt, float (&f)[8]) {
    if (vec) {
        if (BF16) {
            // the operand may alias the target (add_stochastic_(p, p, alpha)): plain loads, not the read-only path
            const uint4 r = ld_u4_plain((const char*)base + i * 2);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                f[2 * k] = __uint_as_float(w[k] << 16);
                f[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
            }
        } else {
            const uint4 a = ld_stream_u4((const char*)base + i * 4), b = ld_stream_u4((const char*)base + i * 4 + 16);
            f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
            f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            f[k] = k < cnt ? (BF16 ? bf16_bits_to_float(((const uint16_t*)base)[i + k]) : ((const float*)base)[i + k]) : 0.0f;
    }
}

__device__ __forceinline__ uint32_t sr16(const Philox4& r, int lane) {
    const uint32_t w = (lane >> 1) == 0 ? r.x : (lane >> 1) == 1 ? r.y : (lane >> 1) == 2 ? r.z : r.w;
    return (lane & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// FL: the arithmetic flavour at compile time (the hot instantiation is OTB_FLAVOR_TORCH_CUDA, what the module passes), or -1 to
// read a.flavor at run time — with the flavour fixed each element carries one division site instead of two.
template <int MODE, bool SRC_BF16, int FL, int SR_UNROLL>
__global__ void __launch_bounds__(SR_THREADS) sr_round_kernel(const otb_sr_args a, const int vec_ok) {
    constexpr int SR_GROUPS_PER_BLOCK = SR_THREADS * SR_UNROLL;
    const uint32_t fl = FL >= 0 ? (uint32_t)FL : a.flavor;
    const int64_t groups = (a.n + 7) >> 3;
    const int64_t g0 = (int64_t)blockIdx.x * SR_GROUPS_PER_BLOCK + threadIdx.x;
    float in[SR_UNROLL][8], s1[SR_UNROLL][8], s2[SR_UNROLL][8];
    // phase 1: every load of this thread's groups is issued before the first store (bytes in flight; and `src` may be `target`)
#pragma unroll
    for (int u = 0; u < SR_UNROLL; ++u) {
        const int64_t g = g0 + (int64_t)u * SR_THREADS;
        if (g < groups) {
            const int64_t i = g << 3;
            const int cnt = (int)((a.n - i) < 8 ? (a.n - i) : 8);
            const bool vec = vec_ok && cnt == 8;
            if (MODE != OTB_SR_COPY) load8<true>(a.target, i, vec, cnt, in[u]);
            load8<SRC_BF16>(a.src, i, vec, cnt, s1[u]);
            if (MODE == OTB_SR_ADDCDIV) load8<SRC_BF16>(a.src2, i, vec, cnt, s2[u]);
        }
    }
    // phase 2: the fp32 result, the integer trick, one 128-bit store per group
#pragma unroll
    for (int u = 0; u < SR_UNROLL; ++u) {
        const int64_t g = g0 + (int64_t)u * SR_THREADS;
        if (g >= groups) continue;
        const int64_t i = g << 3;
        const int cnt = (int)((a.n - i) < 8 ? (a.n - i) : 8);
        const Philox4 rnd = otb_rng(a.seed, a.offset, OTB_STREAM_SR, (uint64_t)((a.rng_base >> 3) + g));
        uint32_t hi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float r;
            if (MODE == OTB_SR_COPY) {
                r = s1[u][k];
            } else if (MODE == OTB_SR_ADD) {                          // result = fp32(other); result.add_(input, alpha=alpha)
                r = (fl & OTB_FLAVOR_FMA_ADD_ALPHA) ? __fmaf_rn(a.alpha, in[u][k], s1[u][k])
                                                          : __fadd_rn(s1[u][k], __fmul_rn(a.alpha, in[u][k]));
            } else {                                                  // result = fp32(input); result.addcdiv_(t1, t2, value=alpha)
                if (fl & OTB_FLAVOR_ADDCDIV_MUL_LAST) {
                    const float q = __fdiv_rn(s1[u][k], s2[u][k]);
                    r = (fl & OTB_FLAVOR_FMA_ADDCDIV) ? __fmaf_rn(a.alpha, q, in[u][k]) : __fadd_rn(in[u][k], __fmul_rn(a.alpha, q));
                } else {
                    r = __fadd_rn(in[u][k], __fdiv_rn(__fmul_rn(a.alpha, s1[u][k]), s2[u][k]));
                }
            }
            hi[k] = (__float_as_uint(r) + sr16(rnd, k)) & 0xFFFF0000u;   // copy_stochastic_: int32 add, mask, keep the high half
        }
        if (vec_ok && cnt == 8) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w[k] = (hi[2 * k] >> 16) | hi[2 * k + 1];
            st_stream_u4((char*)a.target + i * 2, make_uint4(w[0], w[1], w[2], w[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < cnt) ((uint16_t*)a.target)[i + k] = (uint16_t)(hi[k] >> 16);
        }
    }
}

template <int MODE, int FL, int SR_UNROLL>
int launch_sr_unrolled(const otb_sr_args& a, int vec_ok, cudaStream_t st) {
    constexpr int SR_GROUPS_PER_BLOCK = SR_THREADS * SR_UNROLL;
    const int64_t groups = (a.n + 7) >> 3;
    const int64_t blocks = (groups + SR_GROUPS_PER_BLOCK - 1) / SR_GROUPS_PER_BLOCK;
    OTB_CHECK_ARG(blocks <= 0x7FFFFFFF, "too many elements for one launch: %lld", (long long)a.n);
    if (a.src_dtype == OTB_BF16) sr_round_kernel<MODE, true, FL, SR_UNROLL><<<(unsigned)blocks, SR_THREADS, 0, st>>>(a, vec_ok);
    else sr_round_kernel<MODE, false, FL, SR_UNROLL><<<(unsigned)blocks, SR_THREADS, 0, st>>>(a, vec_ok);
    OTB_LAUNCH_CHECK("sr_round_kernel");
    return OTB_OK;
}

// tuning knob for A/B runs on the GPU box (unset in production: the compiled default is used)
template <int MODE, int FL>
int launch_sr_flavor(const otb_sr_args& a, int vec_ok, cudaStream_t st) {
    if constexpr (MODE == OTB_SR_ADDCDIV) {
        const char* v = getenv("OTB_SR_UNROLL");
        if (v && atoi(v) == 2) return launch_sr_unrolled<MODE, FL, 2>(a, vec_ok, st);
    }
    return launch_sr_unrolled<MODE, FL, sr_unroll<MODE>()>(a, vec_ok, st);
}

template <int MODE>
int launch_sr(const otb_sr_args& a, int vec_ok, cudaStream_t st) {
    // only three flavour bits reach this arithmetic; COPY has none
    constexpr uint32_t USED = OTB_FLAVOR_FMA_ADD_ALPHA | OTB_FLAVOR_ADDCDIV_MUL_LAST | OTB_FLAVOR_FMA_ADDCDIV;
    if constexpr (MODE == OTB_SR_COPY) {
        return launch_sr_flavor<MODE, 0>(a, vec_ok, st);
    } else {
        if ((a.flavor & USED) == (OTB_FLAVOR_TORCH_CUDA & USED)) return launch_sr_flavor<MODE, (int)(OTB_FLAVOR_TORCH_CUDA & USED)>(a, vec_ok, st);
        return launch_sr_flavor<MODE, -1>(a, vec_ok, st);
    }
}

}  // namespace

extern "C" int otb_stochastic_round(const otb_sr_args* a, void* stream) {
    OTB_CHECK_ARG(a != nullptr, "null args");
    OTB_CHECK_ARG(a->n >= 0, "negative element count");
    if (a->n == 0) return OTB_OK;
    OTB_CHECK_ARG(a->target && a->src, "null target / src");
    OTB_CHECK_ARG(a->mode == OTB_SR_COPY || a->mode == OTB_SR_ADD || a->mode == OTB_SR_ADDCDIV, "unknown mode %d", a->mode);
    OTB_CHECK_ARG(a->mode != OTB_SR_ADDCDIV || a->src2, "addcdiv needs tensor2");
    OTB_UNSUPPORTED(a->src_dtype != OTB_F32 && a->src_dtype != OTB_BF16, "operands must be fp32 or bf16");
    OTB_CHECK_ARG(a->mode != OTB_SR_COPY || a->src_dtype == OTB_F32, "copy_stochastic_ takes an fp32 source");
    OTB_CHECK_ARG(a->rng_base >= 0 && (a->rng_base & 7) == 0, "rng_base must be a non-negative multiple of 8");
    OTB_CHECK_ARG(((uintptr_t)a->target & 1) == 0 && ((uintptr_t)a->src & (a->src_dtype == OTB_F32 ? 3 : 1)) == 0, "misaligned tensor");
    OTB_BIND(a->target);
    const int vec_ok = (((uintptr_t)a->target | (uintptr_t)a->src | (uintptr_t)a->src2) & 15) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (a->mode == OTB_SR_COPY) return launch_sr<OTB_SR_COPY>(*a, vec_ok, st);
    if (a->mode == OTB_SR_ADD) return launch_sr<OTB_SR_ADD>(*a, vec_ok, st);
    return launch_sr<OTB_SR_ADDCDIV>(*a, vec_ok, st);
}
