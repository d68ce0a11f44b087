// optim_factored.cu — multi-tensor Adafactor and CAME steps.
//
// Reference behaviour restated:
//   step_adafactor_parameter   /root/reference/modules/util/optimizer/adafactor_extensions.py:17-93
//                              (over transformers' Adafactor._get_lr / _rms / _approx_sq_grad)
//   CAME.step_parameter        /root/reference/modules/util/optimizer/CAME.py:80-176
//   add_stochastic_ / copy_stochastic_   modules/util/bf16_stochastic_rounding.py:5-45
//
// Both optimizers keep a FACTORED second moment for every tensor with >= 2 dimensions: a row vector and a column vector
// per [rows, cols] matrix (leading dimensions are a batch of matrices), updated from the row / column means of g^2, and
// the update needs two more tensor-wide reductions (the mean of the row vector, the RMS of the update) before a single
// element can be written.  The reference runs ~25 (Adafactor) / ~45 (CAME) ATen launches with full-size fp32 temporaries
// per tensor, in a python loop over 1444 LoRA tensors, and reads RMS(p) back to the host (`max(eps, RMS)` on a tensor is
// a device sync per tensor).
//
// Here: ONE launch per size class and optimizer step, one thread-block CLUSTER per tensor (1 CTA for LoRA-sized tensors,
// 8 CTAs for the matrices of a full fine-tune; the host sorts the table by size and passes the cluster width).  A cluster
// walks its tensor in phases separated by cluster barriers — every reduction a tensor needs is local to its cluster
// (block sums exchanged through distributed shared memory, added in CTA-rank order), so there is no grid-wide
// synchronisation, no atomics, no temporaries in HBM, and every sum is formed in a fixed order (bit-reproducible for a
// given cluster width):
//   Adafactor   A  RMS(p); row / column means of g^2+eps0 -> factored state (or the unfactored state)
//               B  mean of the row factor per matrix; RMS of u = g * rsqrt(row/mean) * rsqrt(col)
//               C  u / max(1, RMS/clip) * lr_t, first moment, weight decay, p -= u, one rounding to p's dtype
//   CAME        A, B as above;  C  u-hat, first moment m, row / column means of (u-hat - m)^2 + eps1 -> instability factors
//               D  p += -wd*lr*p ; p += -lr * m * rsqrt-factors   (each a bf16 stochastic rounding when asked for)
// The gradient is re-read from L2 / HBM in each phase (4 reads) instead of being kept in shared memory, so tensors of any
// size work (full fine-tuning: 13 M-element matrices); LoRA tensors (<= 41 k elements) stay L2-resident between phases.
// lr_t = max(eps1, RMS(p)) * step size is formed on the device: no host read-back.
#include "common.cuh"

#define FO_THREADS 512
#define FO_WARPS (FO_THREADS / 32)

namespace {


// deterministic block sum: strided per-thread partials -> warp shuffle tree -> warp partials summed in warp order
__device__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();                       // `red` may still be read from the previous reduction
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < FO_WARPS; ++w) t += red[w];
    return t;
}

// ---- cluster plumbing: rank / size of this CTA in its cluster, barrier, remote shared-memory read
__device__ __forceinline__ uint32_t fo_crank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t fo_csize() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t fo_cid() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
// release / acquire at cluster scope: global and shared writes of every CTA before the barrier are visible to every CTA after it
__device__ __forceinline__ void fo_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float fo_ld_remote(const float* slot, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(slot), ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}
// sum over the whole cluster: block sums, exchanged through DSMEM, added in CTA-rank order by every CTA (same bits everywhere)
__device__ float cluster_sum(float v, float* red, float* slot) {
    const float t = block_sum(v, red);
    const uint32_t cs = fo_csize();
    if (cs == 1) return t;
    if (threadIdx.x == 0) *slot = t;
    fo_cluster_sync();
    float s = 0.0f;
    for (uint32_t r = 0; r < cs; ++r) s += fo_ld_remote(slot, r);
    fo_cluster_sync();                     // nobody overwrites its slot while a peer still reads it
    return s;
}
// the phases of a tensor communicate through global memory (factors, first moment) across the CTAs of its cluster
__device__ __forceinline__ void phase_sync() {
    if (fo_csize() == 1) __syncthreads(); else fo_cluster_sync();
}

struct Dims { int batch, rows, cols; int64_t numel; };

// ---- W consecutive elements at a time (W = 4: 16-byte accesses for fp32, 8-byte for bf16 / fp16; W = 1: any shape)
template <int W>
__device__ __forceinline__ void ld_w(const void* p, int64_t i, int dt, float (&f)[W]) {
    if (W == 4) {
        if (dt == OTB_F32) {
            const float4 v = *reinterpret_cast<const float4*>((const float*)p + i);
            f[0] = v.x; f[1 % W] = v.y; f[2 % W] = v.z; f[3 % W] = v.w;
        } else {
            const uint2 v = *reinterpret_cast<const uint2*>((const uint16_t*)p + i);
            if (dt == OTB_BF16) {
                f[0] = __uint_as_float(v.x << 16); f[1 % W] = __uint_as_float(v.x & 0xFFFF0000u);
                f[2 % W] = __uint_as_float(v.y << 16); f[3 % W] = __uint_as_float(v.y & 0xFFFF0000u);
            } else {
                const __half2 h0 = *reinterpret_cast<const __half2*>(&v.x), h1 = *reinterpret_cast<const __half2*>(&v.y);
                f[0] = __low2float(h0); f[1 % W] = __high2float(h0); f[2 % W] = __low2float(h1); f[3 % W] = __high2float(h1);
            }
        }
    } else {
        f[0] = load_as_float(p, i, dt);
    }
}
template <int W>
__device__ __forceinline__ void st_w(void* p, int64_t i, int dt, const float (&f)[W]) {
    if (W == 4 && dt == OTB_F32) {
        *reinterpret_cast<float4*>((float*)p + i) = make_float4(f[0], f[1 % W], f[2 % W], f[3 % W]);
    } else {
#pragma unroll
        for (int k = 0; k < W; ++k) store_from_float(p, i + k, dt, f[k]);
    }
}
template <int W>
__device__ __forceinline__ void ldf_w(const float* p, int64_t i, float (&f)[W]) {
    if (W == 4) {
        const float4 v = *reinterpret_cast<const float4*>(p + i);
        f[0] = v.x; f[1 % W] = v.y; f[2 % W] = v.z; f[3 % W] = v.w;
    } else {
        f[0] = p[i];
    }
}
template <int W>
__device__ __forceinline__ void stf_w(float* p, int64_t i, const float (&f)[W]) {
    if (W == 4) *reinterpret_cast<float4*>(p + i) = make_float4(f[0], f[1 % W], f[2 % W], f[3 % W]);
    else p[i] = f[0];
}

// row / column means of f(i) over every [rows, cols] matrix of the tensor, folded into the running factors:
//   row[b, r] = row * beta + (1 - beta) * mean_c f ,  col[b, c] = col * beta + (1 - beta) * mean_r f
// f(i, out[W]) yields W consecutive elements of one row (W divides cols)
template <int W, typename F>
__device__ void factor_update(const Dims d, float* row, float* col, float beta, float omb, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t cr = fo_crank(), cs = fo_csize();
    const int64_t total_rows = (int64_t)d.batch * d.rows;
    for (int64_t br = cr * FO_WARPS + warp; br < total_rows; br += cs * FO_WARPS) {   // one warp per row: coalesced, shuffle reduce
        const int64_t base = br * d.cols;
        float acc = 0.0f;
        for (int c = lane * W; c < d.cols; c += 32 * W) {
            float v[W];
            f(base + c, v);
#pragma unroll
            for (int k = 0; k < W; ++k) acc += v[k];
        }
        acc = warp_sum(acc);
        if (lane == 0) row[br] = __fmaf_rn(omb, acc / (float)d.cols, __fmul_rn(row[br], beta));
    }
    const int cw = d.cols / W;                                                  // column groups per matrix
    const int64_t total_groups = (int64_t)d.batch * cw;
    for (int64_t bg = cr * FO_THREADS + threadIdx.x; bg < total_groups; bg += cs * FO_THREADS) {   // one thread per W columns
        const int64_t b = bg / cw;
        const int c = (int)(bg - b * cw) * W;
        const int64_t base = b * (int64_t)d.rows * d.cols + c;
        float acc[W];
#pragma unroll
        for (int k = 0; k < W; ++k) acc[k] = 0.0f;
        int r = 0;
        for (; r + 4 <= d.rows; r += 4) {                                       // four independent loads in flight, added in row order
            float v[4][W];
#pragma unroll
            for (int q = 0; q < 4; ++q) f(base + (int64_t)(r + q) * d.cols, v[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < W; ++k) acc[k] += v[q][k];
        }
        for (; r < d.rows; ++r) {
            float v[W];
            f(base + (int64_t)r * d.cols, v);
#pragma unroll
            for (int k = 0; k < W; ++k) acc[k] += v[k];
        }
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const int64_t bc = b * d.cols + c + k;
            col[bc] = __fmaf_rn(omb, acc[k] / (float)d.rows, __fmul_rn(col[bc], beta));
        }
    }
}

// mean of the row factor of every matrix -> scratch[b]  (batch == 1: one block reduction; otherwise a thread per matrix)
__device__ void row_means(const Dims d, const float* row, float* scratch, float* red) {
    if (d.batch == 1) {
        float acc = 0.0f;
        for (int r = threadIdx.x; r < d.rows; r += FO_THREADS) acc += row[r];
        const float t = block_sum(acc, red);
        if (threadIdx.x == 0) scratch[0] = t / (float)d.rows;
    } else {
        for (int b = threadIdx.x; b < d.batch; b += FO_THREADS) {
            float acc = 0.0f;
            for (int r = 0; r < d.rows; ++r) acc += row[(int64_t)b * d.rows + r];
            scratch[b] = acc / (float)d.rows;
        }
    }
    __syncthreads();
}

// _approx_sq_grad for the W consecutive elements starting at i (one row): rsqrt(row / mean(row)) * rsqrt(col).
// 32-bit index arithmetic: every tensor is far below 2^31 elements (checked by the host), and a 64-bit division per
// element was most of the instruction count of these loops.
template <int W>
__device__ __forceinline__ void approx_factor(const Dims d, const float* row, const float* col, const float* mean, int64_t i, float (&a)[W]) {
    const uint32_t mat = (uint32_t)d.rows * (uint32_t)d.cols, ii = (uint32_t)i;
    const uint32_t b = d.batch == 1 ? 0u : ii / mat, rem = ii - b * mat;
    const uint32_t r = rem / (uint32_t)d.cols, c = rem - r * (uint32_t)d.cols;
    const float rf = rsqrtf(row[(int64_t)b * d.rows + r] / mean[b]);
    float cf[W];
    ldf_w<W>(col, (int64_t)b * d.cols + c, cf);
#pragma unroll
    for (int k = 0; k < W; ++k) a[k] = rf * rsqrtf(cf[k]);
}

__device__ __forceinline__ uint32_t sr_rand16(uint64_t seed, uint64_t offset, uint64_t elem) {
    Philox4 r = otb_rng(seed, offset, OTB_STREAM_SR, elem >> 3);
    const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
    const uint32_t w = w4[(elem & 7) >> 1];
    return (elem & 1) ? (w >> 16) : (w & 0xffffu);
}
// copy_stochastic_ (bf16_stochastic_rounding.py:5-30): int32 view + 16 random bits, low half masked off
__device__ __forceinline__ float stochastic_bf16(float x, uint32_t r16) {
    return __int_as_float((__float_as_int(x) + (int)r16) & (int)0xFFFF0000);
}

// W = 4 needs: numel and cols multiples of 4 (a group never straddles a row), every pointer 16-byte aligned (8 for 16-bit p / g)
__device__ __forceinline__ bool can_vec4(const otb_factored_tensor& t) {
    const uintptr_t f32s = (uintptr_t)t.exp_avg | (uintptr_t)t.sq | (uintptr_t)t.sq_col | (uintptr_t)t.res_col;
    const uintptr_t pg = (uintptr_t)t.p | (uintptr_t)t.g;
    return (t.numel & 3) == 0 && (t.cols & 3) == 0 && (f32s & 15) == 0 && (pg & (t.p_dtype == OTB_F32 ? 15 : 7)) == 0 && t.numel < (int64_t)1 << 31;
}

struct FoShared { float* red; float* sh; float* slot; };

template <int W>
__device__ void adafactor_body(const otb_factored_tensor& t, float* scratch, const FoShared S) {
    float* red = S.red; float* sh = S.sh; float* slot = S.slot;
    const Dims d{t.batch, t.rows, t.cols, t.numel};
    const int64_t e0 = ((int64_t)fo_crank() * FO_THREADS + threadIdx.x) * W, estride = (int64_t)fo_csize() * FO_THREADS * W;
    const bool factored = (t.flags & OTB_FACTORED_FACTORED) != 0, first = (t.flags & OTB_FACTORED_FIRST_MOMENT) != 0;
    const int pdt = t.p_dtype;
    // ---- A: RMS(p); second-moment state
    float acc = 0.0f;
    for (int64_t i = e0; i < d.numel; i += estride) {
        float p[W];
        ld_w<W>(t.p, i, pdt, p);
#pragma unroll
        for (int k = 0; k < W; ++k) acc += p[k] * p[k];
    }
    const float sum_p2 = cluster_sum(acc, red, &slot[0]);
    if (threadIdx.x == 0) {
        const float rms_p = sqrtf(sum_p2) / sqrtf((float)d.numel);          // _rms: norm(2) / numel ** 0.5
        if (t.rms_out && fo_crank() == 0) *t.rms_out = rms_p;
        const float scale = (t.flags & OTB_FACTORED_SCALE_PARAMETER) ? fmaxf(t.eps1, rms_p) : 1.0f;
        sh[0] = scale * t.lr;                                                 // _get_lr: param_scale * rel_step_sz
    }
    auto g2 = [&](int64_t i, float (&o)[W]) {
        ld_w<W>(t.g, i, pdt, o);
#pragma unroll
        for (int k = 0; k < W; ++k) o[k] = __fmaf_rn(o[k], o[k], t.eps0);
    };
    if (factored) {
        factor_update<W>(d, t.sq_row, t.sq_col, t.beta2, t.one_minus_beta2, g2);
    } else {
        const float omb = t.one_minus_beta2;
        for (int64_t i = e0; i < d.numel; i += estride) {
            float q[W], s2[W];
            g2(i, q);
            ldf_w<W>(t.sq, i, s2);
#pragma unroll
            for (int k = 0; k < W; ++k) s2[k] = __fmaf_rn(omb, q[k], __fmul_rn(s2[k], t.beta2));
            stf_w<W>(t.sq, i, s2);
        }
    }
    phase_sync();                                 // the factors of every row / column are final, whichever CTA formed them
    // ---- B: RMS of the update
    if (factored) row_means(d, t.sq_row, scratch, red);
    auto upd = [&](int64_t i, float (&u)[W]) {
        float g[W], a[W];
        ld_w<W>(t.g, i, pdt, g);
        if (factored) approx_factor<W>(d, t.sq_row, t.sq_col, scratch, i, a);
        else {
            ldf_w<W>(t.sq, i, a);
#pragma unroll
            for (int k = 0; k < W; ++k) a[k] = rsqrtf(a[k]);
        }
#pragma unroll
        for (int k = 0; k < W; ++k) u[k] = a[k] * g[k];
    };
    acc = 0.0f;
    for (int64_t i = e0; i < d.numel; i += estride) {
        float u[W];
        upd(i, u);
#pragma unroll
        for (int k = 0; k < W; ++k) acc += u[k] * u[k];
    }
    const float sum_u2 = cluster_sum(acc, red, &slot[1]);
    if (threadIdx.x == 0) sh[1] = fmaxf(sqrtf(sum_u2) / sqrtf((float)d.numel) / t.clip_threshold, 1.0f);
    __syncthreads();
    const float lr_t = sh[0], denom = sh[1];
    const float wd_alpha = -t.weight_decay * lr_t;
    // ---- C: apply
    for (int64_t i = e0; i < d.numel; i += estride) {
        float u[W], p[W];
        upd(i, u);
#pragma unroll
        for (int k = 0; k < W; ++k) u[k] = u[k] / denom * lr_t;
        if (first) {
            float m[W];
            ldf_w<W>(t.exp_avg, i, m);
#pragma unroll
            for (int k = 0; k < W; ++k) { m[k] = __fmaf_rn(t.one_minus_beta1, u[k], __fmul_rn(m[k], t.beta1)); u[k] = m[k]; }
            stf_w<W>(t.exp_avg, i, m);
        }
        ld_w<W>(t.p, i, pdt, p);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            if (t.weight_decay != 0.0f) p[k] = __fmaf_rn(wd_alpha, p[k], p[k]);
            p[k] = p[k] - u[k];
        }
        st_w<W>(t.p, i, pdt, p);                   // bf16: round to nearest — the reference overwrites its stochastic rounding (:88-91)
    }
}

__global__ void __launch_bounds__(FO_THREADS)
adafactor_multi_kernel(const otb_factored_tensor* __restrict__ tensors, float* __restrict__ scratch_all) {
    __shared__ float red[FO_WARPS];
    __shared__ float sh[2];
    __shared__ float slot[2];
    const otb_factored_tensor t = tensors[fo_cid()];
    const FoShared S{red, sh, slot};
    if (can_vec4(t)) adafactor_body<4>(t, scratch_all + t.scratch_offset, S);      // block- (and cluster-) uniform
    else adafactor_body<1>(t, scratch_all + t.scratch_offset, S);
}

template <int W>
__device__ void came_body(const otb_factored_tensor& t, float* scratch, const FoShared S, uint64_t seed, uint64_t offset) {
    float* red = S.red; float* sh = S.sh; float* slot = S.slot;
    const Dims d{t.batch, t.rows, t.cols, t.numel};
    const int64_t e0 = ((int64_t)fo_crank() * FO_THREADS + threadIdx.x) * W, estride = (int64_t)fo_csize() * FO_THREADS * W;
    const bool factored = (t.flags & OTB_FACTORED_FACTORED) != 0;
    const bool sr = (t.flags & OTB_FACTORED_SR) != 0 && t.p_dtype == OTB_BF16;
    const int pdt = t.p_dtype;
    // ---- RMS(p) is only recorded (state["RMS"], CAME.py:118)
    if (t.rms_out) {
        float a0 = 0.0f;
        for (int64_t i = e0; i < d.numel; i += estride) {
            float p[W];
            ld_w<W>(t.p, i, pdt, p);
#pragma unroll
            for (int k = 0; k < W; ++k) a0 += p[k] * p[k];
        }
        const float s = cluster_sum(a0, red, &slot[0]);
        if (threadIdx.x == 0 && fo_crank() == 0) *t.rms_out = sqrtf(s) / sqrtf((float)d.numel);
    }
    // ---- A: second moment
    auto g2 = [&](int64_t i, float (&o)[W]) {
        ld_w<W>(t.g, i, pdt, o);
#pragma unroll
        for (int k = 0; k < W; ++k) o[k] = __fmaf_rn(o[k], o[k], t.eps0);
    };
    if (factored) {
        factor_update<W>(d, t.sq_row, t.sq_col, t.beta2, t.one_minus_beta2, g2);
    } else {
        const float omb = t.one_minus_beta2;
        for (int64_t i = e0; i < d.numel; i += estride) {
            float q[W], s2[W];
            g2(i, q);
            ldf_w<W>(t.sq, i, s2);
#pragma unroll
            for (int k = 0; k < W; ++k) s2[k] = __fmaf_rn(omb, q[k], __fmul_rn(s2[k], t.beta2));
            stf_w<W>(t.sq, i, s2);
        }
    }
    phase_sync();
    // ---- B: RMS of the update
    if (factored) row_means(d, t.sq_row, scratch, red);
    auto upd = [&](int64_t i, float (&u)[W]) {
        float g[W], a[W];
        ld_w<W>(t.g, i, pdt, g);
        if (factored) approx_factor<W>(d, t.sq_row, t.sq_col, scratch, i, a);
        else {
            ldf_w<W>(t.sq, i, a);
#pragma unroll
            for (int k = 0; k < W; ++k) a[k] = rsqrtf(a[k]);
        }
#pragma unroll
        for (int k = 0; k < W; ++k) u[k] = a[k] * g[k];
    };
    float acc = 0.0f;
    for (int64_t i = e0; i < d.numel; i += estride) {
        float u[W];
        upd(i, u);
#pragma unroll
        for (int k = 0; k < W; ++k) acc += u[k] * u[k];
    }
    const float sum_u2 = cluster_sum(acc, red, &slot[1]);
    if (threadIdx.x == 0) sh[0] = fmaxf(sqrtf(sum_u2) / sqrtf((float)d.numel) / t.clip_threshold, 1.0f);
    __syncthreads();
    const float denom = sh[0];
    // ---- C: first moment; instability factors from (u-hat - m)^2 + eps1
    for (int64_t i = e0; i < d.numel; i += estride) {
        float u[W], m[W];
        upd(i, u);
        ldf_w<W>(t.exp_avg, i, m);
#pragma unroll
        for (int k = 0; k < W; ++k) m[k] = __fmaf_rn(t.one_minus_beta1, u[k] / denom, __fmul_rn(m[k], t.beta1));
        stf_w<W>(t.exp_avg, i, m);
    }
    phase_sync();                                 // the row / column passes below read first moments other CTAs wrote
    if (factored) {
        auto res = [&](int64_t i, float (&o)[W]) {
            float u[W], m[W];
            upd(i, u);
            ldf_w<W>(t.exp_avg, i, m);
#pragma unroll
            for (int k = 0; k < W; ++k) { const float e = u[k] / denom - m[k]; o[k] = __fmaf_rn(e, e, t.eps1); }
        };
        factor_update<W>(d, t.res_row, t.res_col, t.beta3, t.one_minus_beta3, res);
        phase_sync();                             // (also: nobody overwrites scratch[] while a peer still forms u-hat with it)
        row_means(d, t.res_row, scratch, red);
    }
    // ---- D: apply (CAME.py:160-176)
    const float wd_alpha = -t.weight_decay * t.lr;
    for (int64_t i = e0; i < d.numel; i += estride) {
        float m[W], a[W], p[W];
        ldf_w<W>(t.exp_avg, i, m);
        if (factored) approx_factor<W>(d, t.res_row, t.res_col, scratch, i, a);
        ld_w<W>(t.p, i, pdt, p);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const float u = (factored ? a[k] * m[k] : m[k]) * t.lr;
            float pk = p[k];
            if (t.weight_decay != 0.0f) {
                pk = __fmaf_rn(wd_alpha, pk, pk);                                // add_stochastic_(p, p, alpha): p.float() + alpha * p
                if (sr) pk = stochastic_bf16(pk, sr_rand16(seed, offset, (uint64_t)(t.rng_base + i + k)));
                else if (pdt == OTB_BF16) pk = round_bf16(pk);
            }
            pk = -u + pk;                                                        // add_stochastic_(p, -update)
            if (sr) pk = stochastic_bf16(pk, sr_rand16(seed, offset + (t.weight_decay != 0.0f ? 1 : 0), (uint64_t)(t.rng_base + i + k)));
            p[k] = pk;
        }
        st_w<W>(t.p, i, pdt, p);
    }
}

__global__ void __launch_bounds__(FO_THREADS)
came_multi_kernel(const otb_factored_tensor* __restrict__ tensors, float* __restrict__ scratch_all, uint64_t seed, uint64_t offset) {
    __shared__ float red[FO_WARPS];
    __shared__ float sh[1];
    __shared__ float slot[2];
    const otb_factored_tensor t = tensors[fo_cid()];
    const FoShared S{red, sh, slot};
    if (can_vec4(t)) came_body<4>(t, scratch_all + t.scratch_offset, S, seed, offset);
    else came_body<1>(t, scratch_all + t.scratch_offset, S, seed, offset);
}

int check_launch(const otb_factored_launch* a) {
    OTB_CHECK_ARG(a != nullptr && a->tensors != nullptr && a->n_tensors >= 0, "null / negative factored launch");
    OTB_CHECK_ARG(a->cluster == 0 || a->cluster == 1 || a->cluster == 2 || a->cluster == 4 || a->cluster == 8,
                  "cluster width %d (CTAs per tensor) must be 1, 2, 4 or 8", a->cluster);
    OTB_CHECK_ARG(a->n_tensors == 0 || a->scratch != nullptr, "factored launch needs its scratch buffer (one float per matrix of every tensor)");
    return OTB_OK;
}

// one cluster of `cluster` CTAs per tensor
void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, const otb_factored_launch* a, void* stream) {
    const unsigned cs = a->cluster > 0 ? (unsigned)a->cluster : 1u;
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3((unsigned)a->n_tensors * cs);
    cfg->blockDim = dim3(FO_THREADS);
    cfg->dynamicSmemBytes = 0;
    cfg->stream = (cudaStream_t)stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr; cfg->numAttrs = 1;
}

}  // namespace

extern "C" int otb_adafactor_multi(const otb_factored_launch* a, void* stream) {
    int rc = check_launch(a);
    if (rc != OTB_OK) return rc;
    if (a->n_tensors == 0) return OTB_OK;
    OTB_BIND(a->tensors);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cluster_config(&cfg, attr, a, stream);
    cudaError_t e = cudaLaunchKernelEx(&cfg, adafactor_multi_kernel, a->tensors, a->scratch);
    if (e != cudaSuccess) { otb_set_error("launch of adafactor_multi_kernel failed: %s", cudaGetErrorString(e)); return OTB_ECUDA; }
    return OTB_OK;
}

extern "C" int otb_came_multi(const otb_factored_launch* a, void* stream) {
    int rc = check_launch(a);
    if (rc != OTB_OK) return rc;
    if (a->n_tensors == 0) return OTB_OK;
    OTB_BIND(a->tensors);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cluster_config(&cfg, attr, a, stream);
    cudaError_t e = cudaLaunchKernelEx(&cfg, came_multi_kernel, a->tensors, a->scratch, a->seed, a->offset);
    if (e != cudaSuccess) { otb_set_error("launch of came_multi_kernel failed: %s", cudaGetErrorString(e)); return OTB_ECUDA; }
    return OTB_OK;
}
