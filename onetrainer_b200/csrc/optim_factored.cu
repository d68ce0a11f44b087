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

__device__ __forceinline__ float ldf(const void* p, int64_t i, int dt) { return load_as_float(p, i, dt); }

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

// row / column means of f(i) over every [rows, cols] matrix of the tensor, folded into the running factors:
//   row[b, r] = row * beta + (1 - beta) * mean_c f ,  col[b, c] = col * beta + (1 - beta) * mean_r f
template <typename F>
__device__ void factor_update(const Dims d, float* row, float* col, float beta, float omb, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t cr = fo_crank(), cs = fo_csize();
    const int64_t total_rows = (int64_t)d.batch * d.rows;
    for (int64_t br = cr * FO_WARPS + warp; br < total_rows; br += cs * FO_WARPS) {   // one warp per row: coalesced, shuffle reduce
        const int64_t base = br * d.cols;
        float acc = 0.0f;
        for (int c = lane; c < d.cols; c += 32) acc += f(base + c);
        acc = warp_sum(acc);
        if (lane == 0) row[br] = __fmaf_rn(omb, acc / (float)d.cols, __fmul_rn(row[br], beta));
    }
    const int64_t total_cols = (int64_t)d.batch * d.cols;
    for (int64_t bc = cr * FO_THREADS + threadIdx.x; bc < total_cols; bc += cs * FO_THREADS) {   // one thread per column: coalesced across threads
        const int64_t b = bc / d.cols, c = bc - b * d.cols;
        const int64_t base = b * (int64_t)d.rows * d.cols + c;
        float acc = 0.0f;
        int r = 0;
        for (; r + 8 <= d.rows; r += 8) {                                       // eight independent loads in flight, added in row order
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = f(base + (int64_t)(r + k) * d.cols);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += v[k];
        }
        for (; r < d.rows; ++r) acc += f(base + (int64_t)r * d.cols);
        col[bc] = __fmaf_rn(omb, acc / (float)d.rows, __fmul_rn(col[bc], beta));
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

// _approx_sq_grad at element i: rsqrt(row / mean(row)) * rsqrt(col)
__device__ __forceinline__ float approx_factor(const Dims d, const float* row, const float* col, const float* mean, int64_t i) {
    const int64_t mat = (int64_t)d.rows * d.cols;
    const int64_t b = i / mat, rem = i - b * mat;
    const int64_t r = rem / d.cols, c = rem - r * d.cols;
    const float rf = rsqrtf(row[b * d.rows + r] / mean[b]);
    const float cf = rsqrtf(col[b * d.cols + c]);
    return rf * cf;
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

__global__ void __launch_bounds__(FO_THREADS)
adafactor_multi_kernel(const otb_factored_tensor* __restrict__ tensors, float* __restrict__ scratch_all) {
    __shared__ float red[FO_WARPS];
    __shared__ float sh[2];
    __shared__ float slot[2];
    const otb_factored_tensor t = tensors[fo_cid()];
    const Dims d{t.batch, t.rows, t.cols, t.numel};
    const int64_t e0 = (int64_t)fo_crank() * FO_THREADS + threadIdx.x, estride = (int64_t)fo_csize() * FO_THREADS;
    const bool factored = (t.flags & OTB_FACTORED_FACTORED) != 0, first = (t.flags & OTB_FACTORED_FIRST_MOMENT) != 0;
    float* scratch = scratch_all + t.scratch_offset;
    const int pdt = t.p_dtype;
    // ---- A: RMS(p); second-moment state
    float acc = 0.0f;
    for (int64_t i = e0; i < d.numel; i += estride) { const float p = ldf(t.p, i, pdt); acc += p * p; }
    const float sum_p2 = cluster_sum(acc, red, &slot[0]);
    if (threadIdx.x == 0) {
        const float rms_p = sqrtf(sum_p2) / sqrtf((float)d.numel);          // _rms: norm(2) / numel ** 0.5
        if (t.rms_out && fo_crank() == 0) *t.rms_out = rms_p;
        const float scale = (t.flags & OTB_FACTORED_SCALE_PARAMETER) ? fmaxf(t.eps1, rms_p) : 1.0f;
        sh[0] = scale * t.lr;                                                 // _get_lr: param_scale * rel_step_sz
    }
    auto g2 = [&](int64_t i) { const float g = ldf(t.g, i, pdt); return __fmaf_rn(g, g, t.eps0); };
    if (factored) {
        factor_update(d, t.sq_row, t.sq_col, t.beta2, t.one_minus_beta2, g2);
    } else {
        const float omb = t.one_minus_beta2;
        for (int64_t i = e0; i < d.numel; i += estride) t.sq[i] = __fmaf_rn(omb, g2(i), __fmul_rn(t.sq[i], t.beta2));
    }
    phase_sync();                                 // the factors of every row / column are final, whichever CTA formed them
    // ---- B: RMS of the update
    if (factored) row_means(d, t.sq_row, scratch, red);
    auto upd = [&](int64_t i) {
        const float g = ldf(t.g, i, pdt);
        return (factored ? approx_factor(d, t.sq_row, t.sq_col, scratch, i) : rsqrtf(t.sq[i])) * g;
    };
    acc = 0.0f;
    for (int64_t i = e0; i < d.numel; i += estride) { const float u = upd(i); acc += u * u; }
    const float sum_u2 = cluster_sum(acc, red, &slot[1]);
    if (threadIdx.x == 0) sh[1] = fmaxf(sqrtf(sum_u2) / sqrtf((float)d.numel) / t.clip_threshold, 1.0f);
    __syncthreads();
    const float lr_t = sh[0], denom = sh[1];
    const float wd_alpha = -t.weight_decay * lr_t;
    // ---- C: apply
    for (int64_t i = e0; i < d.numel; i += estride) {
        float u = upd(i) / denom * lr_t;
        if (first) {
            const float m = __fmaf_rn(t.one_minus_beta1, u, __fmul_rn(t.exp_avg[i], t.beta1));
            t.exp_avg[i] = m;
            u = m;
        }
        float p = ldf(t.p, i, pdt);
        if (t.weight_decay != 0.0f) p = __fmaf_rn(wd_alpha, p, p);
        p = p - u;
        store_from_float(t.p, i, pdt, p);          // bf16: round to nearest — the reference overwrites its stochastic rounding (:88-91)
    }
}

__global__ void __launch_bounds__(FO_THREADS)
came_multi_kernel(const otb_factored_tensor* __restrict__ tensors, float* __restrict__ scratch_all, uint64_t seed, uint64_t offset) {
    __shared__ float red[FO_WARPS];
    __shared__ float sh[1];
    __shared__ float slot[2];
    const otb_factored_tensor t = tensors[fo_cid()];
    const Dims d{t.batch, t.rows, t.cols, t.numel};
    const int64_t e0 = (int64_t)fo_crank() * FO_THREADS + threadIdx.x, estride = (int64_t)fo_csize() * FO_THREADS;
    const bool factored = (t.flags & OTB_FACTORED_FACTORED) != 0;
    const bool sr = (t.flags & OTB_FACTORED_SR) != 0 && t.p_dtype == OTB_BF16;
    float* scratch = scratch_all + t.scratch_offset;
    const int pdt = t.p_dtype;
    // ---- RMS(p) is only recorded (state["RMS"], CAME.py:118)
    if (t.rms_out) {
        float a0 = 0.0f;
        for (int64_t i = e0; i < d.numel; i += estride) { const float p = ldf(t.p, i, pdt); a0 += p * p; }
        const float s = cluster_sum(a0, red, &slot[0]);
        if (threadIdx.x == 0 && fo_crank() == 0) *t.rms_out = sqrtf(s) / sqrtf((float)d.numel);
    }
    // ---- A: second moment
    auto g2 = [&](int64_t i) { const float g = ldf(t.g, i, pdt); return __fmaf_rn(g, g, t.eps0); };
    if (factored) {
        factor_update(d, t.sq_row, t.sq_col, t.beta2, t.one_minus_beta2, g2);
    } else {
        const float omb = t.one_minus_beta2;
        for (int64_t i = e0; i < d.numel; i += estride) t.sq[i] = __fmaf_rn(omb, g2(i), __fmul_rn(t.sq[i], t.beta2));
    }
    phase_sync();
    // ---- B: RMS of the update
    if (factored) row_means(d, t.sq_row, scratch, red);
    auto upd = [&](int64_t i) {
        const float g = ldf(t.g, i, pdt);
        return (factored ? approx_factor(d, t.sq_row, t.sq_col, scratch, i) : rsqrtf(t.sq[i])) * g;
    };
    float acc = 0.0f;
    for (int64_t i = e0; i < d.numel; i += estride) { const float u = upd(i); acc += u * u; }
    const float sum_u2 = cluster_sum(acc, red, &slot[1]);
    if (threadIdx.x == 0) sh[0] = fmaxf(sqrtf(sum_u2) / sqrtf((float)d.numel) / t.clip_threshold, 1.0f);
    __syncthreads();
    const float denom = sh[0];
    // ---- C: first moment; instability factors from (u-hat - m)^2 + eps1
    for (int64_t i = e0; i < d.numel; i += estride)
        t.exp_avg[i] = __fmaf_rn(t.one_minus_beta1, upd(i) / denom, __fmul_rn(t.exp_avg[i], t.beta1));
    phase_sync();                                 // the row / column passes below read first moments other CTAs wrote
    if (factored) {
        auto res = [&](int64_t i) { const float e = upd(i) / denom - t.exp_avg[i]; return __fmaf_rn(e, e, t.eps1); };
        factor_update(d, t.res_row, t.res_col, t.beta3, t.one_minus_beta3, res);
        phase_sync();                             // (also: nobody overwrites scratch[] while a peer still forms u-hat with it)
        row_means(d, t.res_row, scratch, red);
    }
    // ---- D: apply (CAME.py:160-176)
    const float wd_alpha = -t.weight_decay * t.lr;
    for (int64_t i = e0; i < d.numel; i += estride) {
        const float m = t.exp_avg[i];
        const float u = (factored ? approx_factor(d, t.res_row, t.res_col, scratch, i) * m : m) * t.lr;
        float p = ldf(t.p, i, pdt);
        if (t.weight_decay != 0.0f) {
            p = __fmaf_rn(wd_alpha, p, p);                                   // add_stochastic_(p, p, alpha): p.float() + alpha * p
            if (sr) p = stochastic_bf16(p, sr_rand16(seed, offset, (uint64_t)(t.rng_base + i)));
            else if (pdt == OTB_BF16) p = round_bf16(p);
        }
        p = -u + p;                                                          // add_stochastic_(p, -update)
        if (sr) p = stochastic_bf16(p, sr_rand16(seed, offset + (t.weight_decay != 0.0f ? 1 : 0), (uint64_t)(t.rng_base + i)));
        store_from_float(t.p, i, pdt, p);
    }
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
