// dora.cu — weight-space kernels of the DoRA adapter.
//
// Reference: DoRAModule.forward  /root/reference/modules/module/LoRAModule.py:385-419
//     V    = W + (B A) * (alpha / rank)                       (make_weight :76-91)
//     norm = || V ||  per INPUT column group (default) or per OUTPUT row (decompose_output_axis), detached, (+ eps)
//     WP   = dora_scale * (V / norm)
//     y    = op(dropout(x), WP, bias)
// The activation-sized work — y = x WP^T and dX = dY WP — runs on K-a (otb_lora_linear with rank 0); the rank-r products
// B A, dV A^T, B^T dV and the weight gradient dY^T x are plain library GEMMs.  What is left is weight-sized, memory-bound
// and fused here: three passes over [N, J] instead of the reference's ~12 ATen launches and five full-size fp32 temporaries:
//   otb_dora_compose   pass 1: squared norms of V per group (fixed-order sums);  pass 2: WP and WP^T in bf16 (the K-major
//                      operands of the two K-a launches), V formed on the fly from W and B A, the norm kept for backward
//   otb_dora_backward  dV = G * dora_scale / norm   and   d(dora_scale) = sum_group G * V / norm     (the norm is a
//                      constant for backpropagation, :396-400), G = dL/dWP as the bf16 GEMM produced it
// A "group" is one input channel: column j of the [N, J] weight matrix belongs to group j / gsz with gsz = kh * kw (1 for
// Linear and 1x1 conv), because the reference reshapes a conv weight to [Cin, N * kh * kw] before taking the norm.
#include "common.cuh"

namespace {

constexpr int DT = 256;

__device__ __forceinline__ float v_at(const void* w, int w_dtype, const float* ba, float scale, int64_t i) {
    return __fmaf_rn(ba[i], scale, load_as_float(w, i, w_dtype));      // orig_weight + make_weight(A, B) * (alpha / rank)
}

// input-axis: per-column sums over the rows, 32 columns per block, 8 row lanes, combined in lane order
__global__ void __launch_bounds__(DT)
dora_colsq_kernel(const void* w, int w_dtype, const float* ba, float scale, int N, int J, float* colsq) {
    __shared__ float part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    float acc = 0.0f;
    if (j < J)
        for (int n = ty; n < N; n += 8) { const float v = v_at(w, w_dtype, ba, scale, (int64_t)n * J + j); acc = __fmaf_rn(v, v, acc); }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && j < J) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][tx];
        colsq[j] = t;
    }
}
// output-axis: one warp per row
__global__ void __launch_bounds__(DT)
dora_rowsq_kernel(const void* w, int w_dtype, const float* ba, float scale, int N, int J, float* rowsq) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * (DT / 32) + warp;
    if (n >= N) return;
    float acc = 0.0f;
    for (int j = lane; j < J; j += 32) { const float v = v_at(w, w_dtype, ba, scale, (int64_t)n * J + j); acc = __fmaf_rn(v, v, acc); }
    acc = warp_sum(acc);
    if (lane == 0) rowsq[n] = acc;
}
// squared sums -> norm (+ eps): groups of gsz adjacent entries (input axis) or one entry per row (output axis)
__global__ void dora_norm_kernel(const float* sq, int groups, int gsz, float eps, float* norm) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    float t = 0.0f;
    for (int k = 0; k < gsz; ++k) t += sq[(int64_t)g * gsz + k];
    norm[g] = sqrtf(t) + eps;
}

// WP = dora_scale * (V / norm) -> bf16, row-major [N, J] and transposed [J, N] (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(DT)
dora_compose_kernel(const void* w, int w_dtype, const float* ba, float scale, const float* m, const float* norm, int N, int J, int gsz,
                    int out_axis, uint16_t* wp, uint16_t* wpt) {
    __shared__ uint16_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, j = j0 + tx;
        uint16_t b = 0;
        if (n < N && j < J) {
            const int g = out_axis ? n : j / gsz;
            const float v = v_at(w, w_dtype, ba, scale, (int64_t)n * J + j);
            b = float_to_bf16_bits_rn(__fmul_rn(m[g], __fdiv_rn(v, norm[g])));
            wp[(int64_t)n * J + j] = b;
        }
        tile[r][tx] = b;
    }
    __syncthreads();
    if (wpt != nullptr)
        for (int r = ty; r < 32; r += 8) {
            const int j = j0 + r, n = n0 + tx;
            if (n < N && j < J) wpt[(int64_t)j * N + n] = tile[tx][r];
        }
}

// dV = G * m / norm ; partial sums of G * V / norm per column (input axis) or per row (output axis)
__global__ void __launch_bounds__(DT)
dora_backward_col_kernel(const uint16_t* G, const void* w, int w_dtype, const float* ba, float scale, const float* m, const float* norm,
                         int N, int J, int gsz, float* dV, float* colsum) {
    __shared__ float part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    float acc = 0.0f;
    if (j < J) {
        const int g = j / gsz;
        const float mg = m[g], ng = norm[g];
        for (int n = ty; n < N; n += 8) {
            const int64_t i = (int64_t)n * J + j;
            const float gr = bf16_bits_to_float(G[i]);
            const float v = v_at(w, w_dtype, ba, scale, i);
            dV[i] = __fdiv_rn(__fmul_rn(gr, mg), ng);
            acc = __fmaf_rn(gr, __fdiv_rn(v, ng), acc);
        }
    }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && j < J) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][tx];
        colsum[j] = t;
    }
}
__global__ void __launch_bounds__(DT)
dora_backward_row_kernel(const uint16_t* G, const void* w, int w_dtype, const float* ba, float scale, const float* m, const float* norm,
                         int N, int J, float* dV, float* dm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * (DT / 32) + warp;
    if (n >= N) return;
    const float mg = m[n], ng = norm[n];
    float acc = 0.0f;
    for (int j = lane; j < J; j += 32) {
        const int64_t i = (int64_t)n * J + j;
        const float gr = bf16_bits_to_float(G[i]);
        const float v = v_at(w, w_dtype, ba, scale, i);
        dV[i] = __fdiv_rn(__fmul_rn(gr, mg), ng);
        acc = __fmaf_rn(gr, __fdiv_rn(v, ng), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) dm[n] = acc;
}
__global__ void dora_group_sum_kernel(const float* colsum, int groups, int gsz, float* dm) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    float t = 0.0f;
    for (int k = 0; k < gsz; ++k) t += colsum[(int64_t)g * gsz + k];
    dm[g] = t;
}

int check(const otb_dora_args* a) {
    OTB_CHECK_ARG(a != nullptr && a->w && a->ba && a->dora_scale && a->norm && a->scratch, "null DoRA argument");
    OTB_CHECK_ARG(a->N > 0 && a->J > 0 && a->group_size > 0 && a->J % a->group_size == 0, "bad DoRA shape N=%d J=%d group=%d", a->N, a->J, a->group_size);
    OTB_CHECK_ARG(a->w_dtype >= OTB_F32 && a->w_dtype <= OTB_F16, "bad base weight dtype");
    return OTB_OK;
}

}  // namespace

extern "C" int otb_dora_compose(const otb_dora_args* a, void* stream_) {
    int rc = check(a);
    if (rc != OTB_OK) return rc;
    OTB_CHECK_ARG(a->wp != nullptr, "otb_dora_compose: wp is required");
    OTB_BIND(a->w);
    cudaStream_t st = (cudaStream_t)stream_;
    const int groups = a->out_axis ? a->N : a->J / a->group_size;
    if (a->out_axis) {
        dora_rowsq_kernel<<<(a->N + DT / 32 - 1) / (DT / 32), DT, 0, st>>>(a->w, a->w_dtype, a->ba, a->scale, a->N, a->J, a->scratch);
        dora_norm_kernel<<<(groups + 255) / 256, 256, 0, st>>>(a->scratch, groups, 1, a->eps, a->norm);
    } else {
        dora_colsq_kernel<<<(a->J + 31) / 32, DT, 0, st>>>(a->w, a->w_dtype, a->ba, a->scale, a->N, a->J, a->scratch);
        dora_norm_kernel<<<(groups + 255) / 256, 256, 0, st>>>(a->scratch, groups, a->group_size, a->eps, a->norm);
    }
    dim3 grid((a->J + 31) / 32, (a->N + 31) / 32);
    dora_compose_kernel<<<grid, DT, 0, st>>>(a->w, a->w_dtype, a->ba, a->scale, a->dora_scale, a->norm, a->N, a->J, a->group_size, a->out_axis,
                                             (uint16_t*)a->wp, (uint16_t*)a->wpt);
    OTB_LAUNCH_CHECK("dora_compose_kernel");
    return OTB_OK;
}

extern "C" int otb_dora_backward(const otb_dora_args* a, void* stream_) {
    int rc = check(a);
    if (rc != OTB_OK) return rc;
    OTB_CHECK_ARG(a->G && a->dV && a->d_scale, "otb_dora_backward: G, dV and d_scale are required");
    OTB_BIND(a->w);
    cudaStream_t st = (cudaStream_t)stream_;
    if (a->out_axis) {
        dora_backward_row_kernel<<<(a->N + DT / 32 - 1) / (DT / 32), DT, 0, st>>>((const uint16_t*)a->G, a->w, a->w_dtype, a->ba, a->scale, a->dora_scale,
                                                                                 a->norm, a->N, a->J, a->dV, a->d_scale);
    } else {
        const int groups = a->J / a->group_size;
        dora_backward_col_kernel<<<(a->J + 31) / 32, DT, 0, st>>>((const uint16_t*)a->G, a->w, a->w_dtype, a->ba, a->scale, a->dora_scale, a->norm,
                                                                  a->N, a->J, a->group_size, a->dV, a->scratch);
        dora_group_sum_kernel<<<(groups + 255) / 256, 256, 0, st>>>(a->scratch, groups, a->group_size, a->d_scale);
    }
    OTB_LAUNCH_CHECK("dora_backward kernels");
    return OTB_OK;
}
