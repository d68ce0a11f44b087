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

// ---- column passes (input axis): the [N, J] matrix is cut into row slabs (blockIdx.y) so that a 10240 x 1280 weight is
// ~600 blocks instead of 40; every block forms the partial sums of 32 * VEC columns over its slab with VEC-wide accesses
// (8 row lanes, combined in lane order) and parks them in scratch[slab * J + j]; the finishing kernels add the slabs in
// slab order — fixed-order sums, bit-reproducible.
constexpr int DORA_MAX_SLABS = 64;

template <int VEC>
__device__ __forceinline__ void v_vec(const void* w, int w_dtype, const float* ba, float scale, int64_t i, float (&v)[VEC]) {
    if (VEC == 4) {
        const float4 b = *reinterpret_cast<const float4*>(ba + i);
        float wv[4];
        if (w_dtype == OTB_F32) {
            const float4 x = *reinterpret_cast<const float4*>((const float*)w + i);
            wv[0] = x.x; wv[1] = x.y; wv[2] = x.z; wv[3] = x.w;
        } else {
            const uint2 x = *reinterpret_cast<const uint2*>((const uint16_t*)w + i);
            if (w_dtype == OTB_BF16) {
                wv[0] = __uint_as_float(x.x << 16); wv[1] = __uint_as_float(x.x & 0xFFFF0000u);
                wv[2] = __uint_as_float(x.y << 16); wv[3] = __uint_as_float(x.y & 0xFFFF0000u);
            } else {
                const __half2 h0 = *reinterpret_cast<const __half2*>(&x.x), h1 = *reinterpret_cast<const __half2*>(&x.y);
                wv[0] = __low2float(h0); wv[1] = __high2float(h0); wv[2] = __low2float(h1); wv[3] = __high2float(h1);
            }
        }
        v[0] = __fmaf_rn(b.x, scale, wv[0]); v[1 % VEC] = __fmaf_rn(b.y, scale, wv[1]);
        v[2 % VEC] = __fmaf_rn(b.z, scale, wv[2]); v[3 % VEC] = __fmaf_rn(b.w, scale, wv[3]);
    } else {
        v[0] = v_at(w, w_dtype, ba, scale, i);
    }
}

template <int VEC>
__global__ void __launch_bounds__(DT)
dora_colsq_kernel(const void* w, int w_dtype, const float* ba, float scale, int N, int J, int slab_rows, float* partial) {
    __shared__ float part[8][32 * VEC + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = (blockIdx.x * 32 + tx) * VEC;
    const int n_lo = blockIdx.y * slab_rows, n_hi = min(N, n_lo + slab_rows);
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.0f;
    if (j < J)
        for (int n = n_lo + ty; n < n_hi; n += 8) {
            float v[VEC];
            v_vec<VEC>(w, w_dtype, ba, scale, (int64_t)n * J + j, v);
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = __fmaf_rn(v[k], v[k], acc[k]);
        }
#pragma unroll
    for (int k = 0; k < VEC; ++k) part[ty][tx * VEC + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < 32 * VEC; c += DT) {
        const int jj = blockIdx.x * 32 * VEC + c;
        if (jj < J) {
            float t = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += part[k][c];
            partial[(int64_t)blockIdx.y * J + jj] = t;
        }
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
// (`slabs` partial tables of `stride` entries each are added in slab order first)
__global__ void dora_norm_kernel(const float* sq, int groups, int gsz, int slabs, int stride, float eps, float* norm) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    float t = 0.0f;
    for (int k = 0; k < gsz; ++k) {
        float c = 0.0f;
        for (int s = 0; s < slabs; ++s) c += sq[(int64_t)s * stride + (int64_t)g * gsz + k];
        t += c;
    }
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

// dV = G * m / norm ; partial sums of G * V / norm per column (input axis; row slabs as above) or per row (output axis)
template <int VEC>
__global__ void __launch_bounds__(DT)
dora_backward_col_kernel(const uint16_t* G, const void* w, int w_dtype, const float* ba, float scale, const float* m, const float* norm,
                         int N, int J, int gsz, int slab_rows, float* dV, float* partial) {
    __shared__ float part[8][32 * VEC + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = (blockIdx.x * 32 + tx) * VEC;
    const int n_lo = blockIdx.y * slab_rows, n_hi = min(N, n_lo + slab_rows);
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.0f;
    if (j < J) {
        float mg[VEC], ng[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { const int g = (j + k) / gsz; mg[k] = m[g]; ng[k] = norm[g]; }
        for (int n = n_lo + ty; n < n_hi; n += 8) {
            const int64_t i = (int64_t)n * J + j;
            float v[VEC], gr[VEC], dv[VEC];
            v_vec<VEC>(w, w_dtype, ba, scale, i, v);
            if (VEC == 4) {
                const uint2 x = *reinterpret_cast<const uint2*>(G + i);
                gr[0] = __uint_as_float(x.x << 16); gr[1 % VEC] = __uint_as_float(x.x & 0xFFFF0000u);
                gr[2 % VEC] = __uint_as_float(x.y << 16); gr[3 % VEC] = __uint_as_float(x.y & 0xFFFF0000u);
            } else {
                gr[0] = bf16_bits_to_float(G[i]);
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                dv[k] = __fdiv_rn(__fmul_rn(gr[k], mg[k]), ng[k]);
                acc[k] = __fmaf_rn(gr[k], __fdiv_rn(v[k], ng[k]), acc[k]);
            }
            if (VEC == 4) *reinterpret_cast<float4*>(dV + i) = make_float4(dv[0], dv[1 % VEC], dv[2 % VEC], dv[3 % VEC]);
            else dV[i] = dv[0];
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) part[ty][tx * VEC + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < 32 * VEC; c += DT) {
        const int jj = blockIdx.x * 32 * VEC + c;
        if (jj < J) {
            float t = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += part[k][c];
            partial[(int64_t)blockIdx.y * J + jj] = t;
        }
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
__global__ void dora_group_sum_kernel(const float* colsum, int groups, int gsz, int slabs, int stride, float* dm) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    float t = 0.0f;
    for (int k = 0; k < gsz; ++k) {
        float c = 0.0f;
        for (int s = 0; s < slabs; ++s) c += colsum[(int64_t)s * stride + (int64_t)g * gsz + k];
        t += c;
    }
    dm[g] = t;
}

// row slabs of a column pass: enough blocks to fill the chip (~4 per SM), at least 8 rows each, at most DORA_MAX_SLABS
struct ColPlan { int vec, blocks_x, slabs, slab_rows; };
ColPlan col_plan(const otb_dora_args* a, bool with_G) {
    ColPlan p;
    const uintptr_t al = (uintptr_t)a->w | (uintptr_t)a->ba | (with_G ? ((uintptr_t)a->G | (uintptr_t)a->dV) : 0);
    const bool w_ok = a->w_dtype == OTB_F32 ? (al & 15) == 0 : (((uintptr_t)a->ba | (with_G ? (uintptr_t)a->dV : 0)) & 15) == 0 && (((uintptr_t)a->w | (with_G ? (uintptr_t)a->G : 0)) & 7) == 0;
    p.vec = (a->J % 4 == 0 && w_ok) ? 4 : 1;
    p.blocks_x = (a->J + 32 * p.vec - 1) / (32 * p.vec);
    int want = (4 * otb_sm_count() + p.blocks_x - 1) / p.blocks_x;
    const int max_by_rows = (a->N + 7) / 8;
    if (want > max_by_rows) want = max_by_rows;
    if (want > DORA_MAX_SLABS) want = DORA_MAX_SLABS;
    if (want < 1) want = 1;
    p.slab_rows = (a->N + want - 1) / want;
    p.slabs = (a->N + p.slab_rows - 1) / p.slab_rows;
    return p;
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
        dora_norm_kernel<<<(groups + 255) / 256, 256, 0, st>>>(a->scratch, groups, 1, 1, 0, a->eps, a->norm);
    } else {
        const ColPlan p = col_plan(a, false);
        const dim3 g((unsigned)p.blocks_x, (unsigned)p.slabs);
        if (p.vec == 4) dora_colsq_kernel<4><<<g, DT, 0, st>>>(a->w, a->w_dtype, a->ba, a->scale, a->N, a->J, p.slab_rows, a->scratch);
        else dora_colsq_kernel<1><<<g, DT, 0, st>>>(a->w, a->w_dtype, a->ba, a->scale, a->N, a->J, p.slab_rows, a->scratch);
        dora_norm_kernel<<<(groups + 255) / 256, 256, 0, st>>>(a->scratch, groups, a->group_size, p.slabs, a->J, a->eps, a->norm);
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
        const ColPlan p = col_plan(a, true);
        const dim3 g((unsigned)p.blocks_x, (unsigned)p.slabs);
        if (p.vec == 4)
            dora_backward_col_kernel<4><<<g, DT, 0, st>>>((const uint16_t*)a->G, a->w, a->w_dtype, a->ba, a->scale, a->dora_scale, a->norm,
                                                          a->N, a->J, a->group_size, p.slab_rows, a->dV, a->scratch);
        else
            dora_backward_col_kernel<1><<<g, DT, 0, st>>>((const uint16_t*)a->G, a->w, a->w_dtype, a->ba, a->scale, a->dora_scale, a->norm,
                                                          a->N, a->J, a->group_size, p.slab_rows, a->dV, a->scratch);
        dora_group_sum_kernel<<<(groups + 255) / 256, 256, 0, st>>>(a->scratch, groups, a->group_size, p.slabs, a->J, a->d_scale);
    }
    OTB_LAUNCH_CHECK("dora_backward kernels");
    return OTB_OK;
}
