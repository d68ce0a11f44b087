// lora_grad.cu — the two skinny weight-gradient contractions of the LoRA backward:
//
//   dA[R,K] += Gs[M,R]^T . x[M,K]          dB[N,R] += dy[M,N]^T . Ts[M,R]
//
// (dA = s (dY B)^T X, dB = s dY^T (X A^T); Gs and Ts are the bf16,
// already-scaled rank-R activations the fused forward / dX kernels park in HBM.)
// Both are "out[R,Q] = U[M,R]^T . V[M,Q]" with a contraction over the LONG dimension M and
// a rank-R (16..64) output: arithmetic intensity is 2R flop per 2 bytes of V, i.e. they are
// bound by streaming V once (HBM/L2), not by the tensor pipe — R=32 needs 32 flop/B x 6.5 TB/s
// = 0.2 PFLOP/s.  So they use warp-level mma.sync (HMMA) fed by cp.async + ldmatrix.trans,
// split M across CTAs to fill the chip, and reduce the partial sums with fp32 red.global.add.
// One launch computes both gradients (grid.z selects the problem).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace {

constexpr int GQ = 128;        // output columns (of V) per CTA
constexpr int GM = 64;         // rows of M per pipeline stage
constexpr int G_STAGES = 5;     // 4 steps (88 KB) of loads in flight per CTA: the kernel is latency-, not bandwidth-bound
constexpr int G_THREADS = 256; // 8 warps x 16 columns
constexpr int V_PITCH = GQ * 2 + 16;   // bytes; +16 keeps ldmatrix rows on distinct bank groups

struct GradProblem {
    const uint16_t* U;   // [M, R]
    const uint16_t* V;   // [M, Q]
    float* out;          // transposed_out ? [Q, R] : [R, Q]
    int Q;
    int q_tiles;
    int transposed_out;
};
struct GradParams {
    GradProblem pr[2];
    int M, R, Rp;        // Rp: rank padded to 16
    int m_per_cta;       // multiple of GM
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;     // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// one 16-byte fp32 vector reduction (sm_90+): four atomic adds, one lane-level operation
__device__ __forceinline__ void red_add_v4(float* addr, const float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// grid = (max q_tiles, m_splits, 2)
__global__ void __launch_bounds__(G_THREADS) lora_grad_kernel(const GradParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    otb_pdl_trigger();
    otb_pdl_wait();          // Gs / Ts come from the dX / forward launches just before this one
    const GradProblem pr = p.pr[blockIdx.z];
    if ((int)blockIdx.x >= pr.q_tiles || pr.out == nullptr) return;
    const int q0 = blockIdx.x * GQ;
    const int m_begin = blockIdx.y * p.m_per_cta;
    int m_end = m_begin + p.m_per_cta;
    if (m_end > p.M) m_end = p.M;
    if (m_begin >= m_end) return;
    const int n_steps = (m_end - m_begin + GM - 1) / GM;

    const int u_pitch = p.Rp * 2 + 16;
    const uint32_t v_stage = GM * V_PITCH, u_stage = GM * u_pitch;
    const uint32_t smem_v = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t smem_u = smem_v + G_STAGES * v_stage;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    auto issue = [&](int step) {
        const int st = step % G_STAGES;
        const int mrow0 = m_begin + step * GM;
        // V tile: GM rows x 16 chunks of 16 B
        for (int c = threadIdx.x; c < GM * (GQ / 8); c += G_THREADS) {
            const int r = c >> 4, ch = c & 15;
            const int m = mrow0 + r, q = q0 + ch * 8;
            const bool ok = (m < m_end) && (q < pr.Q);
            const uint16_t* src = ok ? pr.V + (size_t)m * pr.Q + q : pr.V;
            cp_async16(smem_v + st * v_stage + r * V_PITCH + ch * 16, src, ok);
        }
        // U tile: GM rows x Rp/8 chunks
        const int uch = p.Rp >> 3;
        for (int c = threadIdx.x; c < GM * uch; c += G_THREADS) {
            const int r = c / uch, ch = c % uch;
            const int m = mrow0 + r;
            const bool ok = (m < m_end) && (ch * 8 < p.R);
            const uint16_t* src = ok ? pr.U + (size_t)m * p.R + ch * 8 : pr.U;
            cp_async16(smem_u + st * u_stage + r * u_pitch + ch * 16, src, ok);
        }
    };

    float acc[4][2][4];     // [r tile of 16][n tile of 8][frag]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;

    for (int s = 0; s < G_STAGES - 1; ++s) {
        if (s < n_steps) issue(s);
        cp_async_commit();
    }
    const int r_tiles = p.Rp >> 4;
    for (int step = 0; step < n_steps; ++step) {
        cp_async_wait<G_STAGES - 2>();
        __syncthreads();
        if (step + G_STAGES - 1 < n_steps) issue(step + G_STAGES - 1);
        cp_async_commit();
        const int st = step % G_STAGES;
        const uint32_t vb = smem_v + st * v_stage, ub = smem_u + st * u_stage;
#pragma unroll
        for (int kk = 0; kk < GM / 16; ++kk) {
            // B fragments for this warp's 16 columns: smem is [k][n] (n contiguous) -> .trans
            // matrices: (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15)
            uint32_t b[4];
            {
                const int krow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int ncol = warp * 16 + (lane >> 4) * 8;
                ldsm_x4_trans(vb + krow * V_PITCH + ncol * 2, b[0], b[1], b[2], b[3]);
            }
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) {
                if (rt < r_tiles) {
                    // A = U^T: smem is [k][r] (r contiguous) -> .trans
                    // matrices: (r 0-7, k 0-7), (r 8-15, k 0-7), (r 0-7, k 8-15), (r 8-15, k 8-15)
                    uint32_t a[4];
                    const int krow = kk * 16 + (lane & 7) + (lane >> 4) * 8;
                    const int rcol = rt * 16 + ((lane >> 3) & 1) * 8;
                    ldsm_x4_trans(ub + krow * u_pitch + rcol * 2, a[0], a[1], a[2], a[3]);
                    mma_bf16_16816(acc[rt][0], a, b[0], b[1]);
                    mma_bf16_16816(acc[rt][1], a, b[2], b[3]);
                }
            }
        }
    }
    // partial sums -> global.  The fragments are staged through shared memory (the pipeline buffers are free now) so
    // that every lane issues ONE 16-byte vector reduction on consecutive addresses — a quarter of the lane-level
    // atomic operations of a per-element red, and full 128-byte lines per warp instruction.
    cp_async_wait<0>();
    __syncthreads();
    float* stage = reinterpret_cast<float*>(smem);
    const int g = lane >> 2, t4 = lane & 3;
    // dA: out[R, Q] -> stage[r][q] (pitch GQ + 4);  dB: out[Q, R] -> stage[q][r] (pitch Rp + 4)
    const int pitch = pr.transposed_out ? p.Rp + 4 : GQ + 4;
#pragma unroll
    for (int rt = 0; rt < 4; ++rt) {
        if (rt >= r_tiles) continue;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = rt * 16 + g + (k >> 1) * 8;
                const int ql = warp * 16 + nt * 8 + t4 * 2 + (k & 1);
                stage[pr.transposed_out ? ql * pitch + r : r * pitch + ql] = acc[rt][nt][k];
            }
        }
    }
    __syncthreads();
    if (!pr.transposed_out) {
        // rows of R, 128 columns each: one warp instruction covers one full row (32 lanes x 4 floats)
        for (int r = warp; r < p.R; r += G_THREADS / 32) {
            const int q = q0 + lane * 4;
            if (q < pr.Q) {
                const float4 v = *reinterpret_cast<const float4*>(stage + r * pitch + lane * 4);
                red_add_v4(pr.out + (size_t)r * pr.Q + q, v);
            }
        }
    } else {
        // 128 rows of R floats, contiguous in memory as a whole when the tile is interior
        const int vec_per_row = p.R >> 2;
        for (int i = threadIdx.x; i < GQ * vec_per_row; i += G_THREADS) {
            const int ql = i / vec_per_row, r4 = (i % vec_per_row) * 4;
            const int q = q0 + ql;
            if (q < pr.Q) {
                const float4 v = *reinterpret_cast<const float4*>(stage + ql * pitch + r4);
                red_add_v4(pr.out + (size_t)q * p.R + r4, v);
            }
        }
    }
}

// =====================================================================================
// tcgen05 version: the same contraction on the 5th-gen tensor cores, operands straight from TMA
// =====================================================================================
// ncu on the mma.sync version (profiles/r01b_ncu_kgrad_v2.md): DRAM 20 %, HMMA pipe 26 %, issue 40 %, stalls on
// `wait` / short scoreboard — it is bound by its own cp.async address generation, ldmatrix and HMMA issue, not by
// memory.  Here one elected lane issues TMA boxes and UMMAs and nothing else touches the operands:
//     D[q (128 TMEM lanes), r (64 columns)] += sum_m V[m, q] * U[m, r]
// Both operands are "MN-major" for the tensor core — the M_mma dimension (q, columns of V) and the N_mma dimension
// (r, columns of U) are the contiguous ones in memory, the contraction runs over the ROWS m — which UMMA reads
// directly from the 128B-swizzled TMA boxes through MN-major shared-memory descriptors (idesc a_major = b_major = 1):
// no transposition anywhere.  U is fetched with a 64-column box; columns beyond R are zero-filled by TMA, so one
// descriptor layout serves every rank <= 64 and the MMA's N is always 64 (the tensor pipe is idle anyway).
constexpr int TG_ROWS = 128;          // rows of M per pipeline stage (8 UMMAs of K = 16)
constexpr int TG_STAGES = 4;
constexpr int TG_THREADS = 192;       // warp 0 TMA producer, warp 1 TMEM + MMA issuer, warps 2-5 epilogue
constexpr uint32_t TG_VHALF = TG_ROWS * 128;              // one 64-column box of V
constexpr uint32_t TG_STAGE_BYTES = 3 * TG_VHALF;         // V (two boxes = 128 columns) + U
constexpr int TG_SMEM = TG_STAGES * (int)TG_STAGE_BYTES + 256 + 1024;

struct TcGradProblem {
    float* out;          // transposed_out ? [Q, R] : [R, Q]
    int Q, q_tiles, transposed_out;
};
struct TcGradParams {
    TcGradProblem pr[2];
    int M, R, m_per_cta;     // m_per_cta: multiple of TG_ROWS
};

// MN-major, 128B-swizzled operand: 8 rows (K) x 128 B atoms, 8-row groups `1024` B apart along K (SBO), 64-element
// blocks along MN `lbo` bytes apart (cute::UMMA::make_umma_desc<Major::MN>, LayoutType::B128)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(TG_THREADS, 1)
lora_grad_tc_kernel(const __grid_constant__ CUtensorMap tm_v0, const __grid_constant__ CUtensorMap tm_u0,
                    const __grid_constant__ CUtensorMap tm_v1, const __grid_constant__ CUtensorMap tm_u1,
                    const TcGradParams p) {
    otb_pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prob = (int)blockIdx.x >= p.pr[0].q_tiles ? 1 : 0;
    const int q_tile = (int)blockIdx.x - (prob ? p.pr[0].q_tiles : 0);
    float* out = prob ? p.pr[1].out : p.pr[0].out;
    const int Q = prob ? p.pr[1].Q : p.pr[0].Q;
    const int transposed_out = prob ? p.pr[1].transposed_out : p.pr[0].transposed_out;
    const CUtensorMap* tm_v = prob ? &tm_v1 : &tm_v0;
    const CUtensorMap* tm_u = prob ? &tm_u1 : &tm_u0;
    const int q0 = q_tile * 128;
    const int m_begin = blockIdx.y * p.m_per_cta;
    int m_end = m_begin + p.m_per_cta;
    if (m_end > p.M) m_end = p.M;
    const int n_stages = m_end > m_begin ? (m_end - m_begin + TG_ROWS - 1) / TG_ROWS : 0;   // >= 1 by construction of the grid

    const uint32_t bar_base = smem_base + TG_STAGES * TG_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TG_STAGES + s); };
    const uint32_t d_full = bar_base + 8u * (2 * TG_STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * TG_STAGES + 1);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + TG_STAGES * TG_STAGE_BYTES + 8u * (2 * TG_STAGES + 1));

    if (threadIdx.x == 0) {
        prefetch_tmap(tm_v); prefetch_tmap(tm_u);
        for (int s = 0; s < TG_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(d_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 64u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    otb_pdl_wait();          // Gs / Ts come from the dX / forward launches before this one

    if (warp == 0) {
        int s = 0; uint32_t ph = 0;
        for (int st = 0; st < n_stages; ++st) {
            mbar_wait(empty_bar(s), ph ^ 1);
            if (elect_one()) {
                const uint32_t stg = smem_base + TG_STAGE_BYTES * s;
                const int m = m_begin + st * TG_ROWS;
                mbar_expect_tx(full_bar(s), TG_STAGE_BYTES);
                tma_load_2d(stg, tm_v, full_bar(s), q0, m);
                tma_load_2d(stg + TG_VHALF, tm_v, full_bar(s), q0 + 64, m);
                tma_load_2d(stg + 2 * TG_VHALF, tm_u, full_bar(s), 0, m);
            }
            __syncwarp();
            if (++s == TG_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // c_format f32 [4,6) = 1, a/b format bf16 [7,10)/[10,13) = 1, a_major / b_major MN-major [15]/[16] = 1, N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int s = 0; uint32_t ph = 0;
        for (int st = 0; st < n_stages; ++st) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t stg = smem_base + TG_STAGE_BYTES * s;
                const uint64_t adesc = umma_desc_mn_sw128(stg, TG_VHALF);
                const uint64_t bdesc = umma_desc_mn_sw128(stg + 2 * TG_VHALF, TG_VHALF);
#pragma unroll
                for (int k = 0; k < TG_ROWS / 16; ++k)      // 16 rows = two 8-row swizzle atoms = 2048 B per k-step
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (st | k) != 0);
                umma_commit(empty_bar(s));
                if (st == n_stages - 1) umma_commit(d_full);
            }
            __syncwarp();
            if (++s == TG_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (n_stages > 0) {
        const int quad = warp & 3;
        const int q = q0 + quad * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        mbar_wait(d_full, 0);
        tc_fence_after();
        for (int c = 0; c < p.R; c += 16) {
            uint32_t v[16];
            tmem_ld16(t_lane + (uint32_t)c, v);
            tmem_ld_wait();
            if (q < Q) {
                if (transposed_out) {       // out[q][r]: this lane's 16 values are contiguous
                    float* dst = out + (size_t)q * p.R + c;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        if (c + j < p.R)
                            red_add_v4(dst + j, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
                } else {                    // out[r][q]: consecutive lanes hit consecutive addresses
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c + j < p.R) atomicAdd(out + (size_t)(c + j) * Q + q, __uint_as_float(v[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 64u);
    }
}

int launch_grad_tc(const otb_lora_grad_args* a, void* stream) {
    const bool doA = a->dA != nullptr, doB = a->dB != nullptr;
    TcGradParams p;
    p.M = a->M; p.R = a->R;
    p.pr[0].out = doA ? a->dA : nullptr; p.pr[0].Q = a->K; p.pr[0].q_tiles = doA ? (a->K + 127) / 128 : 0; p.pr[0].transposed_out = 0;
    p.pr[1].out = doB ? a->dB : nullptr; p.pr[1].Q = a->N; p.pr[1].q_tiles = doB ? (a->N + 127) / 128 : 0; p.pr[1].transposed_out = 1;
    const int q_tiles_total = p.pr[0].q_tiles + p.pr[1].q_tiles;
    // one CTA per SM (192 KB of pipeline buffers): split M so that the grid is one wave
    const int sms = otb_sm_count();
    int splits = sms / q_tiles_total;
    const int max_splits = (a->M + TG_ROWS - 1) / TG_ROWS;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int m_per = (a->M + splits - 1) / splits;
    m_per = (m_per + TG_ROWS - 1) / TG_ROWS * TG_ROWS;
    splits = (a->M + m_per - 1) / m_per;
    p.m_per_cta = m_per;
    CUtensorMap tm_v0, tm_u0, tm_v1, tm_u1;
    int rc;
    if (doA) {
        if ((rc = make_tmap(&tm_v0, a->x, a->M, a->K, TG_ROWS, "x")) != OTB_OK) return rc;
        if ((rc = make_tmap(&tm_u0, a->gs, a->M, a->R, TG_ROWS, "gs")) != OTB_OK) return rc;
    }
    if (doB) {
        if ((rc = make_tmap(&tm_v1, a->dy, a->M, a->N, TG_ROWS, "dy")) != OTB_OK) return rc;
        if ((rc = make_tmap(&tm_u1, a->ts, a->M, a->R, TG_ROWS, "ts")) != OTB_OK) return rc;
    }
    if (!doA) { tm_v0 = tm_v1; tm_u0 = tm_u1; }
    if (!doB) { tm_v1 = tm_v0; tm_u1 = tm_u0; }
    static thread_local int attr_dev = -1;
    int dev = 0;
    OTB_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        OTB_CUDA(cudaFuncSetAttribute(lora_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
        attr_dev = dev;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)q_tiles_total, (unsigned)splits, 1);
    cfg.blockDim = dim3(TG_THREADS);
    cfg.dynamicSmemBytes = TG_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = otb_pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, lora_grad_tc_kernel, tm_v0, tm_u0, tm_v1, tm_u1, p);
    if (e != cudaSuccess) {
        otb_set_error("launch of lora_grad_tc_kernel failed: %s", cudaGetErrorString(e));
        return OTB_ECUDA;
    }
    return OTB_OK;
}

}  // namespace

extern "C" int otb_lora_grad_ab(const otb_lora_grad_args* a, void* stream) {
    OTB_CHECK_ARG(a != nullptr, "null args");
    OTB_CHECK_ARG(a->M > 0 && a->R > 0 && a->R <= 64 && a->R % 8 == 0, "bad M=%d R=%d (R multiple of 8, <= 64)", a->M, a->R);
    const bool doA = a->dA != nullptr, doB = a->dB != nullptr;
    if (!doA && !doB) return OTB_OK;
    if (doA) OTB_CHECK_ARG(a->gs && a->x && a->K > 0 && a->K % 8 == 0, "dA needs gs, x and K %% 8 == 0");
    if (doB) OTB_CHECK_ARG(a->ts && a->dy && a->N > 0 && a->N % 8 == 0, "dB needs ts, dy and N %% 8 == 0");
    OTB_CHECK_ARG((((uintptr_t)a->x | (uintptr_t)a->dy | (uintptr_t)a->gs | (uintptr_t)a->ts) & 15) == 0,
                  "all operand pointers must be 16-byte aligned");
    OTB_BIND(doA ? a->x : a->dy);
    {   // tcgen05 path (default); OTB_KGRAD_TC=0 selects the mma.sync kernel below (A/B runs)
        const char* v = getenv("OTB_KGRAD_TC");
        if (!(v && v[0] == '0')) return launch_grad_tc(a, stream);
    }
    GradParams p;
    p.M = a->M; p.R = a->R; p.Rp = (a->R + 15) / 16 * 16;
    p.pr[0].U = (const uint16_t*)a->gs; p.pr[0].V = (const uint16_t*)a->x; p.pr[0].out = doA ? a->dA : nullptr;
    p.pr[0].Q = a->K; p.pr[0].q_tiles = doA ? (a->K + GQ - 1) / GQ : 0; p.pr[0].transposed_out = 0;
    p.pr[1].U = (const uint16_t*)a->ts; p.pr[1].V = (const uint16_t*)a->dy; p.pr[1].out = doB ? a->dB : nullptr;
    p.pr[1].Q = a->N; p.pr[1].q_tiles = doB ? (a->N + GQ - 1) / GQ : 0; p.pr[1].transposed_out = 1;
    const int q_tiles_total = p.pr[0].q_tiles + p.pr[1].q_tiles;
    const int q_tiles_max = p.pr[0].q_tiles > p.pr[1].q_tiles ? p.pr[0].q_tiles : p.pr[1].q_tiles;
    // split M so that the launch is ONE wave of co-resident CTAs (2 per SM with 110 KB of pipeline buffers each),
    // but keep >= GM rows per CTA
    const int sms = otb_sm_count();
    int splits = (2 * sms) / q_tiles_total;
    const int max_splits = (a->M + GM - 1) / GM;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int m_per = (a->M + splits - 1) / splits;
    m_per = (m_per + GM - 1) / GM * GM;
    splits = (a->M + m_per - 1) / m_per;
    p.m_per_cta = m_per;
    const int u_pitch = p.Rp * 2 + 16;
    const size_t smem = (size_t)G_STAGES * GM * (V_PITCH + u_pitch);
    static thread_local int attr_dev = -1;
    int dev = 0;
    OTB_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        OTB_CUDA(cudaFuncSetAttribute(lora_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_dev = dev;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)q_tiles_max, (unsigned)splits, 2);
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = otb_pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, lora_grad_kernel, p);
    if (e != cudaSuccess) {
        otb_set_error("launch of lora_grad_kernel failed: %s", cudaGetErrorString(e));
        return OTB_ECUDA;
    }
    return OTB_OK;
}
