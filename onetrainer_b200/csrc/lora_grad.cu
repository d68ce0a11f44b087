// lora_grad.cu — K-a': the two skinny weight-gradient contractions of the LoRA backward, batched over leaves.
//
//   dA[R,K] += Gs[M,R]^T . x[M,K]          dB[N,R] += dy[M,N]^T . Ts[M,R]
//
// (dA = s (dY B)^T X, dB = s dY^T (X A^T); Gs and Ts are the bf16, already-scaled rank-R activations
// the fused forward / dX kernels park in HBM.  In the reference these are two of autograd's ~7 GEMMs per leaf,
// LoRAModule.forward modules/module/LoRAModule.py:319-323 differentiated.)
// Both are "out[R,Q] = U[M,R]^T . V[M,Q]": a contraction over the LONG dimension M with a rank-R (8..64) output —
// 2R flop per 2 bytes of V, i.e. bound by streaming V (x / dY) once from HBM, not by the tensor pipe.
//
// Design (sm_100a):
//   * ONE persistent launch for up to 32 leaves (64 problems).  dA / dB feed nothing until the optimizer, so the host
//     queues the leaves' operands while backward runs and flushes them in batches: 722 launches of a latency-bound
//     10 us kernel per SDXL step become ~23 launches that each keep every SM streaming.
//   * work unit = (problem, 128 output columns, M-split).  With a full batch there are far more column tiles than SMs,
//     so M is NOT split and every output element has exactly one writer.  Only when a batch is small (the single-leaf
//     entry point, the tail of a backward pass) or M is very long is M split — and then the partial tiles go through a
//     workspace and the LAST CTA to arrive sums them in split order.  Either way the result is a fixed-order sum:
//     bit-identical from run to run (the first version reduced split-M partials with red.global.add).
//   * operands straight from TMA into tcgen05: D[q (128 TMEM lanes), r (64 columns)] += sum_m V[m,q] U[m,r].  Both are
//     "MN-major" for the tensor core — the MMA's M and N dimensions are the contiguous ones in memory and the contraction
//     runs over the ROWS — which UMMA reads from the 128B-swizzled TMA boxes through MN-major shared-memory descriptors
//     (idesc a_major = b_major = 1): no transposition anywhere.  U is fetched with a 64-column box; columns beyond R are
//     zero-filled by TMA, so one descriptor layout serves every rank <= 64.
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer (one elected lane each), warps 2-5 epilogue; two TMEM
//     accumulators, so the epilogue of unit i (TMEM -> registers -> fp32 global) overlaps the stream of unit i+1.
//   * units are dealt to CTAs in snake order over a list sorted by rows per unit (longest first): a static schedule
//     (no atomics in the hot loop) whose per-CTA totals differ by at most one of the shortest units.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <algorithm>

int env_int(const char* name, int dflt);      // lora_gemm.cu

namespace {

constexpr int GB_ROWS = 128;          // rows of M per pipeline stage (8 UMMAs of K = 16)
constexpr int GB_STAGES = 4;
constexpr int GB_THREADS = 192;       // warp 0 TMA producer, warp 1 TMEM + MMA issuer, warps 2-5 epilogue
constexpr uint32_t GB_VHALF = GB_ROWS * 128;              // one 64-column box (V half, or U)
constexpr uint32_t GB_STAGE_BYTES = 3 * GB_VHALF;         // V (two boxes = 128 columns) + U
constexpr int GB_SMEM = GB_STAGES * (int)GB_STAGE_BYTES + 256 + 1024;
constexpr int GB_MAX_PROBLEMS = 2 * OTB_GRAD_BATCH_MAX;   // dA and dB of every leaf
constexpr int GB_COUNTER_BYTES = 64 * 1024;               // head of the workspace: one int32 per (problem, column tile)

struct alignas(64) GbProblem {
    CUtensorMap tm_v;        // [M, Q] bf16 (x or dY), box 128 rows x 64 columns
    CUtensorMap tm_u;        // [M, R] bf16 (Gs or Ts), box 128 rows x 64 columns (columns >= R read as zero)
    float* out;              // transposed_out ? [Q, R] : [R, Q]
    float* partial;          // split tiles: [q_tiles][splits][128][R] fp32 (workspace)
    int32_t* counters;       // split tiles: [q_tiles], zero on entry, reset to zero by the last CTA to arrive
    int32_t M, Q, R;
    int32_t q_tiles, splits, rows_per_split;     // rows_per_split: multiple of GB_ROWS
    int32_t unit_begin;      // first unit of this problem in the launch's unit list
    int32_t transposed_out;
};
struct GbBatch {
    GbProblem pr[GB_MAX_PROBLEMS];
    int32_t n_problems, total_units;
    int32_t chained;         // the launch before this one on the stream is a sibling batch of the same flush (see the host entry)
};
static_assert(sizeof(GbBatch) <= 32 * 1024 - 256, "kernel parameters are limited to 32 KB");

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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the four epilogue warps

// unit `idx` of the launch -> (problem, column tile, split)
struct GbUnit { int p, q_tile, split; };
__device__ __forceinline__ GbUnit decode_unit(const GbBatch& b, int idx) {
    int p = 0;
    while (p + 1 < b.n_problems && idx >= b.pr[p + 1].unit_begin) ++p;
    const int local = idx - b.pr[p].unit_begin;
    GbUnit u;
    u.p = p;
    u.split = local / b.pr[p].q_tiles;          // adjacent CTAs stream adjacent column tiles of the same rows
    u.q_tile = local - u.split * b.pr[p].q_tiles;
    return u;
}
// round r of the snake deal: even rounds left to right, odd rounds right to left
__device__ __forceinline__ int snake_unit(int round, int cta, int ctas) {
    return round * ctas + ((round & 1) ? (ctas - 1 - cta) : cta);
}

__global__ void __launch_bounds__(GB_THREADS, 1)
lora_grad_batched_kernel(const __grid_constant__ GbBatch batch) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = (int)blockIdx.x, ctas = (int)gridDim.x;
    const int rounds = (batch.total_units + ctas - 1) / ctas;

    const uint32_t bar_base = smem_base + GB_STAGES * GB_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (GB_STAGES + s); };
    auto acc_full = [&](int a) { return bar_base + 8u * (2 * GB_STAGES + a); };
    auto acc_empty = [&](int a) { return bar_base + 8u * (2 * GB_STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * GB_STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + GB_STAGES * GB_STAGE_BYTES + 8u * (2 * GB_STAGES + 4));
    volatile uint32_t* last_flag = tmem_slot_gen + 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GB_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128u);            // two accumulators of 64 fp32 columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    // Stream dependency.  A launch that follows a sibling batch of the same flush reads nothing that sibling writes: it
    // skips the wait, so its CTAs start on the SMs the sibling's CTAs vacate and fill the tail of that launch (each launch
    // ends with some CTAs one unit short).  That is only sound because every launch releases its dependents AFTER its own
    // wait: a chained launch can only be running once the head of the chain has seen everything before the flush complete.
    // A chained launch still waits for its predecessor — at its END (below), so that "the last launch of the chain has
    // completed" keeps meaning "the whole chain has", which is what the next kernel on the stream relies on.
    if (!batch.chained) otb_pdl_wait();          // Gs / Ts / dY come from launches before this flush
    otb_pdl_trigger();

    if (warp == 0) {
        // ================= TMA producer =================
        int s = 0; uint32_t ph = 0;
        for (int r = 0; r < rounds; ++r) {
            const int idx = snake_unit(r, cta, ctas);
            if (idx >= batch.total_units) continue;
            const GbUnit u = decode_unit(batch, idx);
            const GbProblem& pr = batch.pr[u.p];
            const int q0 = u.q_tile * 128;
            const int m_begin = u.split * pr.rows_per_split;
            const int m_end = min(pr.M, m_begin + pr.rows_per_split);
            for (int m = m_begin; m < m_end; m += GB_ROWS) {
                mbar_wait(empty_bar(s), ph ^ 1);
                if (elect_one()) {
                    const uint32_t stg = smem_base + GB_STAGE_BYTES * s;
                    mbar_expect_tx(full_bar(s), GB_STAGE_BYTES);
                    tma_load_2d(stg, &pr.tm_v, full_bar(s), q0, m);
                    tma_load_2d(stg + GB_VHALF, &pr.tm_v, full_bar(s), q0 + 64, m);
                    tma_load_2d(stg + 2 * GB_VHALF, &pr.tm_u, full_bar(s), 0, m);
                }
                __syncwarp();
                if (++s == GB_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // c_format f32 [4,6) = 1, a/b format bf16 [7,10)/[10,13) = 1, a_major / b_major MN-major [15]/[16] = 1, N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int s = 0; uint32_t ph = 0;
        int it = 0;
        for (int r = 0; r < rounds; ++r) {
            const int idx = snake_unit(r, cta, ctas);
            if (idx >= batch.total_units) continue;
            const GbUnit u = decode_unit(batch, idx);
            const GbProblem& pr = batch.pr[u.p];
            const int m_begin = u.split * pr.rows_per_split;
            const int m_end = min(pr.M, m_begin + pr.rows_per_split);
            const int as = it & 1;
            mbar_wait(acc_empty(as), (uint32_t)((it >> 1) & 1) ^ 1);      // the epilogue has drained unit it-2
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * 64);
            for (int m = m_begin; m < m_end; m += GB_ROWS) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t stg = smem_base + GB_STAGE_BYTES * s;
                    const uint64_t adesc = umma_desc_mn_sw128(stg, GB_VHALF);
                    const uint64_t bdesc = umma_desc_mn_sw128(stg + 2 * GB_VHALF, GB_VHALF);
#pragma unroll
                    for (int k = 0; k < GB_ROWS / 16; ++k)      // 16 rows = two 8-row swizzle atoms = 2048 B per k-step
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (m != m_begin || k != 0) ? 1u : 0u);
                    umma_commit(empty_bar(s));
                    if (m + GB_ROWS >= m_end) umma_commit(acc_full(as));
                }
                __syncwarp();
                if (++s == GB_STAGES) { s = 0; ph ^= 1; }
            }
            ++it;
        }
    } else {
        // ================= epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =================
        const int quad = warp & 3;
        const int ql = quad * 32 + lane;                  // output column of the tile owned by this thread
        int it = 0;
        for (int r = 0; r < rounds; ++r) {
            const int idx = snake_unit(r, cta, ctas);
            if (idx >= batch.total_units) continue;
            const GbUnit u = decode_unit(batch, idx);
            const GbProblem& pr = batch.pr[u.p];
            const int R = pr.R, Q = pr.Q;
            const int q = u.q_tile * 128 + ql;
            const int as = it & 1;
            mbar_wait(acc_full(as), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t t_lane = tmem_base + (uint32_t)(as * 64) + ((uint32_t)(quad * 32) << 16);
            float v[64];
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
                if (c < R) {
                    uint32_t w[16];
                    tmem_ld16(t_lane + (uint32_t)c, w);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[c + j] = __uint_as_float(w[j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(as));     // the accumulator is in registers: the next unit may overwrite it
            ++it;

            bool write_out = true;
            if (pr.splits > 1) {
                // park this split's partial tile; the last split to arrive sums all of them IN SPLIT ORDER
                float* mine = pr.partial + ((size_t)(u.q_tile * pr.splits + u.split) * 128 + ql) * R;
#pragma unroll
                for (int c = 0; c < 64; c += 4)
                    if (c < R) __stcg(reinterpret_cast<float4*>(mine + c), make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
                __threadfence();
                epi_bar_sync();
                if (threadIdx.x == 64) {
                    const int old = atomicAdd(pr.counters + u.q_tile, 1);
                    *last_flag = (old == pr.splits - 1) ? 1u : 0u;
                }
                epi_bar_sync();
                write_out = *last_flag != 0u;
                if (write_out) {
                    __threadfence();
#pragma unroll
                    for (int c = 0; c < 64; ++c) v[c] = 0.0f;
                    for (int sp = 0; sp < pr.splits; ++sp) {
                        const float* src = pr.partial + ((size_t)(u.q_tile * pr.splits + sp) * 128 + ql) * R;
#pragma unroll
                        for (int c = 0; c < 64; c += 4)
                            if (c < R) {
                                const float4 t = __ldcg(reinterpret_cast<const float4*>(src + c));
                                v[c] += t.x; v[c + 1] += t.y; v[c + 2] += t.z; v[c + 3] += t.w;
                            }
                    }
                    if (threadIdx.x == 64) pr.counters[u.q_tile] = 0;     // ready for the next launch
                }
                epi_bar_sync();          // last_flag is rewritten by the next split unit
            }
            if (write_out && q < Q) {
                // += : the buffers hold zero (first micro-batch of a step) or the running accumulation; ONE writer per element
                if (pr.transposed_out) {        // out[q][r]: this thread's R values are contiguous
                    float* dst = pr.out + (size_t)q * R;
#pragma unroll
                    for (int c = 0; c < 64; c += 4)
                        if (c < R) {
                            float4 t = *reinterpret_cast<const float4*>(dst + c);
                            t.x += v[c]; t.y += v[c + 1]; t.z += v[c + 2]; t.w += v[c + 3];
                            *reinterpret_cast<float4*>(dst + c) = t;
                        }
                } else {                        // out[r][q]: consecutive lanes hit consecutive addresses
#pragma unroll
                    for (int c = 0; c < 64; ++c)
                        if (c < R) pr.out[(size_t)c * Q + q] += v[c];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 128u);
    }
    if (batch.chained) otb_pdl_wait();           // do not complete before the sibling this launch overlapped
}

// ------------------------------- host side ------------------------------------------
struct HostProblem {
    const void* v; const void* u; float* out;
    int M, Q, R, transposed_out, leaf;
    int q_tiles, splits, rows_per_split;
};

int check_leaf(const otb_lora_grad_args* a, int i) {
    OTB_CHECK_ARG(a->M > 0 && a->R > 0 && a->R <= 64 && a->R % 8 == 0, "leaf %d: bad M=%d R=%d (R multiple of 8, <= 64)", i, a->M, a->R);
    const bool doA = a->dA != nullptr, doB = a->dB != nullptr;
    if (doA) OTB_CHECK_ARG(a->gs && a->x && a->K > 0 && a->K % 8 == 0, "leaf %d: dA needs gs, x and K %% 8 == 0", i);
    if (doB) OTB_CHECK_ARG(a->ts && a->dy && a->N > 0 && a->N % 8 == 0, "leaf %d: dB needs ts, dy and N %% 8 == 0", i);
    OTB_CHECK_ARG((((uintptr_t)a->x | (uintptr_t)a->dy | (uintptr_t)a->gs | (uintptr_t)a->ts | (uintptr_t)a->dA | (uintptr_t)a->dB) & 15) == 0,
                  "leaf %d: all pointers must be 16-byte aligned", i);
    return OTB_OK;
}

// Split policy.  Full batches have many more column tiles than SMs and are never split (one writer per output
// element, no workspace traffic).  Small batches are split until there are ~2 units per SM; very long M is always cut
// so that the longest unit (which bounds the tail of the launch) stays near 8192 rows.
int plan_problems(const otb_lora_grad_args* leaves, int count, bool have_workspace, HostProblem* out, int* n_out) {
    int n = 0;
    for (int i = 0; i < count; ++i) {
        const otb_lora_grad_args* a = &leaves[i];
        int rc = check_leaf(a, i);
        if (rc != OTB_OK) return rc;
        if (a->dA) out[n++] = HostProblem{a->x, a->gs, a->dA, a->M, a->K, a->R, 0, i, (a->K + 127) / 128, 1, 0};
        if (a->dB) out[n++] = HostProblem{a->dy, a->ts, a->dB, a->M, a->N, a->R, 1, i, (a->N + 127) / 128, 1, 0};
    }
    int tiles = 0;
    for (int i = 0; i < n; ++i) tiles += out[i].q_tiles;
    const int sms = otb_sm_count();
    const int long_rows = env_int("OTB_KGRAD_SPLIT_ROWS", 8192);
    const int want = tiles > 0 ? (2 * sms + tiles - 1) / tiles : 1;
    for (int i = 0; i < n; ++i) {
        HostProblem& p = out[i];
        const int blocks = (p.M + GB_ROWS - 1) / GB_ROWS;
        int splits = 1;
        if (have_workspace) {
            splits = want;
            const int by_len = (p.M + long_rows - 1) / long_rows;
            if (by_len > splits) splits = by_len;
            const int max_splits = (blocks + 1) / 2;                     // at least 256 rows per split
            if (splits > max_splits) splits = max_splits;
            if (splits < 1) splits = 1;
        }
        const int per = (blocks + splits - 1) / splits;
        p.rows_per_split = per * GB_ROWS;
        p.splits = (blocks + per - 1) / per;
    }
    *n_out = n;
    return OTB_OK;
}

int64_t partial_bytes(const HostProblem* p, int n) {
    int64_t b = 0;
    for (int i = 0; i < n; ++i)
        if (p[i].splits > 1) b += (int64_t)p[i].q_tiles * p[i].splits * 128 * p[i].R * 4;
    return b;
}

}  // namespace

extern "C" int64_t otb_lora_grad_workspace_bytes(const otb_lora_grad_args* leaves, int32_t count) {
    if (leaves == nullptr || count < 1 || count > OTB_GRAD_BATCH_MAX) return -1;
    HostProblem hp[GB_MAX_PROBLEMS];
    int n = 0;
    if (plan_problems(leaves, count, true, hp, &n) != OTB_OK) return -1;
    return GB_COUNTER_BYTES + partial_bytes(hp, n);
}

extern "C" int otb_lora_grad_batched(const otb_lora_grad_args* leaves, int32_t count, void* workspace, int64_t workspace_bytes,
                                     int32_t chained, void* stream) {
    OTB_CHECK_ARG(leaves != nullptr && count >= 1 && count <= OTB_GRAD_BATCH_MAX, "batch of %d leaves (1..%d)", count, OTB_GRAD_BATCH_MAX);
    OTB_CHECK_ARG(workspace == nullptr || ((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    HostProblem hp[GB_MAX_PROBLEMS];
    int n = 0;
    int rc = plan_problems(leaves, count, workspace != nullptr, hp, &n);
    if (rc != OTB_OK) return rc;
    if (n == 0) return OTB_OK;
    if (workspace != nullptr) {
        const int64_t need = GB_COUNTER_BYTES + partial_bytes(hp, n);
        OTB_CHECK_ARG(workspace_bytes >= need, "workspace of %lld bytes, %lld needed (otb_lora_grad_workspace_bytes)",
                      (long long)workspace_bytes, (long long)need);
    }
    OTB_BIND(hp[0].v);
    // longest units first (the snake deal then balances the CTAs to within one of the shortest units)
    int order[GB_MAX_PROBLEMS];
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order, order + n, [&](int a, int b) { return hp[a].rows_per_split > hp[b].rows_per_split; });
    static thread_local GbBatch batch;      // 20 KB: keep it off the stack
    int units = 0, tiles = 0;
    int64_t part_off = GB_COUNTER_BYTES;
    for (int j = 0; j < n; ++j) {
        const HostProblem& p = hp[order[j]];
        GbProblem& g = batch.pr[j];
        if ((rc = make_tmap(&g.tm_v, p.v, p.M, p.Q, GB_ROWS, p.transposed_out ? "dy" : "x")) != OTB_OK) return rc;
        if ((rc = make_tmap(&g.tm_u, p.u, p.M, p.R, GB_ROWS, p.transposed_out ? "ts" : "gs")) != OTB_OK) return rc;
        g.out = p.out;
        g.M = p.M; g.Q = p.Q; g.R = p.R;
        g.q_tiles = p.q_tiles; g.splits = p.splits; g.rows_per_split = p.rows_per_split;
        g.unit_begin = units;
        g.transposed_out = p.transposed_out;
        g.partial = nullptr; g.counters = nullptr;
        if (p.splits > 1) {
            OTB_CHECK_ARG((tiles + p.q_tiles) * 4 <= GB_COUNTER_BYTES, "too many split column tiles for the workspace's counter block");
            g.counters = reinterpret_cast<int32_t*>(workspace) + tiles;
            g.partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + part_off);
            part_off += (int64_t)p.q_tiles * p.splits * 128 * p.R * 4;
            tiles += p.q_tiles;
        }
        units += p.q_tiles * p.splits;
    }
    batch.n_problems = n;
    batch.total_units = units;
    batch.chained = (chained && otb_pdl_enabled()) ? 1 : 0;
    static thread_local int attr_dev = -1;
    int dev = 0;
    OTB_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        OTB_CUDA(cudaFuncSetAttribute(lora_grad_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM));
        attr_dev = dev;
    }
    const int sms = otb_sm_count();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(units < sms ? units : sms));
    cfg.blockDim = dim3(GB_THREADS);
    cfg.dynamicSmemBytes = GB_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = otb_pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, lora_grad_batched_kernel, batch);
    if (e != cudaSuccess) {
        otb_set_error("launch of lora_grad_batched_kernel (%d problems, %d units) failed: %s", n, units, cudaGetErrorString(e));
        return OTB_ECUDA;
    }
    return OTB_OK;
}

// single-leaf form: the batched kernel with one leaf and no workspace (M is not split)
extern "C" int otb_lora_grad_ab(const otb_lora_grad_args* a, void* stream) {
    OTB_CHECK_ARG(a != nullptr, "null args");
    if (a->dA == nullptr && a->dB == nullptr) return OTB_OK;
    return otb_lora_grad_batched(a, 1, nullptr, 0, 0, stream);
}
