// lora_gemm.cu — K-a: LoRA down/up projections fused with the frozen base Linear.
//
//   y[M,N] = x[M,K] . w[N,K]^T  + bias[N] + drop(s * x . down[R,K]^T)_bf16 . up[N,R]^T
//
// Replaces LoRAModule.forward  /root/reference/modules/module/LoRAModule.py:319-323 (three GEMMs,
// dropout, scale and add as separate ATen launches) and, called with the transposed shadows,
// the dX half of its autograd backward.
//
// sm_100a design, common to both variants below (warp-specialised; v2: one CTA per 128 x tile_n tile, v3: a CTA pair
// per 256 x tile_n tile):
//   warp 0     TMA producer: x / w / down tiles (64-wide K slabs, 128B swizzle) into a
//              multi-stage smem ring, mbarrier full/empty pairs
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer.  The w tile and the down
//              tile are adjacent in smem, so ONE UMMA of N' = tile_n + R per 16-wide k-step
//              accumulates both x.w^T (TMEM columns [0,tile_n)) and T = x.down^T (columns
//              [tile_n, tile_n+R)) — the skinny LoRA GEMM costs no extra pass over x.
//   warps 2-5  epilogue: read T from TMEM, scale/dropout, round to bf16, park it in smem in
//              the canonical K-major 128B-swizzled layout; warp 1 then issues R/16 more
//              UMMAs (T . up^T) into the same accumulator; finally the accumulator is read
//              back with tcgen05.ld, bias added, rounded to bf16 and stored.
// Accumulation is fp32 in TMEM; T is rounded to bf16 once (it is an MMA operand) and the
// output once — fewer roundings than the reference's five.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 8;

struct GemmParams {
    int M, N, K;
    int R;             // LoRA rank as stored (t_out leading dimension)
    int Rp;            // rank padded to a multiple of 16 (UMMA N / K granularity); 0 = no LoRA branch
    int tile_n;        // output tile width, multiple of 16, tile_n + Rp <= 256
    int stages;
    int num_kb;        // ceil(K / 64)
    int n_tiles;
    int tmem_cols;     // power of two >= tile_n + Rp
    float scale;
    float dropout_p;
    uint64_t seed, offset;
    const void* bias;
    int bias_dtype;
    uint16_t* y;
    uint16_t* t_out;
};

// =====================================================================================
// v2: persistent, thread-block-cluster TMA multicast, double-buffered TMEM accumulators
// =====================================================================================
// What the first, one-tile-per-CTA version left on the table: one tile per CTA means the epilogue
// (TMEM -> registers -> HBM) and the bf16 T hand-off never overlap the next tile's MMAs, and every
// CTA pulls its own copy of the x and w tiles through L2: 48 KB per 64-wide k-block per CTA is
// ~27 TB/s of L2->SMEM traffic at the tensor peak — the kernel was L2-bound at 32 % tensor-pipe
// utilisation.  v2:
//   * clusters of CM x CN CTAs work on CM x CN neighbouring tiles; each CTA TMA-loads 1/CN of its
//     x tile and 1/CM of its [w; down] tile and MULTICASTS it to the peers that need it, so L2
//     traffic per CTA drops from (x + w) to (x/CN + w/CM);
//   * CTAs are persistent: the smem ring keeps streaming across tiles, the accumulator lives in one
//     of two TMEM stages, and the 4 epilogue warps drain tile i while the MMA warp runs tile i+1;
//   * the rank-R hand-off (T: TMEM -> bf16 -> smem -> UMMA) of tile i is issued two k-blocks INTO
//     tile i+1's mainloop, so the tensor pipe does not idle while the epilogue converts T.
struct GemmParams2 {
    GemmParams g;
    int cm, cn;            // cluster shape (tiles along M, along N)
    int super_m, super_n;  // number of super-tiles (cluster-sized tile groups)
    int m_tiles;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctaid_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// execution barrier only: what it orders is barrier initialisation (published by fence.mbarrier_init.release.cluster) or
// the end of all work of the pair — no data; `.release` here is a MEMBAR that waits for every outstanding store
__device__ __forceinline__ void cluster_sync_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// arrive (once the preceding tcgen05 ops retire) on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"(mask) : "memory");
}

struct SmemLayout2 {
    uint32_t stage_bytes, x_off, w_off, up_off, up_bytes, t_off, bar_off, total;
};
__host__ __device__ inline SmemLayout2 make_layout2(int tile_n, int Rp, int stages) {
    SmemLayout2 L;
    L.x_off = 0;
    L.w_off = BM * 128;
    L.stage_bytes = (uint32_t)(BM + tile_n + Rp) * 128;
    L.up_off = L.stage_bytes * stages;
    L.up_bytes = Rp ? (uint32_t)tile_n * 128 : 0;
    L.t_off = L.up_off + 2 * L.up_bytes;                 // up tile is double-buffered across tiles
    L.bar_off = L.t_off + (Rp ? BM * 128 : 0);
    L.total = L.bar_off + 512;
    return L;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
lora_linear_kernel_v2(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                      const __grid_constant__ CUtensorMap tm_dn, const __grid_constant__ CUtensorMap tm_up,
                      const GemmParams2 q) {
    const GemmParams& p = q.g;
    otb_pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const SmemLayout2 L = make_layout2(p.tile_n, p.Rp, p.stages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool lora = p.Rp > 0;
    const int csize = q.cm * q.cn;
    const uint32_t crank = csize > 1 ? cluster_ctarank() : 0;
    const int rm = (int)crank % q.cm, rn = (int)crank / q.cm;      // position of this CTA's tile inside the super-tile
    const uint32_t cid = csize > 1 ? cluster_id_x() : blockIdx.x;
    const uint32_t ncl = csize > 1 ? cluster_nctaid_x() : gridDim.x;
    const int total_super = q.super_m * q.super_n;

    const uint32_t bar_base = smem_base + L.bar_off;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
    auto misc = [&](int i) { return bar_base + 8u * (2 * MAX_STAGES + i); };
    // misc: 0,1 up_full  2,3 up_empty  4,5 t_full  6 t_ready  8,9 d_full  10,11 acc_empty  12 tmem slot
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L.bar_off + 8u * (2 * MAX_STAGES + 12));

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_x); prefetch_tmap(&tm_w);
        if (lora) { prefetch_tmap(&tm_dn); prefetch_tmap(&tm_up); }
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), (uint32_t)csize); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(misc(0 + i), 1); mbar_init(misc(2 + i), 1); mbar_init(misc(4 + i), 1);
            mbar_init(misc(8 + i), 1); mbar_init(misc(10 + i), 128);
        }
        mbar_init(misc(6), 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(misc(12), 512u);
    tc_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();          // peers' barriers must exist before any multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    constexpr uint32_t ACC_STRIDE = 256;        // TMEM columns per accumulator stage
    otb_pdl_wait();

    const uint16_t all_mask = (uint16_t)((1u << csize) - 1u);
    // CTAs that share my x tile: same rm, all rn ;  CTAs that share my [w; down] tile: same rn, all rm
    uint16_t x_mask = 0, w_mask = 0;
    for (int j = 0; j < q.cn; ++j) x_mask |= (uint16_t)(1u << (rm + j * q.cm));
    for (int j = 0; j < q.cm; ++j) w_mask |= (uint16_t)(1u << (j + rn * q.cm));
    const int x_rows = BM / q.cn, w_rows = p.tile_n / q.cm, d_rows = lora ? p.Rp / q.cm : 0;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (uint32_t st = cid; st < (uint32_t)total_super; st += ncl, ++it) {
                const int sm_i = st / q.super_n, sn_i = st % q.super_n;
                const int m0 = (sm_i * q.cm + rm) * BM, n0 = (sn_i * q.cn + rn) * p.tile_n;
                if (lora) {
                    const int ub = it & 1;
                    mbar_wait(misc(2 + ub), ((it >> 1) & 1) ^ 1);
                    mbar_expect_tx(misc(0 + ub), L.up_bytes);
                    tma_load_2d(smem_base + L.up_off + ub * L.up_bytes, &tm_up, misc(0 + ub), 0, n0);
                }
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1);
                    const uint32_t stg = smem_base + L.stage_bytes * s;
                    mbar_expect_tx(full_bar(s), L.stage_bytes);
                    if (csize == 1) {
                        tma_load_2d(stg + L.x_off, &tm_x, full_bar(s), kb * BK, m0);
                        tma_load_2d(stg + L.w_off, &tm_w, full_bar(s), kb * BK, n0);
                        if (lora) tma_load_2d(stg + L.w_off + (uint32_t)p.tile_n * 128, &tm_dn, full_bar(s), kb * BK, 0);
                    } else {
                        // my slice of each shared tile, delivered to every CTA that needs it
                        tma_load_2d_mc(stg + L.x_off + (uint32_t)(rn * x_rows) * 128, &tm_x, full_bar(s), kb * BK, m0 + rn * x_rows, x_mask);
                        tma_load_2d_mc(stg + L.w_off + (uint32_t)(rm * w_rows) * 128, &tm_w, full_bar(s), kb * BK, n0 + rm * w_rows, w_mask);
                        if (lora)
                            tma_load_2d_mc(stg + L.w_off + (uint32_t)(p.tile_n + rm * d_rows) * 128, &tm_dn, full_bar(s), kb * BK, rm * d_rows, w_mask);
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc_main = umma_idesc_bf16(BM, p.tile_n + p.Rp);
            const uint32_t idesc_up = umma_idesc_bf16(BM, p.tile_n);
            int s = 0; uint32_t ph = 0;
            int it = 0, pending = -1;
            auto issue_up = [&](int j) {      // D[j&1] += T(j) . up(j)^T, then hand the accumulator to the epilogue
                const int as = j & 1;
                mbar_wait(misc(6), (uint32_t)(j & 1));                     // t_ready: epilogue parked bf16 T(j) in smem
                mbar_wait(misc(0 + as), (uint32_t)((j >> 1) & 1));         // up_full
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128(smem_base + L.t_off);
                const uint64_t bdesc = umma_desc_sw128(smem_base + L.up_off + as * L.up_bytes);
                for (int k = 0; k < p.Rp / UMMA_K; ++k)
                    umma_bf16(tmem_base + as * ACC_STRIDE, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_up, 1u);
                umma_commit(misc(8 + as));                                  // d_full
                umma_commit(misc(2 + as));                                  // up_empty
            };
            for (uint32_t st = cid; st < (uint32_t)total_super; st += ncl, ++it) {
                const int as = it & 1;
                mbar_wait(misc(10 + as), ((it >> 1) & 1) ^ 1);             // acc_empty: epilogue drained tile it-2
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
                const int lookahead = p.num_kb < 2 ? p.num_kb : 2;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t stg = smem_base + L.stage_bytes * s;
                    const uint64_t adesc = umma_desc_sw128(stg + L.x_off);
                    const uint64_t bdesc = umma_desc_sw128(stg + L.w_off);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_main, (kb | k) != 0);
                    if (csize == 1) umma_commit(empty_bar(s)); else umma_commit_mc(empty_bar(s), all_mask);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                    if (pending >= 0 && kb == lookahead - 1) { issue_up(pending); pending = -1; }
                }
                if (lora) { umma_commit(misc(4 + as)); pending = it; }     // t_full
                else umma_commit(misc(8 + as));                            // d_full
            }
            if (pending >= 0) issue_up(pending);
        }
    } else {
        // ================= epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const float keep_scale = p.dropout_p > 0.0f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
        uint8_t* t_row = smem_gen + L.t_off + row * 128;
        int it = 0;
        for (uint32_t st = cid; st < (uint32_t)total_super; st += ncl, ++it) {
            const int as = it & 1;
            const uint32_t par = (uint32_t)((it >> 1) & 1);
            const int sm_i = st / q.super_n, sn_i = st % q.super_n;
            const int m_tile = sm_i * q.cm + rm, n_tile = sn_i * q.cn + rn;
            const int m0 = m_tile * BM, n0 = n_tile * p.tile_n;
            const int m = m0 + row;
            const uint32_t t_lane = tmem_base + as * ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
            if (lora) {
                mbar_wait(misc(4 + as), par);                               // t_full
                tc_fence_after();
                for (int c = 0; c < p.Rp; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + (uint32_t)(p.tile_n + c), v);
                    tmem_ld_wait();
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * p.scale;
                    if (p.dropout_p > 0.0f) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const uint64_t e = (uint64_t)m * (uint64_t)p.Rp + (uint64_t)(c + j);
                            Philox4 r = otb_rng(p.seed, p.offset, OTB_STREAM_DROPOUT, e >> 2);
                            f[j + 0] = otb_u01(r.x) >= p.dropout_p ? f[j + 0] * keep_scale : 0.0f;
                            f[j + 1] = otb_u01(r.y) >= p.dropout_p ? f[j + 1] * keep_scale : 0.0f;
                            f[j + 2] = otb_u01(r.z) >= p.dropout_p ? f[j + 2] * keep_scale : 0.0f;
                            f[j + 3] = otb_u01(r.w) >= p.dropout_p ? f[j + 3] * keep_scale : 0.0f;
                        }
                    }
                    uint4 lo = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                    uint4 hi = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
                    const int j0 = c >> 3;
                    *reinterpret_cast<uint4*>(t_row + (((j0 + 0) ^ (row & 7)) << 4)) = lo;
                    *reinterpret_cast<uint4*>(t_row + (((j0 + 1) ^ (row & 7)) << 4)) = hi;
                    if (p.t_out != nullptr && n_tile == 0 && m < p.M) {
                        uint16_t* dst = p.t_out + (size_t)m * p.R + c;
                        if (c + 8 <= p.R) *reinterpret_cast<uint4*>(dst) = lo;
                        if (c + 16 <= p.R) *reinterpret_cast<uint4*>(dst + 8) = hi;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(misc(6));                                       // t_ready
            }
            mbar_wait(misc(8 + as), par);                                   // d_full
            tc_fence_after();
            const bool tile_live = m_tile < q.m_tiles && n0 < p.N;
            for (int c = 0; c < p.tile_n && tile_live; c += 16) {
                const int n = n0 + c;
                if (n >= p.N) break;
                uint32_t v[16];
                tmem_ld16(t_lane + (uint32_t)c, v);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                if (p.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n + j < p.N) f[j] += load_as_float(p.bias, n + j, p.bias_dtype);
                }
                if (m < p.M) {
                    uint16_t* dst = p.y + (size_t)m * p.N + n;
                    if (n + 8 <= p.N)
                        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                    if (n + 16 <= p.N)
                        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
                }
            }
            tc_fence_before();
            mbar_arrive(misc(10 + as));                                     // acc_empty
        }
    }
    tc_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();          // nobody exits while a peer may still multicast into it
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 512u);
    }
}

// =====================================================================================
// v3: CTA pairs — tcgen05.mma.cta_group::2, 256-row tiles
// =====================================================================================
// What v2 left on the table (profiles/r01_ncu_ka_v2.md): with 128 x (256+32) tiles every SM has to ingest
// 52 KB of operands per 64-wide k-block for 576 tensor-pipe cycles — 90 B/clk, but an SM receives
// only ~42.5 B/clk from L2 (the 6300 B/clk chip-wide LTS cap / 148; cuBLAS tops out against the same
// wall).  ncu shows the tensor pipe 47 % active, exactly 42.5 / 90.  TMA multicast does not help: the
// tile still has to LAND in each SM.  What helps is making each SM ingest less per flop:
//   * two CTAs on the two SMs of a TPC form ONE 256 x tile_n tile.  Each CTA stages its own 128 rows
//     of x but only HALF of the [w; down] rows; tcgen05.mma.cta_group::2 reads both halves, each CTA's
//     TMEM receives the accumulators of its own 128 rows.  Ingest per SM per k-block drops from
//     (128 + N')·128 B to (128 + N'/2)·128 B;
//   * tile_n may exceed 256 (two N-chunks per k-step, one 512-column accumulator) when that turns
//     a 2-round problem into a 1-round one (e.g. M=4096, N=1280: 16 x 4 tiles of 256 x 320 on 74 pairs);
//   * otherwise two accumulator stages as in v2: the epilogue of tile i overlaps the mainloop of i+1.
// Roles per CTA: warp 0 TMA producer (both CTAs, each loads its halves, completion is signalled on the
// LEADER's mbarrier), warp 1 MMA issuer (leader CTA only; commits are multicast to both CTAs), warps 2-5
// epilogue (both CTAs; the rank-R hand-off and the accumulator release arrive on the leader's barriers).
// same descriptor as umma_desc_sw128, from a shared-memory address already in 16-byte units (< 2^14)
__device__ __forceinline__ uint64_t umma_desc16(uint32_t addr16) {
    return ((uint64_t)0x40004040u << 32) | (uint64_t)(addr16 | 0x10000u);
}
// TMA store of one swizzled smem box; completion is tracked per issuing thread in bulk async-groups
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// L2-only prefetch of one box of a tensor map (no shared-memory slot, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" :: "l"(map), "r"(c0), "r"(c1) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }

// Up to three same-shape problems in ONE launch (a "group": the q/k/v projections of a self-attention, the k/v
// projections of a cross-attention — or their three dX GEMMs).  Single-leaf launches leave a third of the chip idle
// on the most frequent SDXL shape (64 tiles of 256 x 320 on 74 CTA pairs, no second tile to hide the epilogue
// behind); a group of three is 288 tiles over 74 pairs in ~4 full rounds with the epilogue overlapped.
constexpr int V3_MAX_GROUP = 3;            // the common launch: a leaf, or the q/k/v (k/v) siblings of one attention
constexpr int V3_BIG_GROUP = OTB_LORA_GROUP_MAX;   // the wide launch: every context K/V projection of a network at once (same x)
struct V3Member {
    const void* bias;
    uint16_t* y;
    uint16_t* t_out;
    uint64_t seed, offset;
    float scale;
    int bias_dtype;
};
template <int G>
struct V3Maps {            // one tensor-map set per member of the group
    CUtensorMap x[G], w[G], w2[G], dn[G], up[G], y[G];
};
template <int G>
struct GemmParams3 {
    GemmParams g;          // g.tile_n: width of the pair tile; g.n_tiles (per member); g.stages; shapes common to the group
    int groups;            // members in this launch (1..G)
    V3Member mem[G];
    int chunks;            // MMAs per 16-wide k-step over the fused [w; down] rows (NB = tile_n + Rp, NB / chunks <= 256)
    int up_chunks;         // MMAs per k-step of the rank-R GEMM (tile_n / up_chunks <= 256)
    int acc_stages;        // 2 when NB <= 256, else 1
    int pair_m_tiles;      // ceil(M / 256)
    int prefetch_w;        // weight tiles of the first k-blocks are fetched before griddepcontrol.wait (OTB_LINEAR_STATIC_WEIGHTS)
    int lookahead;         // k-blocks of tile i+1 issued before the rank-R MMAs of tile i (acc_stages == 2); 0 = right after tile i
    int tail_reorder;      // single-accumulator two-chunk tiles: issue the last k-blocks chunk-major (OTB_KA_TAIL=0 switches it off)
    unsigned long long* trace;   // bring-up only (OTB_KA_TRACE=<device address>): %globaltimer stamps of the first tile of clusters 0..7
    int debug;             // bring-up only (OTB_KA_DEBUG): bit 0 skip the TMA loads, bit 1 skip the MMAs, bit 2 skip the epilogue,
                           // bit 3 skip the TMA stores, bit 4 skip the accumulator loads, bit 5 skip the async-proxy fence
};

// bring-up timeline: event e of cluster `cid` (leader CTA, first tile) -> trace[cid * 16 + e], nanoseconds of %globaltimer
__device__ __forceinline__ void ka_stamp(unsigned long long* trace, uint32_t cid, bool leader, int it, int e) {
    if (trace != nullptr && leader && cid < 8 && it == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        trace[cid * 16 + e] = t;
    }
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Arrival on a barrier of the peer (leader) CTA.  Default semantics (.release at CTA scope), as CUTLASS's ClusterBarrier
// does it: what the waiter consumes afterwards is TMEM (read back with tcgen05.ld + wait before the arrive) or this CTA's
// own shared memory through the async proxy (ordered by fence.proxy.async before the arrive) — nothing that needs a
// cluster-scope release.  With `.release.cluster` the arrive is a MEMBAR that waits for every earlier global store of the
// thread: 18 % of the kernel's stall samples on the short-K leaves (`stall_membar`, profiles/r02_ncu_kac640_lines.md).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
// (waits on barriers whose arrivals come from the peer CTA use the plain CTA-scope mbar_wait: what the waiter consumes
// afterwards is TMEM or async-proxy shared memory, as in CUTLASS's ClusterBarrier::wait)
// TMA load issued by either CTA of a pair; the bytes complete on `bar`, a shared::cluster address that may
// live in the peer (leader) CTA
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the preceding tcgen05 ops retire
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}

constexpr uint32_t YSTAGE_WARP_BYTES = 32 * 128;         // one 32-row x 64-column bf16 staging box per epilogue warp
constexpr int V3_EPI_WARPS = 8;                          // two warps per TMEM lane quadrant, alternating 64-column blocks
constexpr int V3_THREADS = 64 + 32 * V3_EPI_WARPS;

struct SmemLayout3 {
    uint32_t stage_bytes, x_off, b_off, up_off, up_bytes, t_off, y_off, bar_off, total;
};
__host__ __device__ inline SmemLayout3 make_layout3(int tile_n, int Rp, int stages, int up_bufs) {
    SmemLayout3 L;
    L.x_off = 0;
    L.b_off = BM * 128;                                   // this CTA's half of the fused [w; down] rows
    L.stage_bytes = L.b_off + (uint32_t)((tile_n + Rp) / 2) * 128;
    L.up_off = L.stage_bytes * stages;
    L.up_bytes = Rp ? (uint32_t)(tile_n / 2) * 128 : 0;
    L.t_off = L.up_off + up_bufs * L.up_bytes;
    L.y_off = L.t_off + (Rp ? BM * 128 : 0);
    L.bar_off = L.y_off + V3_EPI_WARPS * YSTAGE_WARP_BYTES;
    L.total = L.bar_off + 512;
    return L;
}

template <int G>
__global__ void __launch_bounds__(V3_THREADS, 1)
lora_linear_kernel_v3(const __grid_constant__ V3Maps<G> maps, const __grid_constant__ GemmParams3<G> q) {
    const GemmParams& p = q.g;
    otb_pdl_trigger();                           // the next kernel in the stream may begin its own prologue
    if (threadIdx.x == 0) ka_stamp(q.trace, cluster_id_x(), cluster_ctarank() == 0, 0, 0);      // 0: kernel entry
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int up_bufs = q.acc_stages;
    const SmemLayout3 L = make_layout3(p.tile_n, p.Rp, p.stages, up_bufs);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool lora = p.Rp > 0;
    const uint32_t crank = cluster_ctarank();          // 0 = leader
    const bool leader = crank == 0;
    const uint32_t cid = cluster_id_x(), ncl = cluster_nctaid_x();
    const int tiles_per_row = q.groups * p.n_tiles;       // tiles of one 256-row block: members side by side, so a shared x stays hot in L2
    const int total_tiles = q.pair_m_tiles * tiles_per_row;
    const int NB = p.tile_n + p.Rp;                    // fused B rows: [w tile ; down]
    const int nbc = NB / q.chunks, nbh = nbc / 2;      // MMA N per chunk; rows of it held by each CTA
    const int wc = p.tile_n / q.up_chunks, wch = wc / 2;   // same for the rank-R GEMM's `up` operand
    const int acc_stride = q.acc_stages == 2 ? 256 : 512;

    const uint32_t bar_base = smem_base + L.bar_off;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
    auto misc = [&](int i) { return bar_base + 8u * (2 * MAX_STAGES + i); };
    // misc: 0,1 up_full(leader)  2,3 up_empty  4,5 t_full  6 t_ready(leader)  8,9 d_full  10,11 acc_empty(leader)  12 tmem slot
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L.bar_off + 8u * (2 * MAX_STAGES + 12));

    if (threadIdx.x == 0) {
        for (int g = 0; g < q.groups && g < V3_MAX_GROUP; ++g) {
            prefetch_tmap(&maps.x[g]); prefetch_tmap(&maps.w[g]); prefetch_tmap(&maps.y[g]);
            if (lora) { prefetch_tmap(&maps.w2[g]); prefetch_tmap(&maps.dn[g]); prefetch_tmap(&maps.up[g]); }
        }
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(misc(0 + i), 1); mbar_init(misc(2 + i), 1); mbar_init(misc(4 + i), 1);
            mbar_init(misc(8 + i), 1); mbar_init(misc(10 + i), 2 * V3_EPI_WARPS);     // one arrival per epilogue warp of the pair
        }
        mbar_init(misc(6), 2 * V3_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(misc(12), 512u);
    tc_fence_before();
    __syncthreads();
    cluster_sync_relaxed();                     // the peer's barriers and TMEM exist before anything crosses CTAs
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    if (threadIdx.x == 0) ka_stamp(q.trace, cid, leader, 0, 1);                                  // 1: prologue done
    // Everything above overlapped the previous kernel's tail.  griddepcontrol.wait (its outputs are visible afterwards) is
    // issued per role: the producer first fetches the WEIGHT tiles of its first k-blocks when the caller vouches that the
    // preceding launch did not write them — the frozen base and the LoRA shadows come cold from HBM (~2 us to the first
    // byte), the activations x are what the predecessor produced and are read after the wait.
    if (warp != 0) otb_pdl_wait();

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        // The whole warp runs the loop (warp-uniform control flow and values, so addresses and coordinates stay in
        // uniform registers); one elected lane issues.  A lone `if (lane == 0)` thread pays an ELECT + R2UR sequence
        // per UTMALDG / UTCHMMA — measured ~350 cycles per MMA on the issuing thread, 2.7x the tensor pipe's 128.
        int s = 0; uint32_t ph = 0;
        int it = 0;
        const uint32_t stage_tx = 2u * L.stage_bytes, up_tx = 2u * L.up_bytes;
        // the [w; down] halves of k-block kb of a tile into stage `stg` (completion on the leader's barrier `bar`)
        auto load_weights = [&](uint32_t stg, uint32_t bar, int kb, int n0, const CUtensorMap* tm_w, const CUtensorMap* tm_w2, const CUtensorMap* tm_dn) {
            // my half of every chunk of the fused [w; down] rows; only the very last half (peer CTA, last chunk) reaches
            // into `down`, and it holds all Rp rows of it
            for (int c = 0; c < q.chunks; ++c) {
                const uint32_t dst = stg + L.b_off + (uint32_t)(c * nbh) * 128;
                const int row = c * nbc + (int)crank * nbh;      // first fused row of this half
                if (lora && c == q.chunks - 1 && !leader) {
                    if (nbh > p.Rp) tma_load_2d_2sm(dst, tm_w2, bar, kb * BK, n0 + row);
                    tma_load_2d_2sm(dst + (uint32_t)(nbh - p.Rp) * 128, tm_dn, bar, kb * BK, 0);
                } else {
                    tma_load_2d_2sm(dst, tm_w, bar, kb * BK, n0 + row);
                }
            }
        };
        int prefetched = 0;            // k-blocks of this CTA's FIRST tile whose weights are already on their way
        if (q.prefetch_w == 2 && !(q.debug & 1) && cid < (uint32_t)total_tiles) {
            // L2-only variant: the frozen weight tiles of this CTA's first tile are pulled from HBM into L2 while the previous
            // launch still runs; nothing queues in front of the first x box and no ring slot is taken
            const int gm = ((int)cid % tiles_per_row) / p.n_tiles, nt = (int)cid % p.n_tiles;
            const int n0 = nt * p.tile_n;
            if (elect_one()) {
                for (int kb = 0; kb < p.num_kb; ++kb)
                    for (int c = 0; c < q.chunks; ++c) {
                        const int row = c * nbc + (int)crank * nbh;
                        if (lora && c == q.chunks - 1 && !leader) {
                            if (nbh > p.Rp) tma_prefetch_l2_2d(&maps.w2[gm], kb * BK, n0 + row);
                            tma_prefetch_l2_2d(&maps.dn[gm], kb * BK, 0);
                        } else {
                            tma_prefetch_l2_2d(&maps.w[gm], kb * BK, n0 + row);
                        }
                    }
            }
            __syncwarp();
        } else if (q.prefetch_w && !(q.debug & 1) && cid < (uint32_t)total_tiles) {
            const int gm = ((int)cid % tiles_per_row) / p.n_tiles, nt = (int)cid % p.n_tiles;
            const int n0 = nt * p.tile_n;
            prefetched = p.num_kb < p.stages ? p.num_kb : p.stages;
            if (elect_one()) {
                for (int kb = 0; kb < prefetched; ++kb) {            // the ring is empty: stages 0 .. prefetched-1, no wait needed
                    if (leader) mbar_expect_tx(full_bar(kb), stage_tx);
                    load_weights(smem_base + L.stage_bytes * kb, mapa_u32(full_bar(kb), 0), kb, n0, &maps.w[gm], &maps.w2[gm], &maps.dn[gm]);
                }
            }
            __syncwarp();
        }
        otb_pdl_wait();
        if (lane == 0) ka_stamp(q.trace, cid, leader, 0, 2);                                     // 2: previous launch complete
        for (uint32_t st = cid; st < (uint32_t)total_tiles; st += ncl, ++it) {
            const int mt = st / tiles_per_row, gm = (st % tiles_per_row) / p.n_tiles, nt = st % p.n_tiles;
            const int m0 = mt * 2 * BM + (int)crank * BM, n0 = nt * p.tile_n;
            const CUtensorMap* tm_x = &maps.x[gm]; const CUtensorMap* tm_w = &maps.w[gm]; const CUtensorMap* tm_w2 = &maps.w2[gm];
            const CUtensorMap* tm_dn = &maps.dn[gm]; const CUtensorMap* tm_up = &maps.up[gm];
            if (lora) {
                const int ub = it % up_bufs;
                mbar_wait(misc(2 + ub), (uint32_t)((it / up_bufs) & 1) ^ 1);                 // up_empty (own)
                if (elect_one()) {
                    if (leader) mbar_expect_tx(misc(0 + ub), up_tx);
                    const uint32_t bar = mapa_u32(misc(0 + ub), 0);
                    for (int c = 0; c < q.up_chunks; ++c)
                        tma_load_2d_2sm(smem_base + L.up_off + ub * L.up_bytes + (uint32_t)(c * wch) * 128, tm_up, bar, 0,
                                        n0 + c * wc + (int)crank * wch);
                }
                __syncwarp();
            }
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(empty_bar(s), ph ^ 1);
                if (elect_one()) {
                    const uint32_t stg = smem_base + L.stage_bytes * s;
                    if (q.debug & 1) {
                        if (leader) mbar_arrive(full_bar(s));
                    } else {
                        const bool early = it == 0 && kb < prefetched;      // expect_tx and the weight loads were issued before the wait
                        if (leader && !early) mbar_expect_tx(full_bar(s), stage_tx);
                        const uint32_t bar = mapa_u32(full_bar(s), 0);
                        tma_load_2d_2sm(stg + L.x_off, tm_x, bar, kb * BK, m0);
                        if (!early) load_weights(stg, bar, kb, n0, tm_w, tm_w2, tm_dn);
                    }
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only; warp-uniform, one elected lane issues) =================
        if (leader) {
            const uint32_t idesc_main = umma_idesc_bf16(2 * BM, nbc);
            const uint32_t idesc_up = umma_idesc_bf16(2 * BM, wc);
            const uint32_t sb16 = smem_base >> 4;                     // descriptor address unit: 16 B
            const uint32_t stage16 = L.stage_bytes >> 4, b16off = L.b_off >> 4, chunk16 = (uint32_t)(nbh * 128) >> 4;
            int s = 0; uint32_t ph = 0;
            int it = 0, pending = -1;
            auto issue_up = [&](int j) {      // D += T(j) . up(j)^T, then hand the accumulator to the epilogue
                const int as = j % q.acc_stages, ub = j % up_bufs;
                mbar_wait(misc(6), (uint32_t)(j & 1));                                      // t_ready: both CTAs parked T(j)
                mbar_wait(misc(0 + ub), (uint32_t)((j / up_bufs) & 1));                     // up_full
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a16 = sb16 + (L.t_off >> 4), b16 = sb16 + ((L.up_off + ub * L.up_bytes) >> 4);
                    for (int c = 0; c < q.up_chunks; ++c)
                        for (int k = 0; k < p.Rp / UMMA_K; ++k)
                            umma_bf16_2sm(tmem_base + as * acc_stride + c * wc, umma_desc16(a16 + k * 2),
                                          umma_desc16(b16 + c * ((uint32_t)(wch * 128) >> 4) + k * 2), idesc_up, 1u);
                    umma_commit_2sm(misc(8 + as));                                           // d_full
                    umma_commit_2sm(misc(2 + ub));                                           // up_empty
                }
                __syncwarp();
            };
            for (uint32_t st = cid; st < (uint32_t)total_tiles; st += ncl, ++it) {
                const int as = it % q.acc_stages;
                mbar_wait(misc(10 + as), (uint32_t)((it / q.acc_stages) & 1) ^ 1);         // acc_empty
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * acc_stride;
                const int lookahead = p.num_kb < q.lookahead ? p.num_kb : q.lookahead;
                if (pending >= 0 && lookahead == 0) { issue_up(pending); pending = -1; }
                // One accumulator stage, two N-chunks (tiles wider than 224): nothing overlaps the rank-R hand-off of such a
                // tile, so its LAST k-blocks are issued chunk-major — first the chunk that holds the `down` columns, then
                // t_full, then the other chunk — and the epilogue converts T while the tensor pipe still works on the first
                // chunk (tools/ka_trace.py: 0.9 us between "mainloop complete" and "accumulator complete" before)
                int tail = 0;
                if (lora && q.chunks == 2 && q.acc_stages == 1 && q.tail_reorder) {
                    tail = p.stages - 1 < 3 ? p.stages - 1 : 3;
                    if (tail > p.num_kb) tail = p.num_kb;
                }
                for (int kb = 0; kb < p.num_kb - tail; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    if (kb == 0 && lane == 0) ka_stamp(q.trace, cid, true, it, 3);               // 3: first operands landed
                    if (elect_one()) {
                        const uint32_t a16 = sb16 + stage16 * s, b16 = a16 + b16off;
                        if (!(q.debug & 2)) {
                            if (q.chunks == 1) {
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k)
                                    umma_bf16_2sm(d_tmem, umma_desc16(a16 + k * 2), umma_desc16(b16 + k * 2), idesc_main, (kb | k) != 0);
                            } else {
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k) {
                                    umma_bf16_2sm(d_tmem, umma_desc16(a16 + k * 2), umma_desc16(b16 + k * 2), idesc_main, (kb | k) != 0);
                                    umma_bf16_2sm(d_tmem + nbc, umma_desc16(a16 + k * 2), umma_desc16(b16 + chunk16 + k * 2), idesc_main, (kb | k) != 0);
                                }
                            }
                        }
                        umma_commit_2sm(empty_bar(s));
                    }
                    __syncwarp();
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                    if (pending >= 0 && kb == lookahead - 1) { issue_up(pending); pending = -1; }
                }
                if (tail > 0) {
                    const int s0 = s;
                    auto slot = [&](int j) { const int v = s0 + j; return v >= p.stages ? v - p.stages : v; };
                    for (int j = 0; j < tail; ++j) {                                         // every tail slot has landed
                        mbar_wait(full_bar(s), ph);
                        if (++s == p.stages) { s = 0; ph ^= 1; }
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const int kb0 = p.num_kb - tail;
                        for (int j = 0; j < tail && !(q.debug & 2); ++j) {                   // the chunk with the `down` columns first
                            const uint32_t a16 = sb16 + stage16 * slot(j), b16 = a16 + b16off;
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_bf16_2sm(d_tmem + nbc, umma_desc16(a16 + k * 2), umma_desc16(b16 + chunk16 + k * 2), idesc_main, ((kb0 + j) | k) != 0);
                        }
                        umma_commit_2sm(misc(4 + as));                                       // t_full: T is complete
                        for (int j = 0; j < tail; ++j) {
                            const uint32_t a16 = sb16 + stage16 * slot(j), b16 = a16 + b16off;
                            if (!(q.debug & 2)) {
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k)
                                    umma_bf16_2sm(d_tmem, umma_desc16(a16 + k * 2), umma_desc16(b16 + k * 2), idesc_main, ((kb0 + j) | k) != 0);
                            }
                            umma_commit_2sm(empty_bar(slot(j)));
                        }
                    }
                    __syncwarp();
                }
                if (lane == 0) ka_stamp(q.trace, cid, true, it, 4);                              // 4: last k-block issued
                if (lora) {
                    if (tail == 0 && elect_one()) umma_commit_2sm(misc(4 + as));             // t_full
                    __syncwarp();
                    if (q.acc_stages == 2) pending = it; else issue_up(it);
                } else {
                    if (elect_one()) umma_commit_2sm(misc(8 + as));                          // d_full
                    __syncwarp();
                }
            }
            if (pending >= 0) issue_up(pending);
        }
    } else {
        // ================= epilogue (warps 2..9 -> TMEM lane quadrant warp % 4; both CTAs) =================
        // two warps share a lane quadrant: `half` 0 takes the even 64-column blocks (and 16-column T chunks), 1 the odd ones
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const float keep_scale = p.dropout_p > 0.0f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
        uint8_t* t_row = smem_gen + L.t_off + row * 128;
        const uint32_t t_ready_leader = mapa_u32(misc(6), 0);
        // output staging: accumulator -> registers -> bf16 -> 128B-swizzled smem box -> TMA store (coalesced, clipped
        // at the M / N edges by the tensor map); two boxes per warp so a store drains while the next box is packed
        const uint32_t ybox = smem_base + L.y_off + (uint32_t)(warp - 2) * YSTAGE_WARP_BYTES;
        uint8_t* ybox_gen = smem_gen + L.y_off + (uint32_t)(warp - 2) * YSTAGE_WARP_BYTES;
        int it = 0;
        for (uint32_t st = cid; st < (uint32_t)total_tiles; st += ncl, ++it) {
            const int as = it % q.acc_stages;
            const uint32_t par = (uint32_t)((it / q.acc_stages) & 1);
            const int mt = st / tiles_per_row, gm = (st % tiles_per_row) / p.n_tiles, n_tile = st % p.n_tiles;
            const int m0 = mt * 2 * BM + (int)crank * BM, n0 = n_tile * p.tile_n;
            const int m = m0 + row;
            const V3Member& me = q.mem[gm];
            const bool bias_vec = me.bias != nullptr && (me.bias_dtype == OTB_BF16 || me.bias_dtype == OTB_F32) && (((uintptr_t)me.bias) & 15) == 0;
            const uint32_t t_lane = tmem_base + as * acc_stride + ((uint32_t)(quad * 32) << 16);
            // bf16 bias of this warp's 64-column blocks, one 4-byte word (two columns) per lane, requested NOW — before the
            // waits for the MMAs — and broadcast with shuffles in the epilogue: the eight 16-byte loads per block that stood
            // here before were still 5-6 % of the kernel's stall samples after being hoisted above the accumulator load
            uint32_t bw[4] = {0u, 0u, 0u, 0u};
            const bool bias_shfl = bias_vec && me.bias_dtype == OTB_BF16;
            if (bias_shfl) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c0 = half * 64 + k * 128;
                    if (c0 + 64 <= p.tile_n && n0 + c0 + 64 <= p.N)
                        bw[k] = __ldg(reinterpret_cast<const uint32_t*>((const uint16_t*)me.bias + n0 + c0) + lane);
                }
            }
            if (lora) {
                mbar_wait(misc(4 + as), par);                               // t_full
                tc_fence_after();
                if (warp == 2 && lane == 0) ka_stamp(q.trace, cid, leader, it, 5);      // 5: mainloop complete (t_full seen)
                for (int c = half * 16; c < p.Rp; c += 32) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + (uint32_t)(p.tile_n + c), v);
                    tmem_ld_wait();
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * me.scale;
                    if (p.dropout_p > 0.0f) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const uint64_t e = (uint64_t)m * (uint64_t)p.Rp + (uint64_t)(c + j);
                            Philox4 r = otb_rng(me.seed, me.offset, OTB_STREAM_DROPOUT, e >> 2);
                            f[j + 0] = otb_u01(r.x) >= p.dropout_p ? f[j + 0] * keep_scale : 0.0f;
                            f[j + 1] = otb_u01(r.y) >= p.dropout_p ? f[j + 1] * keep_scale : 0.0f;
                            f[j + 2] = otb_u01(r.z) >= p.dropout_p ? f[j + 2] * keep_scale : 0.0f;
                            f[j + 3] = otb_u01(r.w) >= p.dropout_p ? f[j + 3] * keep_scale : 0.0f;
                        }
                    }
                    uint4 lo = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                    uint4 hi = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
                    const int j0 = c >> 3;
                    *reinterpret_cast<uint4*>(t_row + (((j0 + 0) ^ (row & 7)) << 4)) = lo;
                    *reinterpret_cast<uint4*>(t_row + (((j0 + 1) ^ (row & 7)) << 4)) = hi;
                    if (me.t_out != nullptr && n_tile == 0 && m < p.M) {
                        uint16_t* dst = me.t_out + (size_t)m * p.R + c;
                        if (c + 8 <= p.R) *reinterpret_cast<uint4*>(dst) = lo;
                        if (c + 16 <= p.R) *reinterpret_cast<uint4*>(dst + 8) = hi;
                    }
                }
                fence_proxy_async();                                        // every lane: its T rows are visible to the tensor core
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(t_ready_leader);         // t_ready: ONE remote arrival per warp (16 per pair, not 512)
                if (warp == 2 && lane == 0) ka_stamp(q.trace, cid, leader, it, 6);      // 6: T parked
            }
            mbar_wait(misc(8 + as), par);                                   // d_full
            tc_fence_after();
            if (warp == 2 && lane == 0) ka_stamp(q.trace, cid, leader, it, 7);          // 7: accumulator complete (d_full seen)
            for (int c0 = half * 64; c0 < p.tile_n; c0 += 128) {
                const int nb0 = n0 + c0;
                if (nb0 >= p.N || (q.debug & 4)) break;
                const int cols = p.tile_n - c0 < 64 ? p.tile_n - c0 : 64;
                if (cols == 64 && nb0 + 64 <= p.N && (me.bias == nullptr || bias_vec)) {
                    // ---- full 64-column block: one swizzled box, one TMA store ----
                    uint32_t v[64];
                    const int bk = c0 >> 7;                                  // (selects, not an indexed array: keeps bw[] in registers)
                    const uint32_t my_bias = bk == 0 ? bw[0] : bk == 1 ? bw[1] : bk == 2 ? bw[2] : bw[3];   // columns 2*lane, 2*lane + 1
                    if (!(q.debug & 16)) {
                        tmem_ld32<0>(t_lane + (uint32_t)(c0 + 0), v);
                        tmem_ld32<32>(t_lane + (uint32_t)(c0 + 32), v);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 64; ++j) v[j] = 0u;
                    }
                    if (me.bias != nullptr) {
                        if (me.bias_dtype == OTB_BF16) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) {                   // lane e holds the bias of columns 2e, 2e + 1
                                const uint32_t w = __shfl_sync(0xffffffffu, my_bias, e);
                                v[2 * e + 0] = __float_as_uint(__uint_as_float(v[2 * e + 0]) + __uint_as_float(w << 16));
                                v[2 * e + 1] = __float_as_uint(__uint_as_float(v[2 * e + 1]) + __uint_as_float(w & 0xFFFF0000u));
                            }
                        } else {
                            const float4* bp = reinterpret_cast<const float4*>((const float*)me.bias + nb0);
#pragma unroll
                            for (int g = 0; g < 16; ++g) {
                                const float4 b = __ldg(bp + g);
                                v[g * 4 + 0] = __float_as_uint(__uint_as_float(v[g * 4 + 0]) + b.x);
                                v[g * 4 + 1] = __float_as_uint(__uint_as_float(v[g * 4 + 1]) + b.y);
                                v[g * 4 + 2] = __float_as_uint(__uint_as_float(v[g * 4 + 2]) + b.z);
                                v[g * 4 + 3] = __float_as_uint(__uint_as_float(v[g * 4 + 3]) + b.w);
                            }
                        }
                    }
                    if (lane == 0) bulk_wait_read<0>();                    // this warp's previous box has left smem
                    __syncwarp();
                    uint8_t* brow = ybox_gen + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)                              // 16-byte chunk j of row r lives at chunk j ^ (r & 7)
                        *reinterpret_cast<uint4*>(brow + ((j ^ (lane & 7)) << 4)) =
                            make_uint4(pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1])),
                                       pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3])),
                                       pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5])),
                                       pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7])));
                    if (!(q.debug & 32)) fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && !(q.debug & 8)) {
                        tma_store_2d(&maps.y[gm], ybox, nb0, m0 + quad * 32);
                        bulk_commit();
                    }
                } else {
                    // ---- ragged block (N edge, narrow remainder, odd bias): direct predicated stores ----
                    for (int c = c0; c < c0 + cols; c += 16) {
                        const int n = n0 + c;
                        if (n >= p.N) break;
                        uint32_t v[16];
                        // a whole 16-column chunk with an aligned bf16 bias (the 32-column remainder of a 224-wide tile): two
                        // vector loads, requested before the accumulator load — per-element loads made this path 20 % of the
                        // kernel's stall samples on the short-K leaves (profiles/r02_ncu_kac640_lines.md)
                        const bool bias16 = bias_vec && me.bias_dtype == OTB_BF16 && n + 16 <= p.N;
                        uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
                        if (bias16) {
                            const uint4* bp = reinterpret_cast<const uint4*>((const uint16_t*)me.bias + n);
                            b0 = __ldg(bp); b1 = __ldg(bp + 1);
                        }
                        tmem_ld16(t_lane + (uint32_t)c, v);
                        tmem_ld_wait();
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                        if (bias16) {
                            const uint32_t w[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                f[2 * e] += __uint_as_float(w[e] << 16);
                                f[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
                            }
                        } else if (me.bias != nullptr) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (n + j < p.N) f[j] += load_as_float(me.bias, n + j, me.bias_dtype);
                        }
                        if (m < p.M) {
                            uint16_t* dst = me.y + (size_t)m * p.N + n;
                            if (n + 8 <= p.N)
                                *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                            if (n + 16 <= p.N)
                                *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(misc(10 + as), 0));  // acc_empty: one remote arrival per warp
            if (warp == 2 && lane == 0) ka_stamp(q.trace, cid, leader, it, 8);          // 8: this warp's blocks packed and handed to TMA
        }
        if (lane == 0) bulk_wait<0>();            // every TMA store of this thread has been written before the CTA exits
        if (warp == 2 && lane == 0) ka_stamp(q.trace, cid, leader, 0, 9);               // 9: stores written
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_relaxed();                     // neither CTA exits (or frees TMEM) while the pair still works
    if (threadIdx.x == 0) ka_stamp(q.trace, cid, leader, 0, 10);                                 // 10: pair done
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc_2sm(tmem_base, 512u);
    }
}

// ------------------------------- host side ------------------------------------------
constexpr int SMEM_LIMIT = 227 * 1024;

}  // namespace

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// persistent variant of the tile-width choice: `ctas` CTAs each run ceil(tiles / ctas) tiles back to back;
// the epilogue overlaps the next mainloop, so a tile costs max(mainloop, epilogue) + the T hand-off.
int choose_tile_n_v2(int M, int N, int K, int Rp, int ctas, int cm, int cn) {
    const int m_tiles = ((M + BM - 1) / BM + cm - 1) / cm * cm;
    const int kb4 = ((K + BK - 1) / BK) * 4;
    double best = 1e30; int best_tn = 64;
    for (int tn = 32; tn + Rp <= 256; tn += 16) {
        int n_tiles = ((N + tn - 1) / tn + cn - 1) / cn * cn;
        int64_t tiles = (int64_t)m_tiles * n_tiles;
        int64_t rounds = (tiles + ctas - 1) / ctas;
        int np = tn + Rp;
        // operand bytes per k-step that this CTA pulls through L2 (x/cn + [w;down]/cm), at ~64 B/clk/SM sustained
        double l2 = (BM * 32.0 / cn + np * 32.0 / cm) / 64.0;
        double kstep = np / 2.0;
        if (kstep < 32 + np / 4.0) kstep = 32 + np / 4.0;
        if (kstep < l2) kstep = l2;
        double mainloop = kb4 * kstep;
        double epi = 14.0 * tn + 400.0;
        double tile_time = (mainloop > epi ? mainloop : epi) + 300.0;
        double t = rounds * tile_time;
        if (t < best - 1e-9) { best = t; best_tn = tn; }
    }
    return best_tn;
}

int launch_v2(const otb_lora_linear_args* a, void* stream) {
    GemmParams2 q;
    GemmParams& p = q.g;
    int cm = env_int("OTB_KA_CM", 2), cn = env_int("OTB_KA_CN", 1);
    if (cm != 1 && cm != 2) cm = 2;
    if (cn != 1 && cn != 2) cn = 1;
    p.M = a->M; p.N = a->N; p.K = a->K; p.R = a->R;
    p.Rp = (a->R + 15) / 16 * 16;
    const int sms = otb_sm_count();
    const int m_tiles = (a->M + BM - 1) / BM;
    if (m_tiles < cm) cm = 1;                                   // nothing to share along M
    p.tile_n = a->tile_n > 0 ? a->tile_n : choose_tile_n_v2(a->M, a->N, a->K, p.Rp, sms / (cm * cn) * (cm * cn), cm, cn);
    OTB_CHECK_ARG(p.tile_n % 16 == 0 && p.tile_n >= 16 && p.tile_n + p.Rp <= 256, "tile_n %d invalid (multiple of 16, tile_n + rank <= 256)", p.tile_n);
    p.num_kb = (a->K + BK - 1) / BK;
    p.n_tiles = (a->N + p.tile_n - 1) / p.tile_n;
    if (p.n_tiles < cn) cn = 1;
    q.cm = cm; q.cn = cn; q.m_tiles = m_tiles;
    q.super_m = (m_tiles + cm - 1) / cm;
    q.super_n = (p.n_tiles + cn - 1) / cn;
    p.tmem_cols = 512;
    int stages = a->stages > 0 ? a->stages : MAX_STAGES;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    while (stages > 1 && (int)make_layout2(p.tile_n, p.Rp, stages).total + 1024 > SMEM_LIMIT) --stages;
    p.stages = stages;
    const SmemLayout2 L = make_layout2(p.tile_n, p.Rp, stages);
    OTB_CHECK_ARG((int)L.total + 1024 <= SMEM_LIMIT, "tile does not fit shared memory");
    p.scale = a->scale; p.dropout_p = a->dropout_p; p.seed = a->seed; p.offset = a->offset;
    p.bias = a->bias; p.bias_dtype = a->bias_dtype;
    p.y = (uint16_t*)a->y; p.t_out = (uint16_t*)a->t_out;

    CUtensorMap tm_x, tm_w, tm_dn, tm_up;
    int rc;
    if ((rc = make_tmap(&tm_x, a->x, a->M, a->K, BM / cn, "x")) != OTB_OK) return rc;
    if ((rc = make_tmap(&tm_w, a->w, a->N, a->K, p.tile_n / cm, "w")) != OTB_OK) return rc;
    if (p.Rp > 0) {
        if ((rc = make_tmap(&tm_dn, a->down, a->R, a->K, p.Rp / cm, "down")) != OTB_OK) return rc;
        if ((rc = make_tmap(&tm_up, a->up, a->N, a->R, p.tile_n, "up")) != OTB_OK) return rc;
    } else {
        tm_dn = tm_x; tm_up = tm_x;
    }
    static thread_local int smem_set_for_dev = -1;
    int dev = 0;
    OTB_CUDA(cudaGetDevice(&dev));
    if (smem_set_for_dev != dev) {
        OTB_CUDA(cudaFuncSetAttribute(lora_linear_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        smem_set_for_dev = dev;
    }
    const int csize = cm * cn;
    const int total_super = q.super_m * q.super_n;
    int clusters = sms / csize;
    if (clusters > total_super) clusters = total_super;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * csize));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = L.total + 1024;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = otb_pdl_enabled() ? 2 : 1;
    if (csize > 1) {
        // not every SM can host a cluster member (GPC granularity): size the persistent grid to what is co-resident
        static thread_local int max_clusters[5] = {0, 0, 0, 0, 0};
        if (max_clusters[csize] == 0) {
            int n = 0;
            cudaLaunchConfig_t probe = cfg;
            probe.gridDim = dim3((unsigned)(sms / csize * csize));
            if (cudaOccupancyMaxActiveClusters(&n, lora_linear_kernel_v2, &probe) != cudaSuccess || n <= 0) n = sms / csize;
            max_clusters[csize] = n;
            (void)cudaGetLastError();
        }
        if (clusters > max_clusters[csize]) { clusters = max_clusters[csize]; cfg.gridDim = dim3((unsigned)(clusters * csize)); }
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, lora_linear_kernel_v2, tm_x, tm_w, tm_dn, tm_up, q);
    if (e != cudaSuccess) {
        otb_set_error("launch of lora_linear_kernel_v2 (cluster %dx%d, tile_n %d, %d stages) failed: %s", cm, cn, p.tile_n, stages,
                      cudaGetErrorString(e));
        return OTB_ECUDA;
    }
    return OTB_OK;
}

// ---- v3 planning -----------------------------------------------------------------------------------
// One cost model for both variants, in SM cycles.  A k-block (64 of K) costs the slower of the tensor pipe
// (N'/2 cycles per 16-wide k-step for a 128-row CTA tile, the same per CTA of a 256-row pair tile) and the
// operand ingest of ONE SM at ~42.5 B/clk (see the v3 header).  The epilogue drains ~14 cycles per output
// column and overlaps the next mainloop when there are two accumulator stages.
constexpr double INGEST_B_PER_CLK = 37.6;   // 42.5 x the ~88 % the TMA ring sustains (fit to tools/ka_sweep.py, gpurun_out/ka_sweep3.log)

struct KaPlan {
    int version;     // 2 or 3
    int tile_n, chunks, acc_stages;
    double cycles;
};

// epi_per_col: epilogue cycles per output column of the tile (v3's staged TMA store with 8 warps ~15, v2's direct stores ~45)
double plan_cost(int rounds, int num_kb, double ingest_bytes, double mma_cycles_per_kb, int tile_n, bool overlapped, double epi_per_col) {
    double kb = ingest_bytes / INGEST_B_PER_CLK;
    if (kb < mma_cycles_per_kb) kb = mma_cycles_per_kb;
    const double mainloop = num_kb * kb, epi = epi_per_col * tile_n + 500.0;
    if (overlapped) return rounds * ((mainloop > epi ? mainloop : epi) + 300.0) + epi + 3000.0;
    return rounds * (mainloop + epi + 1500.0) + 3000.0;
}

KaPlan plan_v2(int M, int N, int K, int Rp, int ctas) {
    const int m_tiles = (M + BM - 1) / BM, num_kb = (K + BK - 1) / BK;
    KaPlan best{2, 64, 1, 2, 1e30};
    for (int tn = 32; tn + Rp <= 256; tn += 16) {
        const int n_tiles = (N + tn - 1) / tn;
        const int rounds = (int)(((int64_t)m_tiles * n_tiles + ctas - 1) / ctas);
        const int np = tn + Rp;
        const double c = plan_cost(rounds, num_kb, (BM + np) * 128.0, 4.0 * (np / 2.0 > 80.0 ? np / 2.0 : 80.0), tn, true, 45.0);
        if (c < best.cycles - 1e-9) { best.cycles = c; best.tile_n = tn; }
    }
    return best;
}

// legal pair-tile widths: fused rows NB = tn + Rp split into `chunks` MMAs of NB/chunks (multiple of 16, <= 256),
// the rank-R GEMM into `up_chunks` MMAs of tn/up_chunks; the last half-chunk must be able to hold all of `down`
bool v3_shape_ok(int tn, int Rp, int* chunks, int* up_chunks, int* acc_stages) {
    const int NB = tn + Rp;
    if (tn < 16 || NB > 512) return false;
    const int ch = NB > 256 ? 2 : 1, uch = tn > 256 ? 2 : 1;
    if (NB % (16 * ch) != 0 || tn % (16 * uch) != 0) return false;
    if (Rp > NB / ch / 2) return false;
    *chunks = ch; *up_chunks = uch; *acc_stages = NB <= 256 ? 2 : 1;
    return true;
}

KaPlan plan_v3(int M, int N, int K, int Rp, int pairs, int groups = 1) {
    const int m_tiles = (M + 2 * BM - 1) / (2 * BM), num_kb = (K + BK - 1) / BK;
    KaPlan best{3, 64, 1, 2, 1e30};
    for (int tn = 32; tn + Rp <= 512; tn += 16) {
        int chunks, up_chunks, acc_stages;
        if (!v3_shape_ok(tn, Rp, &chunks, &up_chunks, &acc_stages)) continue;
        int stages = MAX_STAGES;
        while (stages > 0 && (int)make_layout3(tn, Rp, stages, acc_stages).total + 1024 > SMEM_LIMIT) --stages;
        if (stages < 3) continue;
        const int n_tiles = (N + tn - 1) / tn;
        const int rounds = (int)(((int64_t)m_tiles * n_tiles * groups + pairs - 1) / pairs);
        const int NB = tn + Rp, nbc = NB / chunks;
        // per CTA per 16-wide k-step: nbc/2 tensor-pipe cycles per chunk, with a measured floor of ~80 per instruction
        const double mma = 4.0 * chunks * (nbc / 2.0 > 80.0 ? nbc / 2.0 : 80.0);
        const double c = plan_cost(rounds, num_kb, (BM + NB / 2) * 128.0, mma, tn, acc_stages == 2, 15.0);
        if (c < best.cycles - 1e-9) { best = KaPlan{3, tn, chunks, acc_stages, c}; }
    }
    return best;
}

// `count` same-shape problems (count == 1: the plain call) as one v3 launch
template <int G>
int launch_v3(const otb_lora_linear_args* a, int count, void* stream, int tile_n, int max_pairs) {
    static thread_local GemmParams3<G> q;          // the wide instantiation is 27 KB with its maps: keep both off the stack
    static thread_local V3Maps<G> maps;
    GemmParams& p = q.g;
    p.M = a->M; p.N = a->N; p.K = a->K; p.R = a->R;
    p.Rp = (a->R + 15) / 16 * 16;
    p.tile_n = tile_n;
    q.groups = count;
    OTB_CHECK_ARG(v3_shape_ok(tile_n, p.Rp, &q.chunks, &q.up_chunks, &q.acc_stages),
                  "pair tile_n %d invalid for rank %d (tile_n + rank a multiple of 16 (32 above 256), <= 512)", tile_n, p.Rp);
    q.debug = env_int("OTB_KA_DEBUG", 0);
    q.lookahead = env_int("OTB_KA_LOOKAHEAD", 2);
    q.tail_reorder = env_int("OTB_KA_TAIL", 1);
    {
        const char* tr = getenv("OTB_KA_TRACE");
        q.trace = tr ? (unsigned long long*)strtoull(tr, nullptr, 0) : nullptr;
    }
    if (q.lookahead < 0) q.lookahead = 0;
    q.prefetch_w = 1;
    for (int g = 0; g < count; ++g)
        if (!(a[g].flags & OTB_LINEAR_STATIC_WEIGHTS)) q.prefetch_w = 0;
    // Off unless asked for: measured on the B200 over one SDXL step's 894 launches, fetching the weights ahead of the wait
    // LOST 0.9 ms (37.2 vs 36.3 ms) — the x tile of the first k-block then queues behind up to `stages` weight boxes in the
    // TMA unit, and the early fetches compete with the predecessor's epilogue stores (profiles/README.md, round 2).
    if (!otb_pdl_enabled() || env_int("OTB_KA_PREFETCH_W", 0) == 0) q.prefetch_w = 0;
    else if (q.prefetch_w) q.prefetch_w = env_int("OTB_KA_PREFETCH_W", 0) == 2 ? 2 : 1;      // 2: into L2 only
    p.num_kb = (a->K + BK - 1) / BK;
    p.n_tiles = (a->N + p.tile_n - 1) / p.tile_n;
    q.pair_m_tiles = (a->M + 2 * BM - 1) / (2 * BM);
    p.tmem_cols = 512;
    int stages = a->stages > 0 ? a->stages : MAX_STAGES;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    while (stages > 1 && (int)make_layout3(p.tile_n, p.Rp, stages, q.acc_stages).total + 1024 > SMEM_LIMIT) --stages;
    p.stages = stages;
    const SmemLayout3 L = make_layout3(p.tile_n, p.Rp, stages, q.acc_stages);
    OTB_CHECK_ARG((int)L.total + 1024 <= SMEM_LIMIT, "pair tile does not fit shared memory");
    p.scale = a->scale; p.dropout_p = a->dropout_p; p.seed = a->seed; p.offset = a->offset;
    p.bias = a->bias; p.bias_dtype = a->bias_dtype;
    p.y = (uint16_t*)a->y; p.t_out = (uint16_t*)a->t_out;

    const int nbh = (p.tile_n + p.Rp) / q.chunks / 2, wch = p.tile_n / q.up_chunks / 2;
    int rc;
    for (int g = 0; g < G; ++g) {
        const otb_lora_linear_args* m = &a[g < count ? g : 0];         // unused slots mirror member 0
        q.mem[g] = V3Member{m->bias, (uint16_t*)m->y, (uint16_t*)m->t_out, m->seed, m->offset, m->scale, m->bias_dtype};
        if (g >= count) {
            maps.x[g] = maps.x[0]; maps.w[g] = maps.w[0]; maps.w2[g] = maps.w2[0]; maps.dn[g] = maps.dn[0]; maps.up[g] = maps.up[0]; maps.y[g] = maps.y[0];
            continue;
        }
        if ((rc = make_tmap(&maps.x[g], m->x, m->M, m->K, BM, "x")) != OTB_OK) return rc;
        if ((rc = make_tmap(&maps.w[g], m->w, m->N, m->K, nbh, "w")) != OTB_OK) return rc;
        if ((rc = make_tmap(&maps.y[g], m->y, m->M, m->N, 32, "y")) != OTB_OK) return rc;
        maps.w2[g] = maps.w[g]; maps.dn[g] = maps.x[g]; maps.up[g] = maps.x[g];
        if (p.Rp > 0) {
            if (nbh > p.Rp && (rc = make_tmap(&maps.w2[g], m->w, m->N, m->K, nbh - p.Rp, "w (last half)")) != OTB_OK) return rc;
            if ((rc = make_tmap(&maps.dn[g], m->down, m->R, m->K, p.Rp, "down")) != OTB_OK) return rc;
            if ((rc = make_tmap(&maps.up[g], m->up, m->N, m->R, wch, "up")) != OTB_OK) return rc;
        }
    }
    const int total_tiles = q.pair_m_tiles * p.n_tiles * count;
    int clusters = max_pairs < total_tiles ? max_pairs : total_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * 2));
    cfg.blockDim = dim3(V3_THREADS);
    cfg.dynamicSmemBytes = L.total + 1024;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = otb_pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, lora_linear_kernel_v3<G>, maps, q);
    if (e != cudaSuccess) {
        otb_set_error("launch of lora_linear_kernel_v3 (%d x pair tile 256x%d, %d chunks, %d stages) failed: %s", count, p.tile_n, q.chunks,
                      stages, cudaGetErrorString(e));
        return OTB_ECUDA;
    }
    return OTB_OK;
}

// how many CTA pairs can be co-resident (GPC granularity may strand a few SMs); cached per device
int v3_max_pairs(int sms) {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return sms / 2;
    if (cached_dev == dev) return cached;
    (void)cudaFuncSetAttribute(lora_linear_kernel_v3<V3_MAX_GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    (void)cudaFuncSetAttribute(lora_linear_kernel_v3<V3_BIG_GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    cudaLaunchConfig_t probe = {};
    probe.gridDim = dim3((unsigned)(sms / 2 * 2));
    probe.blockDim = dim3(V3_THREADS);
    probe.dynamicSmemBytes = SMEM_LIMIT;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    probe.attrs = attr; probe.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lora_linear_kernel_v3<V3_MAX_GROUP>, &probe) != cudaSuccess || n <= 0) n = sms / 2;
    (void)cudaGetLastError();
    cached_dev = dev; cached = n;
    return n;
}

static int check_lora_linear_args(const otb_lora_linear_args* a) {
    OTB_CHECK_ARG(a != nullptr, "null args");
    OTB_CHECK_ARG(a->x && a->w && a->y, "null x / w / y");
    OTB_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
    OTB_CHECK_ARG(a->R >= 0, "negative rank");
    OTB_UNSUPPORTED(a->K % 8 != 0 || a->N % 8 != 0, "K (%d) and N (%d) must be multiples of 8 (16-byte TMA strides / vector stores)", a->K, a->N);
    OTB_UNSUPPORTED(a->R % 8 != 0 || a->R > 64, "LoRA rank %d unsupported: must be a multiple of 8 and <= 64", a->R);
    OTB_CHECK_ARG(a->R == 0 || (a->down && a->up), "rank > 0 needs down / up");
    OTB_CHECK_ARG(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "dropout_p must be in [0,1)");
    OTB_CHECK_ARG((((uintptr_t)a->x | (uintptr_t)a->w | (uintptr_t)a->y | (uintptr_t)a->down | (uintptr_t)a->up |
                    (uintptr_t)a->t_out) & 15) == 0, "all tensor pointers must be 16-byte aligned");
    return OTB_OK;
}

extern "C" int otb_lora_linear(const otb_lora_linear_args* a, void* stream) {
    int rc = check_lora_linear_args(a);
    if (rc != OTB_OK) return rc;
    OTB_BIND(a->x);
    {
        // variant choice: the cost model picks between single-CTA tiles (v2) and CTA-pair tiles (v3);
        // OTB_KA_VER=2|3 forces one (A/B runs), an explicit tile_n > 256 needs v3
        const int sms = otb_sm_count();
        const int Rp = (a->R + 15) / 16 * 16;
        const int force = env_int("OTB_KA_VER", 0);
        const int pairs = v3_max_pairs(sms);
        KaPlan p3 = plan_v3(a->M, a->N, a->K, Rp, pairs);
        int ch, uch, as;
        const bool explicit3 = a->tile_n > 0 && v3_shape_ok(a->tile_n, Rp, &ch, &uch, &as);
        bool use3;
        if (force == 3 || a->tile_n > 256) use3 = true;
        else if (force == 2 || a->M <= BM || (a->tile_n > 0 && !explicit3)) use3 = false;
        else use3 = p3.cycles < plan_v2(a->M, a->N, a->K, Rp, sms).cycles;
        if (use3) return launch_v3<V3_MAX_GROUP>(a, 1, stream, a->tile_n > 0 ? a->tile_n : p3.tile_n, pairs);
        return launch_v2(a, stream);
    }

}

extern "C" int otb_lora_linear_grouped(const otb_lora_linear_args* members, int32_t count, void* stream) {
    OTB_CHECK_ARG(members != nullptr && count >= 1 && count <= V3_BIG_GROUP, "group of %d (1..%d members)", count, V3_BIG_GROUP);
    for (int g = 0; g < count; ++g) {
        int rc = check_lora_linear_args(&members[g]);
        if (rc != OTB_OK) return rc;
        const otb_lora_linear_args& a = members[g], &b = members[0];
        OTB_CHECK_ARG(a.M == b.M && a.N == b.N && a.K == b.K && a.R == b.R && a.dropout_p == b.dropout_p,
                      "members of a group must share M, N, K, rank and dropout_p");
    }
    if (count == 1) return otb_lora_linear(members, stream);
    OTB_BIND(members[0].x);
    const int Rp = (members[0].R + 15) / 16 * 16;
    const int pairs = v3_max_pairs(otb_sm_count());
    const KaPlan plan = plan_v3(members[0].M, members[0].N, members[0].K, Rp, pairs, count);
    const int tn = members[0].tile_n > 0 ? members[0].tile_n : plan.tile_n;
    if (count <= V3_MAX_GROUP) return launch_v3<V3_MAX_GROUP>(members, count, stream, tn, pairs);
    return launch_v3<V3_BIG_GROUP>(members, count, stream, tn, pairs);
}
