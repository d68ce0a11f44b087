// exchange.cu — K-r: the data-parallel gradient exchange fused with the global gradient norm, over NVLink peer memory.
//
// The path has exactly one exchange step: the sum of the flat fp32 gradient arena over ranks, followed
// by the global-norm clip (GenericTrainer.py:712-713) and the optimizer.  The arena lives in symmetric memory: every
// replica is mapped into every process (NVLink loads / stores) and all of them together as ONE multicast object
// (NVSwitch NVLS).  Rank r owns the r-th slice of the arena and, in one kernel:
//     v   = sum over ranks of replica[slice]     P2P: one 16-byte load per peer, added in rank order
//                                                NVLS: multimem.ld_reduce.add, the switch sums on the way in
//     acc += (prescale * v)^2                    the squared norm rides along: no second pass over the gradients
//     replica[slice] = v on every rank           P2P: one store per peer;  NVLS: multimem.st, the switch replicates
// i.e. a reduce-scatter and an all-gather in one launch, one rounding sequence per element, every replica
// bit-identical; the per-block partial norms land in every rank's partial table the same way.  Two peer barriers (flag
// words in peer-mapped memory) bracket it: all ranks' backward passes have written their gradients / all slices have
// landed everywhere.  No NCCL call, no staging copy, no host sync.
// Per-GPU NVLink traffic for an arena of S bytes on n ranks: P2P (n-1)/n * S each way in n-1 request streams; NVLS
// (1 + 1/n) * S each way (the requester's own replica also travels to the switch and its own store comes back) in
// one.  Measured, S = 340 MB, against ncclAllReduce of the same buffer (which does not compute the norm):
//     2 ranks  P2P 524 us   NVLS 889 us   NCCL 650 us
//     4 ranks  P2P 806 us   NVLS 780 us   NCCL 855 us
//     8 ranks  P2P 961 us   NVLS 757 us   NCCL 885 us        -> host default: P2P on 2 ranks, NVLS from 4.
#include "common.cuh"

int env_int(const char* name, int dflt);      // lora_gemm.cu

namespace {

constexpr int EX_THREADS = 256;
constexpr int EX_UNROLL_DEFAULT = 4;
constexpr long long PEER_WAIT_LIMIT_CYCLES = 120000000000ll;   // ~60 s: a missing rank traps instead of hanging the GPU

struct PeerFlags {
    uint32_t* ptr[OTB_MAX_RANKS];
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st(float* mc, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(mc), "f"(v) : "memory");
}

__device__ __forceinline__ float4 ld_sys(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(float* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_sys(float* p, float v) {
    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

struct PeerPtrs {
    float* grads[OTB_MAX_RANKS];
    float* partials[OTB_MAX_RANKS];
};

__device__ __forceinline__ void spin_until(uint32_t* addr, uint32_t expect, uint32_t desired) {
    long long t0 = 0;
    uint32_t spins = 0;
    while (atomicCAS_system(addr, expect, desired) != expect) {
        if ((++spins & 255u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > PEER_WAIT_LIMIT_CYCLES) __trap();
        }
    }
}

// one thread per peer: raise my flag in the peer's table, then consume the peer's flag in mine
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerFlags f, int rank, int world, int phase) {
    const int peer = threadIdx.x;
    if (peer >= world) return;
    __threadfence_system();
    spin_until(f.ptr[peer] + phase * OTB_MAX_RANKS + rank, 0u, 1u);
    spin_until(f.ptr[rank] + phase * OTB_MAX_RANKS + peer, 1u, 0u);
    __threadfence_system();
}

template <int EX_UNROLL>
__global__ void __launch_bounds__(EX_THREADS) grad_exchange_nvls_kernel(float* __restrict__ mc, int64_t v_begin, int64_t v_end,
                                                                        float prescale, float* __restrict__ mc_partials, int rank) {
    __shared__ float sm[EX_THREADS / 32];
    float acc = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * EX_THREADS;
    int64_t i = v_begin + (int64_t)blockIdx.x * EX_THREADS + threadIdx.x;
    for (; i + (EX_UNROLL - 1) * stride < v_end; i += EX_UNROLL * stride) {
        float4 v[EX_UNROLL];
#pragma unroll
        for (int u = 0; u < EX_UNROLL; ++u) v[u] = multimem_ld_reduce_add(mc + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < EX_UNROLL; ++u) {
            const float a = v[u].x * prescale, b = v[u].y * prescale, c = v[u].z * prescale, d = v[u].w * prescale;
            acc += a * a + b * b + c * c + d * d;
            multimem_st(mc + 4 * (i + u * stride), v[u]);
        }
    }
    for (; i < v_end; i += stride) {
        const float4 v = multimem_ld_reduce_add(mc + 4 * i);
        const float a = v.x * prescale, b = v.y * prescale, c = v.z * prescale, d = v.w * prescale;
        acc += a * a + b * b + c * c + d * d;
        multimem_st(mc + 4 * i, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = 0.0f;
#pragma unroll
        for (int w = 0; w < EX_THREADS / 32; ++w) r += sm[w];
        multimem_st(mc_partials + (int64_t)rank * gridDim.x + blockIdx.x, r);     // into every rank's partial table
    }
    __threadfence_system();
}

template <int WORLD, int EX_UNROLL>
__global__ void __launch_bounds__(EX_THREADS) grad_exchange_p2p_kernel(PeerPtrs pp, int64_t v_begin, int64_t v_end, float prescale, int rank) {
    __shared__ float sm[EX_THREADS / 32];
    float acc = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * EX_THREADS;
    auto reduce_store = [&](const float4 (&in)[WORLD]) {
        float4 v = in[0];                                   // fixed rank order: the sum does not depend on who reduces
#pragma unroll
        for (int r = 1; r < WORLD; ++r) { v.x += in[r].x; v.y += in[r].y; v.z += in[r].z; v.w += in[r].w; }
        const float a = v.x * prescale, b = v.y * prescale, c = v.z * prescale, d = v.w * prescale;
        acc += a * a + b * b + c * c + d * d;
        return v;
    };
    int64_t i = v_begin + (int64_t)blockIdx.x * EX_THREADS + threadIdx.x;
    for (; i + (EX_UNROLL - 1) * stride < v_end; i += EX_UNROLL * stride) {
        float4 in[EX_UNROLL][WORLD];
#pragma unroll
        for (int u = 0; u < EX_UNROLL; ++u)
#pragma unroll
            for (int r = 0; r < WORLD; ++r) in[u][r] = ld_sys(pp.grads[r] + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < EX_UNROLL; ++u) {
            const float4 v = reduce_store(in[u]);
#pragma unroll
            for (int r = 0; r < WORLD; ++r) st_sys(pp.grads[r] + 4 * (i + u * stride), v);
        }
    }
    for (; i < v_end; i += stride) {
        float4 in[WORLD];
#pragma unroll
        for (int r = 0; r < WORLD; ++r) in[r] = ld_sys(pp.grads[r] + 4 * i);
        const float4 v = reduce_store(in);
#pragma unroll
        for (int r = 0; r < WORLD; ++r) st_sys(pp.grads[r] + 4 * i, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < EX_THREADS / 32; ++w) t += sm[w];
#pragma unroll
        for (int r = 0; r < WORLD; ++r) st_sys(pp.partials[r] + (int64_t)rank * gridDim.x + blockIdx.x, t);
    }
    __threadfence_system();
}

template <int WORLD>
void launch_p2p(const otb_grad_exchange_args* a, const PeerPtrs& pp, int64_t v0, int64_t v1, cudaStream_t s) {
    // registers: WORLD * unroll float4 in flight per thread
    if (WORLD <= 2) grad_exchange_p2p_kernel<WORLD, 4><<<a->blocks, EX_THREADS, 0, s>>>(pp, v0, v1, a->grad_prescale, a->rank);
    else if (WORLD <= 4) grad_exchange_p2p_kernel<WORLD, 2><<<a->blocks, EX_THREADS, 0, s>>>(pp, v0, v1, a->grad_prescale, a->rank);
    else grad_exchange_p2p_kernel<WORLD, 1><<<a->blocks, EX_THREADS, 0, s>>>(pp, v0, v1, a->grad_prescale, a->rank);
}


// =====================================================================================================================
// K-z: the sharded form of the exchange (reduce-scatter -> sharded K-c -> all-gather).
// Full fine-tuning (config C4) moves 4.4 GB of bf16 gradients; summing them into EVERY replica (K-r, ncclAllReduce) only
// to run the same 2.2 B-element AdamW + EMA n times is wasted fabric and HBM traffic.  Here the flat gradient arena
// and the flat parameter arena are both symmetric memory, every contiguous tensor-aligned range ("bucket") has ONE
// owner rank, and
//   shard_reduce   the owner sums the range over all replicas (P2P loads in rank order with fp32 accumulation, or
//                  multimem.ld_reduce...acc::f32 inside the switch), rounds once, keeps the sum in ITS replica only and
//                  publishes the block partials of (prescale * sum)^2 to every rank: each rank can finish the global
//                  norm (otb_grad_norm_from_partials) without a collective;
//   [K-c]          runs on the owner's tensors only: 1/n of the optimizer state, EMA and HBM traffic per rank;
//   shard_gather   the owner pushes its updated parameter range into every replica (P2P stores, or ONE multimem.st
//                  stream the switch replicates).
// Hand-shakes are flag words in peer-mapped memory as in K-r.  `signal` / `wait` (stage 2) let a rank tell a bucket's
// owner "my backward has produced this bucket", so the reduce overlaps the rest of backward bucket by bucket.
struct ShardPeers {
    char* base[OTB_MAX_RANKS];
    float* partials[OTB_MAX_RANKS];
};

__device__ __forceinline__ uint4 ld_sys_u4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u4(void* p, uint4 v) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc) {
    uint4 v;           // eight bf16 sums, accumulated in fp32 inside the switch, rounded once
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_u4(void* mc, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t bf16_pack_rn(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// squared norm of one 16-byte vector of the (already rounded) sum
template <bool BF16>
__device__ __forceinline__ float sq_of(uint4 v, float prescale) {
    float acc = 0.0f;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (BF16) {
            const float a = bf16_lo(w[k]) * prescale, b = bf16_hi(w[k]) * prescale;
            acc = fmaf(a, a, acc); acc = fmaf(b, b, acc);
        } else {
            const float a = __uint_as_float(w[k]) * prescale;
            acc = fmaf(a, a, acc);
        }
    }
    return acc;
}

template <bool BF16, int WORLD>
__device__ __forceinline__ uint4 sum_in_rank_order(const uint4 (&in)[WORLD]) {
    uint4 out;
    uint32_t* o = &out.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (BF16) {
            float lo = 0.0f, hi = 0.0f;
#pragma unroll
            for (int r = 0; r < WORLD; ++r) { const uint32_t w = (&in[r].x)[k]; lo += bf16_lo(w); hi += bf16_hi(w); }
            o[k] = bf16_pack_rn(lo, hi);
        } else {
            float a = __uint_as_float((&in[0].x)[k]);
#pragma unroll
            for (int r = 1; r < WORLD; ++r) a += __uint_as_float((&in[r].x)[k]);
            o[k] = __float_as_uint(a);
        }
    }
    return out;
}

__device__ __forceinline__ float block_total(float acc, float* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < EX_THREADS / 32; ++w) t += sm[w];
    }
    return t;
}

// v_begin / v_end: 16-byte vector indices into the arena.  The sum stays in the OWNER's replica (mine).
template <bool BF16, int WORLD, int UNROLL>
__global__ void __launch_bounds__(EX_THREADS) shard_reduce_p2p_kernel(ShardPeers pp, int64_t v_begin, int64_t v_end, float prescale, int rank,
                                                                       int partial_slot) {
    __shared__ float sm[EX_THREADS / 32];
    float acc = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * EX_THREADS;
    int64_t i = v_begin + (int64_t)blockIdx.x * EX_THREADS + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < v_end; i += UNROLL * stride) {
        uint4 in[UNROLL][WORLD];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int r = 0; r < WORLD; ++r) in[u][r] = ld_sys_u4(pp.base[r] + 16 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint4 v = sum_in_rank_order<BF16, WORLD>(in[u]);
            acc += sq_of<BF16>(v, prescale);
            *reinterpret_cast<uint4*>(pp.base[rank] + 16 * (i + u * stride)) = v;
        }
    }
    for (; i < v_end; i += stride) {
        uint4 in[WORLD];
#pragma unroll
        for (int r = 0; r < WORLD; ++r) in[r] = ld_sys_u4(pp.base[r] + 16 * i);
        const uint4 v = sum_in_rank_order<BF16, WORLD>(in);
        acc += sq_of<BF16>(v, prescale);
        *reinterpret_cast<uint4*>(pp.base[rank] + 16 * i) = v;
    }
    const float t = block_total(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int r = 0; r < WORLD; ++r) st_sys(pp.partials[r] + partial_slot + blockIdx.x, t);
    }
    __threadfence_system();
}

template <bool BF16, int UNROLL>
__global__ void __launch_bounds__(EX_THREADS) shard_reduce_nvls_kernel(const char* __restrict__ mc, char* __restrict__ mine, int64_t v_begin,
                                                                        int64_t v_end, float prescale, float* __restrict__ mc_partials,
                                                                        int partial_slot) {
    __shared__ float sm[EX_THREADS / 32];
    float acc = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * EX_THREADS;
    auto load = [&](int64_t v) -> uint4 {
        if (BF16) return multimem_ld_reduce_bf16x8(mc + 16 * v);
        const float4 f = multimem_ld_reduce_add(reinterpret_cast<const float*>(mc + 16 * v));
        return make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w));
    };
    int64_t i = v_begin + (int64_t)blockIdx.x * EX_THREADS + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < v_end; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = load(i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            acc += sq_of<BF16>(v[u], prescale);
            *reinterpret_cast<uint4*>(mine + 16 * (i + u * stride)) = v[u];
        }
    }
    for (; i < v_end; i += stride) {
        const uint4 v = load(i);
        acc += sq_of<BF16>(v, prescale);
        *reinterpret_cast<uint4*>(mine + 16 * i) = v;
    }
    const float t = block_total(acc, sm);
    if (threadIdx.x == 0) multimem_st(mc_partials + partial_slot + blockIdx.x, t);
    __threadfence_system();
}

// push MY replica's range into every replica
template <int WORLD, int UNROLL>
__global__ void __launch_bounds__(EX_THREADS) shard_gather_p2p_kernel(ShardPeers pp, int64_t v_begin, int64_t v_end, int rank) {
    const int64_t stride = (int64_t)gridDim.x * EX_THREADS;
    int64_t i = v_begin + (int64_t)blockIdx.x * EX_THREADS + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < v_end; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = *reinterpret_cast<const uint4*>(pp.base[rank] + 16 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int r = 0; r < WORLD; ++r)
                if (r != rank) st_sys_u4(pp.base[r] + 16 * (i + u * stride), v[u]);
    }
    for (; i < v_end; i += stride) {
        const uint4 v = *reinterpret_cast<const uint4*>(pp.base[rank] + 16 * i);
#pragma unroll
        for (int r = 0; r < WORLD; ++r)
            if (r != rank) st_sys_u4(pp.base[r] + 16 * i, v);
    }
    __threadfence_system();
}

template <int UNROLL>
__global__ void __launch_bounds__(EX_THREADS) shard_gather_nvls_kernel(char* __restrict__ mc, const char* __restrict__ mine, int64_t v_begin, int64_t v_end) {
    const int64_t stride = (int64_t)gridDim.x * EX_THREADS;
    int64_t i = v_begin + (int64_t)blockIdx.x * EX_THREADS + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < v_end; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = *reinterpret_cast<const uint4*>(mine + 16 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) multimem_st_u4(mc + 16 * (i + u * stride), v[u]);
    }
    for (; i < v_end; i += stride) multimem_st_u4(mc + 16 * i, *reinterpret_cast<const uint4*>(mine + 16 * i));
    __threadfence_system();
}

// one-directional hand-shake of stage 2: every non-owner raises its word in the OWNER's table (waiting until last
// step's has been consumed); the owner consumes one word per peer
__global__ void __launch_bounds__(32) peer_signal_kernel(uint32_t* owner_word) {
    if (threadIdx.x != 0) return;
    __threadfence_system();
    spin_until(owner_word, 0u, 1u);
}
__global__ void __launch_bounds__(32) peer_wait_kernel(uint32_t* my_words, int rank, int world) {
    const int peer = threadIdx.x;
    if (peer >= world || peer == rank) return;
    spin_until(my_words + peer, 1u, 0u);
    __threadfence_system();
}

template <bool BF16, int WORLD>
void launch_shard_reduce_p2p(const otb_shard_exchange_args* a, const ShardPeers& pp, int64_t v0, int64_t v1, cudaStream_t s) {
    if (WORLD <= 2) shard_reduce_p2p_kernel<BF16, WORLD, 4><<<a->blocks, EX_THREADS, 0, s>>>(pp, v0, v1, a->grad_prescale, a->rank, a->partial_slot);
    else if (WORLD <= 4) shard_reduce_p2p_kernel<BF16, WORLD, 2><<<a->blocks, EX_THREADS, 0, s>>>(pp, v0, v1, a->grad_prescale, a->rank, a->partial_slot);
    else shard_reduce_p2p_kernel<BF16, WORLD, 1><<<a->blocks, EX_THREADS, 0, s>>>(pp, v0, v1, a->grad_prescale, a->rank, a->partial_slot);
}

int check_shard_args(const otb_shard_exchange_args* a, bool need_partials) {
    OTB_CHECK_ARG(a != nullptr, "null args");
    OTB_CHECK_ARG(a->world >= 2 && a->world <= OTB_MAX_RANKS && a->rank >= 0 && a->rank < a->world, "bad rank %d of %d", a->rank, a->world);
    OTB_CHECK_ARG(a->mode == OTB_EXCHANGE_P2P || a->mode == OTB_EXCHANGE_NVLS, "unknown exchange mode %d", a->mode);
    OTB_CHECK_ARG(a->dtype == OTB_BF16 || a->dtype == OTB_F32, "arena dtype must be bf16 or fp32");
    const int64_t per_vec = a->dtype == OTB_BF16 ? 8 : 4;
    OTB_CHECK_ARG(a->begin >= 0 && a->end >= a->begin && a->begin % per_vec == 0 && a->end % per_vec == 0,
                  "range [%lld, %lld) must be whole 16-byte vectors", (long long)a->begin, (long long)a->end);
    OTB_CHECK_ARG(a->blocks >= 1 && a->blocks <= 4096, "blocks %d out of range", a->blocks);
    for (int r = 0; r < a->world; ++r) {
        OTB_CHECK_ARG(a->peer_base[r] != nullptr && (((uintptr_t)a->peer_base[r]) & 15) == 0, "arena of rank %d missing or not 16-byte aligned", r);
        OTB_CHECK_ARG(!need_partials || a->mode == OTB_EXCHANGE_NVLS || a->peer_partials[r] != nullptr, "partial table of rank %d missing", r);
    }
    OTB_CHECK_ARG(a->mode != OTB_EXCHANGE_NVLS || (a->mc_base != nullptr && (!need_partials || a->mc_partials != nullptr)),
                  "NVLS mode needs the multicast addresses of the arena (and of the partial table)");
    OTB_UNSUPPORTED(a->mode == OTB_EXCHANGE_P2P && a->world != 2 && a->world != 4 && a->world != 8,
                    "P2P mode is instantiated for 2, 4 and 8 ranks (got %d); use OTB_EXCHANGE_NVLS", a->world);
    return OTB_OK;
}

}  // namespace

extern "C" int otb_shard_reduce(const otb_shard_exchange_args* a, void* stream) {
    int rc = check_shard_args(a, true);
    if (rc != OTB_OK) return rc;
    OTB_BIND(a->peer_base[a->rank]);
    cudaStream_t s = (cudaStream_t)stream;
    if (a->end == a->begin) return OTB_OK;
    const bool bf16 = a->dtype == OTB_BF16;
    const int64_t per_vec = bf16 ? 8 : 4, v0 = a->begin / per_vec, v1 = a->end / per_vec;
    ShardPeers pp;
    for (int r = 0; r < a->world; ++r) { pp.base[r] = (char*)a->peer_base[r]; pp.partials[r] = a->peer_partials[r]; }
    if (a->mode == OTB_EXCHANGE_NVLS) {
        if (bf16) shard_reduce_nvls_kernel<true, 4><<<a->blocks, EX_THREADS, 0, s>>>((const char*)a->mc_base, pp.base[a->rank], v0, v1, a->grad_prescale, a->mc_partials, a->partial_slot);
        else shard_reduce_nvls_kernel<false, 4><<<a->blocks, EX_THREADS, 0, s>>>((const char*)a->mc_base, pp.base[a->rank], v0, v1, a->grad_prescale, a->mc_partials, a->partial_slot);
        OTB_LAUNCH_CHECK("shard_reduce_nvls_kernel");
    } else {
        if (bf16) {
            if (a->world == 2) launch_shard_reduce_p2p<true, 2>(a, pp, v0, v1, s);
            else if (a->world == 4) launch_shard_reduce_p2p<true, 4>(a, pp, v0, v1, s);
            else launch_shard_reduce_p2p<true, 8>(a, pp, v0, v1, s);
        } else {
            if (a->world == 2) launch_shard_reduce_p2p<false, 2>(a, pp, v0, v1, s);
            else if (a->world == 4) launch_shard_reduce_p2p<false, 4>(a, pp, v0, v1, s);
            else launch_shard_reduce_p2p<false, 8>(a, pp, v0, v1, s);
        }
        OTB_LAUNCH_CHECK("shard_reduce_p2p_kernel");
    }
    return OTB_OK;
}

extern "C" int otb_shard_gather(const otb_shard_exchange_args* a, void* stream) {
    int rc = check_shard_args(a, false);
    if (rc != OTB_OK) return rc;
    OTB_BIND(a->peer_base[a->rank]);
    cudaStream_t s = (cudaStream_t)stream;
    if (a->end == a->begin) return OTB_OK;
    const int64_t per_vec = a->dtype == OTB_BF16 ? 8 : 4, v0 = a->begin / per_vec, v1 = a->end / per_vec;
    ShardPeers pp;
    for (int r = 0; r < a->world; ++r) { pp.base[r] = (char*)a->peer_base[r]; pp.partials[r] = nullptr; }
    if (a->mode == OTB_EXCHANGE_NVLS) {
        shard_gather_nvls_kernel<4><<<a->blocks, EX_THREADS, 0, s>>>((char*)a->mc_base, pp.base[a->rank], v0, v1);
        OTB_LAUNCH_CHECK("shard_gather_nvls_kernel");
    } else {
        if (a->world == 2) shard_gather_p2p_kernel<2, 4><<<a->blocks, EX_THREADS, 0, s>>>(pp, v0, v1, a->rank);
        else if (a->world == 4) shard_gather_p2p_kernel<4, 4><<<a->blocks, EX_THREADS, 0, s>>>(pp, v0, v1, a->rank);
        else shard_gather_p2p_kernel<8, 2><<<a->blocks, EX_THREADS, 0, s>>>(pp, v0, v1, a->rank);
        OTB_LAUNCH_CHECK("shard_gather_p2p_kernel");
    }
    return OTB_OK;
}

extern "C" int otb_peer_barrier(uint32_t* const* peer_flags, int32_t rank, int32_t world, int32_t phase, void* stream) {
    OTB_CHECK_ARG(peer_flags != nullptr && world >= 2 && world <= OTB_MAX_RANKS && rank >= 0 && rank < world, "bad rank %d of %d", rank, world);
    OTB_CHECK_ARG(phase >= 0 && phase < OTB_PEER_PHASES, "phase %d outside 0..%d", phase, OTB_PEER_PHASES - 1);
    PeerFlags f;
    for (int r = 0; r < world; ++r) {
        OTB_CHECK_ARG(peer_flags[r] != nullptr, "null flag table of rank %d", r);
        f.ptr[r] = peer_flags[r];
    }
    OTB_BIND(peer_flags[rank]);
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, world, phase);
    OTB_LAUNCH_CHECK("peer_barrier_kernel");
    return OTB_OK;
}

extern "C" int otb_peer_signal(uint32_t* owner_word, void* stream) {
    OTB_CHECK_ARG(owner_word != nullptr, "null flag word");
    OTB_BIND(owner_word);
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(owner_word);
    OTB_LAUNCH_CHECK("peer_signal_kernel");
    return OTB_OK;
}

extern "C" int otb_peer_wait(uint32_t* my_words, int32_t rank, int32_t world, void* stream) {
    OTB_CHECK_ARG(my_words != nullptr && world >= 2 && world <= OTB_MAX_RANKS && rank >= 0 && rank < world, "bad rank %d of %d", rank, world);
    OTB_BIND(my_words);
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(my_words, rank, world);
    OTB_LAUNCH_CHECK("peer_wait_kernel");
    return OTB_OK;
}

extern "C" int otb_grad_exchange_blocks(int64_t n_elems, int32_t world) {
    if (n_elems <= 0 || world <= 0) return 0;
    const int64_t per = ((n_elems / 4) + world - 1) / world;
    const int64_t want = (per + (int64_t)EX_THREADS * EX_UNROLL_DEFAULT - 1) / ((int64_t)EX_THREADS * EX_UNROLL_DEFAULT);
    // measured on 2 B200s (profiles/r01d_logs/nvls_sweep3.log): flat from 1 block per SM upwards, unroll 2..8 alike
    const int64_t cap = env_int("OTB_KR_BLOCKS", otb_sm_count() * 2);      // every rank must agree
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

extern "C" int otb_grad_exchange(const otb_grad_exchange_args* a, void* stream) {
    OTB_CHECK_ARG(a != nullptr, "null args");
    OTB_CHECK_ARG(a->world >= 2 && a->world <= OTB_MAX_RANKS && a->rank >= 0 && a->rank < a->world, "bad rank %d of %d", a->rank, a->world);
    OTB_CHECK_ARG(a->mode == OTB_EXCHANGE_P2P || a->mode == OTB_EXCHANGE_NVLS, "unknown exchange mode %d", a->mode);
    OTB_CHECK_ARG(a->n_elems > 0 && a->n_elems % 4 == 0, "arena must be a multiple of 4 floats");
    OTB_CHECK_ARG(a->blocks == otb_grad_exchange_blocks(a->n_elems, a->world), "blocks must be otb_grad_exchange_blocks(n_elems, world)");
    PeerFlags f;
    PeerPtrs pp;
    for (int r = 0; r < a->world; ++r) {
        OTB_CHECK_ARG(a->peer_flags[r] != nullptr, "null flag table of rank %d", r);
        f.ptr[r] = a->peer_flags[r];
        pp.grads[r] = a->peer_grads[r];
        pp.partials[r] = a->peer_partials[r];
        OTB_CHECK_ARG(a->mode != OTB_EXCHANGE_P2P || (pp.grads[r] != nullptr && pp.partials[r] != nullptr && (((uintptr_t)pp.grads[r]) & 15) == 0),
                      "P2P exchange needs the 16-byte aligned arena and the partial table of rank %d", r);
    }
    OTB_CHECK_ARG(a->mode != OTB_EXCHANGE_NVLS || (a->mc_grads != nullptr && a->mc_partials != nullptr && (((uintptr_t)a->mc_grads) & 15) == 0),
                  "NVLS exchange needs the multicast addresses of the arena and the partial table");
    OTB_UNSUPPORTED(a->mode == OTB_EXCHANGE_P2P && a->world != 2 && a->world != 4 && a->world != 8,
                    "P2P exchange is instantiated for 2, 4 and 8 ranks (got %d); use OTB_EXCHANGE_NVLS", a->world);
    OTB_BIND(a->peer_flags[a->rank]);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nvec = a->n_elems / 4, per = (nvec + a->world - 1) / a->world;
    const int64_t v0 = per * a->rank < nvec ? per * a->rank : nvec, v1 = v0 + per < nvec ? v0 + per : nvec;
    peer_barrier_kernel<<<1, 32, 0, s>>>(f, a->rank, a->world, 0);                 // every rank's gradients are complete
    OTB_LAUNCH_CHECK("peer_barrier_kernel");
    if (a->mode == OTB_EXCHANGE_NVLS) {
        grad_exchange_nvls_kernel<4><<<a->blocks, EX_THREADS, 0, s>>>((float*)a->mc_grads, v0, v1, a->grad_prescale, a->mc_partials, a->rank);
        OTB_LAUNCH_CHECK("grad_exchange_nvls_kernel");
    } else {
        if (a->world == 2) launch_p2p<2>(a, pp, v0, v1, s);
        else if (a->world == 4) launch_p2p<4>(a, pp, v0, v1, s);
        else launch_p2p<8>(a, pp, v0, v1, s);
        OTB_LAUNCH_CHECK("grad_exchange_p2p_kernel");
    }
    peer_barrier_kernel<<<1, 32, 0, s>>>(f, a->rank, a->world, 1);                 // every slice has landed in every replica
    OTB_LAUNCH_CHECK("peer_barrier_kernel");
    return OTB_OK;
}
