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

}  // namespace

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
