// common.cuh — shared device helpers for libotb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/otb200.h"

// ---------------------------------------------------------------------------------------
// error plumbing (thread-local message, C ABI returns a code)
// ---------------------------------------------------------------------------------------
void otb_set_error(const char* fmt, ...);

#define OTB_CHECK_ARG(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            otb_set_error(__VA_ARGS__);                           \
            return OTB_EINVAL;                                    \
        }                                                         \
    } while (0)

#define OTB_UNSUPPORTED(cond, ...)                                \
    do {                                                          \
        if (cond) {                                               \
            otb_set_error(__VA_ARGS__);                           \
            return OTB_EUNSUPPORTED;                              \
        }                                                         \
    } while (0)

#define OTB_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            otb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),      \
                          __FILE__, __LINE__);                                          \
            return OTB_ECUDA;                                                           \
        }                                                                               \
    } while (0)

#define OTB_LAUNCH_CHECK(name)                                                          \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            otb_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));    \
            return OTB_ECUDA;                                                           \
        }                                                                               \
    } while (0)

int otb_sm_count();   // cached; 148 on B200

// Make sure the calling thread has a current CUDA context: the one that owns `device_ptr`.
// autograd's backward thread may reach this library before torch has bound a context to it; driver
// entry points (cuTensorMapEncodeTiled) then fail with CUDA_ERROR_INVALID_CONTEXT, and a runtime call
// would silently create a context on device 0.  No-op when a context is already current.
int otb_bind_context(const void* device_ptr);
#define OTB_BIND(ptr)                                             \
    do {                                                          \
        int rc__ = otb_bind_context(ptr);                         \
        if (rc__ != OTB_OK) return rc__;                          \
    } while (0)

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Restated bit-for-bit in oracle/otb_oracle.c.
// counter = (group_lo, group_hi, offset_lo, stream<<24 | offset_hi[0:24)), key = seed.
// ---------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t otb_mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 otb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                               uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = otb_mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = otb_mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o{c0, c1, c2, c3};
    return o;
}

__host__ __device__ __forceinline__ Philox4 otb_rng(uint64_t seed, uint64_t offset, uint32_t stream,
                                                     uint64_t group) {
    return otb_philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)offset,
                             (stream << 24) | ((uint32_t)(offset >> 32) & 0x00FFFFFFu),
                             (uint32_t)seed, (uint32_t)(seed >> 32));
}

// uniform in [0,1) with 24 bits (same construction torch's CPU/CUDA rand uses for fp32)
__host__ __device__ __forceinline__ float otb_u01(uint32_t x) {
    return (float)(x >> 8) * 5.9604644775390625e-08f;  // 2^-24
}

// Box-Muller on two 32-bit words -> two N(0,1) fp32 values.
//   u1 = ((a>>8)+0.5) * 2^-24  in (0,1);   th = (((int)b>>8)+0.5) * 2^-23 * pi  in (-pi,pi)
//   r  = sqrt(-2 ln u1);  z0 = r cos th;  z1 = r sin th
// The bulk-noise kernel evaluates this with one MUFU op per transcendental (lg2 / sqrt / sin / cos
// .approx, abs err ~1e-6) and single-rounding FMAs for the affine maps; the oracle
// evaluates the same definition with libm in fp64 (tests: 2e-5 absolute).
__device__ __forceinline__ float otb_mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float otb_mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float otb_mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float otb_mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void otb_box_muller_fast(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u1 = fmaf((float)(a >> 8), 5.9604644775390625e-08f, 2.98023223876953125e-08f);          // (v + 0.5) * 2^-24
    const float th = fmaf((float)((int32_t)b >> 8), 3.7450702829239286e-07f, 1.8725351414619643e-07f);   // (v + 0.5) * 2^-23 * pi
    const float r = otb_mufu_sqrt(otb_mufu_lg2(u1) * -1.3862943611198906f);                               // sqrt(-2 ln u1)
    z0 = r * otb_mufu_cos(th);
    z1 = r * otb_mufu_sin(th);
}

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the
// stream is still draining; otb_pdl_wait() blocks until that predecessor has completed and
// its writes are visible, so everything BEFORE it (barrier init, TMEM allocation, tensor-map
// prefetch) overlaps the predecessor's tail.  otb_pdl_trigger() lets the successor do the
// same to us.  Both are no-ops for a launch without the attribute.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void otb_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void otb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
int otb_pdl_enabled();   // OTB_PDL=0 disables the launch attribute (A/B runs)

// ---------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }
__device__ __forceinline__ uint16_t float_to_bf16_bits_rn(float f) {
    return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
// two floats -> packed bf16x2 (lo in bits [0,16)), round-to-nearest-even, ONE F2FP.BF16.PACK_AB
__device__ __forceinline__ uint32_t pack2_bf16_rn(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// round a float to the nearest bf16 and return it as float (emulates a bf16 store+load)
__device__ __forceinline__ float round_bf16(float f) { return bf16_bits_to_float(float_to_bf16_bits_rn(f)); }

__device__ __forceinline__ float load_as_float(const void* p, int64_t i, int dtype) {
    if (dtype == OTB_F32) return ((const float*)p)[i];
    if (dtype == OTB_BF16) return bf16_bits_to_float(((const uint16_t*)p)[i]);
    return __half2float(((const __half*)p)[i]);
}
__device__ __forceinline__ void store_from_float(void* p, int64_t i, int dtype, float v) {
    if (dtype == OTB_F32) ((float*)p)[i] = v;
    else if (dtype == OTB_BF16) ((uint16_t*)p)[i] = float_to_bf16_bits_rn(v);
    else ((__half*)p)[i] = __float2half_rn(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit accesses (inputs read once, outputs written once)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u2(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
