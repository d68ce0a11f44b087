// api.cu — error plumbing, device probing and small utilities of the C ABI.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

void otb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* otb_last_error(void) { return g_err; }

int otb_pdl_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* v = getenv("OTB_PDL");
        cached = (v && v[0] == '0') ? 0 : 1;
    }
    return cached;
}
extern "C" int otb_abi_version(void) { return OTB_ABI_VERSION; }

#include <cuda.h>

namespace {
template <typename F>
F driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return (F)p;
}
}  // namespace

int otb_bind_context(const void* device_ptr) {
    typedef CUresult (*GetCur)(CUcontext*);
    typedef CUresult (*PtrAttr)(void*, CUpointer_attribute, CUdeviceptr);
    typedef CUresult (*Retain)(CUcontext*, CUdevice);
    typedef CUresult (*SetCur)(CUcontext);
    static GetCur get_cur = driver_fn<GetCur>("cuCtxGetCurrent");
    static PtrAttr ptr_attr = driver_fn<PtrAttr>("cuPointerGetAttribute");
    static Retain retain = driver_fn<Retain>("cuDevicePrimaryCtxRetain");
    static SetCur set_cur = driver_fn<SetCur>("cuCtxSetCurrent");
    if (!get_cur || !ptr_attr || !retain || !set_cur) {
        otb_set_error("CUDA driver entry points not available");
        return OTB_ECUDA;
    }
    CUcontext ctx = nullptr;
    if (get_cur(&ctx) == CUDA_SUCCESS && ctx != nullptr) return OTB_OK;
    if (device_ptr == nullptr) {
        otb_set_error("no CUDA context is current on this thread and no device pointer to find one from");
        return OTB_ECUDA;
    }
    int ordinal = -1;
    CUresult r = ptr_attr(&ordinal, CU_POINTER_ATTRIBUTE_DEVICE_ORDINAL, (CUdeviceptr)(uintptr_t)device_ptr);
    if (r != CUDA_SUCCESS || ordinal < 0) {
        otb_set_error("cuPointerGetAttribute(%p) failed with CUresult %d: not a device pointer?", device_ptr, (int)r);
        return OTB_EINVAL;
    }
    r = retain(&ctx, (CUdevice)ordinal);
    if (r == CUDA_SUCCESS) r = set_cur(ctx);
    if (r != CUDA_SUCCESS) {
        otb_set_error("binding the primary context of device %d failed with CUresult %d", ordinal, (int)r);
        return OTB_ECUDA;
    }
    return OTB_OK;
}

int otb_sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

extern "C" int otb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0, n = 0, maj = 0, min = 0;
    OTB_CUDA(cudaGetDevice(&dev));
    OTB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    OTB_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    OTB_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (maj != 10) {
        otb_set_error("libotb200 is built for sm_100a only; device reports compute capability %d.%d", maj, min);
        return OTB_EARCH;
    }
    return OTB_OK;
}

extern "C" int otb_struct_sizes(int32_t* out) {
    OTB_CHECK_ARG(out != nullptr, "null output");
    out[0] = (int32_t)sizeof(otb_optim_tensor);
    out[1] = (int32_t)sizeof(otb_optim_launch);
    out[2] = (int32_t)sizeof(otb_shadow_tensor);
    out[3] = (int32_t)sizeof(otb_noise_args);
    out[4] = (int32_t)sizeof(otb_loss_args);
    out[5] = (int32_t)sizeof(otb_lora_linear_args);
    out[6] = (int32_t)sizeof(otb_lora_grad_args);
    out[7] = (int32_t)sizeof(otb_grad_exchange_args);
    out[8] = (int32_t)sizeof(otb_factored_tensor);
    out[9] = (int32_t)sizeof(otb_factored_launch);
    out[10] = (int32_t)sizeof(otb_dora_args);
    out[11] = (int32_t)sizeof(otb_shard_exchange_args);
    out[12] = (int32_t)sizeof(otb_peer_store);
    out[13] = (int32_t)sizeof(otb_sr_args);
    return OTB_OK;
}

__global__ void philox_dump_kernel(uint64_t seed, uint64_t offset, uint32_t stream_id, int64_t first, int64_t n, uint32_t* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Philox4 r = otb_rng(seed, offset, stream_id, (uint64_t)(first + i));
    out[4 * i + 0] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

extern "C" int otb_philox_dump(uint64_t seed, uint64_t offset, uint32_t stream_id, int64_t first_group,
                               int64_t n_groups, uint32_t* out4, void* stream) {
    OTB_CHECK_ARG(out4 != nullptr && n_groups >= 0, "bad philox dump args");
    if (n_groups == 0) return OTB_OK;
    OTB_BIND(out4);
    philox_dump_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, offset, stream_id, first_group, n_groups, out4);
    OTB_LAUNCH_CHECK("philox_dump_kernel");
    return OTB_OK;
}
