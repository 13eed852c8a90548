// vix_runtime.cu -- context, error reporting and pointer staging of libvindex_b200.
#include "vix_common.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace vix {

static thread_local char g_err[512] = "";
static thread_local Ctx g_ctx;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return VIX_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return VIX_ERR_NO_DEVICE;
    return VIX_ERR_CUDA;
}

Ctx& ctx() { return g_ctx; }

int ensure_device() {
    static thread_local int ok = 0;
    if (ok) return VIX_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("libvindex_b200 needs a CUDA device (sm_100a); none is usable (%s). "
                  "There is no CPU fallback.", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return VIX_ERR_NO_DEVICE;
    }
    int dev = 0;
    VIX_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    VIX_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        set_error("libvindex_b200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
        return VIX_ERR_NO_DEVICE;
    }
    // keep freed scratch memory in the stream-ordered pool: with the default release threshold (0) every
    // stream synchronisation hands it back to the driver and the next call pays for mapping it again
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
        unsigned long long keep = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    ok = 1;
    return VIX_OK;
}

int num_sms() {
    static thread_local int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    }
    return sms;
}

bool is_device_ptr(const void* p) {
    if (p == nullptr) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// A tensor-core pipeline that gives up on a barrier raises this flag.  It lives in mapped pinned host memory, so the
// host can look at it without a device-to-host copy (and therefore without synchronising the stream): it is
// checked when a call that used the pipeline synchronises anyway, and at the start of the next such call.
static int* g_pipe_flag = nullptr;

int* pipeline_error_flag() {
    if (!g_pipe_flag) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        memset(p, 0, 64);
        g_pipe_flag = static_cast<int*>(p);
    }
    return g_pipe_flag;
}

int check_pipeline_error() {
    if (g_pipe_flag && *reinterpret_cast<volatile int*>(g_pipe_flag) != 0) {
        *reinterpret_cast<volatile int*>(g_pipe_flag) = 0;
        set_error("a device pipeline gave up (tensor-core barrier time-out, or a scan launch whose shared-memory layout did "
                  "not fit); the results of that call are invalid");
        return VIX_ERR_CUDA;
    }
    return VIX_OK;
}

int finish(bool any_host_output) {
    if (any_host_output || !ctx().async) {
        VIX_CUDA(cudaStreamSynchronize(ctx().stream));
        return check_pipeline_error();
    }
    return VIX_OK;
}

}  // namespace vix

using namespace vix;

namespace vix { long long tc_scan_launches(); }

extern "C" {

int vix_version(void) { return 100; }   // 0.1.0

const char* vix_last_error(void) { return g_err; }

void vix_clear_error(void) { g_err[0] = 0; }

int vix_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int vix_set_device(int device) {
    VIX_CUDA(cudaSetDevice(device));
    return VIX_OK;
}

int vix_set_stream(void* cuda_stream) {
    ctx().stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return VIX_OK;
}

int vix_set_async(int enabled) {
    ctx().async = enabled != 0;
    return VIX_OK;
}

int vix_get_async(void) { return ctx().async ? 1 : 0; }

int vix_synchronize(void) {
    VIX_CUDA(cudaStreamSynchronize(ctx().stream));
    return VIX_OK;
}

int64_t vix_kernel_launches(int reset) {
    int64_t v = ctx().launches;
    if (reset) ctx().launches = 0;
    return v;
}

int64_t vix_scan_tc_launches(void) { return (int64_t)vix::tc_scan_launches(); }

}  // extern "C"
