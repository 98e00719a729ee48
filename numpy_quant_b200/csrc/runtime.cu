// Library-wide state: per-thread error string, device properties.
#include "common.cuh"

namespace nq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace nq

extern "C" int nq_version(void) { return 100; }

extern "C" const char* nq_last_error(void) { return nq::g_err; }

extern "C" int nq_device_info(int* props_host) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return nq::cuda_fail(e, "cudaGetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return nq::cuda_fail(e, "cudaGetDeviceProperties");
    props_host[0] = prop.multiProcessorCount;
    props_host[1] = prop.major;
    props_host[2] = prop.minor;
    props_host[3] = (int)prop.sharedMemPerBlockOptin;
    return NQ_OK;
}

extern "C" int nq_memset_async(void* ptr, int value, int64_t bytes, void* stream) {
    if (bytes <= 0) return NQ_OK;
    cudaError_t e = cudaMemsetAsync(ptr, value, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) return nq::cuda_fail(e, "nq_memset_async");
    return NQ_OK;
}
