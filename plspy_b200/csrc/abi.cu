// Error reporting, ABI version and launch accounting for libplsb200.
#include "common.cuh"

namespace plsb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

char* err_buf() { return g_err; }

void set_err(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace plsb

extern "C" int plsb200_abi_version(void) { return PLSB200_ABI_VERSION; }
extern "C" const char* plsb200_last_error(void) { return plsb::err_buf(); }
extern "C" int64_t plsb200_launch_count(void) { return (int64_t)plsb::g_launches.load(std::memory_order_relaxed); }

// Strided (2-D) host -> device copy on a stream: `height` rows of `width_bytes`, row pitches in bytes.  Used for the
// pipelined upload of X in voxel ranges (column blocks of a row-major matrix); the host buffer must be pinned for
// the copy to be asynchronous.
extern "C" int plsb200_copy2d_h2d(void* dst, size_t dst_pitch, const void* src_host, size_t src_pitch, size_t width_bytes,
                                  size_t height, void* stream) {
    PLSB_CHECK_ARG(dst && src_host, "copy2d_h2d: null pointer");
    PLSB_CHECK_ARG(width_bytes <= dst_pitch && width_bytes <= src_pitch, "copy2d_h2d: width exceeds pitch");
    PLSB_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src_host, src_pitch, width_bytes, height, cudaMemcpyHostToDevice,
                                (cudaStream_t)stream));
    return PLSB200_OK;
}

// float32 -> float64 widening of a device array (X stored as float32 on the host: half the PCIe traffic; the images a
// NIfTI file holds are float32 / int16, so nothing is lost relative to the float64 array the reference computes on)
namespace plsb {
__global__ void widen_f32_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        *reinterpret_cast<double2*>(dst + i) = make_double2((double)v.x, (double)v.y);
        *reinterpret_cast<double2*>(dst + i + 2) = make_double2((double)v.z, (double)v.w);
    } else {
        for (; i < n; ++i) dst[i] = (double)src[i];
    }
}
}  // namespace plsb

extern "C" int plsb200_widen_f32_f64(const float* src, double* dst, int64_t n, void* stream) {
    PLSB_CHECK_ARG(src && dst && n >= 0, "widen_f32_f64: bad arguments");
    PLSB_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                   "widen_f32_f64: pointers must be 16-byte aligned");
    if (n == 0) return PLSB200_OK;
    const long long nthreads = (n + 3) / 4;
    plsb::widen_f32_f64_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    PLSB_LAUNCH_CHECK("widen_f32_f64_kernel");
    return PLSB200_OK;
}
