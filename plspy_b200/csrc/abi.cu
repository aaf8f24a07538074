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
