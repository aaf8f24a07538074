// FP64 calibration microbenchmarks for B200 (sm_100a): peak DFMA and DMMA issue rates.
// Not part of the product path; used to choose the bootstrap-kernel instruction path
// and to provide the FP64 roofline denominator.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* d, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double d[CHAINS][2];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { d[c][0] = threadIdx.x; d[c][1] = c; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) dmma884(d[c][0], d[c][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS, int SHAPE>
__global__ void __launch_bounds__(256) k_dmma16(double* out, int iters, double a, double b) {
    double d[CHAINS][4];
    double av[8], bv[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] = a + j;
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = b + j;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { d[c][0] = threadIdx.x; d[c][1] = c; d[c][2] = 1; d[c][3] = 2; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (SHAPE == 4) dmma1684(d[c], av, bv[0]);
            if (SHAPE == 8) dmma1688(d[c], av, bv);
            if (SHAPE == 16) dmma16816(d[c], av, bv);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 8));
    const int iters = 4096;
    for (int wps = 1; wps <= 4; wps *= 2) {           // CTAs per SM (256 thr each => 2,4,8 warps per SMSP)
        int grid = sms * wps;
        printf("--- %d CTA(s) of 256 threads per SM ---\n", wps);
        {
            float ms = time_ms([&] { k_dfma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 8 * iters * 256.0 * grid;
            printf("DFMA  chains=8          : %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dmma884<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 4 * iters * 8.0 * grid;
            printf("DMMA m8n8k4 chains=4    : %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dmma884<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 8 * iters * 8.0 * grid;
            printf("DMMA m8n8k4 chains=8    : %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dmma884<1><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 1 * iters * 8.0 * grid;
            printf("DMMA m8n8k4 chains=1    : %8.3f ms  %7.2f TFLOP/s (latency-bound: %.1f ns/mma)\n", ms, fl / ms * 1e-9, ms * 1e6 / iters);
        }
        {
            float ms = time_ms([&] { k_dmma16<4, 4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * 8 * 4 * 4 * iters * 8.0 * grid;
            printf("DMMA m16n8k4 chains=4   : %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dmma16<4, 8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * 8 * 8 * 4 * iters * 8.0 * grid;
            printf("DMMA m16n8k8 chains=4   : %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dmma16<4, 16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * 8 * 16 * 4 * iters * 8.0 * grid;
            printf("DMMA m16n8k16 chains=4  : %8.3f ms  %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        }
    }
    // sustained: run m8n8k4 for ~2 s and report
    {
        int grid = sms * 2;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0;
        for (; n < 400; ++n) k_dmma884<8><<<grid, 256>>>(out, iters * 4, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 256 * 8 * (iters * 4.0) * 8.0 * grid * n;
        printf("DMMA m8n8k4 sustained %.0f ms : %7.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    CK(cudaFree(out));
    return 0;
}
