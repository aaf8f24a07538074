// tcgen05 / TMEM probe for B200 (sm_100a): one CTA issues kind::tf32 MMAs on a 128 x N x KB tile whose
// operands sit in shared memory in a chosen canonical layout, reads the accumulator back with tcgen05.ld
// and compares with a host reference.  Development aid used to pin the shared-memory descriptor and
// instruction-descriptor encodings before they were used in plspy_b200/csrc/boot_tf32.cu; not part of
// the product path.
//   mode 0: K-major, no swizzle ("interleave"): [k-chunk of 4][row][16 B], LBO = rows*16, SBO = 128
//   mode 1: K-major, SWIZZLE_64B  (16 fp32 per row, 8-row atoms of 512 B)
//   mode 2: K-major, SWIZZLE_128B (32 fp32 per row, 8-row atoms of 1024 B)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ inline int elem_offset(int mode, int rows, int r, int k) {   // in floats
    if (mode == 0) return ((k >> 2) * rows + r) * 4 + (k & 3);
    if (mode == 1) return r * 16 + ((((k >> 2) ^ ((r >> 1) & 3))) << 2) + (k & 3);
    return r * 32 + ((((k >> 2) ^ (r & 7))) << 2) + (k & 3);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int mode, int rows) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    uint32_t lbo, sbo, layout;
    if (mode == 0) { lbo = (uint32_t)rows * 16u; sbo = 128u; layout = 0; }
    else if (mode == 1) { lbo = 16u; sbo = 512u; layout = 4; }
    else { lbo = 16u; sbo = 1024u; layout = 2; }
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;
    return d;
}

template <int N>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                             int mode, int KB) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* sA = reinterpret_cast<float*>(sm);
    float* sB = sA + 128 * KB;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * KB; i += 128) sA[i] = A[i];
    for (int i = tid; i < N * KB; i += 128) sB[i] = B[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // make generic-proxy smem writes visible to the async (tensor core) proxy
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const int nks = KB / 8;
        for (int ks = 0; ks < nks; ++ks) {
            uint32_t offA, offB;   // byte offset of k-step ks
            if (mode == 0) { offA = (uint32_t)ks * 2u * 128u * 16u; offB = (uint32_t)ks * 2u * N * 16u; }
            else { offA = offB = (uint32_t)ks * 32u; }
            const uint64_t da = make_desc(smem_u32(sA) + offA, mode, 128);
            const uint64_t db = make_desc(smem_u32(sB) + offB, mode, N);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tbase),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar))
                     : "memory");
    }
    // everyone waits for the MMAs
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(ok)
                : "r"(smem_u32(&bar)), "r"(0u)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int e = 0; e < 16; ++e) D[tid * N + c0 + e] = __uint_as_float(v[e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tbase));
}

// ---- issue-rate calibration: one thread issues `iters` x 6 kind::tf32 MMAs (128 x N x 8) on resident shared-memory
// operands, all CTAs of the grid at once (power/clock as in a real kernel); cycles per MMA from clock64.
template <int N>
__global__ void __launch_bounds__(128) issue_rate(long long* cycles, int iters) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* sA = reinterpret_cast<float*>(sm);
    float* sB = sA + 128 * 16;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + N) * 16; i += 128) sA[i] = 1.0f + (float)(i % 7) * 0.125f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da0 = make_desc(smem_u32(sA), 1, 128), da1 = make_desc(smem_u32(sA) + 32, 1, 128);
        const uint64_t db0 = make_desc(smem_u32(sB), 1, N), db1 = make_desc(smem_u32(sB) + 32, 1, N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tbase + (uint32_t)(it & 1) * 256u;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const uint64_t da = (j & 1) ? da1 : da0, db = (j & 1) ? db1 : db0;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                    "l"(da), "l"(db), "r"(idesc), "r"(1u)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar))
                     : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(ok)
                : "r"(smem_u32(&bar)), "r"(0u)
                : "memory");
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tbase));
}

template <int N>
static void run_issue_rate(int grid) {
    long long* d; CK(cudaMalloc(&d, grid * sizeof(long long)));
    const size_t smem = (size_t)(128 + N) * 16 * 4 + 1024;
    CK(cudaFuncSetAttribute(issue_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 20000;
    issue_rate<N><<<grid, 128, smem>>>(d, 1000);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    issue_rate<N><<<grid, 128, smem>>>(d, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double cyc = 0; for (auto c : h) cyc += (double)c; cyc /= grid;
    const double flops = 2.0 * 128 * N * 8 * 6.0 * iters * grid;
    printf("issue rate: grid %d N %d: %.1f cycles per 128xNx8 TF32 MMA (ideal %d), %.1f TFLOP/s over %.2f ms, %.3f GHz\n", grid, N,
           cyc / (6.0 * iters), 128 * N / 256, flops / (ms * 1e-3) * 1e-12, ms, cyc / (ms * 1e-3) * 1e-9);
    cudaFree(d);
}

static float tf32_trunc(float x) {
    uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x;
}

template <int N>
static int run(int mode) {
    const int KB = mode == 2 ? 32 : 16;
    std::vector<float> A(128 * KB), B(N * KB), Al(128 * KB), Bl(N * KB), D(128 * N);
    srand(1 + mode);
    for (int r = 0; r < 128; ++r) for (int k = 0; k < KB; ++k) A[r * KB + k] = tf32_trunc((float)(rand() % 2001 - 1000) / 512.f);
    for (int r = 0; r < N; ++r) for (int k = 0; k < KB; ++k) B[r * KB + k] = tf32_trunc((float)(rand() % 2001 - 1000) / 512.f);
    for (int r = 0; r < 128; ++r) for (int k = 0; k < KB; ++k) Al[elem_offset(mode, 128, r, k)] = A[r * KB + k];
    for (int r = 0; r < N; ++r) for (int k = 0; k < KB; ++k) Bl[elem_offset(mode, N, r, k)] = B[r * KB + k];
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, Al.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bl.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, D.size() * 4));
    const size_t smem = (size_t)(128 + N) * KB * 4 + 1024;
    CK(cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<N><<<1, 128, smem>>>(dA, dB, dD, mode, KB);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < KB; ++k) ref += (double)A[m * KB + k] * B[n * KB + k];
            double e = fabs(ref - D[m * N + n]);
            if (!(e <= 1e-3)) ++bad;
            if (e > maxerr || e != e) maxerr = e;
        }
    printf("mode %d N %d KB %d: max abs err %.3g, mismatches %d / %d  (D[0][0]=%g D[5][7]=%g)\n", mode, N, KB, maxerr, bad,
           128 * N, D[0], D[5 * N + 7]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad;
}

int main() {
    int bad = 0;
    for (int mode = 0; mode < 3; ++mode) { bad += run<240>(mode); bad += run<256>(mode); bad += run<16>(mode); }
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    run_issue_rate<240>(1); run_issue_rate<240>(148); run_issue_rate<256>(148); run_issue_rate<128>(148);
    return bad != 0;
}
