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
__global__ void __launch_bounds__(128) issue_rate(long long* cycles, int iters, int commit_each) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* sA = reinterpret_cast<float*>(sm);
    float* sB = sA + 128 * 16;
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + N) * 16; i += 128) sA[i] = 1.0f + (float)(i % 7) * 0.125f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar2)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar3)));   // never arrived on: parity-1 waits pass
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
            if (commit_each)      // one commit per group of six MMAs, as a pipeline stage of the real kernel does
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                                 smem_u32(&bar2))
                             : "memory");
            if (commit_each >= 2) {   // ... followed by the wait on the next stage's (already complete) full barrier + fence
                uint32_t ok2 = 0;
                while (!ok2) {
                    asm volatile(
                        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                        : "=r"(ok2)
                        : "r"(smem_u32(&bar3)), "r"(1u)
                        : "memory");
                }
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
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
static void run_issue_rate(int grid, int commit_each = 0) {
    long long* d; CK(cudaMalloc(&d, grid * sizeof(long long)));
    const size_t smem = (size_t)(128 + N) * 16 * 4 + 1024;
    CK(cudaFuncSetAttribute(issue_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 20000;
    issue_rate<N><<<grid, 128, smem>>>(d, 1000, commit_each);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    issue_rate<N><<<grid, 128, smem>>>(d, iters, commit_each);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double cyc = 0; for (auto c : h) cyc += (double)c; cyc /= grid;
    const double flops = 2.0 * 128 * N * 8 * 6.0 * iters * grid;
    printf("issue rate (commit per 6 MMAs: %d): grid %d N %d: %.1f cycles per 128xNx8 TF32 MMA (ideal %d), %.1f TFLOP/s over %.2f ms, %.3f GHz\n", commit_each, grid, N,
           cyc / (6.0 * iters), 128 * N / 256, flops / (ms * 1e-3) * 1e-12, ms, cyc / (ms * 1e-3) * 1e-9);
    cudaFree(d);
}


// ---- does the epilogue's tcgen05.ld traffic slow the MMAs down?  Thread 0 issues MMAs into accumulator 0 while warps
// 2-5 read accumulator 1 with tcgen05.ld at the real kernel's ratio (`lds_per_mma` x16 loads per MMA, per warp) in
// chunks of LDN columns.
template <int N, int LDN>
__global__ void __launch_bounds__(192) mma_vs_ld(long long* cycles, int iters, int ld_per_114) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* sA = reinterpret_cast<float*>(sm);
    float* sB = sA + 128 * 16;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + N) * 16; i += 192) sA[i] = 1.0f + (float)(i % 7) * 0.125f;
    if (tid == 0) {
        done = 0;
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
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const uint64_t da = (j & 1) ? da1 : da0, db = (j & 1) ? db1 : db0;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tbase),
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
        done = 1;
    } else if (warp >= 2 && ld_per_114 > 0) {
        // total x16-equivalent loads this warp should issue over the run, spread evenly with nanosleep pacing
        const long long total = (long long)iters * 6 * ld_per_114 / 114;
        const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
        float accv = 0.f;
        for (long long i = 0; i < total && !done; i += LDN / 16) {
            if constexpr (LDN == 16) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr + (uint32_t)((i * 16) % 224)));
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                accv += __uint_as_float(v[0]) + __uint_as_float(v[15]);
            } else {
                uint32_t v[64];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                    "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                      "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                      "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
                      "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
                      "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
                      "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                    : "r"(taddr + (uint32_t)((i * 16) % 192)));
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                accv += __uint_as_float(v[0]) + __uint_as_float(v[63]);
            }
            __nanosleep(400);
        }
        if (accv == 12345.f) cycles[0] = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tbase));
}

template <int N, int LDN>
static void run_mma_vs_ld(int ld_per_114) {
    const int grid = 148, iters = 20000;
    long long* d; CK(cudaMalloc(&d, grid * sizeof(long long)));
    const size_t smem = (size_t)(128 + N) * 16 * 4 + 1024;
    CK(cudaFuncSetAttribute(mma_vs_ld<N, LDN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_vs_ld<N, LDN><<<grid, 192, smem>>>(d, 1000, ld_per_114);
    CK(cudaDeviceSynchronize());
    mma_vs_ld<N, LDN><<<grid, 192, smem>>>(d, iters, ld_per_114);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double cyc = 0; for (auto c : h) cyc += (double)c; cyc /= grid;
    printf("MMA vs tcgen05.ld: %d x16-loads per warp per 114 MMAs, issued as .x%d: %.1f cycles per MMA (ideal %d)\n", ld_per_114, LDN,
           cyc / (6.0 * iters), 128 * N / 256);
    cudaFree(d);
}


// ---- MMAs whose operands cycle through a 4-slot ring of 46 KB stages (Xh, Xl 128x16 and Ch, Cl 240x16, as in the real
// kernel: three products x two k-steps per stage), optionally with a second thread streaming bulk copies from global
// memory into the ring at the real kernel's rate (no hazard tracking: the values do not matter here).
__global__ void __launch_bounds__(128) mma_ring(long long* cycles, int iters, const float* __restrict__ gsrc, int with_tma) {
    constexpr int N = 240;
    constexpr uint32_t A_PLANE = 128 * 16 * 4, B_PLANE = N * 16 * 4, STAGE = 2 * A_PLANE + 2 * B_PLANE;   // 47104 B
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar, tbar[4];
    __shared__ uint32_t tmem_base;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* f = reinterpret_cast<float*>(sm);
    for (int i = tid; i < 4 * (int)STAGE / 4; i += 128) f[i] = 1.0f + (float)(i % 7) * 0.125f;
    if (tid == 0) {
        done = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&tbar[i])));
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
    const uint32_t ring = smem_u32(sm);
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t sa = ring + (uint32_t)(it & 3) * STAGE, sb = sa + 2 * A_PLANE;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint64_t ah = make_desc(sa + ks * 32, 1, 128), al = make_desc(sa + A_PLANE + ks * 32, 1, 128);
                const uint64_t bh = make_desc(sb + ks * 32, 1, N), bl = make_desc(sb + B_PLANE + ks * 32, 1, N);
                const uint64_t da[3] = {ah, ah, al}, db[3] = {bl, bh, bh};
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tbase),
                        "l"(da[j]), "l"(db[j]), "r"(idesc), "r"(1u)
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
        done = 1;
    } else if (tid == 32 && with_tma) {
        // one 46 KB stage per ~720 MMA cycles: issue a stage, wait for it, repeat (self-paced by the copy latency)
        uint32_t ph[4] = {0, 0, 0, 0};
        const char* src = reinterpret_cast<const char*>(gsrc) + (size_t)blockIdx.x * 4 * STAGE;
        for (int it = 0; !done; ++it) {
            const int sl = it & 3;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&tbar[sl])), "r"(STAGE)
                         : "memory");
            for (uint32_t off = 0; off < STAGE; off += 11776u)
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                        ring + sl * STAGE + off),
                    "l"(src + sl * STAGE + off), "r"(11776u), "r"(smem_u32(&tbar[sl]))
                    : "memory");
            if (it >= 3) {      // keep three stages in flight: wait for the one issued three iterations ago
                const int w = (it - 3) & 3;
                uint32_t ok = 0;
                while (!ok && !done) {
                    asm volatile(
                        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                        : "=r"(ok)
                        : "r"(smem_u32(&tbar[w])), "r"(ph[w])
                        : "memory");
                }
                ph[w] ^= 1u;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tbase));
}

static void run_mma_ring(int with_tma) {
    const int grid = 148, iters = 20000;
    long long* d; CK(cudaMalloc(&d, grid * sizeof(long long)));
    float* g; CK(cudaMalloc(&g, (size_t)grid * 4 * 47104)); CK(cudaMemset(g, 0, (size_t)grid * 4 * 47104));
    const size_t smem = 4 * 47104 + 1024;
    CK(cudaFuncSetAttribute(mma_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_ring<<<grid, 128, smem>>>(d, 1000, g, with_tma);
    CK(cudaDeviceSynchronize());
    mma_ring<<<grid, 128, smem>>>(d, iters, g, with_tma);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double cyc = 0; for (auto c : h) cyc += (double)c; cyc /= grid;
    printf("MMA on a 4-slot operand ring (3 products x 2 k-steps per 46 KB stage), concurrent bulk copies: %d -> %.1f cycles per MMA (ideal 120)\n",
           with_tma, cyc / (6.0 * iters));
    cudaFree(d); cudaFree(g);
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
    run_issue_rate<240>(148, 1); run_issue_rate<240>(148, 2);
    run_mma_vs_ld<240, 16>(0); run_mma_vs_ld<240, 16>(15); run_mma_vs_ld<240, 16>(60); run_mma_vs_ld<240, 64>(15); run_mma_vs_ld<240, 64>(60);
    run_mma_ring(0); run_mma_ring(1);
    return bad != 0;
}
