// K2/K6: per-resample contractions in N-space, permutation counters, U_hat, column std.
//
// For one resample with index vector idx (row i of the resampled matrix is row idx[i] of X) the
// reference rebuilds the K x p cross-block matrix and projects it (bootstrap_permutation.py:384-405,
// 602-634).  With G = X X^T and E = Lop^T U (N x K, constant) the same numbers are
//     H = G . C,  C[idx[i], :] += E[i, :]       =>  H[j, k] = sum_i G[idx[i], j] . E[i, k]
//     ||VS[:, k]||^2 = C[:, k]^T H[:, k]         =   sum_i E[i, k] . H[idx[i], k]
//     Tdistrib      = Lmat . H . diag(1/||VS||)                     (cell means of X @ normalize(VS))
// i.e. row gathers of the L2-resident G; no scatter, no atomics, bit-reproducible.
#include "common.cuh"

namespace plsb {

// One CTA per resample.  Thread j owns row j of H for KT columns at a time.
template <int KT>
__global__ void __launch_bounds__(512) nspace_kernel(const double* __restrict__ G, int N, const double* __restrict__ E,
                                                      long long e_stride, int K, int ldk, int koff,
                                                      const int32_t* __restrict__ idx,
                                                      const double* __restrict__ Lmat, int Kt,
                                                      double* __restrict__ d2, double* __restrict__ T,
                                                      double* __restrict__ Bfull) {
    extern __shared__ __align__(16) double sm[];
    double* Es = sm;                      // [N][K]
    double* Hs = Es + (size_t)N * K;      // [N][K]
    double* dn = Hs + (size_t)N * K;      // [K] squared norms, then 1/sqrt
    int* ids = reinterpret_cast<int*>(dn + K);   // [N]
    const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    // e_stride != 0: explicit per-resample coefficient matrices (behaviour PLS); idx == NULL: identity
    const double* Er = E + (size_t)r * e_stride;
    for (int i = tid; i < N * K; i += nt) Es[i] = Er[(size_t)(i / K) * ldk + koff + i % K];
    for (int i = tid; i < N; i += nt) ids[i] = idx ? idx[(size_t)r * N + i] : i;
    __syncthreads();

    for (int j = tid; j < N; j += nt) {
        for (int k0 = 0; k0 < K; k0 += KT) {
            double acc[KT];
#pragma unroll
            for (int t = 0; t < KT; ++t) acc[t] = 0.0;
            const int kn = min(KT, K - k0);
            if (kn == KT && (K & 1) == 0) {
                // full column group, even row stride: the broadcast reads of E go out as 16-byte loads (one LDS per two
                // FMAs); the gathers of G rows are latency-bound (L2), so eight are kept in flight per thread
#pragma unroll 8
                for (int i = 0; i < N; ++i) {
                    const double g = __ldg(G + (size_t)ids[i] * N + j);
                    const double2* e2 = reinterpret_cast<const double2*>(Es + i * K + k0);
#pragma unroll
                    for (int t = 0; t < KT / 2; ++t) {
                        const double2 ev = e2[t];
                        acc[2 * t] = fma(g, ev.x, acc[2 * t]);
                        acc[2 * t + 1] = fma(g, ev.y, acc[2 * t + 1]);
                    }
                }
            } else if (kn == KT) {
#pragma unroll 4
                for (int i = 0; i < N; ++i) {
                    const double g = __ldg(G + (size_t)ids[i] * N + j);
                    const double* e = Es + i * K + k0;
#pragma unroll
                    for (int t = 0; t < KT; ++t) acc[t] = fma(g, e[t], acc[t]);
                }
            } else {
                for (int i = 0; i < N; ++i) {
                    const double g = __ldg(G + (size_t)ids[i] * N + j);
                    const double* e = Es + i * K + k0;
#pragma unroll
                    for (int t = 0; t < KT; ++t)
                        if (t < kn) acc[t] = fma(g, e[t], acc[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < KT; ++t)
                if (t < kn) Hs[j * K + k0 + t] = acc[t];
        }
    }
    __syncthreads();

    // d2[k] = sum_i E[i,k] * H[idx[i], k] : one warp per column, fixed-order shuffle reduction
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int k = warp; k < K; k += nw) {
        double s = 0.0;
        for (int i = lane; i < N; i += 32) s = fma(Es[i * K + k], Hs[ids[i] * K + k], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            d2[(size_t)r * ldk + koff + k] = s;
            dn[k] = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
        }
    }
    // optional: the whole K x K matrix B = C^T G C = C^T H (the Gram matrix of the resampled cross-block matrix when
    // E carries the builder's rows instead of projected weights): input of the per-resample SVD mode
    if (Bfull != nullptr) {
        for (int o = tid; o < K * K; o += nt) {
            const int k1 = o / K, k2 = o % K;
            double s = 0.0;
            for (int i = 0; i < N; ++i) s = fma(Es[i * K + k1], Hs[ids[i] * K + k2], s);
            Bfull[(size_t)r * K * K + o] = s;
        }
    }
    if (T == nullptr) return;
    __syncthreads();
    // T[c, k] = (sum_j Lmat[c, j] H[j, k]) / ||VS_k||
    for (int o = tid; o < Kt * K; o += nt) {
        const int c = o / K, k = o % K;
        const double* l = Lmat + (size_t)c * N;
        double s = 0.0;
        for (int j = 0; j < N; ++j) s = fma(__ldg(l + j), Hs[j * K + k], s);
        T[((size_t)r * Kt + c) * ldk + koff + k] = s * dn[k];
    }
}

// One thread per resample; block-level integer reduction, int64 atomics (exact, order-independent).
__global__ void perm_count_kernel(const double* __restrict__ d2, int R, int K, const double* __restrict__ s_ref,
                                  const double* __restrict__ totcov_ref, double thresh,
                                  const double* __restrict__ mb_total, unsigned long long* __restrict__ counts,
                                  double* __restrict__ s_hat) {
    extern __shared__ int cnt[];  // [2K]
    for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) {
        const double* d = d2 + (size_t)r * K;
        double* sh = s_hat + (size_t)r * K;
        double q4 = 0.0;
        if (mb_total) {
            for (int k = 0; k < K; ++k) { const double v = fmax(d[k], 0.0); q4 += v * v; }
        }
        for (int k = 0; k < K; ++k) {
            double v = fmax(d[k], 0.0);
            double s = sqrt(v);
            if (mb_total) s = sqrt(v * v / q4 * mb_total[r]);   // s_hat^4 / sum s_hat^4 * total SS
            if (thresh > 0.0 && fabs(s) < thresh) s = 0.0;
            sh[k] = s;
            if (s >= s_ref[k]) atomicAdd(&cnt[k], 1);
        }
        double tail = 0.0;
        for (int k = K - 1; k >= 0; --k) {
            tail += sh[k] * sh[k];
            if (tail >= totcov_ref[k]) atomicAdd(&cnt[K + k], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * K; i += blockDim.x)
        if (cnt[i]) atomicAdd(counts + i, (unsigned long long)cnt[i]);
}

// U_hat_r = Lop (Ku x N) . XL[idx_r, :]   one CTA per resample
__global__ void uhat_kernel(const double* __restrict__ XL, long long xl_stride, int N, int K, int ldk, int koff,
                            const double* __restrict__ Lop, int Ku, const int32_t* __restrict__ idx,
                            double* __restrict__ Uhat) {
    extern __shared__ __align__(16) double smu[];
    double* Xs = smu;   // gathered XL rows [N][K] (K = width of this column chunk, ldk = full width)
    const int r = blockIdx.x;
    const double* xl = XL + (size_t)r * xl_stride;       // xl_stride != 0: one latent matrix per resample
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        const int src = idx ? idx[(size_t)r * N + i / K] : i / K;
        Xs[i] = xl[(size_t)src * ldk + koff + i % K];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < Ku * K; o += blockDim.x) {
        const int c = o / K, k = o % K;
        const double* l = Lop + (size_t)c * N;
        double s = 0.0;
        for (int i = 0; i < N; ++i) s = fma(__ldg(l + i), Xs[i * K + k], s);
        Uhat[((size_t)r * Ku + c) * ldk + koff + k] = s;
    }
}

// population std over axis 0 of A[R][M], two-pass like numpy (mean, then mean of squared deviations).
// A CTA owns CS_COLS columns; its 1024 threads form CS_COLS x 128 (column, row-lane) pairs that stride over the
// rows, and the row-lane partials are combined by a fixed-order tree in shared memory (deterministic).
constexpr int CS_COLS = 8;
constexpr int CS_ROWS = 128;
__device__ __forceinline__ double colstd_block_sum(double v, double* red, int col, int rl) {
    red[rl * CS_COLS + col] = v;
    __syncthreads();
    for (int h = CS_ROWS / 2; h > 0; h >>= 1) {
        if (rl < h) red[rl * CS_COLS + col] += red[(rl + h) * CS_COLS + col];
        __syncthreads();
    }
    const double t = red[col];
    __syncthreads();
    return t;
}
__global__ void __launch_bounds__(CS_COLS * CS_ROWS) colstd_kernel(const double* __restrict__ A, int R, long long M,
                                                                 double* __restrict__ out) {
    __shared__ double red[CS_ROWS * CS_COLS];
    const int col = threadIdx.x % CS_COLS, rl = threadIdx.x / CS_COLS;
    const long long m = (long long)blockIdx.x * CS_COLS + col;
    const bool ok = m < M;
    double s = 0.0;
#pragma unroll 4
    for (int r = rl; r < R; r += CS_ROWS) s += ok ? __ldg(A + (size_t)r * M + m) : 0.0;
    const double mu = colstd_block_sum(s, red, col, rl) / R;
    double q = 0.0;
#pragma unroll 4
    for (int r = rl; r < R; r += CS_ROWS) {
        const double d = ok ? __ldg(A + (size_t)r * M + m) - mu : 0.0;
        q = fma(d, d, q);
    }
    const double t = colstd_block_sum(q, red, col, rl);
    if (rl == 0 && ok) out[m] = sqrt(t / R);
}

template <int KT>
static int launch_nspace(const double* G, int N, const double* E, long long e_stride, int K, int ldk, int koff,
                         const int32_t* idx, int R,
                         const double* Lmat, int Kt, double* d2, double* T, double* Bfull, cudaStream_t st) {
    size_t smem = ((size_t)2 * N * K + K) * sizeof(double) + (size_t)N * sizeof(int);
    PLSB_CUDA(cudaFuncSetAttribute(nspace_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = (int)cdiv(N, 32) * 32;
    if (threads > 512) threads = 512;
    if (threads < 128) threads = 128;
    nspace_kernel<KT><<<R, threads, smem, st>>>(G, N, E, e_stride, K, ldk, koff, idx, Lmat, Kt, d2, T, Bfull);
    PLSB_LAUNCH_CHECK("nspace_kernel");
    return PLSB200_OK;
}

}  // namespace plsb

using namespace plsb;

static int nspace_dispatch(const double* G, int N, const double* E, long long e_stride, int K, const int32_t* idx,
                           int R, const double* Lmat, int Kt, double* d2, double* T, cudaStream_t st,
                           double* Bfull = nullptr) {
    // columns are independent: chunk K so that E and H chunks (2 * N * Kc doubles) fit in shared memory
    const size_t budget = 200 * 1024;
    int kc = K;
    while (kc > 1 && ((size_t)2 * N * kc + kc) * sizeof(double) + (size_t)N * sizeof(int) > budget) kc = (kc + 1) / 2;
    if (((size_t)2 * N * kc + kc) * sizeof(double) + (size_t)N * sizeof(int) > budget) {
        set_err("nspace_f64: N=%d too large for shared memory", N);
        return PLSB200_EUNSUPPORTED;
    }
    if (Bfull != nullptr && kc < K) {
        set_err("nspace_gram_f64: N=%d x K=%d does not fit in shared memory in one column chunk", N, K);
        return PLSB200_EUNSUPPORTED;
    }
    for (int k0 = 0; k0 < K; k0 += kc) {
        const int kw = K - k0 < kc ? K - k0 : kc;
        int rc;
        if (kw <= 4) rc = launch_nspace<4>(G, N, E, e_stride, kw, K, k0, idx, R, Lmat, Kt, d2, T, Bfull, st);
        else if (kw <= 8) rc = launch_nspace<8>(G, N, E, e_stride, kw, K, k0, idx, R, Lmat, Kt, d2, T, Bfull, st);
        else if (kw <= 12) rc = launch_nspace<12>(G, N, E, e_stride, kw, K, k0, idx, R, Lmat, Kt, d2, T, Bfull, st);
        else if (kw <= 16) rc = launch_nspace<16>(G, N, E, e_stride, kw, K, k0, idx, R, Lmat, Kt, d2, T, Bfull, st);
        else rc = launch_nspace<24>(G, N, E, e_stride, kw, K, k0, idx, R, Lmat, Kt, d2, T, Bfull, st);
        if (rc != PLSB200_OK) return rc;
    }
    return PLSB200_OK;
}

extern "C" int plsb200_nspace_f64(const double* G, int N, const double* E, int K, const int32_t* idx, int R,
                                  const double* Lmat, int Kt, double* d2, double* T, void* stream) {
    PLSB_CHECK_ARG(G && E && idx && d2, "nspace_f64: null pointer");
    PLSB_CHECK_ARG(N > 0 && K > 0 && R >= 0, "nspace_f64: bad shape N=%d K=%d R=%d", N, K, R);
    PLSB_CHECK_ARG((T == nullptr) || (Lmat != nullptr && Kt > 0), "nspace_f64: T requested without Lmat");
    if (R == 0) return PLSB200_OK;
    return nspace_dispatch(G, N, E, 0, K, idx, R, Lmat, Kt, d2, T, (cudaStream_t)stream);
}

extern "C" int plsb200_nspace_gram_f64(const double* G, int N, const double* E, int K, const int32_t* idx, int R,
                                       double* d2, double* B, void* stream) {
    PLSB_CHECK_ARG(G && E && idx && d2 && B, "nspace_gram_f64: null pointer");
    PLSB_CHECK_ARG(N > 0 && K > 0 && R >= 0, "nspace_gram_f64: bad shape N=%d K=%d R=%d", N, K, R);
    if (R == 0) return PLSB200_OK;
    return nspace_dispatch(G, N, E, 0, K, idx, R, nullptr, 0, d2, nullptr, (cudaStream_t)stream, B);
}

extern "C" int plsb200_nspace_coef_f64(const double* G, int N, const double* C, int K, int R, const double* Lmat,
                                       int Kt, double* d2, double* T, void* stream) {
    PLSB_CHECK_ARG(G && C && d2, "nspace_coef_f64: null pointer");
    PLSB_CHECK_ARG(N > 0 && K > 0 && R >= 0, "nspace_coef_f64: bad shape N=%d K=%d R=%d", N, K, R);
    PLSB_CHECK_ARG((T == nullptr) || (Lmat != nullptr && Kt > 0), "nspace_coef_f64: T requested without Lmat");
    if (R == 0) return PLSB200_OK;
    return nspace_dispatch(G, N, C, (long long)N * K, K, nullptr, R, Lmat, Kt, d2, T, (cudaStream_t)stream);
}

extern "C" int plsb200_nspace_coef_gram_f64(const double* G, int N, const double* C, int K, int R, double* d2, double* B,
                                            void* stream) {
    PLSB_CHECK_ARG(G && C && d2 && B, "nspace_coef_gram_f64: null pointer");
    PLSB_CHECK_ARG(N > 0 && K > 0 && R >= 0, "nspace_coef_gram_f64: bad shape N=%d K=%d R=%d", N, K, R);
    if (R == 0) return PLSB200_OK;
    return nspace_dispatch(G, N, C, (long long)N * K, K, nullptr, R, nullptr, 0, d2, nullptr, (cudaStream_t)stream, B);
}

extern "C" int plsb200_perm_count_f64(const double* d2, int R, int K, const double* s_ref, const double* totcov_ref,
                                      double thresh, const double* mb_total, int64_t* counts, double* s_hat,
                                      void* stream) {
    PLSB_CHECK_ARG(d2 && s_ref && totcov_ref && counts && s_hat, "perm_count_f64: null pointer");
    PLSB_CHECK_ARG(K > 0 && R >= 0, "perm_count_f64: bad shape");
    if (R == 0) return PLSB200_OK;
    const int threads = 128;
    perm_count_kernel<<<(int)cdiv(R, threads), threads, 2 * K * sizeof(int), (cudaStream_t)stream>>>(
        d2, R, K, s_ref, totcov_ref, thresh, mb_total, reinterpret_cast<unsigned long long*>(counts), s_hat);
    PLSB_LAUNCH_CHECK("perm_count_kernel");
    return PLSB200_OK;
}

extern "C" int plsb200_uhat_f64(const double* XL, int64_t xl_stride, int N, int K, const double* Lop, int Ku,
                                const int32_t* idx, int R, double* Uhat, void* stream) {
    PLSB_CHECK_ARG(XL && Lop && Uhat, "uhat_f64: null pointer");
    PLSB_CHECK_ARG(N > 0 && K > 0 && Ku > 0 && R >= 0, "uhat_f64: bad shape");
    if (R == 0) return PLSB200_OK;
    int kc = K;
    while (kc > 1 && (size_t)N * kc * sizeof(double) > 200 * 1024) kc = (kc + 1) / 2;
    if ((size_t)N * kc * sizeof(double) > 200 * 1024) {
        set_err("uhat_f64: N=%d too large for shared memory", N);
        return PLSB200_EUNSUPPORTED;
    }
    for (int k0 = 0; k0 < K; k0 += kc) {       // columns are independent
        const int kw = K - k0 < kc ? K - k0 : kc;
        size_t smem = (size_t)N * kw * sizeof(double);
        PLSB_CUDA(cudaFuncSetAttribute(uhat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uhat_kernel<<<R, 256, smem, (cudaStream_t)stream>>>(XL, xl_stride, N, kw, K, k0, Lop, Ku, idx, Uhat);
        PLSB_LAUNCH_CHECK("uhat_kernel");
    }
    return PLSB200_OK;
}

extern "C" int plsb200_colstd_f64(const double* A, int R, int64_t M, double* out, void* stream) {
    PLSB_CHECK_ARG(A && out, "colstd_f64: null pointer");
    PLSB_CHECK_ARG(R > 0 && M > 0, "colstd_f64: bad shape");
    colstd_kernel<<<(unsigned)cdiv(M, CS_COLS), CS_COLS * CS_ROWS, 0, (cudaStream_t)stream>>>(A, R, M, out);
    PLSB_LAUNCH_CHECK("colstd_kernel");
    return PLSB200_OK;
}
