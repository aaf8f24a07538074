// K3 + split-half: K x K Gram blocks of the two half-sample cross-block matrices through G, a
// warp-per-matrix one-sided Jacobi eigensolver with shuffle-based rotations, and the test-train /
// split-half reproducibility outputs.
//
// Reference (split_half_resampling.py:119-262, 537-683 and their null loops): per split gather the rows
// of both halves, rebuild M1, M2 (K x p), run np.linalg.svd on them (class_functions.py:122) and form
//     pls_s_train = s1                   pls_s_test = V1^T M2^T U1                       (:195-196)
//     pls_u_repro = V1^T V2              pls_v_repro = U1^T U2                           (:682-683)
// For the task methods M_h = A_h X[idx_h] with fixed K x n_h operators, so
//     M1 M1^T = A1 G[idx1,idx1] A1^T = S11,  M1 M2^T = S12,  M2 M2^T = S22      (K x K each)
//     S11 = U1 diag(s1^2) U1^T  ->  V1 = M1^T U1 diag(1/s1)
//     pls_s_test = diag(1/s1) U1^T S12 U1,   pls_u_repro = diag(1/s1) U1^T S12 U2 diag(1/s2)
// and no N x p work is left per split.
#include <float.h>

#include "common.cuh"

namespace plsb {

// ------------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi on a symmetric PSD matrix, one warp per matrix.  Lane j owns column j
// of W (initially A) and of V (initially I); in each of the n-1 steps of a round-robin tournament
// every lane fetches its partner's columns with shuffles and both lanes of a pair apply the same
// plane rotation, so n/2 rotations run in parallel.  On exit W = A V has orthogonal columns:
// eigenvalue_j = ||w_j||, eigenvector_j = v_j.  Returns the rank (0 = largest) of this lane's pair.
template <int KM>
__device__ __forceinline__ int jacobi_warp(double (&w)[KM], double (&v)[KM], double& lambda, int lane) {
    constexpr int n = KM;          // even
    const double tol = 4.0 * DBL_EPSILON;
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
#pragma unroll 1
        for (int r = 0; r < n - 1; ++r) {
            int partner;
            if (lane == n - 1) partner = r;
            else {
                partner = (2 * r - lane) % (n - 1);
                if (partner < 0) partner += n - 1;
                if (partner == lane) partner = n - 1;
            }
            const bool active = lane < n;
            const int src = active ? partner : lane;
            // pass 1: the three dot products (partner column streamed through shuffles, not stored)
            double own = 0.0, oth = 0.0, gam = 0.0;
#pragma unroll
            for (int e = 0; e < KM; ++e) {
                const double wp = __shfl_sync(0xffffffffu, w[e], src);
                own = fma(w[e], w[e], own);
                oth = fma(wp, wp, oth);
                gam = fma(w[e], wp, gam);
            }
            const bool lo = lane < partner;
            const double alpha = lo ? own : oth, beta = lo ? oth : own;   // (i < j) orientation
            const bool rot = active && gam != 0.0 && fabs(gam) > tol * sqrt(alpha * beta);
            double co = 1.0, so = 0.0;
            if (rot) {
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gam);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                co = c; so = lo ? -sn : sn;      // w_i <- c w_i - s w_j ; w_j <- s w_i + c w_j
            }
            // pass 2: both lanes of a pair apply the same rotation to their own W and V columns
            if (__any_sync(0xffffffffu, rot)) {
#pragma unroll
                for (int e = 0; e < KM; ++e) {
                    const double wp = __shfl_sync(0xffffffffu, w[e], src);
                    const double vp = __shfl_sync(0xffffffffu, v[e], src);
                    w[e] = fma(so, wp, co * w[e]);
                    v[e] = fma(so, vp, co * v[e]);
                }
            }
        }
        if (!__any_sync(0xffffffffu, rotated)) break;
    }
    double nn = 0.0;
#pragma unroll
    for (int e = 0; e < KM; ++e) nn = fma(w[e], w[e], nn);
    lambda = lane < n ? sqrt(nn) : -1.0;
    int rank = 0;
    for (int l = 0; l < n; ++l) {
        const double o = __shfl_sync(0xffffffffu, lambda, l);
        if (o > lambda || (o == lambda && l < lane)) ++rank;
    }
    return rank;
}

// load column `lane` of the K x K matrix A (row-major, symmetric) zero-padded to KM
template <int KM>
__device__ __forceinline__ void load_cols(const double* A, int K, int lane, double (&w)[KM], double (&v)[KM]) {
#pragma unroll
    for (int e = 0; e < KM; ++e) {
        w[e] = (lane < K && e < K) ? A[e * K + lane] : 0.0;
        v[e] = (e == lane) ? 1.0 : 0.0;
    }
}

template <int KM>
__global__ void __launch_bounds__(128) sym_eig_kernel(const double* __restrict__ A, int K, int B,
                                                     double* __restrict__ evals, double* __restrict__ evecs) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    double w[KM], v[KM], lam;
    load_cols<KM>(A + (size_t)warp * K * K, K, lane, w, v);
    const int rank = jacobi_warp<KM>(w, v, lam, lane);
    if (lane < KM && rank < K) {
        evals[(size_t)warp * K + rank] = lam;
#pragma unroll
        for (int e = 0; e < KM; ++e)
            if (e < K) evecs[((size_t)warp * K + e) * K + rank] = v[e];
    }
}

// ------------------------------------------------------------------------------------------------
// Gram blocks of one split: S11 = A1 G[i1,i1] A1^T, S12 = A1 G[i1,i2] A2^T, S22 = A2 G[i2,i2] A2^T.
__global__ void __launch_bounds__(512) split_gram_kernel(const double* __restrict__ G, int N,
                                                        const int32_t* __restrict__ idx1, int n1,
                                                        const int32_t* __restrict__ idx2, int n2,
                                                        const double* __restrict__ A1, const double* __restrict__ A2,
                                                        int K, double* __restrict__ S11, double* __restrict__ S12,
                                                        double* __restrict__ S22) {
    extern __shared__ __align__(16) double smg[];
    double* A1t = smg;                        // [n1][K]
    double* A2t = A1t + (size_t)n1 * K;       // [n2][K]
    double* T11 = A2t + (size_t)n2 * K;       // [n1][K]
    double* T12 = T11 + (size_t)n1 * K;       // [n1][K]
    double* T22 = T12 + (size_t)n1 * K;       // [n2][K]
    int* i1 = reinterpret_cast<int*>(T22 + (size_t)n2 * K);
    int* i2 = i1 + n1;
    const int s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n1 * K; i += nt) A1t[i] = A1[(size_t)(i % K) * n1 + i / K];
    for (int i = tid; i < n2 * K; i += nt) A2t[i] = A2[(size_t)(i % K) * n2 + i / K];
    for (int i = tid; i < n1; i += nt) i1[i] = idx1[(size_t)s * n1 + i];
    for (int i = tid; i < n2; i += nt) i2[i] = idx2[(size_t)s * n2 + i];
    __syncthreads();
    // step 1: T = G[rows, cols] . At  -- one (block, row) task per thread, K accumulators in chunks of 8
    const int ntask = 2 * n1 + n2;
    for (int t = tid; t < ntask; t += nt) {
        int row, nc; const int* cols; const double* At; double* T;
        if (t < n1)          { row = i1[t];          cols = i1; nc = n1; At = A1t; T = T11 + (size_t)t * K; }
        else if (t < 2 * n1) { row = i1[t - n1];     cols = i2; nc = n2; At = A2t; T = T12 + (size_t)(t - n1) * K; }
        else                 { row = i2[t - 2 * n1]; cols = i2; nc = n2; At = A2t; T = T22 + (size_t)(t - 2 * n1) * K; }
        const double* g = G + (size_t)row * N;
        for (int k0 = 0; k0 < K; k0 += 8) {
            double acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = 0.0;
            for (int j = 0; j < nc; ++j) {
                const double gv = __ldg(g + cols[j]);
                const double* a = At + (size_t)j * K + k0;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (k0 + u < K) acc[u] = fma(gv, a[u], acc[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < K) T[k0 + u] = acc[u];
        }
    }
    __syncthreads();
    // step 2: S = A . T
    for (int o = tid; o < 3 * K * K; o += nt) {
        const int which = o / (K * K), a = (o % (K * K)) / K, b = o % K;
        const double* At = which == 2 ? A2t : A1t;
        const double* T = which == 0 ? T11 : (which == 1 ? T12 : T22);
        const int n = which == 2 ? n2 : n1;
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc = fma(At[(size_t)i * K + a], T[(size_t)i * K + b], acc);
        double* out = which == 0 ? S11 : (which == 1 ? S12 : S22);
        out[((size_t)s * K + a) * K + b] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Per split: eigendecompose S11 (warp 0) and S22 (warp 1), then the K x K outputs.
template <int KM>
__global__ void __launch_bounds__(64) split_svd_kernel(const double* __restrict__ S11, const double* __restrict__ S12,
                                                      const double* __restrict__ S22, int K,
                                                      double* __restrict__ s_train, double* __restrict__ s_test,
                                                      double* __restrict__ u_repro, double* __restrict__ v_repro,
                                                      double* __restrict__ s2_out) {
    __shared__ double U[2][KM * KM];      // eigenvectors as columns, sorted: U[h][r*KM + c]
    __shared__ double sv[2][KM];          // singular values sqrt(lambda)
    __shared__ double P[KM * KM];         // S12 . U1 or S12 . U2
    const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const double* A = (warp == 0 ? S11 : S22) + (size_t)s * K * K;
        double w[KM], v[KM], lam;
        load_cols<KM>(A, K, lane, w, v);
        const int rank = jacobi_warp<KM>(w, v, lam, lane);
        if (lane < KM && rank < K) {
            sv[warp][rank] = sqrt(lam);
#pragma unroll
            for (int e = 0; e < KM; ++e)
                if (e < K) U[warp][e * KM + rank] = v[e];
        }
    }
    __syncthreads();
    const double* C = S12 + (size_t)s * K * K;
    const int tid = threadIdx.x;
    if (s_train)
        for (int k = tid; k < K; k += 64) s_train[(size_t)s * K + k] = sv[0][k];
    if (s2_out)
        for (int k = tid; k < K; k += 64) s2_out[(size_t)s * K + k] = sv[1][k];
    if (v_repro)    // U1^T U2
        for (int o = tid; o < K * K; o += 64) {
            const int i = o / K, j = o % K;
            double acc = 0.0;
            for (int r = 0; r < K; ++r) acc = fma(U[0][r * KM + i], U[1][r * KM + j], acc);
            v_repro[(size_t)s * K * K + o] = acc;
        }
    for (int pass = 0; pass < 2; ++pass) {
        double* out = pass == 0 ? s_test : u_repro;
        if (!out) continue;
        const int h = pass;     // right factor: U1 for the test-train output, U2 for split-half
        __syncthreads();
        for (int o = tid; o < K * K; o += 64) {        // P = S12 . U_h
            const int r = o / K, j = o % K;
            double acc = 0.0;
            for (int c = 0; c < K; ++c) acc = fma(C[r * K + c], U[h][c * KM + j], acc);
            P[r * KM + j] = acc;
        }
        __syncthreads();
        for (int o = tid; o < K * K; o += 64) {        // diag(1/s1) U1^T P [diag(1/s2)]
            const int i = o / K, j = o % K;
            double acc = 0.0;
            for (int r = 0; r < K; ++r) acc = fma(U[0][r * KM + i], P[r * KM + j], acc);
            double scale = sv[0][i] > 0.0 ? 1.0 / sv[0][i] : 0.0;
            if (pass == 1) scale *= sv[1][j] > 0.0 ? 1.0 / sv[1][j] : 0.0;
            out[(size_t)s * K * K + o] = acc * scale;
        }
    }
}

}  // namespace plsb

using namespace plsb;

extern "C" int plsb200_sym_eig_f64(const double* A, int K, int B, double* evals, double* evecs, void* stream) {
    PLSB_CHECK_ARG(A && evals && evecs, "sym_eig_f64: null pointer");
    PLSB_CHECK_ARG(K > 0 && B >= 0, "sym_eig_f64: bad shape");
    if (K > 32) {
        set_err("sym_eig_f64: K=%d > 32 not supported by the warp-per-matrix solver", K);
        return PLSB200_EUNSUPPORTED;
    }
    if (B == 0) return PLSB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)cdiv((int64_t)B * 32, 128);
    if (K <= 8) sym_eig_kernel<8><<<grid, 128, 0, st>>>(A, K, B, evals, evecs);
    else if (K <= 16) sym_eig_kernel<16><<<grid, 128, 0, st>>>(A, K, B, evals, evecs);
    else if (K <= 24) sym_eig_kernel<24><<<grid, 128, 0, st>>>(A, K, B, evals, evecs);
    else sym_eig_kernel<32><<<grid, 128, 0, st>>>(A, K, B, evals, evecs);
    PLSB_LAUNCH_CHECK("sym_eig_kernel");
    return PLSB200_OK;
}

extern "C" int plsb200_split_gram_f64(const double* G, int N, const int32_t* idx1, int n1, const int32_t* idx2, int n2,
                                      int S, const double* A1, const double* A2, int K, double* S11, double* S12,
                                      double* S22, void* stream) {
    PLSB_CHECK_ARG(G && idx1 && idx2 && A1 && A2 && S11 && S12 && S22, "split_gram_f64: null pointer");
    PLSB_CHECK_ARG(N > 0 && n1 > 0 && n2 > 0 && K > 0 && S >= 0, "split_gram_f64: bad shape");
    if (S == 0) return PLSB200_OK;
    size_t smem = ((size_t)(3 * n1 + 2 * n2) * K) * sizeof(double) + (size_t)(n1 + n2) * sizeof(int);
    if (smem > 220 * 1024) {
        set_err("split_gram_f64: halves too large for shared memory (n1=%d n2=%d K=%d)", n1, n2, K);
        return PLSB200_EUNSUPPORTED;
    }
    PLSB_CUDA(cudaFuncSetAttribute(split_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = (int)cdiv(2 * n1 + n2, 32) * 32;
    if (threads > 512) threads = 512;
    if (threads < 128) threads = 128;
    split_gram_kernel<<<S, threads, smem, (cudaStream_t)stream>>>(G, N, idx1, n1, idx2, n2, A1, A2, K, S11, S12, S22);
    PLSB_LAUNCH_CHECK("split_gram_kernel");
    return PLSB200_OK;
}

extern "C" int plsb200_split_svd_f64(const double* S11, const double* S12, const double* S22, int K, int S,
                                     double* s_train, double* s_test, double* u_repro, double* v_repro, double* s2,
                                     void* stream) {
    PLSB_CHECK_ARG(S11 && S12 && S22, "split_svd_f64: null pointer");
    PLSB_CHECK_ARG(K > 0 && S >= 0, "split_svd_f64: bad shape");
    if (K > 32) {
        set_err("split_svd_f64: K=%d > 32 not supported", K);
        return PLSB200_EUNSUPPORTED;
    }
    if (S == 0) return PLSB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 8) split_svd_kernel<8><<<S, 64, 0, st>>>(S11, S12, S22, K, s_train, s_test, u_repro, v_repro, s2);
    else if (K <= 16) split_svd_kernel<16><<<S, 64, 0, st>>>(S11, S12, S22, K, s_train, s_test, u_repro, v_repro, s2);
    else if (K <= 24) split_svd_kernel<24><<<S, 64, 0, st>>>(S11, S12, S22, K, s_train, s_test, u_repro, v_repro, s2);
    else split_svd_kernel<32><<<S, 64, 0, st>>>(S11, S12, S22, K, s_train, s_test, u_repro, v_repro, s2);
    PLSB_LAUNCH_CHECK("split_svd_kernel");
    return PLSB200_OK;
}
