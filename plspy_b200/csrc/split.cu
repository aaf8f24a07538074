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
// Sweep limit of both Jacobi solvers.  Cyclic one-sided Jacobi converges quadratically; the K x K Gram blocks of this
// path need 6-10 sweeps.  A matrix that has not converged within the limit is reported through the `status` output
// of the entry points (1 = not converged) and the Python layer raises -- results are never consumed unchecked.
constexpr int JACOBI_MAX_SWEEPS = 60;

// One-sided (Hestenes) Jacobi on a symmetric PSD matrix, one warp per matrix.  Lane j owns column j
// of W (initially A) and of V (initially I); in each of the n-1 steps of a round-robin tournament
// every lane fetches its partner's columns with shuffles and both lanes of a pair apply the same
// plane rotation, so n/2 rotations run in parallel.  On exit W = A V has orthogonal columns:
// eigenvalue_j = ||w_j||, eigenvector_j = v_j.  Returns the rank (0 = largest) of this lane's pair.
template <int KM>
__device__ __forceinline__ int jacobi_warp(double (&w)[KM], double (&v)[KM], double& lambda, int lane,
                                           bool& converged) {
    constexpr int n = KM;          // even
    const double tol = 4.0 * DBL_EPSILON;
    converged = false;
    // Columns whose norm has fallen to the rounding level of the matrix (||w_j|| = lambda_j <= n eps lambda_max: the
    // null space of a rank-deficient Gram block, e.g. mctype 3 or repeated rows) are pure noise: their cosines with the
    // other columns never drop below `tol`, so they are left alone instead of being rotated for ever.
    double nul;
    {
        double own = 0.0;
#pragma unroll
        for (int e = 0; e < KM; ++e) own = fma(w[e], w[e], own);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) own = fmax(own, __shfl_xor_sync(0xffffffffu, own, o));
        nul = own * ((double)n * DBL_EPSILON) * ((double)n * DBL_EPSILON);
    }
    for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS; ++sweep) {
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
            const bool rot = active && gam != 0.0 && alpha > nul && beta > nul && fabs(gam) > tol * sqrt(alpha * beta);
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
        if (!__any_sync(0xffffffffu, rotated)) { converged = true; break; }
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
                                                     double* __restrict__ evals, double* __restrict__ evecs,
                                                     int32_t* __restrict__ status) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    double w[KM], v[KM], lam;
    bool conv;
    load_cols<KM>(A + (size_t)warp * K * K, K, lane, w, v);
    const int rank = jacobi_warp<KM>(w, v, lam, lane, conv);
    if (status && lane == 0 && !conv) status[warp] = 1;          // raised only: the entry point zeroes the array
    if (lane < KM && rank < K) {
        evals[(size_t)warp * K + rank] = lam;
#pragma unroll
        for (int e = 0; e < KM; ++e)
            if (e < K) evecs[((size_t)warp * K + e) * K + rank] = v[e];
    }
}

// ------------------------------------------------------------------------------------------------
// The same one-sided Jacobi for 32 < K <= JACOBI_CTA_KMAX, one CTA per matrix: the columns of W and V live in shared
// memory (column-major), the n/2 disjoint column pairs of a tournament round are rotated by n/2 warps in parallel
// (lanes stride the column, dot products by xor-shuffles), one __syncthreads per round.  Same rotation formula,
// threshold and ordering of the eigenpairs as jacobi_warp.  (behaviour designs: K = groups x conditions x behaviours,
// e.g. 3 x 4 x 4 = 48; multiblock: groups x (conditions + |bscan| x behaviours).)
constexpr int JACOBI_CTA_KMAX = 112;       // 2 n^2 doubles of shared memory: 196 KB at n = 112

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

__global__ void __launch_bounds__(1024) sym_eig_cta_kernel(const double* __restrict__ A, int K, int B,
                                                          double* __restrict__ evals, double* __restrict__ evecs,
                                                          int32_t* __restrict__ status) {
    extern __shared__ __align__(16) double smj[];
    const int n = (K + 1) & ~1;
    double* W = smj;                       // [n][n]: column c at W + c*n
    double* V = W + (size_t)n * n;
    double* lam = V + (size_t)n * n;       // [n]
    __shared__ int any_rot;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const double tol = 4.0 * DBL_EPSILON;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const double* Ab = A + (size_t)b * K * K;
        for (int i = tid; i < n * n; i += nt) {
            const int c = i / n, e = i % n;
            W[i] = (c < K && e < K) ? Ab[(size_t)e * K + c] : 0.0;
            V[i] = (c == e) ? 1.0 : 0.0;
        }
        __syncthreads();
        // rounding-level columns are left alone (see jacobi_warp)
        for (int c = warp; c < n; c += nw) {
            double nn = 0.0;
            for (int e = lane; e < n; e += 32) nn = fma(W[(size_t)c * n + e], W[(size_t)c * n + e], nn);
            nn = warp_sum(nn);
            if (lane == 0) lam[c] = nn;
        }
        __syncthreads();
        double nul = 0.0;
        for (int c = 0; c < n; ++c) nul = fmax(nul, lam[c]);
        nul *= ((double)n * DBL_EPSILON) * ((double)n * DBL_EPSILON);
        __syncthreads();
        bool conv = false;
        for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS && !conv; ++sweep) {
            if (tid == 0) any_rot = 0;
            __syncthreads();
            for (int r = 0; r < n - 1; ++r) {
                for (int q = warp; q < n / 2; q += nw) {
                    int i, j;
                    if (q == 0) { i = r; j = n - 1; }
                    else {
                        i = (r + q) % (n - 1);
                        j = (r - q + (n - 1)) % (n - 1);
                        if (i > j) { const int t = i; i = j; j = t; }
                    }
                    double* wi = W + (size_t)i * n; double* wj = W + (size_t)j * n;
                    double alpha = 0.0, beta = 0.0, gam = 0.0;
                    for (int e = lane; e < n; e += 32) {
                        const double a = wi[e], c = wj[e];
                        alpha = fma(a, a, alpha); beta = fma(c, c, beta); gam = fma(a, c, gam);
                    }
                    alpha = warp_sum(alpha); beta = warp_sum(beta); gam = warp_sum(gam);
                    if (gam != 0.0 && alpha > nul && beta > nul && fabs(gam) > tol * sqrt(alpha * beta)) {
                        const double zeta = (beta - alpha) / (2.0 * gam);
                        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                        double* vi = V + (size_t)i * n; double* vj = V + (size_t)j * n;
                        for (int e = lane; e < n; e += 32) {      // w_i <- c w_i - s w_j ; w_j <- s w_i + c w_j
                            const double a = wi[e], d = wj[e];
                            wi[e] = fma(-sn, d, c * a); wj[e] = fma(sn, a, c * d);
                            const double x = vi[e], y = vj[e];
                            vi[e] = fma(-sn, y, c * x); vj[e] = fma(sn, x, c * y);
                        }
                        if (lane == 0) any_rot = 1;
                    }
                }
                __syncthreads();
            }
            conv = any_rot == 0;
            __syncthreads();
        }
        if (status && tid == 0 && !conv) status[b] = 1;              // raised only
        for (int c = warp; c < n; c += nw) {
            double nn = 0.0;
            for (int e = lane; e < n; e += 32) nn = fma(W[(size_t)c * n + e], W[(size_t)c * n + e], nn);
            nn = warp_sum(nn);
            if (lane == 0) lam[c] = sqrt(nn);
        }
        __syncthreads();
        for (int c = warp; c < n; c += nw) {
            const double l = lam[c];
            int rank = 0;
            for (int o = lane; o < n; o += 32) {
                const double x = lam[o];
                if (x > l || (x == l && o < c)) ++rank;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
            if (rank < K) {
                if (lane == 0) evals[(size_t)b * K + rank] = l;
                for (int e = lane; e < K; e += 32) evecs[((size_t)b * K + e) * K + rank] = V[(size_t)c * n + e];
            }
        }
        __syncthreads();
    }
}

static int launch_sym_eig_cta(const double* A, int K, int B, double* evals, double* evecs, int32_t* status,
                              cudaStream_t st) {
    const int n = (K + 1) & ~1;
    const size_t smem = ((size_t)2 * n * n + n) * sizeof(double);
    PLSB_CUDA(cudaFuncSetAttribute(sym_eig_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int warps = n / 2; if (warps > 32) warps = 32;
    const int grid = B < 4 * num_sms() ? B : 4 * num_sms();
    sym_eig_cta_kernel<<<grid, warps * 32, smem, st>>>(A, K, B, evals, evecs, status);
    PLSB_LAUNCH_CHECK("sym_eig_cta_kernel");
    return PLSB200_OK;
}

// K x K outputs of a split from the eigenpairs of S11 and S22 (the K > 32 path of plsb200_split_svd_f64):
// lam_h = eigenvalues of S_hh (= squared singular values), U_h = eigenvectors in columns (row-major K x K).
__global__ void __launch_bounds__(256) split_out_kernel(const double* __restrict__ S12, const double* __restrict__ lam1,
                                                       const double* __restrict__ U1, const double* __restrict__ lam2,
                                                       const double* __restrict__ U2, int K,
                                                       double* __restrict__ s_train, double* __restrict__ s_test,
                                                       double* __restrict__ u_repro, double* __restrict__ v_repro,
                                                       double* __restrict__ s2_out) {
    extern __shared__ __align__(16) double smo[];
    double* P = smo;                      // [K][K]
    double* sv1 = P + (size_t)K * K; double* sv2 = sv1 + K;
    const int s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double* C = S12 + (size_t)s * K * K;
    const double* Ua = U1 + (size_t)s * K * K; const double* Ub = U2 + (size_t)s * K * K;
    for (int k = tid; k < K; k += nt) {
        sv1[k] = sqrt(fmax(lam1[(size_t)s * K + k], 0.0));      // singular value of M_h = sqrt(eigenvalue of S_hh)
        sv2[k] = sqrt(fmax(lam2[(size_t)s * K + k], 0.0));
    }
    __syncthreads();
    if (s_train) for (int k = tid; k < K; k += nt) s_train[(size_t)s * K + k] = sv1[k];
    if (s2_out) for (int k = tid; k < K; k += nt) s2_out[(size_t)s * K + k] = sv2[k];
    if (v_repro)
        for (int o = tid; o < K * K; o += nt) {
            const int i = o / K, j = o % K;
            double acc = 0.0;
            for (int r = 0; r < K; ++r) acc = fma(Ua[(size_t)r * K + i], Ub[(size_t)r * K + j], acc);
            v_repro[(size_t)s * K * K + o] = acc;
        }
    for (int pass = 0; pass < 2; ++pass) {
        double* out = pass == 0 ? s_test : u_repro;
        if (!out) continue;
        const double* Uh = pass == 0 ? Ua : Ub;
        __syncthreads();
        for (int o = tid; o < K * K; o += nt) {
            const int r = o / K, j = o % K;
            double acc = 0.0;
            for (int c = 0; c < K; ++c) acc = fma(C[(size_t)r * K + c], Uh[(size_t)c * K + j], acc);
            P[o] = acc;
        }
        __syncthreads();
        for (int o = tid; o < K * K; o += nt) {
            const int i = o / K, j = o % K;
            double acc = 0.0;
            for (int r = 0; r < K; ++r) acc = fma(Ua[(size_t)r * K + i], P[(size_t)r * K + j], acc);
            double scale = sv1[i] > 0.0 ? 1.0 / sv1[i] : 0.0;
            if (pass == 1) scale *= sv2[j] > 0.0 ? 1.0 / sv2[j] : 0.0;
            out[(size_t)s * K * K + o] = acc * scale;
        }
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
                                                      double* __restrict__ s2_out, int32_t* __restrict__ status) {
    __shared__ double U[2][KM * KM];      // eigenvectors as columns, sorted: U[h][r*KM + c]
    __shared__ double sv[2][KM];          // singular values sqrt(lambda)
    __shared__ double P[KM * KM];         // S12 . U1 or S12 . U2
    const int s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const double* A = (warp == 0 ? S11 : S22) + (size_t)s * K * K;
        double w[KM], v[KM], lam;
        bool conv;
        load_cols<KM>(A, K, lane, w, v);
        const int rank = jacobi_warp<KM>(w, v, lam, lane, conv);
        if (status && lane == 0 && !conv) atomicOr(status + s, 1);
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

extern "C" int plsb200_sym_eig_f64(const double* A, int K, int B, double* evals, double* evecs, int32_t* status,
                                   void* stream) {
    PLSB_CHECK_ARG(A && evals && evecs, "sym_eig_f64: null pointer");
    PLSB_CHECK_ARG(K > 0 && B >= 0, "sym_eig_f64: bad shape");
    if (K > JACOBI_CTA_KMAX) {
        set_err("sym_eig_f64: K=%d > %d not supported (columns of W and V must fit in shared memory)", K,
                JACOBI_CTA_KMAX);
        return PLSB200_EUNSUPPORTED;
    }
    if (B == 0) return PLSB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (status) PLSB_CUDA(cudaMemsetAsync(status, 0, (size_t)B * sizeof(int32_t), st));
    if (K > 32) return launch_sym_eig_cta(A, K, B, evals, evecs, status, st);
    const int grid = (int)cdiv((int64_t)B * 32, 128);
    if (K <= 8) sym_eig_kernel<8><<<grid, 128, 0, st>>>(A, K, B, evals, evecs, status);
    else if (K <= 16) sym_eig_kernel<16><<<grid, 128, 0, st>>>(A, K, B, evals, evecs, status);
    else if (K <= 24) sym_eig_kernel<24><<<grid, 128, 0, st>>>(A, K, B, evals, evecs, status);
    else sym_eig_kernel<32><<<grid, 128, 0, st>>>(A, K, B, evals, evecs, status);
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

extern "C" size_t plsb200_split_svd_f64_workspace(int K, int S) {
    if (K <= 32 || S <= 0) return 16;
    return (size_t)2 * S * ((size_t)K * K + K) * sizeof(double);      // eigenpairs of S11 and S22
}

extern "C" int plsb200_split_svd_f64(const double* S11, const double* S12, const double* S22, int K, int S,
                                     double* s_train, double* s_test, double* u_repro, double* v_repro, double* s2,
                                     int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    PLSB_CHECK_ARG(S11 && S12 && S22, "split_svd_f64: null pointer");
    PLSB_CHECK_ARG(K > 0 && S >= 0, "split_svd_f64: bad shape");
    if (K > JACOBI_CTA_KMAX) {
        set_err("split_svd_f64: K=%d > %d not supported", K, JACOBI_CTA_KMAX);
        return PLSB200_EUNSUPPORTED;
    }
    if (S == 0) return PLSB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (status) PLSB_CUDA(cudaMemsetAsync(status, 0, (size_t)S * sizeof(int32_t), st));
    if (K > 32) {
        // CTA-per-matrix eigensolver on S11 and S22, then the outputs from the eigenpairs
        if (!workspace || workspace_bytes < plsb200_split_svd_f64_workspace(K, S)) {
            set_err("split_svd_f64: workspace too small (%zu < %zu bytes)", workspace_bytes,
                    plsb200_split_svd_f64_workspace(K, S));
            return PLSB200_EWORKSPACE;
        }
        double* U1 = static_cast<double*>(workspace);
        double* U2 = U1 + (size_t)S * K * K;
        double* l1 = U2 + (size_t)S * K * K;
        double* l2 = l1 + (size_t)S * K;
        int rc = launch_sym_eig_cta(S11, K, S, l1, U1, status, st);      // status[s] is only ever raised to 1,
        if (rc) return rc;                                               // so both solves report into it
        rc = launch_sym_eig_cta(S22, K, S, l2, U2, status, st);
        if (rc) return rc;
        const size_t smem = ((size_t)K * K + 2 * K) * sizeof(double);
        PLSB_CUDA(cudaFuncSetAttribute(split_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        split_out_kernel<<<S, 256, smem, st>>>(S12, l1, U1, l2, U2, K, s_train, s_test, u_repro, v_repro, s2);
        PLSB_LAUNCH_CHECK("split_out_kernel");
        return PLSB200_OK;
    }
    if (K <= 8) split_svd_kernel<8><<<S, 64, 0, st>>>(S11, S12, S22, K, s_train, s_test, u_repro, v_repro, s2, status);
    else if (K <= 16) split_svd_kernel<16><<<S, 64, 0, st>>>(S11, S12, S22, K, s_train, s_test, u_repro, v_repro, s2, status);
    else if (K <= 24) split_svd_kernel<24><<<S, 64, 0, st>>>(S11, S12, S22, K, s_train, s_test, u_repro, v_repro, s2, status);
    else split_svd_kernel<32><<<S, 64, 0, st>>>(S11, S12, S22, K, s_train, s_test, u_repro, v_repro, s2, status);
    PLSB_LAUNCH_CHECK("split_svd_kernel");
    return PLSB200_OK;
}
