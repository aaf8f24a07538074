// Split-half Gram blocks of the behaviour / multiblock family, windowed version (split_half_resampling.py:198-262,
// 315-383, 615-683, 734-802).  Same quantity as half_gram_kernel (rb.cu):
//     M_h[k](v) = sum_seg sc_seg(v) sum_{pos in seg} x[ids[pos], v] Q_h[pos, k],     S11 = M1 M1^T, S12 = M1 M2^T, S22 = M2 M2^T
// but it uses the structure of Q: the positions of one block only feed a few adjacent rows of the cross-block
// matrix (a block's correlations with the nb behaviours; a block's mean), so every segment carries a column window
// [col0, col0 + width), width <= 8, and phase 1 costs `width` FMAs per position instead of K.  Phase 2, the
// 2K x 2K Gram of the tile's rows, runs on the FP64 tensor cores (DMMA.8x8x4) out of shared memory.
//
// CTA = 128 voxels (thread per voxel in phase 1, 4 warps of DMMA row-tasks in phase 2); splits are looped inside the
// CTA so the tile's rows of X stay hot in L2; per (tile, split) partials are summed in a fixed order afterwards.
#include "common.cuh"
#include <math.h>

namespace plsb {

constexpr int HG_VT = 128;       // voxels per CTA
constexpr int HG_U = 8;          // positions in flight per thread
constexpr int HG_SEG = 6;        // ints per segment: pos_begin, pos_end, col0, width, unit, offset of its packed coefficients

// One block of positions for one voxel: P[j] += sum_pos x[pos] q[pos][j], block moments m1 = sum x, m2 = sum x^2
// (standardised blocks only).  Even / odd positions go to separate accumulators (shorter dependency chains).
template <int W, bool UNIT>
__device__ __forceinline__ void hg_segment(const char* __restrict__ srcv, bool ok,
                                           const long long* __restrict__ soff, const double* __restrict__ qw, int b,
                                           int e, double (&P)[8], double& m1, double& m2) {
    double PA[W], PB[W], m1b = 0.0, m2b = 0.0;
#pragma unroll
    for (int j = 0; j < W; ++j) PA[j] = PB[j] = 0.0;
    int pos = b;
    for (; pos + HG_U <= e; pos += HG_U) {
        double x[HG_U];
#pragma unroll
        for (int u = 0; u < HG_U; ++u) x[u] = ok ? __ldg(reinterpret_cast<const double*>(srcv + soff[pos + u])) : 0.0;
#pragma unroll
        for (int u = 0; u < HG_U; ++u) {
            const double* q = qw + (size_t)(pos + u - b) * W;
            if (!UNIT) {
                if (u & 1) { m1b += x[u]; m2b = fma(x[u], x[u], m2b); }
                else { m1 += x[u]; m2 = fma(x[u], x[u], m2); }
            }
            if (W == 1) {
                if (u & 1) PB[0] = fma(x[u], q[0], PB[0]);
                else PA[0] = fma(x[u], q[0], PA[0]);
            } else {
#pragma unroll
                for (int j = 0; j < W; j += 2) {
                    const double2 qq = *reinterpret_cast<const double2*>(q + j);
                    if (u & 1) { PB[j] = fma(x[u], qq.x, PB[j]); PB[j + 1] = fma(x[u], qq.y, PB[j + 1]); }
                    else { PA[j] = fma(x[u], qq.x, PA[j]); PA[j + 1] = fma(x[u], qq.y, PA[j + 1]); }
                }
            }
        }
    }
    if (pos < e) {
        double x[HG_U];
#pragma unroll
        for (int u = 0; u < HG_U; ++u)
            x[u] = (ok && pos + u < e) ? __ldg(reinterpret_cast<const double*>(srcv + soff[pos + u])) : 0.0;
#pragma unroll
        for (int u = 0; u < HG_U; ++u) {
            if (pos + u < e) {
                const double* q = qw + (size_t)(pos + u - b) * W;
                if (!UNIT) { m1 += x[u]; m2 = fma(x[u], x[u], m2); }
#pragma unroll
                for (int j = 0; j < W; ++j) PA[j] = fma(x[u], q[j], PA[j]);
            }
        }
    }
    m1 += m1b; m2 += m2b;
#pragma unroll
    for (int j = 0; j < W; ++j) P[j] += PA[j] + PB[j];
}

template <int W>
__device__ __forceinline__ void hg_segment_u(const char* __restrict__ srcv, bool ok,
                                             const long long* __restrict__ soff, const double* __restrict__ qw, int b,
                                             int e, bool unit, double (&P)[8], double& m1, double& m2) {
    if (unit) hg_segment<W, true>(srcv, ok, soff, qw, b, e, P, m1, m2);
    else hg_segment<W, false>(srcv, ok, soff, qw, b, e, P, m1, m2);
}

// doubles per position in the packed window coefficients: the width rounded up to 1, 2, 4 or 8
__host__ __device__ __forceinline__ int hg_wq(int w) { return w <= 1 ? 1 : (w == 2 ? 2 : (w <= 4 ? 4 : 8)); }

template <int KC>
__global__ void __launch_bounds__(HG_VT) half_gram_win_kernel(const double* __restrict__ Xstd,
                                                             const double* __restrict__ Xlin, long long p,
                                                             const int32_t* __restrict__ ids,
                                                             const double* __restrict__ Q,
                                                             const int32_t* __restrict__ segs, int nseg, int nq,
                                                             int nmax, int K, int s0, int ns,
                                                             double* __restrict__ part) {
    constexpr int VTP = HG_VT + 4;       // pitch of one row of M over the tile's voxels: phase-1 stores (lane = voxel) and
    constexpr int NB8 = KC / 8;          // the DMMA fragment loads (8 rows x 4 voxels) are both bank-conflict-free
    extern __shared__ __align__(16) double smw[];
    double* Rs = smw;                                              // [2][KC][VTP]
    double* Qs = Rs + 2 * KC * VTP;                                // [nq] packed window coefficients of one half
    long long* soff = reinterpret_cast<long long*>(Qs + nq);       // [nmax] byte offsets of the rows ids * p * 8
    int* sg = reinterpret_cast<int*>(soff + nmax);                 // [2][nseg][HG_SEG]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long v = (long long)blockIdx.x * HG_VT + tid;
    const bool ok = v < p;
    for (int i = tid; i < 2 * nseg * HG_SEG; i += HG_VT) sg[i] = segs[i];
    // blockIdx.y = group of splits (the host picks the group count so that tiles x groups fills whole waves of CTAs)
    const int per = (ns + (int)gridDim.y - 1) / (int)gridDim.y;
    const int ss_end = min(ns, ((int)blockIdx.y + 1) * per);
    for (int ss = (int)blockIdx.y * per; ss < ss_end; ++ss) {
        const int sp = s0 + ss;
        for (int h = 0; h < 2; ++h) {
            __syncthreads();             // previous users of Qs / soff (and, for h == 0, of Rs) are done
            const double* q = Q + ((size_t)(sp * 2 + h) * nmax) * K;
            const int* sh = sg + h * nseg * HG_SEG;
            for (int s = 0; s < nseg; ++s) {
                const int b = sh[s * HG_SEG], e = sh[s * HG_SEG + 1], c0 = sh[s * HG_SEG + 2], w = sh[s * HG_SEG + 3];
                double* dst = Qs + sh[s * HG_SEG + 5];
                const int wq = hg_wq(w), sh_wq = wq == 1 ? 0 : (wq == 2 ? 1 : (wq == 4 ? 2 : 3));
                for (int i = tid; i < (e - b) * wq; i += HG_VT) {
                    const int pos = b + (i >> sh_wq), j = i & (wq - 1);
                    dst[i] = j < w ? q[(size_t)pos * K + c0 + j] : 0.0;
                }
            }
            for (int i = tid; i < nmax; i += HG_VT)
                soff[i] = (long long)ids[(size_t)(sp * 2 + h) * nmax + i] * p * (long long)sizeof(double);
            double* row = Rs + (size_t)h * KC * VTP + tid;          // row[k * VTP] = M_h[k](v)
#pragma unroll
            for (int k = 0; k < KC; ++k) row[k * VTP] = 0.0;
            __syncthreads();
            for (int s = 0; s < nseg; ++s) {
                const int b = sh[s * HG_SEG], e = sh[s * HG_SEG + 1], c0 = sh[s * HG_SEG + 2], w = sh[s * HG_SEG + 3];
                if (e <= b || w <= 0) continue;
                const bool unit = sh[s * HG_SEG + 4] != 0;
                const char* srcv = reinterpret_cast<const char*>((unit ? Xlin : Xstd) + (ok ? v : 0));
                const double* qw = Qs + sh[s * HG_SEG + 5];
                double P[8], m1 = 0.0, m2 = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) P[j] = 0.0;
                if (w == 1) hg_segment_u<1>(srcv, ok, soff, qw, b, e, unit, P, m1, m2);
                else if (w == 2) hg_segment_u<2>(srcv, ok, soff, qw, b, e, unit, P, m1, m2);
                else if (w <= 4) hg_segment_u<4>(srcv, ok, soff, qw, b, e, unit, P, m1, m2);
                else hg_segment_u<8>(srcv, ok, soff, qw, b, e, unit, P, m1, m2);
                double sc = 1.0;
                if (!unit) {
                    const double n = (double)(e - b), rn = 1.0 / n;
                    m1 *= rn; m2 *= rn;
                    const double var = m2 - m1 * m1;
                    sc = (var > 1e-13 * m2 && var > 0.0) ? rsqrt(var * n) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < w) row[(c0 + j) * VTP] += sc * P[j];
            }
        }
        __syncthreads();
        // phase 2: row-task (ha, a0) = 8 rows of M_ha against every needed 8-column block; R = [R1 | R2]
        double* out = part + ((size_t)blockIdx.x * ns + ss) * 3 * K * K;
        for (int task = warp; task < 2 * NB8; task += HG_VT / 32) {
            const int ha = task / NB8, a0 = (task % NB8) * 8;
            double acc[2 * NB8][2];
#pragma unroll
            for (int bb = 0; bb < 2 * NB8; ++bb) acc[bb][0] = acc[bb][1] = 0.0;
            const double* Ra = Rs + (size_t)(ha * KC + a0 + (lane >> 2)) * VTP + (lane & 3);
            const double* Rb = Rs + (size_t)(lane >> 2) * VTP + (lane & 3);
#pragma unroll 2
            for (int ks = 0; ks < HG_VT / 4; ++ks) {
                const double a = Ra[ks * 4];
#pragma unroll
                for (int bb = 0; bb < 2 * NB8; ++bb) {
                    if (bb >= ha * NB8) {
                        const double bv = Rb[((bb / NB8) * KC + (bb % NB8) * 8) * VTP + ks * 4];
                        dmma884(acc[bb][0], acc[bb][1], a, bv);
                    }
                }
            }
            const int ar = a0 + (lane >> 2);
#pragma unroll
            for (int bb = 0; bb < 2 * NB8; ++bb) {
                if (bb >= ha * NB8) {
                    const int which = ha ? 2 : (bb < NB8 ? 0 : 1);
                    const int bc = (bb % NB8) * 8 + 2 * (lane & 3);
                    if (ar < K) {
                        if (bc < K) out[(size_t)which * K * K + ar * K + bc] = acc[bb][0];
                        if (bc + 1 < K) out[(size_t)which * K * K + ar * K + bc + 1] = acc[bb][1];
                    }
                }
            }
        }
    }
}

// out[j] = sum over tiles of part[tile][j]  (fixed order).  A block takes 32 consecutive outputs (coalesced rows of the
// partials) and splits the tiles over 8 thread rows, whose sums are combined in a fixed order: with one thread per
// output and all tiles in sequence only half of the SMs had work (19 008 outputs at cfg 4) and each thread walked 1563
// rows alone.
constexpr int HGR_Y = 8;
__global__ void __launch_bounds__(32 * HGR_Y) hg_reduce_kernel(const double* __restrict__ part, int ntile, long long n,
                                                                double* __restrict__ out) {
    __shared__ double red[HGR_Y][32];
    const long long j = (long long)blockIdx.x * 32 + threadIdx.x;
    const int y = threadIdx.y;
    const int per = (ntile + HGR_Y - 1) / HGR_Y;
    const int t0 = y * per, t1 = min(ntile, t0 + per);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (j < n) {
        int t = t0;
        for (; t + 4 <= t1; t += 4) {
            a0 += part[(size_t)t * n + j];
            a1 += part[(size_t)(t + 1) * n + j];
            a2 += part[(size_t)(t + 2) * n + j];
            a3 += part[(size_t)(t + 3) * n + j];
        }
        for (; t < t1; ++t) a0 += part[(size_t)t * n + j];
    }
    red[y][threadIdx.x] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (y == 0 && j < n) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < HGR_Y; ++k) s += red[k][threadIdx.x];
        out[j] = s;
    }
}


}  // namespace plsb

using namespace plsb;

extern "C" size_t plsb200_half_gram_win_f64_workspace(int64_t p, int K, int ns) {
    if (p <= 0 || K <= 0 || ns <= 0) return 0;
    return (size_t)cdiv(p, HG_VT) * ns * 3 * K * K * sizeof(double);
}

extern "C" int plsb200_half_gram_win_f64(const double* Xstd, const double* Xlin, int64_t p, const int32_t* ids,
                                         const double* Q, const int32_t* segs, int nseg, int nq, int nmax, int K,
                                         int s0, int ns, double* S3, void* workspace, size_t workspace_bytes,
                                         void* stream) {
    PLSB_CHECK_ARG(Xstd && Xlin && ids && Q && segs && S3 && workspace, "half_gram_win_f64: null pointer");
    PLSB_CHECK_ARG(p > 0 && nseg > 0 && nq > 0 && (nq & 1) == 0 && nmax > 0 && K > 0 && ns > 0, "half_gram_win_f64: bad shape");
    if (K > 96) {      // the 2 x K rows of a 128-voxel tile must fit in shared memory (203 KB at K = 96)
        set_err("half_gram_win_f64: K=%d > 96 not supported", K);
        return PLSB200_EUNSUPPORTED;
    }
    const int ntile = (int)cdiv(p, HG_VT);
    const size_t need = (size_t)ntile * ns * 3 * K * K * sizeof(double);
    if (workspace_bytes < need) {
        set_err("half_gram_win_f64: workspace %zu < %zu bytes", workspace_bytes, need);
        return PLSB200_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
#define PLSB_HGW_LAUNCH(KCV)                                                                                      \
    do {                                                                                                          \
        size_t smem = ((size_t)2 * KCV * (HG_VT + 4) + (size_t)nq + (size_t)nmax) * sizeof(double) +        \
                      (size_t)2 * nseg * HG_SEG * sizeof(int);                                                    \
        if (smem > 227 * 1024) { set_err("half_gram_win_f64: halves too large for shared memory"); return PLSB200_EUNSUPPORTED; } \
        PLSB_CUDA(cudaFuncSetAttribute(half_gram_win_kernel<KCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        int per_sm = 1;                                                                                           \
        PLSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, half_gram_win_kernel<KCV>, HG_VT, smem)); \
        /* groups of splits: a CTA keeps its voxel tile for a whole group; 1563 tiles on 592 slots are 2.64 waves, */ \
        /* i.e. a third wave at 64 % -- with g groups the grid is g times longer and the tail g times shorter      */ \
        const double slots = (double)num_sms() * (per_sm > 0 ? per_sm : 1);                                       \
        int groups = 1;                                                                                           \
        double best = -1.0;                                                                                       \
        for (int c = 1; c <= ns && c <= 16; ++c) {                                                                \
            const double waves = (double)ntile * c / slots, eff = waves / ceil(waves);                            \
            if (eff > best + 0.02) { best = eff; groups = c; }                                                    \
        }                                                                                                         \
        half_gram_win_kernel<KCV><<<dim3(ntile, groups), HG_VT, smem, st>>>(Xstd, Xlin, p, ids, Q, segs, nseg, nq, \
                                                                            nmax, K, s0, ns, (double*)workspace); \
    } while (0)
    if (K <= 8) PLSB_HGW_LAUNCH(8);
    else if (K <= 16) PLSB_HGW_LAUNCH(16);
    else if (K <= 24) PLSB_HGW_LAUNCH(24);
    else if (K <= 32) PLSB_HGW_LAUNCH(32);
    else if (K <= 48) PLSB_HGW_LAUNCH(48);
    else if (K <= 64) PLSB_HGW_LAUNCH(64);
    else PLSB_HGW_LAUNCH(96);
#undef PLSB_HGW_LAUNCH
    PLSB_LAUNCH_CHECK("half_gram_win_kernel");
    const long long n = (long long)ns * 3 * K * K;
    hg_reduce_kernel<<<(unsigned)cdiv(n, 32), dim3(32, HGR_Y), 0, st>>>((const double*)workspace, ntile, n,
                                                             S3 + (size_t)s0 * 3 * K * K);
    PLSB_LAUNCH_CHECK("hg_reduce_kernel");
    return PLSB200_OK;
}

