// K4: bootstrap salience moments -- the dominant kernel of the path (>99% of the algorithmic flops).
//
// Reference (bootstrap_permutation.py:557-626, 695): per bootstrap, gather N rows of X, rebuild the
// K x p cross-block matrix, VS_hat = permuted.T @ U (p x K) is stored in right_sv_sampled[B, p, K] and
// std_errs = np.std(right_sv_sampled, axis=0) at the end.  Here VS_r = X^T . C_r with
// C_r = scatter(E, idx_r) (N x K), so all resamples form ONE skinny GEMM
//       VS[v, (r,k)] = sum_i X[i, v] . C[i, (r,k)]          M = p, N = R*K, K-dim = N rows
// whose output is reduced over r on the fly: sum_r (VS - pivot), sum_r (VS - pivot)^2.
//
// Mapping onto the FP64 tensor path (DMMA.8x8x4, 37 TFLOP/s measured on B200):
//   * M (8 per MMA) = voxels.  A warp owns 8 voxels and keeps ALL its A-fragments -- X[0..N, v0..v0+8) --
//     in registers for the whole kernel (N/4 doubles per thread: 75 for N = 300).  X is read from HBM
//     exactly once per CTA row-tile and never staged in shared memory.
//   * N (8 per MMA) = flattened (resample, k) columns.  The coefficient tensor is pre-packed in the
//     exact B-fragment order (lane-major 256-byte blocks), so the TMA engine streams it into shared
//     memory with plain 1-D bulk copies (cp.async.bulk / UBLKCP, mbarrier completion) and every
//     fragment load is one conflict-free LDS.64.  It is shared by the 8 warps of the CTA.
//   * A "period" is 8*NBLK columns = nb whole resamples (K=12 -> 24 columns = 2 resamples, no padding).
//     Within a period the NBLK accumulator chains are independent, which hides the DMMA latency.
//   * After N/4 chained MMAs a D fragment holds VS - pivot (the chain starts from -pivot) for 8 voxels
//     x 8 columns; it is folded into per-thread running sums and discarded.  Column -> k is fixed per
//     thread for the whole kernel because the period is a multiple of K.
// Shared memory: nstage x (N/4 * NBLK * 256 B) ring (57.6 KB per stage at N=300, K=12 -> 3 stages).
#include "common.cuh"

namespace plsb {

constexpr int BM_WARPS = 8;          // consumer warps per CTA
constexpr int BM_THREADS = BM_WARPS * 32;
constexpr int BM_VOX = BM_WARPS * 8; // voxels per CTA

struct BootPlan {
    int Kp;        // padded K (divides 8*nblk)
    int nblk;      // 8-column blocks per period
    int nb;        // resamples per period
    int nks;       // k-steps of 4 rows (multiple of 4)
    int maxks;     // template bucket
    int nper;      // periods
    int nstage;
    size_t stage_doubles;
    size_t smem_bytes;
    int nsplit;        // splits of the period range (grid.y)
    int per_per_split;
};

static bool boot_plan(int N, int K, int R, int64_t p, BootPlan& b) {
    if (K < 1 || K > 24 || N < 1 || R < 1) return false;
    // smallest padded K that divides a 24-, 16- or 8-column period; prefer wider periods (more ILP)
    int best_kp = 0, best_blk = 0;
    for (int blk = 3; blk >= 1; --blk) {
        const int cols = 8 * blk;
        for (int kp = K; kp <= cols; ++kp) {
            if (cols % kp == 0) {
                if (best_kp == 0 || kp < best_kp) { best_kp = kp; best_blk = blk; }
                break;
            }
        }
    }
    if (!best_kp) return false;
    b.Kp = best_kp; b.nblk = best_blk; b.nb = 8 * best_blk / best_kp;
    b.nks = (int)cdiv(N, 16) * 4;
    if (b.nks > 80) return false;
    b.maxks = b.nks;
    b.nper = (int)cdiv(R, b.nb);
    b.stage_doubles = (size_t)b.nks * b.nblk * 32;
    const size_t red = (size_t)BM_WARPS * 8 * b.Kp * 2 * sizeof(double);
    const size_t avail = 227 * 1024 - red - 256;
    int ns = (int)(avail / (b.stage_doubles * sizeof(double)));
    if (ns > 4) ns = 4;
    if (ns < 2) return false;
    b.nstage = ns;
    b.smem_bytes = ns * b.stage_doubles * sizeof(double) + red + 256;
    // split the resample range so that tiles*nsplit fills whole waves of SMs
    const int64_t tiles = cdiv(p > 0 ? p : 1, BM_VOX);
    const int nsm = num_sms();
    int best = 1; double best_cost = 1e30;
    for (int n = 1; n <= 8; ++n) {
        if (n > 1 && b.nper / n < 8) break;
        const double waves = (double)tiles * n / nsm;
        const double cost = ceil(waves) / waves + 0.004 * (n - 1);
        if (cost < best_cost - 1e-12) { best_cost = cost; best = n; }
    }
    b.nsplit = best;
    b.per_per_split = (int)cdiv(b.nper, best);
    b.nsplit = (int)cdiv(b.nper, b.per_per_split);
    return true;
}

// ------------------------------------------------------------------------------------------------
// coefficient packing: C_r = scatter(E, idx_r) written in B-fragment order.
// offset(per, s, jb, lane) = ((per*nks + s)*nblk + jb)*32 + lane,  lane = 4*n + q,
// column-in-period c = (r % nb)*Kp + k -> jb = c/8, n = c%8 ; row i -> s = i/4, q = i%4.
// One CTA per PERIOD (nb resamples): thread i scans idx for sources of target row i in index order (deterministic),
// the period's block is assembled in shared memory in fragment order and written out as one contiguous, coalesced
// piece.  (Round 1 wrote every coefficient straight to its fragment position: 8-byte stores at a stride of
// 256 nblk bytes, 25 % sector efficiency -- 0.42 ms for the 146 MB of the bench workload.)
__global__ void __launch_bounds__(256) boot_coef_pack_kernel(const double* __restrict__ E, int N, int K,
                                                            const int32_t* __restrict__ idx, int R, int Kp, int nblk,
                                                            int nb, int nks, double* __restrict__ coef) {
    extern __shared__ __align__(16) double smp[];
    const int stage_doubles = nks * nblk * 32;
    double* blk = smp;                                  // [nks][nblk][32]
    int* ids = reinterpret_cast<int*>(blk + stage_doubles);
    int* start = ids + N; int* cur = start + N + 1; int* list = cur + N;
    const int per = blockIdx.x;
    for (int i = threadIdx.x; i < stage_doubles; i += blockDim.x) blk[i] = 0.0;
    for (int slot = 0; slot < nb; ++slot) {
        const int r = per * nb + slot;
        if (r >= R) break;                              // ragged last period: the remaining slots stay zero
        __syncthreads();                                // previous users of the lists are done (and blk is zeroed)
        build_source_lists(idx + (size_t)r * N, N, ids, start, cur, list);
        const int cbase = slot * Kp;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            double acc[24];
#pragma unroll
            for (int k = 0; k < 24; ++k) acc[k] = 0.0;
            for (int t = start[i]; t < start[i + 1]; ++t) {
                const int src = list[t];
#pragma unroll
                for (int k = 0; k < 24; ++k)
                    if (k < K) acc[k] += __ldg(E + (size_t)src * K + k);
            }
            const int s = i >> 2, q = i & 3;
#pragma unroll
            for (int k = 0; k < 24; ++k) {
                if (k < K) {
                    const int c = cbase + k;
                    blk[(s * nblk + (c >> 3)) * 32 + 4 * (c & 7) + q] = acc[k];
                }
            }
        }
    }
    __syncthreads();
    double2* out = reinterpret_cast<double2*>(coef + (size_t)per * stage_doubles);
    const double2* in = reinterpret_cast<const double2*>(blk);
    for (int i = threadIdx.x; i < stage_doubles / 2; i += blockDim.x) out[i] = in[i];
}

// ------------------------------------------------------------------------------------------------
// NKS (k-steps of 4 rows) is a compile-time constant so that the whole period is one straight-line
// stream of NKS*NBLK {LDS.64, DMMA} pairs that ptxas can software-pipeline; kernels are instantiated
// for every multiple of 4 up to 80 (N <= 320).
template <int NKS, int NBLK>
__global__ void __launch_bounds__(BM_THREADS, 1)
boot_moments_kernel(const double* __restrict__ X, long long ldx, int N, long long p,
                    const double* __restrict__ coef, int nper, int per_per_split, int nstage,
                    int R, int Kp, int K, const double* __restrict__ pivot,
                    double* __restrict__ osum, double* __restrict__ osumsq) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int stage_doubles = NKS * NBLK * 32;
    constexpr uint32_t stage_bytes = (uint32_t)stage_doubles * 8u;
    double* ring = reinterpret_cast<double*>(smraw);
    double* red = ring + (size_t)nstage * stage_doubles;              // [8 warps][2][8][Kp]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + BM_WARPS * 16 * Kp);
    uint64_t* empty = full + nstage;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = lane & 3, vr = lane >> 2;
    const long long v = (long long)blockIdx.x * BM_VOX + warp * 8 + vr;
    const int per0 = blockIdx.y * per_per_split;
    const int per1 = min(nper, per0 + per_per_split);
    const int nit = per1 - per0;

    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, BM_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int it, int slot) {   // thread 0 only: stream period per0+it into `slot`
        mbar_expect_tx(full + slot, stage_bytes);
        const char* src = reinterpret_cast<const char*>(coef + (size_t)(per0 + it) * stage_doubles);
        char* dst = reinterpret_cast<char*>(ring + (size_t)slot * stage_doubles);
#pragma unroll 1
        for (uint32_t off = 0; off < stage_bytes; off += 16384u) {
            const uint32_t n = min(16384u, stage_bytes - off);
            bulk_g2s(dst + off, src + off, n, full + slot);
        }
    };
    if (tid == 0)
        for (int it = 0; it < min(nstage, nit); ++it) issue(it, it);

    // ---- A fragments: this warp's 8 voxels x all rows, resident in registers
    double a[NKS];
#pragma unroll
    for (int s = 0; s < NKS; ++s) {
        const int row = 4 * s + q;
        a[s] = (row < N && v < p) ? __ldg(X + (long long)row * ldx + v) : 0.0;
    }

    // ---- per-thread column bookkeeping (fixed for the whole kernel)
    double piv[NBLK][2], s1[NBLK][2], s2[NBLK][2];
#pragma unroll
    for (int j = 0; j < NBLK; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = 8 * j + 2 * q + e, k = c % Kp;
            piv[j][e] = (pivot != nullptr && k < K && v < p) ? __ldg(pivot + v * K + k) : 0.0;
            s1[j][e] = 0.0; s2[j][e] = 0.0;
        }
    const int nb = 8 * NBLK / Kp;

    // The two warps that share an SM sub-partition (w and w+4) run half a period apart, so that one of
    // them always has a full MMA stream in flight while the other folds its accumulators and waits on
    // the next stage (otherwise all eight warps drain the DMMA pipe at the same instant).
    if (warp >= 4) __nanosleep((unsigned)(NKS * NBLK * 8));

    int slot = 0, prev_slot = 0;
    uint32_t phase = 0, prev_phase = 0;
    for (int it = 0; it < nit; ++it) {
        if (tid == 0 && it > 0) {
            // refill the slot drained in the previous iteration with period it-1+nstage
            const int nx = it - 1 + nstage;
            if (nx < nit) {
                mbar_wait(empty + prev_slot, prev_phase);
                issue(nx, prev_slot);
            }
        }
        __syncwarp();
        mbar_wait(full + slot, phase);

        double d[NBLK][2];
#pragma unroll
        for (int j = 0; j < NBLK; ++j) { d[j][0] = -piv[j][0]; d[j][1] = -piv[j][1]; }
        // volatile: ptxas must keep these loads in program order (s-major, chains round-robin), which
        // keeps the NBLK accumulator chains interleaved; without it the straight-line stream is
        // re-ordered chain by chain (fewest live registers) and every DMMA waits on its predecessor.
        const volatile double* bs = ring + (size_t)slot * stage_doubles + lane;
#pragma unroll
        for (int s = 0; s < NKS; ++s) {
#pragma unroll
            for (int j = 0; j < NBLK; ++j) {
                const double b = bs[(s * NBLK + j) * 32];
                dmma884(d[j][0], d[j][1], a[s], b);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);

        if (it + 1 < nit || (per0 + it + 1) * nb <= R) {      // every slot of the period is a real resample
#pragma unroll
            for (int j = 0; j < NBLK; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    s1[j][e] += d[j][e];
                    s2[j][e] = fma(d[j][e], d[j][e], s2[j][e]);
                }
        } else {                                              // ragged last period: mask the padding slots
            const int rbase = (per0 + it) * nb;
#pragma unroll
            for (int j = 0; j < NBLK; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (rbase + (8 * j + 2 * q + e) / Kp < R) {
                        s1[j][e] += d[j][e];
                        s2[j][e] = fma(d[j][e], d[j][e], s2[j][e]);
                    }
                }
        }
        prev_slot = slot; prev_phase = phase;
        if (++slot == nstage) { slot = 0; phase ^= 1u; }
    }

    // ---- fold the nb resample slots of a period onto k, in slot order (deterministic)
    double* r1 = red + warp * (16 * Kp);
    double* r2 = r1 + 8 * Kp;
    for (int i = lane; i < 16 * Kp; i += 32) r1[i] = 0.0;
    __syncwarp();
    for (int round = 0; round < nb; ++round) {
#pragma unroll
        for (int j = 0; j < NBLK; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * j + 2 * q + e;
                if (c / Kp == round) {
                    r1[vr * Kp + c % Kp] += s1[j][e];
                    r2[vr * Kp + c % Kp] += s2[j][e];
                }
            }
        __syncwarp();
    }
    const long long vbase = (long long)blockIdx.x * BM_VOX + warp * 8;
    double* o1 = osum + (size_t)blockIdx.y * p * K;
    double* o2 = osumsq + (size_t)blockIdx.y * p * K;
    for (int i = lane; i < 8 * K; i += 32) {
        const int rr = i / K, k = i % K;
        if (vbase + rr < p) {
            o1[(vbase + rr) * K + k] = r1[rr * Kp + k];
            o2[(vbase + rr) * K + k] = r2[rr * Kp + k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The N-space pass (K2/K6) on the same machinery: H = G^T [C_1 | C_2 | ...] is the bootstrap GEMM with the Gram
// matrix as the data matrix (rows i, "voxels" j) and the same packed coefficient stream, and
//     d2[r,k] = C_r[:,k]^T H[:,(r,k)],      T[r,c,k] = (Lmat G C_r)[c,k] / sqrt(d2[r,k])
// come out of the epilogue: the data matrix is W = [G | 0 | (Lmat G)^T] (N rows x (Npad + Kt) columns, Npad = N rounded
// up to 8), so a warp's 8 "voxels" are either rows of H -- its D fragment is multiplied element-wise with the
// coefficients of those rows (read from the stage in shared memory), reduced over the 8 rows with shuffles and written
// as one partial per (voxel group, resample, k) -- or rows of T, which are stored as they are.  Partials are summed in
// a fixed order by nspace_finish_kernel (deterministic), which also applies 1/sqrt(d2) to T.
// Replaces nspace_kernel (FMA, bound by the shared-memory pipe: profiles/ncu_nspace_r02.md) for N <= 320.
struct NsArgs {
    const double* W; long long ldw; int N, Nv, Npad;
    const double* coef; int nper, per_per_split, nstage, R, Kp, K, Kt;
    double* dpart;      // [ngH][R][K]
    double* Traw;       // [R][Kt][K] or NULL
};

template <int NKS, int NBLK>
__global__ void __launch_bounds__(BM_THREADS, 1) nspace_dmma_kernel(const NsArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int stage_doubles = NKS * NBLK * 32;
    constexpr uint32_t stage_bytes = (uint32_t)stage_doubles * 8u;
    double* ring = reinterpret_cast<double*>(smraw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)a.nstage * stage_doubles);
    uint64_t* empty = full + a.nstage;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = lane & 3, vr = lane >> 2;
    const int group = blockIdx.x * BM_WARPS + warp;          // 8 consecutive columns of W
    const int v = group * 8 + vr;
    const int per0 = blockIdx.y * a.per_per_split;
    const int per1 = min(a.nper, per0 + a.per_per_split);
    const int nit = per1 - per0;

    if (tid == 0) {
        for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, BM_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int it, int slot) {
        mbar_expect_tx(full + slot, stage_bytes);
        const char* src = reinterpret_cast<const char*>(a.coef + (size_t)(per0 + it) * stage_doubles);
        char* dst = reinterpret_cast<char*>(ring + (size_t)slot * stage_doubles);
#pragma unroll 1
        for (uint32_t off = 0; off < stage_bytes; off += 16384u)
            bulk_g2s(dst + off, src + off, min(16384u, stage_bytes - off), full + slot);
    };
    if (tid == 0)
        for (int it = 0; it < min(a.nstage, nit); ++it) issue(it, it);

    double x[NKS];
#pragma unroll
    for (int s = 0; s < NKS; ++s) {
        const int row = 4 * s + q;
        x[s] = (row < a.N && v < a.Nv) ? __ldg(a.W + (long long)row * a.ldw + v) : 0.0;
    }
    const bool is_h = group * 8 < a.Npad;                    // warp-uniform: rows of H, else rows of T
    const int nb = 8 * NBLK / a.Kp;
    if (warp >= 4) __nanosleep((unsigned)(NKS * NBLK * 8));

    int slot = 0, prev_slot = 0;
    uint32_t phase = 0, prev_phase = 0;
    for (int it = 0; it < nit; ++it) {
        if (tid == 0 && it > 0) {
            const int nx = it - 1 + a.nstage;
            if (nx < nit) {
                mbar_wait(empty + prev_slot, prev_phase);
                issue(nx, prev_slot);
            }
        }
        __syncwarp();
        mbar_wait(full + slot, phase);
        double d[NBLK][2];
#pragma unroll
        for (int j = 0; j < NBLK; ++j) d[j][0] = d[j][1] = 0.0;
        const volatile double* bs = ring + (size_t)slot * stage_doubles + lane;
#pragma unroll
        for (int s = 0; s < NKS; ++s) {
#pragma unroll
            for (int j = 0; j < NBLK; ++j) {
                const double b = bs[(s * NBLK + j) * 32];
                dmma884(d[j][0], d[j][1], x[s], b);
            }
        }
        // the coefficients of this thread's own row (v) for its columns: C[v, c] sits at k-step v/4, element v%4
        double cv[NBLK][2];
        if (is_h) {
            const double* cs = ring + (size_t)slot * stage_doubles + (size_t)(v >> 2) * NBLK * 32 + (v & 3);
#pragma unroll
            for (int j = 0; j < NBLK; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) cv[j][e] = v < a.N ? cs[j * 32 + 4 * (2 * q + e)] : 0.0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);

        const int rbase = (per0 + it) * nb;
        if (is_h) {
#pragma unroll
            for (int j = 0; j < NBLK; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double pr = cv[j][e] * d[j][e];
                    pr += __shfl_xor_sync(0xffffffffu, pr, 4);
                    pr += __shfl_xor_sync(0xffffffffu, pr, 8);
                    pr += __shfl_xor_sync(0xffffffffu, pr, 16);
                    const int c = 8 * j + 2 * q + e, r = rbase + c / a.Kp, k = c % a.Kp;
                    if (vr == 0 && k < a.K && r < a.R) a.dpart[((size_t)group * a.R + r) * a.K + k] = pr;
                }
        } else if (a.Traw != nullptr) {
            const int ci = v - a.Npad;
#pragma unroll
            for (int j = 0; j < NBLK; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = 8 * j + 2 * q + e, r = rbase + c / a.Kp, k = c % a.Kp;
                    if (ci < a.Kt && k < a.K && r < a.R) a.Traw[((size_t)r * a.Kt + ci) * a.K + k] = d[j][e];
                }
        }
        prev_slot = slot; prev_phase = phase;
        if (++slot == a.nstage) { slot = 0; phase ^= 1u; }
    }
}

// W = [G | 0 | (Lmat G)^T]: N rows, ldw columns
__global__ void nspace_w_kernel(const double* __restrict__ G, int N, const double* __restrict__ Lmat, int Kt, int Npad,
                                long long ldw, double* __restrict__ W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * ldw) return;
    const int row = (int)(i / ldw), col = (int)(i % ldw);
    double val = 0.0;
    if (col < N) val = G[(size_t)row * N + col];
    else if (col >= Npad && col - Npad < Kt) {
        const double* l = Lmat + (size_t)(col - Npad) * N;
        for (int t = 0; t < N; ++t) val = fma(l[t], G[(size_t)t * N + row], val);
    }
    W[i] = val;
}

// d2[r,k] = sum over the voxel groups (fixed order); T[r,c,k] *= 1/sqrt(d2[r,k])
__global__ void nspace_finish_kernel(const double* __restrict__ dpart, int ngroups, int R, int K, int Kt,
                                     double* __restrict__ d2, double* __restrict__ T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)R * K) return;
    double s = 0.0;
    for (int g = 0; g < ngroups; ++g) s += dpart[(size_t)g * R * K + i];
    d2[i] = s;
    if (T != nullptr) {
        const double dn = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
        const int r = (int)(i / K), k = (int)(i % K);
        for (int c = 0; c < Kt; ++c) T[((size_t)r * Kt + c) * K + k] *= dn;
    }
}

__global__ void moments_reduce_kernel(const double* __restrict__ part1, const double* __restrict__ part2, int nsplit,
                                      long long n, double* __restrict__ sum, double* __restrict__ sumsq) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = 0.0, b = 0.0;
    for (int s = 0; s < nsplit; ++s) { a += part1[(size_t)s * n + i]; b += part2[(size_t)s * n + i]; }
    sum[i] = a; sumsq[i] = b;
}

__global__ void boot_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, long long n,
                                     double R, const double* __restrict__ numer, double* __restrict__ std_errs,
                                     double* __restrict__ ratios) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = sum[i] / R;
    double var = sumsq[i] / R - m * m;
    if (var < 0.0) var = 0.0;
    const double sd = sqrt(var);
    std_errs[i] = sd;
    if (ratios) ratios[i] = numer[i] / sd;
}

// explicit saliences VS[r][v][k] = sum_i X[idx_r[i], v] . E[i, k]  (reference order of operations;
// small problems / debugging only -- writes R*p*K doubles)
// SERIES: output [v][k][r] (the R samples of an element contiguous: the layout percentile_kernel sorts fastest) with the
// resample as the FAST grid index, so that the CTAs running at one time share their rows of X in L2 and write
// neighbouring 8-byte slots of the same sectors.
template <int KT, bool SERIES>
__global__ void __launch_bounds__(128) salience_kernel(const double* __restrict__ X, int N, long long p, long long ldx,
                                                      const double* __restrict__ E, int K,
                                                      const int32_t* __restrict__ idx, int R, double* __restrict__ VS) {
    extern __shared__ __align__(16) double sms[];
    double* Es = sms;
    int* ids = reinterpret_cast<int*>(Es + (size_t)N * K);
    const int r = SERIES ? blockIdx.x : blockIdx.y;
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) Es[i] = E[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) ids[i] = idx[(size_t)r * N + i];
    __syncthreads();
    const long long v = (long long)(SERIES ? blockIdx.y : blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= p) return;
    for (int k0 = 0; k0 < K; k0 += KT) {
        double acc[KT];
#pragma unroll
        for (int t = 0; t < KT; ++t) acc[t] = 0.0;
        for (int i = 0; i < N; ++i) {
            const double x = __ldg(X + (long long)ids[i] * ldx + v);
#pragma unroll
            for (int t = 0; t < KT; ++t)
                if (k0 + t < K) acc[t] = fma(x, Es[i * K + k0 + t], acc[t]);
        }
#pragma unroll
        for (int t = 0; t < KT; ++t)
            if (k0 + t < K) {
                if (SERIES) VS[((size_t)v * K + k0 + t) * R + r] = acc[t];
                else VS[((size_t)r * p + v) * K + k0 + t] = acc[t];
            }
    }
}

// XL = X . V : each CTA takes a chunk of XV_CHUNK voxels; the V chunk is staged transposed in shared memory
// ([k][voxel]: conflict-free), a warp owns XV_RG rows at a time so that every V value fetched from shared
// memory feeds XV_RG FMAs (2: 24 accumulators keep the kernel at three CTAs per SM), lanes stride over the voxels (coalesced 256-B row segments of X), and the per-lane
// partial sums are combined by shuffles in a fixed order (deterministic).  HBM-bound: X is read once.
constexpr int XV_CHUNK = 512;
constexpr int XV_RG = 2;
template <int KT>
__global__ void __launch_bounds__(256, 3) xv_partial_kernel(const double* __restrict__ X, int N, long long p, long long ldx,
                                                        const double* __restrict__ V, int K,
                                                        double* __restrict__ part) {
    extern __shared__ __align__(16) double smv[];   // [KT][XV_CHUNK]
    const long long v0 = (long long)blockIdx.x * XV_CHUNK;
    const int nv = (int)min((long long)XV_CHUNK, p - v0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int k0 = 0; k0 < K; k0 += KT) {
        const int kn = min(KT, K - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < XV_CHUNK * KT; i += blockDim.x) {
            const int c = i / KT, t = i % KT;
            smv[t * XV_CHUNK + c] = (c < nv && t < kn) ? V[(v0 + c) * K + k0 + t] : 0.0;
        }
        __syncthreads();
        for (int r0 = warp * XV_RG; r0 < N; r0 += nw * XV_RG) {
            double acc[XV_RG][KT];
#pragma unroll
            for (int r = 0; r < XV_RG; ++r)
#pragma unroll
                for (int t = 0; t < KT; ++t) acc[r][t] = 0.0;
            const double* x = X + (long long)r0 * ldx + v0;
#pragma unroll 2
            for (int c = lane; c < nv; c += 32) {
                double xr[XV_RG];
#pragma unroll
                for (int r = 0; r < XV_RG; ++r) xr[r] = (r0 + r < N) ? __ldg(x + (long long)r * ldx + c) : 0.0;
#pragma unroll
                for (int t = 0; t < KT; ++t) {
                    const double vv = smv[t * XV_CHUNK + c];
#pragma unroll
                    for (int r = 0; r < XV_RG; ++r) acc[r][t] = fma(xr[r], vv, acc[r][t]);
                }
            }
#pragma unroll
            for (int r = 0; r < XV_RG; ++r)
#pragma unroll
                for (int t = 0; t < KT; ++t) {
                    double sacc = acc[r][t];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
                    if (lane == 0 && r0 + r < N && t < kn)
                        part[((size_t)blockIdx.x * N + r0 + r) * K + k0 + t] = sacc;
                }
        }
    }
}

__global__ void xv_reduce_kernel(const double* __restrict__ part, int nchunk, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int c = 0; c < nchunk; ++c) s += part[(size_t)c * n + i];
    out[i] = s;
}

template <int NKS, int NBLK>
static int launch_moments(const BootPlan& b, const double* X, int N, int64_t p, int64_t ldx, const double* coef, int K,
                          int R, const double* pivot, double* o1, double* o2, cudaStream_t st) {
    PLSB_CUDA(cudaFuncSetAttribute(boot_moments_kernel<NKS, NBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)b.smem_bytes));
    dim3 grid((unsigned)cdiv(p, BM_VOX), (unsigned)b.nsplit);
    boot_moments_kernel<NKS, NBLK><<<grid, BM_THREADS, b.smem_bytes, st>>>(
        X, ldx, N, p, coef, b.nper, b.per_per_split, b.nstage, R, b.Kp, K, pivot, o1, o2);
    PLSB_LAUNCH_CHECK("boot_moments_kernel");
    return PLSB200_OK;
}

template <int NBLK>
static int dispatch_nks(const BootPlan& b, const double* X, int N, int64_t p, int64_t ldx, const double* coef, int K,
                        int R, const double* pivot, double* o1, double* o2, cudaStream_t st) {
    switch (b.nks) {
#define PLSB_CASE(n) case n: return launch_moments<n, NBLK>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
        PLSB_CASE(4) PLSB_CASE(8) PLSB_CASE(12) PLSB_CASE(16) PLSB_CASE(20) PLSB_CASE(24) PLSB_CASE(28)
        PLSB_CASE(32) PLSB_CASE(36) PLSB_CASE(40) PLSB_CASE(44) PLSB_CASE(48) PLSB_CASE(52) PLSB_CASE(56)
        PLSB_CASE(60) PLSB_CASE(64) PLSB_CASE(68) PLSB_CASE(72) PLSB_CASE(76) PLSB_CASE(80)
#undef PLSB_CASE
        default:
            set_err("boot_moments_f64: no kernel for nks=%d", b.nks);
            return PLSB200_EUNSUPPORTED;
    }
}

}  // namespace plsb

using namespace plsb;

// Tall designs (N > 320): the output-stationary kernel (boot_os.cu) unless PLSB200_TALL=rs selects the row-split one
// (boot_rs.cu, N <= 1280).  Read once per process: the pack and the GEMM must agree on the coefficient layout.
static bool tall_os() {
    static const char* const env = getenv("PLSB200_TALL");
    return !(env && env[0] == 'r');
}
// N <= 320: the register-resident kernel of this file unless PLSB200_EXACT=os routes every design through boot_os.cu
static bool short_os() {
    static const char* const env = getenv("PLSB200_EXACT");
    return env && env[0] == 'o';
}
static bool use_other(int N) { return N > 320 || short_os(); }
static bool use_os(int N) { return N > 320 ? tall_os() : true; }

extern "C" size_t plsb200_boot_coef_bytes(int N, int K, int R) {
    if (use_other(N)) return use_os(N) ? boot_os_coef_bytes(N, K, R) : boot_rs_coef_bytes(N, K, R);
    BootPlan b;
    if (!boot_plan(N, K, R, 1, b)) return 0;
    return (size_t)b.nper * b.stage_doubles * sizeof(double);
}

extern "C" int plsb200_boot_coef_pack_f64(const double* E, int N, int K, const int32_t* idx, int R, double* coef,
                                          void* stream) {
    PLSB_CHECK_ARG(E && idx && coef, "boot_coef_pack_f64: null pointer");
    if (use_other(N))
        return use_os(N) ? boot_os_pack(E, N, K, idx, R, coef, (cudaStream_t)stream)
                         : boot_rs_pack(E, N, K, idx, R, coef, (cudaStream_t)stream);
    BootPlan b;
    if (!boot_plan(N, K, R, 1, b)) {
        set_err("boot_coef_pack_f64: unsupported shape N=%d K=%d R=%d (need K<=24, N<=320)", N, K, R);
        return PLSB200_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    size_t smem = b.stage_doubles * sizeof(double) + (size_t)(4 * N + 1) * sizeof(int);
    PLSB_CUDA(cudaFuncSetAttribute(boot_coef_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    boot_coef_pack_kernel<<<b.nper, 256, smem, st>>>(E, N, K, idx, R, b.Kp, b.nblk, b.nb, b.nks, coef);
    PLSB_LAUNCH_CHECK("boot_coef_pack_kernel");
    return PLSB200_OK;
}

extern "C" size_t plsb200_boot_moments_f64_workspace(int N, int64_t p, int K, int R) {
    if (use_other(N)) return use_os(N) ? boot_os_workspace(N, p, K, R) : boot_rs_workspace(N, p, K, R);
    BootPlan b;
    if (!boot_plan(N, K, R, p, b)) return 0;
    return b.nsplit > 1 ? (size_t)2 * b.nsplit * p * K * sizeof(double) : 16;
}

extern "C" int plsb200_boot_moments_f64(const double* X, int N, int64_t p, int64_t ldx, const double* coef, int K, int R,
                                        const double* pivot, double* sum, double* sumsq, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    PLSB_CHECK_ARG(X && coef && sum && sumsq, "boot_moments_f64: null pointer");
    PLSB_CHECK_ARG(p > 0 && ldx >= p, "boot_moments_f64: bad shape p=%lld ldx=%lld", (long long)p, (long long)ldx);
    if (use_other(N))
        return use_os(N) ? boot_os_moments(X, N, p, ldx, coef, K, R, pivot, sum, sumsq, workspace, workspace_bytes,
                                           (cudaStream_t)stream)
                         : boot_rs_moments(X, N, p, ldx, coef, K, R, pivot, sum, sumsq, workspace, workspace_bytes,
                                           (cudaStream_t)stream);
    BootPlan b;
    if (!boot_plan(N, K, R, p, b)) {
        set_err("boot_moments_f64: unsupported shape N=%d K=%d R=%d (need K<=24, N<=320)", N, K, R);
        return PLSB200_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double *o1 = sum, *o2 = sumsq;
    if (b.nsplit > 1) {
        size_t need = (size_t)2 * b.nsplit * p * K * sizeof(double);
        if (!workspace || workspace_bytes < need) {
            set_err("boot_moments_f64: workspace %zu < %zu bytes", workspace_bytes, need);
            return PLSB200_EWORKSPACE;
        }
        o1 = (double*)workspace;
        o2 = o1 + (size_t)b.nsplit * p * K;
    }
    int rc;
    switch (b.nblk) {
        case 1: rc = dispatch_nks<1>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st); break;
        case 2: rc = dispatch_nks<2>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st); break;
        default: rc = dispatch_nks<3>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st); break;
    }
    if (rc != PLSB200_OK) return rc;
    if (b.nsplit > 1) {
        const long long n = (long long)p * K;
        moments_reduce_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(o1, o2, b.nsplit, n, sum, sumsq);
        PLSB_LAUNCH_CHECK("moments_reduce_kernel");
    }
    return PLSB200_OK;
}

extern "C" int plsb200_boot_finalize_f64(const double* sum, const double* sumsq, int64_t p, int K, int64_t R_total,
                                         const double* numer, double* std_errs, double* boot_ratios, void* stream) {
    PLSB_CHECK_ARG(sum && sumsq && std_errs, "boot_finalize_f64: null pointer");
    PLSB_CHECK_ARG(p > 0 && K > 0 && R_total > 0, "boot_finalize_f64: bad shape");
    PLSB_CHECK_ARG(boot_ratios == nullptr || numer != nullptr, "boot_finalize_f64: ratios requested without numerator");
    const long long n = (long long)p * K;
    boot_finalize_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(sum, sumsq, n, (double)R_total, numer,
                                                                                  std_errs, boot_ratios);
    PLSB_LAUNCH_CHECK("boot_finalize_kernel");
    return PLSB200_OK;
}

static int salience_launch(const double* X, int N, int64_t p, int64_t ldx, const double* E, int K, const int32_t* idx, int R,
                           double* VS, bool series, void* stream);

extern "C" int plsb200_salience_f64(const double* X, int N, int64_t p, int64_t ldx, const double* E, int K,
                                    const int32_t* idx, int R, double* VS, void* stream) {
    return salience_launch(X, N, p, ldx, E, K, idx, R, VS, false, stream);
}

extern "C" int plsb200_salience_series_f64(const double* X, int N, int64_t p, int64_t ldx, const double* E, int K,
                                           const int32_t* idx, int R, double* VS, void* stream) {
    return salience_launch(X, N, p, ldx, E, K, idx, R, VS, true, stream);
}

static int salience_launch(const double* X, int N, int64_t p, int64_t ldx, const double* E, int K, const int32_t* idx, int R,
                           double* VS, bool series, void* stream) {
    PLSB_CHECK_ARG(X && E && idx && VS, "salience_f64: null pointer");
    PLSB_CHECK_ARG(N > 0 && p > 0 && K > 0 && R >= 0 && ldx >= p, "salience_f64: bad shape");
    if (R == 0) return PLSB200_OK;
    size_t smem = (size_t)N * K * sizeof(double) + (size_t)N * sizeof(int);
    if (smem > 200 * 1024) {
        set_err("salience_f64: N*K too large");
        return PLSB200_EUNSUPPORTED;
    }
    if (series) {
        PLSB_CHECK_ARG(cdiv(p, 128) <= 65535, "salience_series_f64: more than 65535 voxel blocks per call");
        PLSB_CUDA(cudaFuncSetAttribute(salience_kernel<12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)R, (unsigned)cdiv(p, 128));
        salience_kernel<12, true><<<grid, 128, smem, (cudaStream_t)stream>>>(X, N, p, ldx, E, K, idx, R, VS);
    } else {
        PLSB_CHECK_ARG(R <= 65535, "salience_f64: more than 65535 resamples per call");
        PLSB_CUDA(cudaFuncSetAttribute(salience_kernel<12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)cdiv(p, 128), (unsigned)R);
        salience_kernel<12, false><<<grid, 128, smem, (cudaStream_t)stream>>>(X, N, p, ldx, E, K, idx, R, VS);
    }
    PLSB_LAUNCH_CHECK("salience_kernel");
    return PLSB200_OK;
}

extern "C" size_t plsb200_xv_f64_workspace(int N, int64_t p, int K) {
    if (N <= 0 || p <= 0 || K <= 0) return 0;
    return (size_t)cdiv(p, XV_CHUNK) * N * K * sizeof(double);
}

extern "C" int plsb200_xv_f64(const double* X, int N, int64_t p, int64_t ldx, const double* V, int K, double* XL,
                              void* workspace, size_t workspace_bytes, void* stream) {
    PLSB_CHECK_ARG(X && V && XL && workspace, "xv_f64: null pointer");
    PLSB_CHECK_ARG(N > 0 && p > 0 && K > 0 && ldx >= p, "xv_f64: bad shape");
    const int nchunk = (int)cdiv(p, XV_CHUNK);
    size_t need = (size_t)nchunk * N * K * sizeof(double);
    if (workspace_bytes < need) {
        set_err("xv_f64: workspace %zu < %zu bytes", workspace_bytes, need);
        return PLSB200_EWORKSPACE;
    }
    size_t smem = (size_t)XV_CHUNK * 12 * sizeof(double);
    cudaStream_t st = (cudaStream_t)stream;
    PLSB_CUDA(cudaFuncSetAttribute(xv_partial_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xv_partial_kernel<12><<<nchunk, 256, smem, st>>>(X, N, p, ldx, V, K, (double*)workspace);
    PLSB_LAUNCH_CHECK("xv_partial_kernel");
    xv_reduce_kernel<<<(unsigned)cdiv((int64_t)N * K, 256), 256, 0, st>>>((const double*)workspace, nchunk, N * K, XL);
    PLSB_LAUNCH_CHECK("xv_reduce_kernel");
    return PLSB200_OK;
}

// ---- N-space pass on DMMA (see nspace_dmma_kernel) -----------------------------------------------------------
namespace plsb {
struct NsLayout { int Npad, Nv, ngH; long long ldw; size_t off_w, off_dpart, total; };
static NsLayout ns_layout(int N, int K, int Kt, int R) {
    NsLayout L;
    L.Npad = (N + 7) / 8 * 8;
    L.Nv = L.Npad + Kt;
    L.ldw = (L.Nv + 7) / 8 * 8;
    L.ngH = L.Npad / 8;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    L.off_w = 0;
    L.off_dpart = al((size_t)N * L.ldw * sizeof(double));
    L.total = L.off_dpart + al((size_t)L.ngH * R * K * sizeof(double));
    return L;
}

template <int NKS, int NBLK>
static int ns_launch(const BootPlan& b, const NsArgs& a, int tiles, cudaStream_t st) {
    PLSB_CUDA(cudaFuncSetAttribute(nspace_dmma_kernel<NKS, NBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)b.smem_bytes));
    dim3 grid((unsigned)tiles, (unsigned)b.nsplit);
    nspace_dmma_kernel<NKS, NBLK><<<grid, BM_THREADS, b.smem_bytes, st>>>(a);
    PLSB_LAUNCH_CHECK("nspace_dmma_kernel");
    return PLSB200_OK;
}
template <int NBLK>
static int ns_dispatch(const BootPlan& b, const NsArgs& a, int tiles, cudaStream_t st) {
    switch (b.nks) {
#define PLSB_CASE(n) case n: return ns_launch<n, NBLK>(b, a, tiles, st);
        PLSB_CASE(4) PLSB_CASE(8) PLSB_CASE(12) PLSB_CASE(16) PLSB_CASE(20) PLSB_CASE(24) PLSB_CASE(28)
        PLSB_CASE(32) PLSB_CASE(36) PLSB_CASE(40) PLSB_CASE(44) PLSB_CASE(48) PLSB_CASE(52) PLSB_CASE(56)
        PLSB_CASE(60) PLSB_CASE(64) PLSB_CASE(68) PLSB_CASE(72) PLSB_CASE(76) PLSB_CASE(80)
#undef PLSB_CASE
        default: set_err("nspace_dmma_f64: no kernel for nks=%d", b.nks); return PLSB200_EUNSUPPORTED;
    }
}
}  // namespace plsb

extern "C" size_t plsb200_nspace_dmma_f64_workspace(int N, int K, int Kt, int R) {
    BootPlan b;
    if (N > 320 || short_os() || Kt < 0 || !boot_plan(N, K, R, 1, b)) return 0;
    return ns_layout(N, K, Kt, R).total;
}

extern "C" int plsb200_nspace_dmma_f64(const double* G, int N, const double* Lmat, int Kt, const double* coef, int K,
                                       int R, double* d2, double* T, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    PLSB_CHECK_ARG(G && coef && d2 && workspace, "nspace_dmma_f64: null pointer");
    PLSB_CHECK_ARG((T == nullptr && Kt == 0) || (T != nullptr && Lmat != nullptr && Kt > 0),
                   "nspace_dmma_f64: T, Lmat and Kt go together");
    PLSB_CHECK_ARG(N > 0 && K > 0 && R > 0, "nspace_dmma_f64: bad shape");
    BootPlan b;
    if (N > 320 || short_os() || !boot_plan(N, K, R, 1, b)) {
        set_err("nspace_dmma_f64: unsupported shape N=%d K=%d (need N<=320, K<=24): use nspace_f64", N, K);
        return PLSB200_EUNSUPPORTED;
    }
    const NsLayout L = ns_layout(N, K, Kt, R);
    if (workspace_bytes < L.total) {
        set_err("nspace_dmma_f64: workspace %zu < %zu bytes", workspace_bytes, L.total);
        return PLSB200_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double* W = (double*)((char*)workspace + L.off_w);
    double* dpart = (double*)((char*)workspace + L.off_dpart);
    {
        const long long n = (long long)N * L.ldw;
        nspace_w_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(G, N, Lmat, Kt, L.Npad, L.ldw, W);
        PLSB_LAUNCH_CHECK("nspace_w_kernel");
    }
    // the grid: voxel tiles of 64 columns of W x splits of the period range that fill the SMs
    const int tiles = (int)cdiv(L.Nv, BM_VOX);
    {
        const int nsm = num_sms();
        int best = 1; double best_cost = 1e30;
        for (int n = 1; n <= 64; ++n) {
            if (n > 1 && b.nper / n < 4) break;
            const double waves = (double)tiles * n / nsm;
            const double cost = ceil(waves) / waves + 0.002 * (n - 1);
            if (cost < best_cost - 1e-12) { best_cost = cost; best = n; }
        }
        b.per_per_split = (int)cdiv(b.nper, best);
        b.nsplit = (int)cdiv(b.nper, b.per_per_split);
    }
    NsArgs a;
    a.W = W; a.ldw = L.ldw; a.N = N; a.Nv = L.Nv; a.Npad = L.Npad; a.coef = coef; a.nper = b.nper;
    a.per_per_split = b.per_per_split; a.nstage = b.nstage; a.R = R; a.Kp = b.Kp; a.K = K; a.Kt = Kt;
    a.dpart = dpart; a.Traw = T;
    int rc;
    switch (b.nblk) {
        case 1: rc = ns_dispatch<1>(b, a, tiles, st); break;
        case 2: rc = ns_dispatch<2>(b, a, tiles, st); break;
        default: rc = ns_dispatch<3>(b, a, tiles, st); break;
    }
    if (rc != PLSB200_OK) return rc;
    const long long n = (long long)R * K;
    nspace_finish_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(dpart, L.ngH, R, K, Kt, d2, T);
    PLSB_LAUNCH_CHECK("nspace_finish_kernel");
    return PLSB200_OK;
}
