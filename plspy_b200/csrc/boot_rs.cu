// K4 for tall designs (N > 320 rows): row-split variant of boot_moments_kernel.
//
// The A fragments of 8 voxels x N rows no longer fit one warp's registers, so RS (2 or 4) warps share a
// voxel group and each keeps N/RS rows resident (NKS k-steps of 4 rows).  As in boot_moments_kernel a warp works
// on a whole PERIOD of 8*NACC columns at a time, i.e. NACC independent accumulator chains that are kept
// round-robin so that the DMMA pipe always has NACC MMAs in flight per warp (a single chain per warp -- the first
// version of this kernel, one 8-column block per stage -- ran at 0.63 of the DGEMM peak, bound by the DMMA latency).
// A period's coefficients for all rows (RS*NKS*NACC*256 B: 230 KB at N = 1200, K = 24) exceed shared memory, so
// a period is streamed as NSUB = 4 pipeline stages of CK = NKS/4 k-steps per row chunk; the accumulators live
// across the four stages.  Once per period the RS partial D fragments of a voxel group are exchanged through shared
// memory (double-buffered, one named barrier per period and voxel group) and the warp with row-chunk 0 folds
// (VS - pivot) into the running moments.
//
// L2 traffic: with the A fragments in registers an SM holds only 16 (RS = 4) or 32 (RS = 2) voxels, so every SM
// re-reads the whole coefficient stream for very few voxels: 246 KB per 3.9 us period and SM = 9.3 TB/s over the chip
// at N = 1200 -- the L2 cannot deliver that (ncu: 12 % of the warp samples wait for a stage to land, DMMA pipe 77 %).
// The kernel therefore runs as thread-block CLUSTERS of CL (default 2) CTAs on adjacent voxel tiles that share
// one coefficient stream: every CTA fetches 1/CL of each stage and MULTICASTS it into the shared memory of all CTAs
// of the cluster (cp.async.bulk ... .multicast::cluster), each CTA's `full` barrier collecting the bytes of all CL
// shares; a slot is refilled once every warp of every CTA of the cluster has drained it (remote mbarrier arrives).
#include "common.cuh"

namespace plsb {

static int rs_nsub() { return 4; }       // pipeline stages per period

struct RsPlan {
    int Kp, nacc, nb, rs, nks, ck, nsub, nper, nstage, nsplit, per_per_split, vox, cl;
    size_t stage_doubles, smem_bytes;
};

// CTAs per cluster.  Measured at N = 1200 (RS = 4), 1250 bootstraps x 10^6 voxels per GPU, TFLOP/s at (SM clock, board
// power): no cluster 23.4-24.8 (1822 MHz under sw_power_cap, 1006 W) and 28.0 on a cool box; pairs 24.7-27.1 (1965 MHz,
// 859 W); clusters of four 24.7 (1965 MHz, 681 W).  Multicast takes a third of the board power out (the L2 reads) but
// the SM-side ingest of a stage stays on the critical path; pairs are the default.  PLSB200_RS_CLUSTER=1|2|4 overrides.
static int rs_cluster_size(int rs) {
    static const char* const env = getenv("PLSB200_RS_CLUSTER");
    if (env && (env[0] == '1' || env[0] == '2' || env[0] == '4')) return env[0] - '0';
    (void)rs;
    return 2;
}

static bool rs_plan(int N, int K, int R, int64_t p, RsPlan& b) {
    if (K < 1 || K > 24 || N <= 320 || R < 1) return false;
    int best_kp = 0, best_blk = 0;
    for (int blk = 3; blk >= 1; --blk) {
        const int cols = 8 * blk;
        for (int kp = K; kp <= cols; ++kp)
            if (cols % kp == 0) { if (best_kp == 0 || kp < best_kp) { best_kp = kp; best_blk = blk; } break; }
    }
    if (!best_kp) return false;
    b.Kp = best_kp; b.nacc = best_blk; b.nb = 8 * best_blk / best_kp;
    const int ksteps = (int)cdiv(N, 4);
    b.rs = ksteps <= 160 ? 2 : 4;
    int nks = (int)cdiv(cdiv(ksteps, b.rs), 8) * 8;      // buckets of 8 k-steps: 48, 56, 64, 72, 80
    if (nks < 48) nks = 48;
    if (nks > 80) return false;                          // N > 1280
    b.nks = nks; b.nsub = rs_nsub(); b.ck = nks / b.nsub;
    b.nper = (int)cdiv(R, b.nb);
    b.stage_doubles = (size_t)b.rs * b.ck * b.nacc * 32;
    b.vox = 8 * (8 / b.rs);
    const size_t extra = (size_t)(8 / b.rs) * 16 * b.Kp * 8 +
                         (size_t)2 * (8 / b.rs) * (b.rs - 1) * 64 * b.nacc * 8 + 256;
    int ns = (int)((227 * 1024 - extra) / (b.stage_doubles * 8));
    if (ns > 8) ns = 8;
    if (ns < 2) return false;
    b.nstage = ns;
    b.smem_bytes = ns * b.stage_doubles * 8 + extra;
    b.cl = rs_cluster_size(b.rs);
    const int64_t tiles = cdiv(cdiv(p > 0 ? p : 1, b.vox), b.cl) * b.cl;     // whole clusters
    const int nsm = num_sms();
    int best = 1; double best_cost = 1e30;
    for (int n = 1; n <= 8; ++n) {
        if (n > 1 && b.nper / n < 8) break;
        const double waves = (double)tiles * n / nsm;
        const double cost = ceil(waves) / waves + 0.004 * (n - 1);
        if (cost < best_cost - 1e-12) { best_cost = cost; best = n; }
    }
    b.per_per_split = (int)cdiv(b.nper, best);
    b.nsplit = (int)cdiv(b.nper, b.per_per_split);
    return true;
}

// packed layout: offset(per, sub, rc, s_in, jb, lane) = ((((per*NSUB + sub)*rs + rc)*ck + s_in)*nacc + jb)*32 + lane
// column-in-period c = (r % nb)*Kp + k -> jb = c/8, n = c%8 ; row i -> chunk rc = i / (4*nks), k-step of the chunk
// s = (i % (4*nks))/4 -> sub = s / ck, s_in = s % ck ; q = i%4 ; lane = 4n + q
__global__ void __launch_bounds__(256) boot_rs_pack_kernel(const double* __restrict__ E, int N, int K,
                                                          const int32_t* __restrict__ idx, int Kp, int nacc, int nb,
                                                          int rs, int nks, int ck, int nsub, double* __restrict__ coef) {
    extern __shared__ int ids[];          // E (N x K) stays in global memory: L1/L2-resident, read via __ldg
    int* start = ids + N; int* cur = start + N + 1; int* list = cur + N;
    const int r = blockIdx.x;
    build_source_lists(idx + (size_t)r * N, N, ids, start, cur, list);
    const int per = r / nb, cbase = (r % nb) * Kp;
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
        const int rc = i / (4 * nks), s = (i % (4 * nks)) >> 2, q = i & 3;
        const int sub = s / ck, s_in = s % ck;
        const size_t base = ((((size_t)per * nsub + sub) * rs + rc) * ck + s_in) * nacc;
#pragma unroll
        for (int k = 0; k < 24; ++k)
            if (k < K) {
                const int c = cbase + k;
                coef[(base + (c >> 3)) * 32 + 4 * (c & 7) + q] = acc[k];
            }
    }
}

template <int NKS, int NACC, int RS, int NSUB>
__global__ void __launch_bounds__(256, 1)
boot_moments_rs_kernel(const double* __restrict__ X, long long ldx, int N, long long p,
                       const double* __restrict__ coef, int nper, int per_per_split, int nstage, int R, int Kp, int K,
                       const double* __restrict__ pivot, double* __restrict__ osum, double* __restrict__ osumsq, int cl) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int NVG = 8 / RS;                    // voxel groups per CTA
    constexpr int CK = NKS / NSUB;              // k-steps per row chunk and stage
    constexpr int stage_doubles = RS * CK * NACC * 32;
    constexpr uint32_t stage_bytes = (uint32_t)stage_doubles * 8u;
    double* ring = reinterpret_cast<double*>(smraw);
    double* red = ring + (size_t)nstage * stage_doubles;                 // [NVG][2][8][Kp]
    double* exch = red + NVG * 16 * Kp;                                  // [2][NVG][RS-1][NACC][32][2]
    uint64_t* full = reinterpret_cast<uint64_t*>(exch + 2 * NVG * (RS - 1) * 64 * NACC);
    uint64_t* empty = full + nstage;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int vg = warp / RS, rc = warp % RS;
    const int q = lane & 3, vr = lane >> 2;
    const long long v = (long long)blockIdx.x * (NVG * 8) + vg * 8 + vr;
    const int per0 = blockIdx.y * per_per_split;
    const int per1 = min(nper, per0 + per_per_split);
    const int nit = (per1 - per0) * NSUB;       // pipeline stages of this CTA

    const uint32_t crank = cl > 1 ? cluster_ctarank() : 0u;
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8 * cl); }
        mbar_fence_init();
    }
    if (cl > 1) cluster_sync_all(); else __syncthreads();      // every CTA's barriers exist before a peer touches them
    // thread 0: this CTA's share (1 / cl) of stage `it`, delivered to every CTA of the cluster; the CTA's own `full`
    // barrier expects the whole stage (its own share plus the peers' multicasts)
    const uint32_t share = stage_bytes / (uint32_t)cl;
    const uint16_t mask = (uint16_t)((1u << cl) - 1u);
    auto issue = [&](int it, int slot) {
        mbar_expect_tx(full + slot, stage_bytes);
        const char* src = reinterpret_cast<const char*>(coef + ((size_t)per0 * NSUB + it) * stage_doubles) + crank * share;
        char* dst = reinterpret_cast<char*>(ring + (size_t)slot * stage_doubles) + crank * share;
#pragma unroll 1
        for (uint32_t off = 0; off < share; off += 16384u) {
            const uint32_t n = min(16384u, share - off);
            if (cl > 1) bulk_g2s_multicast(dst + off, src + off, n, full + slot, mask);
            else bulk_g2s(dst + off, src + off, n, full + slot);
        }
    };
    if (tid == 0)
        for (int it = 0; it < min(nstage, nit); ++it) issue(it, it);
    // Refills: thread 0 (also a consumer: the register file holds exactly eight warps of resident fragments, there is no
    // room for a producer warp) refills, at the start of each of its stages, the slot drained in the previous stage,
    // after waiting for every warp of the cluster to have left it.  Measured alternatives at N = 1200 (TFLOP/s, same
    // box): a non-blocking probe of the `empty` barrier at stage boundaries 24.6, also in mid-stage 22.8, refill by
    // whichever warp drains the slot last 27.3, this scheme 28.0; stages of half the size (ring twice as deep)
    // 19.9-22.7: what is on the critical path is the refill latency of a 61 KB stage, not lock-step between warps.

    double a[NKS];
#pragma unroll
    for (int s = 0; s < NKS; ++s) {
        const int row = (rc * NKS + s) * 4 + q;
        a[s] = (row < N && v < p) ? __ldg(X + (long long)row * ldx + v) : 0.0;
    }
    double piv[NACC][2], s1[NACC][2], s2[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = 8 * j + 2 * q + e, k = c % Kp;
            piv[j][e] = (rc == 0 && pivot != nullptr && k < K && v < p) ? __ldg(pivot + v * K + k) : 0.0;
            s1[j][e] = 0.0; s2[j][e] = 0.0;
        }
    const int nb = 8 * NACC / Kp;

    // the two warps of an SM sub-partition (w and w + 4) run about one stage apart, so that one of them has its MMA
    // stream in flight while the other exchanges / folds its accumulators at a period boundary
    if (warp >= 4) __nanosleep((unsigned)(CK * NACC * 8));

    int slot = 0, prev_slot = 0, it = 0;
    uint32_t phase = 0, prev_phase = 0;
    const uint32_t empty_base = smem_u32(empty);
    for (int per = per0; per < per1; ++per) {
        double d[NACC][2];
#pragma unroll
        for (int j = 0; j < NACC; ++j) { d[j][0] = -piv[j][0]; d[j][1] = -piv[j][1]; }
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub, ++it) {
            if (tid == 0 && it > 0) {
                const int nx = it - 1 + nstage;
                if (nx < nit) {
                    mbar_wait(empty + prev_slot, prev_phase);
                    issue(nx, prev_slot);
                }
            }
            __syncwarp();
            mbar_wait(full + slot, phase);
            // volatile: keeps the loads in program order (k-step major, chains round-robin), see boot_moments_kernel
            const volatile double* bs = ring + (size_t)slot * stage_doubles + (size_t)rc * CK * NACC * 32 + lane;
#pragma unroll
            for (int s = 0; s < CK; ++s) {
#pragma unroll
                for (int j = 0; j < NACC; ++j) {
                    const double b = bs[(s * NACC + j) * 32];
                    dmma884(d[j][0], d[j][1], a[sub * CK + s], b);
                }
            }
            __syncwarp();
            if (cl == 1) {
                if (lane == 0) mbar_arrive(empty + slot);
            } else if (lane < cl) {                                  // one arrive per CTA of the cluster
                mbar_arrive_cluster(mapa_u32(empty_base + (uint32_t)slot * 8u, (uint32_t)lane));
            }
            prev_slot = slot; prev_phase = phase;
            if (++slot == nstage) { slot = 0; phase ^= 1u; }
        }
        // combine the RS partial fragments of this voxel group, once per period
        double* ex = exch + (size_t)(((per - per0) & 1) * NVG + vg) * (RS - 1) * 64 * NACC;
        if (rc > 0) {
#pragma unroll
            for (int j = 0; j < NACC; ++j)
                *reinterpret_cast<double2*>(ex + ((size_t)((rc - 1) * NACC + j) * 32 + lane) * 2) =
                    make_double2(d[j][0], d[j][1]);
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + vg), "r"(RS * 32) : "memory");
        if (rc == 0) {
#pragma unroll
            for (int c = 0; c < RS - 1; ++c)
#pragma unroll
                for (int j = 0; j < NACC; ++j) {
                    const double2 o = *reinterpret_cast<const double2*>(ex + ((size_t)(c * NACC + j) * 32 + lane) * 2);
                    d[j][0] += o.x; d[j][1] += o.y;
                }
            const int rbase = per * nb;
            if ((per + 1) * nb <= R) {                        // every slot of the period is a real resample
#pragma unroll
                for (int j = 0; j < NACC; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) { s1[j][e] += d[j][e]; s2[j][e] = fma(d[j][e], d[j][e], s2[j][e]); }
            } else {                                          // ragged last period: mask the padding slots
#pragma unroll
                for (int j = 0; j < NACC; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (rbase + (8 * j + 2 * q + e) / Kp < R) {
                            s1[j][e] += d[j][e]; s2[j][e] = fma(d[j][e], d[j][e], s2[j][e]);
                        }
            }
        }
    }
    if (rc == 0) {
    double* r1 = red + vg * (16 * Kp);
    double* r2 = r1 + 8 * Kp;
    for (int i = lane; i < 16 * Kp; i += 32) r1[i] = 0.0;
    __syncwarp();
    for (int round = 0; round < nb; ++round) {
#pragma unroll
        for (int j = 0; j < NACC; ++j)
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
    const long long vbase = (long long)blockIdx.x * (NVG * 8) + vg * 8;
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
    // no CTA may exit while a peer can still multicast into its shared memory or arrive on its barriers
    if (cl > 1) cluster_sync_all();
}

__global__ void rs_moments_reduce_kernel(const double* __restrict__ p1, const double* __restrict__ p2, int nsplit,
                                         long long n, double* __restrict__ sum, double* __restrict__ sumsq) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = 0.0, b = 0.0;
    for (int s = 0; s < nsplit; ++s) { a += p1[(size_t)s * n + i]; b += p2[(size_t)s * n + i]; }
    sum[i] = a; sumsq[i] = b;
}

template <int NKS, int NACC, int RS, int NSUB>
static int rs_launch(const RsPlan& b, const double* X, int N, int64_t p, int64_t ldx, const double* coef, int K, int R,
                     const double* pivot, double* o1, double* o2, cudaStream_t st) {
    PLSB_CUDA(cudaFuncSetAttribute(boot_moments_rs_kernel<NKS, NACC, RS, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)b.smem_bytes));
    const unsigned gx = (unsigned)(cdiv(cdiv(p, b.vox), b.cl) * b.cl);           // whole clusters (surplus CTAs: v >= p)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, (unsigned)b.nsplit); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = b.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)b.cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = b.cl > 1 ? 1 : 0;
    const long long ldx_ll = (long long)ldx, p_ll = (long long)p;
    PLSB_CUDA(cudaLaunchKernelEx(&cfg, boot_moments_rs_kernel<NKS, NACC, RS, NSUB>, X, ldx_ll, N, p_ll, coef, b.nper,
                                 b.per_per_split, b.nstage, R, b.Kp, K, pivot, o1, o2, b.cl));
    PLSB_LAUNCH_CHECK("boot_moments_rs_kernel");
    return PLSB200_OK;
}

template <int NACC, int RS>
static int rs_dispatch_nks(const RsPlan& b, const double* X, int N, int64_t p, int64_t ldx, const double* coef, int K,
                           int R, const double* pivot, double* o1, double* o2, cudaStream_t st) {
    switch (b.nks) {
        case 48: return rs_launch<48, NACC, RS, 4>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
        case 56: return rs_launch<56, NACC, RS, 4>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
        case 64: return rs_launch<64, NACC, RS, 4>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
        case 72: return rs_launch<72, NACC, RS, 4>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
        case 80: return rs_launch<80, NACC, RS, 4>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
        default: set_err("boot_moments_f64: no row-split kernel for nks=%d", b.nks); return PLSB200_EUNSUPPORTED;
    }
}

template <int RS>
static int rs_dispatch_nacc(const RsPlan& b, const double* X, int N, int64_t p, int64_t ldx, const double* coef, int K,
                            int R, const double* pivot, double* o1, double* o2, cudaStream_t st) {
    switch (b.nacc) {
        case 1: return rs_dispatch_nks<1, RS>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
        case 2: return rs_dispatch_nks<2, RS>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
        default: return rs_dispatch_nks<3, RS>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
    }
}

size_t boot_rs_coef_bytes(int N, int K, int R) {
    RsPlan b;
    if (!rs_plan(N, K, R, 1, b)) return 0;
    return (size_t)b.nper * b.nsub * b.stage_doubles * sizeof(double);
}

size_t boot_rs_workspace(int N, int64_t p, int K, int R) {
    RsPlan b;
    if (!rs_plan(N, K, R, p, b)) return 0;
    return b.nsplit > 1 ? (size_t)2 * b.nsplit * p * K * sizeof(double) : 16;
}

int boot_rs_pack(const double* E, int N, int K, const int32_t* idx, int R, double* coef, cudaStream_t st) {
    RsPlan b;
    if (!rs_plan(N, K, R, 1, b)) {
        set_err("boot_coef_pack_f64: unsupported shape N=%d K=%d R=%d (need K<=24, N<=1280)", N, K, R);
        return PLSB200_EUNSUPPORTED;
    }
    PLSB_CUDA(cudaMemsetAsync(coef, 0, (size_t)b.nper * b.nsub * b.stage_doubles * sizeof(double), st));
    size_t smem = (size_t)(4 * N + 1) * sizeof(int);
    boot_rs_pack_kernel<<<R, 256, smem, st>>>(E, N, K, idx, b.Kp, b.nacc, b.nb, b.rs, b.nks, b.ck, b.nsub, coef);
    PLSB_LAUNCH_CHECK("boot_rs_pack_kernel");
    return PLSB200_OK;
}

int boot_rs_moments(const double* X, int N, int64_t p, int64_t ldx, const double* coef, int K, int R, const double* pivot,
                    double* sum, double* sumsq, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    RsPlan b;
    if (!rs_plan(N, K, R, p, b)) {
        set_err("boot_moments_f64: unsupported shape N=%d K=%d R=%d (need K<=24, N<=1280)", N, K, R);
        return PLSB200_EUNSUPPORTED;
    }
    double *o1 = sum, *o2 = sumsq;
    if (b.nsplit > 1) {
        const size_t need = (size_t)2 * b.nsplit * p * K * sizeof(double);
        if (!workspace || workspace_bytes < need) {
            set_err("boot_moments_f64: workspace %zu < %zu bytes", workspace_bytes, need);
            return PLSB200_EWORKSPACE;
        }
        o1 = (double*)workspace; o2 = o1 + (size_t)b.nsplit * p * K;
    }
    const int rc = b.rs == 2 ? rs_dispatch_nacc<2>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st)
                             : rs_dispatch_nacc<4>(b, X, N, p, ldx, coef, K, R, pivot, o1, o2, st);
    if (rc != PLSB200_OK) return rc;
    if (b.nsplit > 1) {
        const long long n = (long long)p * K;
        rs_moments_reduce_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(o1, o2, b.nsplit, n, sum, sumsq);
        PLSB_LAUNCH_CHECK("moments_reduce_kernel");
    }
    return PLSB200_OK;
}

}  // namespace plsb
