// Phase B, bulk-copy pipeline (the fast path for the reference's own model shape: Potts
// compatibility, <= 8 neighbour slots, K <= 40 states).  Same arithmetic as kernels_b.cu
// (reference: phylo_hmrf.py:311-314, 334-468).
//
// One CTA per SM, 4 + P warps (P = 8; 12 for small state x feature tiles):
//   warps 0-3   CONSUMERS (one per SM sub-partition): hold the K x F sufficient-statistic
//               accumulators and issue nothing but  S[k][f] += e[n][k] * y[n][f]  as DMMA.8x8x4
//               (mma.sync m8n8k4 f64; on B200 it shares the DFMA datapath, but one instruction
//               carries 256 FMAs and takes one operand per lane).
//   warps 4..   PRODUCERS: the per-node work for tiles of 32 nodes, one lane per node.
// A producer owns one shared-memory slot of two regions, both [row][32 nodes]:
//   region 1 [KP]       the log-likelihood tile, fetched by ONE cp.async.bulk (the HBM layout
//                       is tile-major and pre-swizzled, common.cuh lp_index) onto an mbarrier;
//                       the soft-max terms e_k = exp(logp_k - shift) * G_k are written in
//                       place, 8 states at a time, so no per-state register row exists;
//   region 2            first the neighbour products G_k = exp(sum of beta*w over the
//                       neighbours labelled k) (built multiplicatively, slot by slot), then
//                       -- G is dead once e is written -- the feature rows y_f / sum(e),
//                       y = (1, x, x (x) x packed).
// Every row is XOR-swizzled by (row % 8) * 4 columns: lane-per-node accesses along a row and
// the mma operand fragments (8 rows x 4 nodes) are both bank-conflict free
// (tools/swizzle_check.py), and a consumer addresses every operand of a step from two
// registers: base ^ (step << 5) plus an immediate per 8-row tile.
// exp() is a 256-entry table (2^(j/256), shared memory) times a cubic: 8 FP64 instructions.
#include "estep_common.cuh"

namespace phmrf {

using namespace estep;

namespace {

constexpr int kCons = 4;
constexpr int kTile = 32;
constexpr int kExpTab = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// one contiguous run global -> shared through the bulk-copy engine, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D(8x8) += A(8x4) * B(4x8), FP64.  Lane (g = lane/4, t = lane%4) supplies A[g][t], B[t][g]
// and holds D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// N independent exponentials in lock step: t = (n + j/256) ln2 + r, |r| <= ln2/512,
// exp(t) = 2^n * tab[j] * (1 + r + r^2/2 + r^3/6)   (truncation r^4/24 < 1.5e-13 relative).
// The argument is clamped from below at about -704 with one integer minimum on the high word
// (results that would be below exp(-705) come out in [exp(-705), exp(-704)] instead of 0: for
// soft-max terms whose sum is at least 1 an absolute error below 1e-306); it must not exceed
// +700.  8 FP64-pipe instructions per value (the Taylor form it replaces took 13).
template <int N>
__device__ __forceinline__ void exp_tab(double (&t)[N], const double *tab) {
    const double kMagic = 6755399441055744.0;
    double sft[N], r[N], p[N];
#pragma unroll
    for (int u = 0; u < N; ++u)
        t[u] = __hiloint2double((int)min((unsigned)__double2hiint(t[u]), 0xC0860000u), __double2loint(t[u]));
#pragma unroll
    for (int u = 0; u < N; ++u) sft[u] = fma(t[u], 369.3299304675746, kMagic);
#pragma unroll
    for (int u = 0; u < N; ++u) {
        const double fn = sft[u] - kMagic;
        r[u] = fma(fn, -0x1.62e42fe000000p-9, t[u]);   // ln2/256, leading 29 bits (fn * this is exact)
        r[u] = fma(fn, -0x1.f473de6af278fp-38, r[u]);  // ln2/256, remainder
    }
#pragma unroll
    for (int u = 0; u < N; ++u) p[u] = fma(r[u], 1.66666666666666657e-01, 0.5);
#pragma unroll
    for (int u = 0; u < N; ++u) p[u] = fma(p[u], r[u], 1.0);
#pragma unroll
    for (int u = 0; u < N; ++u) p[u] = fma(p[u], r[u], 1.0);
#pragma unroll
    for (int u = 0; u < N; ++u) {
        const int m = __double2loint(sft[u]);
        const double v = tab[m & (kExpTab - 1)] * p[u];
        t[u] = __hiloint2double(__double2hiint(v) + ((m >> 8) << 20), __double2loint(v));
    }
}

// feature f of the node's row: 1/sum, x_j/sum, x_a*x_b/sum (packed upper triangle), 0 on padding
template <int D, int POS>
__device__ __forceinline__ double y_feature(const double (&x)[D], const double (&xs)[D], double inv) {
    constexpr int F = n_stat_features(D);
    if constexpr (POS >= F) {
        return 0.0;
    } else if constexpr (POS == 0) {
        return inv;
    } else if constexpr (POS <= D) {
        return xs[POS - 1];
    } else {
        constexpr int r = POS - 1 - D;
        return xs[tri_row_of(r, D)] * x[tri_col_of(r, D)];
    }
}
template <int D, int RY, int... Fs>
__device__ __forceinline__ void write_y_rows(double *const (&col)[8], const double (&x)[D], const double (&xs)[D],
                                             double inv, std::integer_sequence<int, Fs...>) {
    ((col[(RY + Fs) & 7][(RY + Fs) * 32] = y_feature<D, Fs>(x, xs, inv)), ...);
}

template <int D, int NK8>
struct BulkCfg {
    static constexpr int F = n_stat_features(D);
    static constexpr int NT = (F + 7) / 8;
    static constexpr int KP = 8 * NK8, FP = 8 * NT;
    static constexpr int R2 = KP > FP ? KP : FP;     // rows of region 2: G rows, then the feature rows
    static constexpr int ROWS = KP + R2;
    static constexpr int SLOT_BYTES = ROWS * 256;
    static constexpr size_t smem_bytes(int P) {
        return (size_t)P * SLOT_BYTES + kExpTab * 8 + 3 * (size_t)P * 8 + 256;
    }
    // 12 producer warps where the accumulator tile is small (the consumers hardly load the FP64
    // pipe and every warp fits 128 registers), else 8 at 168 registers
    static constexpr int P = (NK8 * NT <= 12 && NK8 <= 3 && smem_bytes(12) <= 227 * 1024) ? 12 : 8;
};

template <int D, int NK8, int P>
__global__ void __launch_bounds__(32 * (kCons + P), 1) estep_bulk_kernel(EstepArgs a) {
    using C = BulkCfg<D, NK8>;
    constexpr int F = C::F, NT = C::NT, KP = C::KP, FP = C::FP, SLOT_BYTES = C::SLOT_BYTES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // rows must start on 256-byte boundaries (the swizzle is an XOR on address bits 5-7)
    unsigned char *sbase = smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u);
    double *tab = reinterpret_cast<double *>(sbase + (size_t)P * SLOT_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(tab + kExpTab);
    uint64_t *full = bars, *empty = bars + P, *landed = bars + 2 * P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int p = 0; p < P; ++p) {
            mbar_init(full + p, 1);
            mbar_init(empty + p, 1);
            mbar_init(landed + p, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int j = threadIdx.x; j < kExpTab; j += blockDim.x) tab[j] = exp((double)j * (0.693147180559945309417 / kExpTab));
    __syncthreads();

    const int K = a.K, W = a.W;
    const int64_t n = a.n, ld = a.ld;
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int64_t tile_stride_g = (int64_t)gridDim.x * P;
    const int KF = K * F;

    if (warp >= kCons) {
        // =============================== PRODUCER ===============================
        const int p = warp - kCons;
        unsigned char *sb = sbase + (size_t)p * SLOT_BYTES;
        // column of this lane in a row r: lane ^ ((r % 8) * 4); one base pointer per r % 8
        double *col[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) col[q] = reinterpret_cast<double *>(sb + ((lane ^ (q << 2)) << 3));
        auto dyn_row = [&](int r) -> double * {
            return reinterpret_cast<double *>(sb + (r << 8) + ((lane ^ ((r & 7) << 2)) << 3));
        };
        const bool weighted = a.estimate_type == 3;
        const double beta = a.beta;
        double c_pair = 0.0, c_pwn = 0.0, c_un = 0.0;
        int bad_any = 0;
        int64_t j = 0;
        for (int64_t T = (int64_t)blockIdx.x * P + p; T < n_tiles; T += tile_stride_g, ++j) {
            const int64_t i_raw = T * kTile + lane;
            const bool valid = i_raw < n;
            const int64_t i = valid ? i_raw : n - 1;
            if (lane == 0) {
                const int64_t T2 = T + tile_stride_g;
                if (T2 < n_tiles) bulk_prefetch_l2(a.logp + T2 * (KP * 32), KP * 256);
            }
            {   // pull this producer's next tile of the per-node arrays towards L2
                const int64_t T2 = T + tile_stride_g;
                if (T2 < n_tiles) {
                    const int64_t i2 = T2 * kTile;
                    if (lane < 2 * D) prefetch_l2(a.X_soa + (lane >> 1) * ld + i2 + (lane & 1) * 16);
                    if (lane < 2 * W) prefetch_l2(a.nbr_w + (lane >> 1) * ld + i2 + (lane & 1) * 16);
                    if (lane >= 16 && lane - 16 < 2 * W)
                        prefetch_l2(a.nbr_g + ((lane - 16) >> 1) * ld + i2 + (lane & 1) * 16);
                    if (lane < W) prefetch_l2(a.nbr_id + lane * ld + i2);
                    if (lane == 31) prefetch_l2(a.rowmax + i2);
                }
            }
            // ---- neighbour phase: global loads and what depends only on them, while the
            // consumer still reads this slot's previous tile
            int lab[kFastSlots];
            double sw[kFastSlots], gw[kFastSlots];  // w_s and g_s = exp(beta*w_s) (precomputed, 1 if empty)
            int li;
            {
                int jid[kFastSlots];
                const int32_t *pid = a.nbr_id + i;
#pragma unroll
                for (int s = 0; s < kFastSlots; ++s) {
                    jid[s] = s < W ? *pid : -1;
                    pid += ld;
                }
                li = a.labels[a.own_offset + i];
                const double *pw = a.nbr_w + i;
                const double *pg = a.nbr_g + i;
#pragma unroll
                for (int s = 0; s < kFastSlots; ++s) {
                    sw[s] = (s < W && weighted) ? *pw : 0.0;
                    gw[s] = s < W ? *pg : 1.0;
                    pw += ld;
                    pg += ld;
                }
#pragma unroll
                for (int s = 0; s < kFastSlots; ++s) lab[s] = jid[s] >= 0 ? a.labels[jid[s]] : -1;
            }
            const double rmax = a.rowmax[i];
            double x[D];
            {
                const double *px = a.X_soa + i;
#pragma unroll
                for (int jx = 0; jx < D; ++jx) {
                    x[jx] = *px;
                    px += ld;
                }
            }
            int all_neg = -1;  // sign bit stays set while no slot holds a neighbour
            double pc = 0.0;   // sum over the incident edges of V[l_nbr, l_i] * w
#pragma unroll
            for (int s = 0; s < kFastSlots; ++s) {
                // an empty slot has label -1 and stored weight 0; unweighted estimates count 1 per edge
                const double ws = weighted ? sw[s] : (lab[s] >= 0 ? 1.0 : 0.0);
                all_neg &= lab[s];
                pc += lab[s] != li ? ws : 0.0;
            }
            pc *= beta;
            if (all_neg < 0) {  // isolated node: pp = V[label] unweighted (phylo_hmrf.py:421-423)
                lab[0] = li;
                gw[0] = a.exp_beta;
            }
            // ---- the slot is free once the consumer has released the previous tile: fetch the
            // log-likelihood tile into region 1
            if (j > 0) mbar_wait(empty + p, (uint32_t)((j - 1) & 1));
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(landed + p, KP * 256);
                bulk_g2s(sb, a.logp + T * (KP * 32), KP * 256, landed + p);
            }
            // ---- G_k in region 2: ones, then g_s multiplied into G[label_s] slot by slot
            // (exp(a)exp(b) = exp(a+b): no duplicate-label bookkeeping).  The host only selects
            // this kernel when |beta| * W * max|w| < 100, so the products stay finite.
#pragma unroll
            for (int q = 0; q < KP; ++q) col[q & 7][(KP + q) * 32] = 1.0;
            double qs = 0.0;  // sum_k G_k - K
#pragma unroll
            for (int s = 0; s < kFastSlots; ++s) {
                if (lab[s] >= 0) {
                    double *gp = dyn_row(KP + lab[s]);
                    const double g_old = *gp;
                    const double g_new = g_old * gw[s];
                    *gp = g_new;
                    qs += g_new - g_old;
                }
            }
            const double g_li = *dyn_row(KP + li);
            const double qsum = (double)K + qs;
            // soft-max of -pp at the node's own label: exp(S_li) / sum_k exp(S_k)
            const double pwn_log = log(fma(g_li, fast_rcp(qsum), 1e-16));
            // ---- the log-likelihood tile has landed
            mbar_wait(landed + p, (uint32_t)(j & 1));
            const double lp_li = *dyn_row(li);
            // soft-max shift = max(logp_li, max_k logp_k - 598): overflow-free for any labels
            const double shift = valid ? fmax(lp_li, rmax - 598.0) : 1.0e300;
            double esum = 0.0;
#pragma unroll
            for (int c = 0; c < NK8; ++c) {
                double tb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) tb[u] = col[u][(8 * c + u) * 32] - shift;
                exp_tab<8>(tb, tab);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const double v = tb[u] * col[u][(KP + 8 * c + u) * 32];
                    esum += v;
                    col[u][(8 * c + u) * 32] = v;
                }
            }
            const bool bad = !(esum <= DBL_MAX) || !(qsum <= DBL_MAX) || !(esum > 0.0);
            bad_any |= (bad && valid) ? 1 : 0;
            const double inv = valid ? fast_rcp(esum) : 0.0;
            // ---- feature rows over the dead G rows
            {
                double xs[D];
#pragma unroll
                for (int jx = 0; jx < D; ++jx) xs[jx] = x[jx] * inv;
                write_y_rows<D, KP>(col, x, xs, inv, std::make_integer_sequence<int, FP>{});
            }
            if (valid) {
                c_pair += all_neg < 0 ? 0.0 : pc;
                c_un += lp_li;
                c_pwn += pwn_log;
            }
            if (a.post_soa != nullptr) {
                if (valid) {
#pragma unroll
                    for (int q = 0; q < KP; ++q)
                        if (q < K) a.post_soa[q * ld + i] = col[q & 7][q * 32] * inv;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(full + p);
        }
        if (bad_any) atomicOr(a.flags, 1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_pair += __shfl_xor_sync(0xffffffffu, c_pair, o);
            c_pwn += __shfl_xor_sync(0xffffffffu, c_pwn, o);
            c_un += __shfl_xor_sync(0xffffffffu, c_un, o);
        }
        __syncthreads();  // (A) every slot consumed
        double *red = reinterpret_cast<double *>(sbase);
        for (int e0 = threadIdx.x; e0 < KF + 3; e0 += blockDim.x) red[e0] = 0.0;
        __syncthreads();  // (B)
        for (int w = 0; w < kCons; ++w) __syncthreads();  // consumers add their tiles in order
        for (int w = 0; w < P; ++w) {
            if (p == w && lane == 0) {
                red[KF + 0] += c_pair;
                red[KF + 1] += c_pwn;
                red[KF + 2] += c_un;
            }
            __syncthreads();
        }
    } else {
        // =============================== CONSUMER ===============================
        const int c = warp;
        const int g = lane >> 2, t = lane & 3;
        double acc[NK8][NT][2];
#pragma unroll
        for (int kt = 0; kt < NK8; ++kt)
#pragma unroll
            for (int ft = 0; ft < NT; ++ft) acc[kt][ft][0] = acc[kt][ft][1] = 0.0;
        // byte offset of (row 8*tile + g, node 4*ns + t) in a slot:
        //   ((g*256 + t*8) ^ (g << 5)) ^ (ns << 5)  +  tile * 2048      (KP is a multiple of 8)
        const uint32_t pk = (uint32_t)((g << 8) + (t << 3)) ^ (uint32_t)(g << 5);
        constexpr int PPC = P / kCons;
        int64_t cnt[PPC];
#pragma unroll
        for (int q = 0; q < PPC; ++q) {
            const int64_t gidx = (int64_t)blockIdx.x * P + (c + q * kCons);
            cnt[q] = n_tiles > gidx ? (n_tiles - gidx - 1) / tile_stride_g + 1 : 0;
        }
        for (int64_t j = 0; j < cnt[0]; ++j) {  // cnt[0] >= cnt[q] for every q
#pragma unroll
            for (int q = 0; q < PPC; ++q) {
                if (j < cnt[q]) {
                    const int p = c + q * kCons;
                    const unsigned char *sb = sbase + (size_t)p * SLOT_BYTES;
                    mbar_wait(full + p, (uint32_t)(j & 1));
#pragma unroll
                    for (int ns = 0; ns < kTile / 4; ++ns) {
                        double av[NK8], bv[NT];
                        const unsigned char *pa = sb + (pk ^ (uint32_t)(ns << 5));
#pragma unroll
                        for (int kt = 0; kt < NK8; ++kt) av[kt] = *reinterpret_cast<const double *>(pa + kt * 2048);
#pragma unroll
                        for (int ft = 0; ft < NT; ++ft)
                            bv[ft] = *reinterpret_cast<const double *>(pa + (KP / 8 + ft) * 2048);
#pragma unroll
                        for (int kt = 0; kt < NK8; ++kt)
#pragma unroll
                            for (int ft = 0; ft < NT; ++ft) dmma_8x8x4(acc[kt][ft][0], acc[kt][ft][1], av[kt], bv[ft]);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + p);
                }
            }
        }
        __syncthreads();  // (A)
        double *red = reinterpret_cast<double *>(sbase);
        for (int e0 = threadIdx.x; e0 < KF + 3; e0 += blockDim.x) red[e0] = 0.0;
        __syncthreads();  // (B)
        for (int w = 0; w < kCons; ++w) {
            if (c == w) {
#pragma unroll
                for (int kt = 0; kt < NK8; ++kt) {
                    const int k = 8 * kt + g;
#pragma unroll
                    for (int ft = 0; ft < NT; ++ft)
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            const int f = 8 * ft + 2 * t + jj;
                            if (k < K && f < F) red[k * F + f] += acc[kt][ft][jj];
                        }
                }
            }
            __syncthreads();
        }
        for (int w = 0; w < P; ++w) __syncthreads();
    }
    double *out = a.partials + (size_t)blockIdx.x * (KF + 3);
    const double *red = reinterpret_cast<const double *>(sbase);
    for (int e0 = threadIdx.x; e0 < KF + 3; e0 += blockDim.x) out[e0] = red[e0];
}

template <int D, int NK8, int P>
int launch_bulk_p(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    using C = BulkCfg<D, NK8>;
    size_t smem = C::smem_bytes(P);
    const size_t red_bytes = ((size_t)a.K * C::F + 3) * sizeof(double) + 256;
    if (smem < red_bytes) smem = red_bytes;
    if (smem > 227 * 1024) return PHMRF_OK;
    const int64_t n_tiles = (a.n + kTile - 1) / kTile;
    int64_t want = (n_tiles + P - 1) / P;
    const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    auto kern = estep_bulk_kernel<D, NK8, P>;
    PHMRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 32 * (kCons + P), smem, s>>>(a);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    *handled = true;
    return launch_estep_finalize(a.partials, grid, a.K, D, a.stats_out, s);
}

template <int D, int NK8>
int launch_bulk(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    using C = BulkCfg<D, NK8>;
    if constexpr (NK8 * C::NT * 2 > 64) {
        return PHMRF_OK;  // accumulator tiles would not fit the consumer's registers
    } else {
        return launch_bulk_p<D, NK8, C::P>(a, sm_count, s, handled);
    }
}

template <int D>
int launch_bulk_d(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    switch ((a.K + 7) / 8) {
        case 1: return launch_bulk<D, 1>(a, sm_count, s, handled);
        case 2: return launch_bulk<D, 2>(a, sm_count, s, handled);
        case 3: return launch_bulk<D, 3>(a, sm_count, s, handled);
        case 4: return launch_bulk<D, 4>(a, sm_count, s, handled);
        case 5: return launch_bulk<D, 5>(a, sm_count, s, handled);
        default: return PHMRF_OK;  // K > 40: general kernel
    }
}

}  // namespace

int launch_estep_bulk(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    *handled = false;
    if (!a.potts || a.W > kFastSlots || a.pp_soa != nullptr || a.n == 0 || a.nbr_g == nullptr || a.rowmax == nullptr)
        return PHMRF_OK;
    if (!(fabs(a.s_bound) < 100.0)) return PHMRF_OK;  // exp(S) * exp(600) must stay finite without range checks
    switch (a.D) {
#define PHMRF_CASE(DD) \
    case DD:           \
        return launch_bulk_d<DD>(a, sm_count, s, handled);
        PHMRF_CASE(1) PHMRF_CASE(2) PHMRF_CASE(3) PHMRF_CASE(4) PHMRF_CASE(5) PHMRF_CASE(6)
        PHMRF_CASE(7) PHMRF_CASE(8) PHMRF_CASE(9) PHMRF_CASE(10) PHMRF_CASE(11) PHMRF_CASE(12)
#undef PHMRF_CASE
    }
    return PHMRF_OK;
}

}  // namespace phmrf
