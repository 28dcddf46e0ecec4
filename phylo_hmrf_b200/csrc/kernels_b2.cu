// Phase B, warp-specialised pipeline (the fast path for the reference's own model shape:
// Potts compatibility, <= 8 neighbour slots, K within one register-tile pass).
//
// One CTA per SM, 16 warps:
//   warps 0-3   CONSUMERS (one per SM sub-partition, 232 registers): hold the K x F
//               sufficient-statistic accumulators as TK x TF register tiles and do nothing
//               but  S[k][f] += e[n][k] * y[n][f]  from shared-memory rows (LDS.128 + DFMA).
//   warps 4-15  PRODUCERS (88 registers): the latency-bound per-node work -- neighbour
//               gather, duplicate-label fold, soft-max terms, cost scalars, feature row --
//               for tiles of 16 nodes, two lanes per node (each lane takes half the states).
// A producer owns one shared-memory slot (16 P rows + 16 Y rows); producer p feeds consumer
// p % 4 through a full/empty mbarrier pair, so the FP64 pipe of every sub-partition always
// has the consumer's independent DFMA stream to issue while producers wait on memory.
// Same arithmetic as kernels_b.cu (reference: phylo_hmrf.py:311-314, 334-468).
#include "estep_common.cuh"

namespace phmrf {

using namespace estep;

namespace {

constexpr int kConsumers = 4;
constexpr int kProducers = 12;
constexpr int kPipeThreads = 32 * (kConsumers + kProducers);
constexpr int kTileNodes = 16;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// producers are not latency critical: back off between polls so the spin does not eat the
// issue slots of the consumer on the same sub-partition
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) __nanosleep(200);
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

template <int D, int TF, int TFs, int C0, int... Cs>
__device__ __forceinline__ void write_y_chunks(double *Yrow, const double (&x)[D], const double (&xs)[D], double inv,
                                               std::integer_sequence<int, Cs...>) {
    ((*reinterpret_cast<double2 *>(Yrow + 2 * (C0 + Cs)) =
          make_double2(y_at<D, TF, TFs, 2 * (C0 + Cs)>(x, xs, inv), y_at<D, TF, TFs, 2 * (C0 + Cs) + 1>(x, xs, inv))),
     ...);
}

template <int D, int TK, int TF, int KTH>
__global__ void __launch_bounds__(kPipeThreads, 1) estep_pipe_kernel(EstepArgs a, int nkt_total, int rsp) {
    using C = Cfg<D, TK, TF>;
    constexpr int F = C::F, TKs = C::TKs, TFs = C::TFs, NFT = C::NFT, RSY = C::RSY;
    // KTH = k tiles per half-lane of a producer (template: sizes the register row)
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot_doubles = kTileNodes * (rsp + RSY);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)kProducers * slot_doubles);
    uint64_t *full = bars, *empty = bars + kProducers;
    if (threadIdx.x == 0) {
        for (int p = 0; p < kProducers; ++p) {
            mbar_init(full + p, 1);
            mbar_init(empty + p, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int K = a.K, W = a.W;
    const int64_t n = a.n, ld = a.ld;
    const int64_t n_tiles = (n + kTileNodes - 1) / kTileNodes;
    const int64_t tile_stride_g = (int64_t)gridDim.x * kProducers;
    const int KF = K * F;

    if (warp >= kConsumers) {
        // =============================== PRODUCER ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        const int p = warp - kConsumers;
        double *slot = smem + (size_t)p * slot_doubles;
        const int nd = lane & 15, hs = lane >> 4;
        double *Prow = slot + nd * rsp;
        double *Yrow = slot + kTileNodes * rsp + nd * RSY;
        const bool weighted = a.estimate_type == 3;
        const double beta = a.beta;
        double c_pair = 0.0, c_pwn = 0.0, c_un = 0.0;
        int bad_any = 0;
        int64_t j = 0;
        // lane-constant offsets of this half-lane's first state
        const int k_first = hs * KTH * TK;
        const int split = KTH * TK;  // states >= split belong to the upper half-lane
        for (int64_t T = (int64_t)blockIdx.x * kProducers + p; T < n_tiles; T += tile_stride_g, ++j) {
            const int64_t i_raw = T * kTileNodes + nd;
            const bool valid = i_raw < n;
            const int64_t i = valid ? i_raw : n - 1;
            {   // pull this producer's next tile towards L2 (one 128-byte line per row and tile)
                const int64_t T2 = T + tile_stride_g;
                if (T2 < n_tiles) {
                    const int64_t i2 = T2 * kTileNodes;
                    const double *row = a.logp + lane * ld + i2;
                    if (lane < K) prefetch_l2(row);
                    if (lane + 32 < K) prefetch_l2(row + 32 * ld);
                    if (lane < D) prefetch_l2(a.X_soa + lane * ld + i2);
                    if (lane < W) prefetch_l2(a.nbr_w + lane * ld + i2);
                    if (lane < W) prefetch_l2(a.nbr_id + lane * ld + i2);
                }
            }
            // ---- loads first: neighbour ids and own label, then what depends on them.
            int lab[kFastSlots];
            double sw[kFastSlots];
            int li;
            double lp_li;
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
#pragma unroll
                for (int s = 0; s < kFastSlots; ++s) {
                    sw[s] = (s < W && weighted) ? *pw : 1.0;
                    pw += ld;
                }
#pragma unroll
                for (int s = 0; s < kFastSlots; ++s) lab[s] = jid[s] >= 0 ? a.labels[jid[s]] : -1;
                lp_li = a.logp[li * ld + i];
            }
            // this lane's share of the log-likelihood row (registers; issued before the
            // neighbour arithmetic so that its latency overlaps)
            double e[KTH][TK];
            {
                const double *pk = a.logp + (int64_t)k_first * ld + i;
#pragma unroll
                for (int q = 0; q < KTH; ++q)
#pragma unroll
                    for (int ii = 0; ii < TK; ++ii) {
                        const int k = k_first + q * TK + ii;
                        e[q][ii] = k < K ? *pk : -1.0e6;
                        pk += ld;
                    }
            }
            int all_neg = -1;  // sign bit stays set while no slot holds a neighbour
            double pc = 0.0;   // sum over the incident edges of V[l_nbr, l_i] * w
#pragma unroll
            for (int s = 0; s < kFastSlots; ++s) {
                sw[s] = lab[s] >= 0 ? beta * sw[s] : 0.0;
                all_neg &= lab[s];
                pc += lab[s] != li ? sw[s] : 0.0;  // an empty slot has weight 0
            }
            if (all_neg < 0) {  // isolated node: pp = V[label] unweighted (phylo_hmrf.py:421-423)
                lab[0] = li;
                sw[0] = beta;
            }
            // g_s = exp(beta*w_s) per slot, in lock step.  The host only selects this kernel when
            // beta * W * max|w| < 100, so no range checks.  (A slot without neighbour gives 1.)
            exp_batch<kFastSlots, false>(sw);

            // ---- soft-max shift = max(logp_li, ~max_k logp_k - 600).  The maximum only has to be
            // right to about one unit, so it is taken on the order-preserving integer image of the
            // high words (3 integer instructions per value instead of an FP64 compare/select).
            int kmax = (int)0x80000000;
#pragma unroll
            for (int q = 0; q < KTH; ++q)
#pragma unroll
                for (int ii = 0; ii < TK; ++ii) {
                    const int h = __double2hiint(e[q][ii]);
                    kmax = max(kmax, h ^ ((h >> 31) & 0x7fffffff));
                }
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, 16));
            const double lpmax = __hiloint2double(kmax ^ ((kmax >> 31) & 0x7fffffff), 0);
            const double shift = fmax(lp_li, lpmax - 598.0);
#pragma unroll
            for (int q = 0; q < KTH; ++q) {
#pragma unroll
                for (int ii = 0; ii < TK; ++ii) e[q][ii] -= shift;
                exp_batch<TK, true>(e[q]);
            }

            // ---- slot: wait until the consumer released it.  The P row first holds
            // G_k = exp(S_k), S_k = sum of beta*w over the neighbours labelled k, built by
            // multiplying g_s into G[label_s] slot by slot (no duplicate-label bookkeeping:
            // exp(a)exp(b) = exp(a+b)); each half-lane owns the states of its own half.
            if (j > 0) mbar_wait_relaxed(empty + p, (uint32_t)((j - 1) & 1));
#pragma unroll
            for (int q = 0; q < KTH; ++q) {
                const int ktile = hs * KTH + q;
                if (ktile < nkt_total) {
#pragma unroll
                    for (int c = 0; c < even_up(TK); c += 2)
                        *reinterpret_cast<double2 *>(Prow + ktile * TKs + c) = make_double2(1.0, 1.0);
                }
            }
#pragma unroll
            for (int s = 0; s < kFastSlots; ++s) {
                if (lab[s] >= 0 && (lab[s] >= split) == (hs != 0)) {
                    const int kt_s = lab[s] / TK;
                    double *g = Prow + kt_s * (TKs - TK) + lab[s];
                    *g = *g * sw[s];
                }
            }
            // e_k = exp(logp_k - shift) * G_k, written over G; Q = sum_k G_k on the way
            double esum = 0.0, qsum = 0.0, g_li = 0.0;
#pragma unroll
            for (int q = 0; q < KTH; ++q) {
                const int ktile = hs * KTH + q;
                if (ktile < nkt_total) {
                    double g[even_up(TK)];
#pragma unroll
                    for (int c = 0; c < even_up(TK); c += 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(Prow + ktile * TKs + c);
                        g[c] = v.x;
                        g[c + 1] = v.y;
                    }
#pragma unroll
                    for (int ii = 0; ii < TK; ++ii) {
                        const int k = ktile * TK + ii;
                        if (k < K) qsum += g[ii];
                        g_li = k == li ? g[ii] : g_li;
                        g[ii] *= e[q][ii];
                        esum += g[ii];
                    }
                    if (TK < even_up(TK)) g[even_up(TK) - 1] = 0.0;
#pragma unroll
                    for (int c = 0; c < even_up(TK); c += 2)
                        *reinterpret_cast<double2 *>(Prow + ktile * TKs + c) = make_double2(g[c], g[c + 1]);
                }
            }
            esum += __shfl_xor_sync(0xffffffffu, esum, 16);
            qsum += __shfl_xor_sync(0xffffffffu, qsum, 16);
            g_li += __shfl_xor_sync(0xffffffffu, g_li, 16);
            // soft-max of -pp at the node's own label: exp(S_li) / sum_k exp(S_k)
            const double pwn_log = log(g_li / qsum + 1e-16);
            const bool bad = !(esum <= DBL_MAX) || !(qsum <= DBL_MAX) || !(esum > 0.0);
            bad_any |= bad ? 1 : 0;
            const double inv = valid ? 1.0 / esum : 0.0;
            {
                double x[D], xs[D];
                const double *px = a.X_soa + i;
#pragma unroll
                for (int jx = 0; jx < D; ++jx) {
                    x[jx] = *px;
                    px += ld;
                }
#pragma unroll
                for (int jx = 0; jx < D; ++jx) xs[jx] = x[jx] * inv;
                constexpr int NCH = RSY / 2, H0 = NCH / 2;
                if (hs == 0)
                    write_y_chunks<D, TF, TFs, 0>(Yrow, x, xs, inv, std::make_integer_sequence<int, H0>{});
                else
                    write_y_chunks<D, TF, TFs, H0>(Yrow, x, xs, inv, std::make_integer_sequence<int, NCH - H0>{});
            }
            if (valid && hs == 0) {
                c_pair += all_neg < 0 ? 0.0 : pc;
                c_un += lp_li;
                c_pwn += pwn_log;
            }
            if (a.post_soa != nullptr) {
                __syncwarp();
                if (valid) {
#pragma unroll
                    for (int q = 0; q < KTH; ++q)
#pragma unroll
                        for (int ii = 0; ii < TK; ++ii) {
                            const int ktile = hs * KTH + q;
                            const int k = ktile * TK + ii;
                            if (k < K) a.post_soa[k * ld + i] = Prow[ktile * TKs + ii] * inv;
                        }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(full + p);
        }
        if (bad_any) atomicOr(a.flags, 1);
        // cost sums of this producer -> shared scratch after the pipeline drained (below)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c_pair += __shfl_xor_sync(0xffffffffu, c_pair, o);
            c_pwn += __shfl_xor_sync(0xffffffffu, c_pwn, o);
            c_un += __shfl_xor_sync(0xffffffffu, c_un, o);
        }
        __syncthreads();  // (A) every slot consumed
        double *red = smem;
        for (int e0 = threadIdx.x; e0 < KF + 3; e0 += blockDim.x) red[e0] = 0.0;
        __syncthreads();  // (B)
        for (int w = 0; w < kConsumers; ++w) __syncthreads();  // consumers add their tiles in order
        for (int w = 0; w < kProducers; ++w) {
            if (p == w && lane == 0) {
                red[KF + 0] += c_pair;
                red[KF + 1] += c_pwn;
                red[KF + 2] += c_un;
            }
            __syncthreads();
        }
    } else {
        // =============================== CONSUMER ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int c = warp;
        const int tiles = nkt_total * NFT;
        const int NS = 32 / tiles;
        const int sub = lane / tiles;
        const int tl = lane - sub * tiles;
        const int kt = tl / NFT, ft = tl - kt * NFT;
        const bool lane_active = sub < NS;
        double acc[TK][TF];
#pragma unroll
        for (int i = 0; i < TK; ++i)
#pragma unroll
            for (int jj = 0; jj < TF; ++jj) acc[i][jj] = 0.0;
        constexpr int PPC = kProducers / kConsumers;
        int64_t cnt[PPC];
#pragma unroll
        for (int q = 0; q < PPC; ++q) {
            const int64_t g = (int64_t)blockIdx.x * kProducers + (c + q * kConsumers);
            cnt[q] = n_tiles > g ? (n_tiles - g - 1) / tile_stride_g + 1 : 0;
        }
        for (int64_t j = 0; j < cnt[0]; ++j) {  // cnt[0] >= cnt[q] for every q
#pragma unroll
            for (int q = 0; q < PPC; ++q) {
                if (j < cnt[q]) {
                    const int p = c + q * kConsumers;
                    const double *slot = smem + (size_t)p * slot_doubles;
                    mbar_wait(full + p, (uint32_t)(j & 1));
                    if (lane_active) {
                        const double *pb = slot + kt * TKs;
                        const double *yb = slot + kTileNodes * rsp + ft * TFs;
#pragma unroll 2
                        for (int nn = sub; nn < kTileNodes; nn += NS) {
                            double pv[even_up(TK)], yv[even_up(TF)];
#pragma unroll
                            for (int cc = 0; cc < even_up(TK); cc += 2) {
                                const double2 v = *reinterpret_cast<const double2 *>(pb + nn * rsp + cc);
                                pv[cc] = v.x;
                                pv[cc + 1] = v.y;
                            }
#pragma unroll
                            for (int cc = 0; cc < even_up(TF); cc += 2) {
                                const double2 v = *reinterpret_cast<const double2 *>(yb + nn * RSY + cc);
                                yv[cc] = v.x;
                                yv[cc + 1] = v.y;
                            }
#pragma unroll
                            for (int ii = 0; ii < TK; ++ii)
#pragma unroll
                                for (int jj = 0; jj < TF; ++jj) acc[ii][jj] = fma(pv[ii], yv[jj], acc[ii][jj]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty + p);
                }
            }
        }
        __syncthreads();  // (A)
        double *red = smem;
        for (int e0 = threadIdx.x; e0 < KF + 3; e0 += blockDim.x) red[e0] = 0.0;
        __syncthreads();  // (B)
        for (int w = 0; w < kConsumers; ++w) {
            for (int s = 0; s < NS; ++s) {
                if (c == w && sub == s && lane_active) {
#pragma unroll
                    for (int ii = 0; ii < TK; ++ii) {
                        const int k = kt * TK + ii;
#pragma unroll
                        for (int jj = 0; jj < TF; ++jj) {
                            const int f = ft * TF + jj;
                            if (k < K && f < F) red[k * F + f] += acc[ii][jj];
                        }
                    }
                }
                __syncwarp();
            }
            __syncthreads();
        }
        for (int w = 0; w < kProducers; ++w) __syncthreads();
    }
    double *out = a.partials + (size_t)blockIdx.x * (KF + 3);
    const double *red = smem;
    for (int e0 = threadIdx.x; e0 < KF + 3; e0 += blockDim.x) out[e0] = red[e0];
}

template <int D>
int launch_pipe_d(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    constexpr TileChoice tc = tile_for(D);
    using C = Cfg<D, tc.tk, tc.tf>;
    const int nkt_total = (a.K + tc.tk - 1) / tc.tk;
    if (nkt_total > C::NKT_MAX) return PHMRF_OK;  // needs the multi-pass general kernel
    const int rsp = pad_row(nkt_total * C::TKs);
    const size_t slot = (size_t)kTileNodes * (rsp + C::RSY) * sizeof(double);
    size_t smem = kProducers * slot + 2 * kProducers * sizeof(uint64_t);
    const size_t red_bytes = ((size_t)a.K * C::F + 3) * sizeof(double);
    if (smem < red_bytes) smem = red_bytes;
    if (smem > 227 * 1024) return PHMRF_OK;
    const int64_t n_tiles = (a.n + kTileNodes - 1) / kTileNodes;
    int64_t want = (n_tiles + kProducers - 1) / kProducers;
    const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    const int kth = (nkt_total + 1) / 2;
    void (*kern)(EstepArgs, int, int) = nullptr;
    switch (kth) {
        case 1: kern = estep_pipe_kernel<D, tc.tk, tc.tf, 1>; break;
        case 2: kern = estep_pipe_kernel<D, tc.tk, tc.tf, 2>; break;
        case 3: kern = estep_pipe_kernel<D, tc.tk, tc.tf, 3>; break;
        case 4: kern = estep_pipe_kernel<D, tc.tk, tc.tf, 4>; break;
        default: return PHMRF_OK;  // more than 8 k tiles: general kernel
    }
    PHMRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kPipeThreads, smem, s>>>(a, nkt_total, rsp);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    *handled = true;
    return launch_estep_finalize(a.partials, grid, a.K, D, a.stats_out, s);
}

}  // namespace

int launch_estep_pipe(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    *handled = false;
    if (!a.potts || a.W > kFastSlots || a.pp_soa != nullptr || a.n == 0) return PHMRF_OK;
    if (!(a.s_bound < 100.0)) return PHMRF_OK;  // exp(S) * exp(600) must stay finite without range checks
    switch (a.D) {
#define PHMRF_CASE(DD) \
    case DD:           \
        return launch_pipe_d<DD>(a, sm_count, s, handled);
        PHMRF_CASE(1) PHMRF_CASE(2) PHMRF_CASE(3) PHMRF_CASE(4) PHMRF_CASE(5) PHMRF_CASE(6)
        PHMRF_CASE(7) PHMRF_CASE(8) PHMRF_CASE(9) PHMRF_CASE(10) PHMRF_CASE(11) PHMRF_CASE(12)
#undef PHMRF_CASE
    }
    return PHMRF_OK;
}

}  // namespace phmrf
