// Phase B, warp-specialised pipeline (the fast path for the reference's own model shape:
// Potts compatibility, <= 8 neighbour slots, K <= 40 states).
//
// One CTA per SM, 4 + P warps (P = 8, or 12 for small accumulator tiles):
//   warps 0-3   CONSUMERS (one per SM sub-partition): hold the K x F sufficient-statistic
//               accumulators and do nothing but  S[k][f] += e[n][k] * y[n][f].  This is a
//               dense (K x n)(n x F) FP64 product, issued as DMMA.8x8x4 (mma.sync m8n8k4
//               f64): on B200 the FP64 mma shares the DFMA datapath (same measured peak,
//               36.6-37.1 TFLOP/s), but one instruction carries 256 FMAs and takes ONE
//               operand per lane, so the stat phase needs 8x fewer issue slots and 4.5x fewer
//               shared-memory wavefronts than the DFMA formulation it replaced (ncu: LSU
//               wavefronts were at 72 % of peak and the sub-partition issue port was the
//               limiter, see profiles/r1_history.md).
//   warps 4..   PRODUCERS: the per-node work -- neighbour gather (per-slot factors exp(beta*w)
//               are precomputed per region), soft-max terms, cost scalars, feature row -- for
//               tiles of 32 nodes, one lane per node.
// A producer owns one shared-memory slot (32 P rows + 32 Y rows); producer p feeds consumer
// p % 4 through a full/empty mbarrier pair.
// Same arithmetic as kernels_b.cu (reference: phylo_hmrf.py:311-314, 334-468).
#include "estep_common.cuh"

namespace phmrf {

using namespace estep;

namespace {

constexpr int kConsumers = 4;
// producer warps per CTA: 8 (two per consumer) for the big accumulator tiles, 12 when the state x
// feature tile is small enough that the consumers hardly load the FP64 pipe and a producer fits
// 128 registers
constexpr int pipe_threads(int P) { return 32 * (kConsumers + P); }
constexpr int kTileNodes = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// producers are not latency critical: let the hardware suspend the warp on the barrier
// instead of spinning next to the consumer that shares its sub-partition
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

// D(8x8) += A(8x4) * B(4x8), FP64.  Lane (g = lane/4, t = lane%4) supplies A[g][t], B[t][g]
// and holds D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// dense feature row: position f holds feature f (1, x, x (x) x packed), zero beyond F
template <int D, int POS>
__device__ __forceinline__ double yd_at(const double (&x)[D], const double (&xs)[D], double inv) {
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
template <int D, int C0, int... Cs>
__device__ __forceinline__ void write_yd_chunks(double *Yrow, const double (&x)[D], const double (&xs)[D], double inv,
                                                std::integer_sequence<int, Cs...>) {
    ((*reinterpret_cast<double2 *>(Yrow + 2 * (C0 + Cs)) =
          make_double2(yd_at<D, 2 * (C0 + Cs)>(x, xs, inv), yd_at<D, 2 * (C0 + Cs) + 1>(x, xs, inv))),
     ...);
}

// NK8 = number of 8-state tiles (K <= 8*NK8).
template <int D, int NK8, int P>
__global__ void __launch_bounds__(pipe_threads(P), 1) estep_mma_kernel(EstepArgs a) {
    constexpr int kProducers = P;
    constexpr int F = n_stat_features(D);
    constexpr int NT = (F + 7) / 8;        // 8-feature tiles
    constexpr int KP = 8 * NK8, FP = 8 * NT;
    constexpr int RSP = KP + 4, RSY = FP + 4;  // == 4 (mod 8): the DMMA operand loads are conflict free
    constexpr int EB = 8;                   // exponentials evaluated in lock step
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int slot_doubles = kTileNodes * (RSP + RSY);
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
        if constexpr (P == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 168;");
        const int p = warp - kConsumers;
        double *slot = smem + (size_t)p * slot_doubles;
        double *Prow = slot + lane * RSP;
        double *Yrow = slot + kTileNodes * RSP + lane * RSY;
        const bool weighted = a.estimate_type == 3;
        const double beta = a.beta;
        double c_pair = 0.0, c_pwn = 0.0, c_un = 0.0;
        int bad_any = 0;
        int64_t j = 0;
        for (int64_t T = (int64_t)blockIdx.x * kProducers + p; T < n_tiles; T += tile_stride_g, ++j) {
            const int64_t i_raw = T * kTileNodes + lane;
            const bool valid = i_raw < n;
            const int64_t i = valid ? i_raw : n - 1;
            {   // pull this producer's next tile towards L2 (two 128-byte lines per row and tile)
                const int64_t T2 = T + tile_stride_g;
                if (T2 < n_tiles) {
                    const int64_t i2 = T2 * kTileNodes;
                    for (int q = lane; q < 2 * K; q += 32) prefetch_l2(a.logp + (T2 * KP + (q >> 1)) * 32 + (q & 1) * 16);
                    if (lane < 2 * D) prefetch_l2(a.X_soa + (lane >> 1) * ld + i2 + (lane & 1) * 16);
                    if (lane < 2 * W) prefetch_l2(a.nbr_w + (lane >> 1) * ld + i2 + (lane & 1) * 16);
                    if (lane >= 16 && lane - 16 < 2 * W)
                        prefetch_l2(a.nbr_g + ((lane - 16) >> 1) * ld + i2 + (lane & 1) * 16);
                    if (lane < W) prefetch_l2(a.nbr_id + lane * ld + i2);
                }
            }
            // ---- loads first: neighbour ids and own label, then what depends on them.
            int lab[kFastSlots];
            double sw[kFastSlots], gw[kFastSlots];  // w_s and g_s = exp(beta*w_s) (precomputed, 1 if empty)
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
                lp_li = a.logp[lp_index(li, i, KP)];
            }
            // the node's log-likelihood row (registers; issued before the neighbour arithmetic so
            // that its latency overlaps)
            double e[KP];
            {
                const double *pk = a.logp + (i >> 5) * (KP * 32);
                const int col = (int)(i & 31);
#pragma unroll
                for (int q = 0; q < KP; ++q)  // rows K..KP-1 hold kLogpPad (phmrf_region_create)
                    e[q] = pk[q * 32 + (col ^ ((q & 7) << 2))];
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
            // (the host only selects this kernel when |beta| * W * max|w| < 100: the products of
            // the g_s below stay finite)

            // ---- soft-max shift = max(logp_li, ~max_k logp_k - 598).  The maximum only has to be
            // right to about one unit, so it is taken on the order-preserving integer image of the
            // high words (3 integer instructions per value instead of an FP64 compare/select).
            int kmax = (int)0x80000000;
#pragma unroll
            for (int q = 0; q < KP; ++q) {
                const int h = __double2hiint(e[q]);
                kmax = max(kmax, h ^ ((h >> 31) & 0x7fffffff));
            }
            const double lpmax = __hiloint2double(kmax ^ ((kmax >> 31) & 0x7fffffff), 0);
            const double shift = fmax(lp_li, lpmax - 598.0);
#pragma unroll
            for (int q0 = 0; q0 < KP; q0 += EB) {
                double tb[EB];
#pragma unroll
                for (int u = 0; u < EB; ++u) tb[u] = e[q0 + u] - shift;
                exp_batch<EB, 2>(tb);
#pragma unroll
                for (int u = 0; u < EB; ++u) e[q0 + u] = tb[u];
            }

            // ---- slot: wait until the consumer released it.  The P row first holds
            // G_k = exp(S_k), S_k = sum of beta*w over the neighbours labelled k, built by
            // multiplying g_s into G[label_s] slot by slot (no duplicate-label bookkeeping:
            // exp(a)exp(b) = exp(a+b)).
            if (j > 0) mbar_wait_relaxed(empty + p, (uint32_t)((j - 1) & 1));
#pragma unroll
            for (int c = 0; c < KP; c += 2) *reinterpret_cast<double2 *>(Prow + c) = make_double2(1.0, 1.0);
#pragma unroll
            for (int s = 0; s < kFastSlots; ++s) {
                if (lab[s] >= 0) Prow[lab[s]] *= gw[s];
            }
            const double g_li = Prow[li];
            // e_k = exp(logp_k - shift) * G_k, written over G; Q = sum_k G_k on the way
            double esum = 0.0, qsum = 0.0;
#pragma unroll
            for (int c = 0; c < KP; c += 2) {
                double2 v = *reinterpret_cast<const double2 *>(Prow + c);
                qsum += v.x + v.y;
                v.x *= e[c];
                v.y *= e[c + 1];
                esum += v.x + v.y;
                *reinterpret_cast<double2 *>(Prow + c) = v;
            }
            // soft-max of -pp at the node's own label: exp(S_li) / sum_k exp(S_k); the padding
            // states have no neighbour, G = 1 exactly
            qsum -= (double)(KP - K);
            const double pwn_log = log(fma(g_li, fast_rcp(qsum), 1e-16));
            const bool bad = !(esum <= DBL_MAX) || !(qsum <= DBL_MAX) || !(esum > 0.0);
            bad_any |= bad ? 1 : 0;
            const double inv = valid ? fast_rcp(esum) : 0.0;
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
                write_yd_chunks<D, 0>(Yrow, x, xs, inv, std::make_integer_sequence<int, FP / 2>{});
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
                        if (q < K) a.post_soa[q * ld + i] = Prow[q] * inv;
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
        if constexpr (P == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        const int c = warp;
        const int g = lane >> 2, t = lane & 3;
        double acc[NK8][NT][2];
#pragma unroll
        for (int kt = 0; kt < NK8; ++kt)
#pragma unroll
            for (int ft = 0; ft < NT; ++ft) acc[kt][ft][0] = acc[kt][ft][1] = 0.0;
        constexpr int PPC = kProducers / kConsumers;
        int64_t cnt[PPC];
#pragma unroll
        for (int q = 0; q < PPC; ++q) {
            const int64_t gidx = (int64_t)blockIdx.x * kProducers + (c + q * kConsumers);
            cnt[q] = n_tiles > gidx ? (n_tiles - gidx - 1) / tile_stride_g + 1 : 0;
        }
        for (int64_t j = 0; j < cnt[0]; ++j) {  // cnt[0] >= cnt[q] for every q
#pragma unroll
            for (int q = 0; q < PPC; ++q) {
                if (j < cnt[q]) {
                    const int p = c + q * kConsumers;
                    const double *slot = smem + (size_t)p * slot_doubles;
                    const double *pa = slot + t * RSP + g;
                    const double *pb = slot + kTileNodes * RSP + t * RSY + g;
                    mbar_wait(full + p, (uint32_t)(j & 1));
#pragma unroll
                    for (int ns = 0; ns < kTileNodes / 4; ++ns) {
                        double av[NK8], bv[NT];
#pragma unroll
                        for (int kt = 0; kt < NK8; ++kt) av[kt] = pa[ns * 4 * RSP + 8 * kt];
#pragma unroll
                        for (int ft = 0; ft < NT; ++ft) bv[ft] = pb[ns * 4 * RSY + 8 * ft];
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
        double *red = smem;
        for (int e0 = threadIdx.x; e0 < KF + 3; e0 += blockDim.x) red[e0] = 0.0;
        __syncthreads();  // (B)
        for (int w = 0; w < kConsumers; ++w) {
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
        for (int w = 0; w < kProducers; ++w) __syncthreads();
    }
    double *out = a.partials + (size_t)blockIdx.x * (KF + 3);
    const double *red = smem;
    for (int e0 = threadIdx.x; e0 < KF + 3; e0 += blockDim.x) out[e0] = red[e0];
}

template <int D, int NK8, int P>
int launch_mma_p(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    constexpr int F = n_stat_features(D);
    constexpr int NT = (F + 7) / 8;
    constexpr int RSP = 8 * NK8 + 4, RSY = 8 * NT + 4;
    const size_t slot = (size_t)kTileNodes * (RSP + RSY) * sizeof(double);
    size_t smem = P * slot + 2 * P * sizeof(uint64_t);
    const size_t red_bytes = ((size_t)a.K * F + 3) * sizeof(double);
    if (smem < red_bytes) smem = red_bytes;
    if (smem > 227 * 1024) return PHMRF_OK;
    const int64_t n_tiles = (a.n + kTileNodes - 1) / kTileNodes;
    int64_t want = (n_tiles + P - 1) / P;
    const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    auto kern = estep_mma_kernel<D, NK8, P>;
    PHMRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, pipe_threads(P), smem, s>>>(a);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    *handled = true;
    return launch_estep_finalize(a.partials, grid, a.K, D, a.stats_out, s);
}

template <int D, int NK8>
int launch_mma(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    constexpr int F = n_stat_features(D);
    constexpr int NT = (F + 7) / 8;
    if constexpr (NK8 * NT * 2 > 64) {
        return PHMRF_OK;  // accumulator tiles would not fit the consumer's registers
    } else if constexpr (NK8 * NT <= 12 && NK8 <= 3) {
        return launch_mma_p<D, NK8, 12>(a, sm_count, s, handled);
    } else {
        return launch_mma_p<D, NK8, 8>(a, sm_count, s, handled);
    }
}
template <int D>
int launch_pipe_d(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    switch ((a.K + 7) / 8) {
        case 1: return launch_mma<D, 1>(a, sm_count, s, handled);
        case 2: return launch_mma<D, 2>(a, sm_count, s, handled);
        case 3: return launch_mma<D, 3>(a, sm_count, s, handled);
        case 4: return launch_mma<D, 4>(a, sm_count, s, handled);
        case 5: return launch_mma<D, 5>(a, sm_count, s, handled);
        default: return PHMRF_OK;  // K > 40: general kernel
    }
}

__global__ void nbr_g_kernel(const int32_t *__restrict__ nbr_id, const double *__restrict__ nbr_w,
                             double *__restrict__ nbr_g, int64_t count, double beta, int weighted) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        nbr_g[i] = nbr_id[i] >= 0 ? exp(beta * (weighted ? nbr_w[i] : 1.0)) : 1.0;
}

}  // namespace

int launch_nbr_g(const int32_t *nbr_id, const double *nbr_w, double *nbr_g, int64_t count, double beta, int weighted,
                 cudaStream_t s) {
    if (count <= 0) return PHMRF_OK;
    const int64_t blocks = (count + 255) / 256;
    nbr_g_kernel<<<(int)(blocks < 2368 ? blocks : 2368), 256, 0, s>>>(nbr_id, nbr_w, nbr_g, count, beta, weighted);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

int launch_estep_pipe(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    *handled = false;
    if (!a.potts || a.W > kFastSlots || a.pp_soa != nullptr || a.n == 0 || a.nbr_g == nullptr) return PHMRF_OK;
    if (!(fabs(a.s_bound) < 100.0)) return PHMRF_OK;  // exp(S) * exp(600) must stay finite without range checks
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
