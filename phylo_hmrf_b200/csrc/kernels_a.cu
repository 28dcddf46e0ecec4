// Phase A of the Phylo-HMRF E-step on sm_100a: emission log-likelihood, the deterministic
// max|logp| reduction behind pygco's down_weight_factor, and the integer cost arrays.
//
// Reference arithmetic: sklearn-0.18 _log_multivariate_normal_density_full behind
// phylo_hmrf.py:266-268; unary = -logprob (phylo_hmrf.py:490); pygco's float->int
// conversion behind phylo_hmrf.py:496-498.
#include "common.cuh"

namespace phmrf {

// ------------------------------------------------------------------------------------
// layout helpers (one-off per region / per host read-back; not on the per-iteration path)
// ------------------------------------------------------------------------------------
__global__ void aos_to_soa_kernel(const double *__restrict__ aos, double *__restrict__ soa, int64_t n, int D,
                                  int64_t ld) {
    int64_t total = n * D;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / D;
        int j = (int)(e - i * D);
        soa[j * ld + i] = aos[e];
    }
}

int launch_aos_to_soa(const double *X_aos, double *X_soa, int64_t n, int D, int64_t ld, cudaStream_t s) {
    if (n == 0) return PHMRF_OK;
    int64_t total = n * D;
    int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    aos_to_soa_kernel<<<grid, 256, 0, s>>>(X_aos, X_soa, n, D, ld);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

// host order [n,K] <-> tiled log-likelihood (lp_index): one warp per 32-node tile, staged
// through shared memory so that both sides stay coalesced (one-off per host read-back / upload)
template <bool TO_AOS>
__global__ void logp_convert_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t n, int K) {
    extern __shared__ double cv_tile[];  // [warps][32*K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *tile = cv_tile + (size_t)warp * 32 * K;
    const int KP = logp_rows(K);
    const int64_t n_tiles = (n + 31) >> 5;
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; t < n_tiles;
         t += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t base = t << 5;
        const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
        if (TO_AOS) {
            for (int k = 0; k < K; ++k)
                if (lane < cnt) tile[lane * K + k] = src[lp_index(k, base + lane, KP)];
            __syncwarp();
            for (int e = lane; e < cnt * K; e += 32) dst[base * K + e] = tile[e];
        } else {
            for (int e = lane; e < cnt * K; e += 32) tile[e] = src[base * K + e];
            __syncwarp();
            for (int k = 0; k < K; ++k)
                if (lane < cnt) dst[lp_index(k, base + lane, KP)] = tile[lane * K + k];
        }
        __syncwarp();
    }
}

template <bool TO_AOS>
static int launch_logp_convert(const double *src, double *dst, int64_t n, int K, cudaStream_t s) {
    if (n == 0) return PHMRF_OK;
    int warps = 8;
    while (warps > 1 && (size_t)warps * 32 * K * 8 > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * 32 * K * 8;
    if (smem > 200 * 1024) {
        set_error("n_states too large for the layout conversion tile");
        return PHMRF_E_UNSUPPORTED;
    }
    auto kern = logp_convert_kernel<TO_AOS>;
    if (smem > 48 * 1024) PHMRF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = ((n + 31) / 32 + warps - 1) / warps;
    kern<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), warps * 32, smem, s>>>(src, dst, n, K);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

int launch_logp_from_aos(const double *aos, double *logp, int64_t n, int K, cudaStream_t s) {
    return launch_logp_convert<false>(aos, logp, n, K, s);
}
int launch_logp_to_aos(const double *logp, double *aos, int64_t n, int K, cudaStream_t s) {
    return launch_logp_convert<true>(logp, aos, n, K, s);
}

// [K][ld] -> [n][K] through a 32x32 shared-memory tile so both sides stay coalesced.
__global__ void soa_to_aos_kernel(const double *__restrict__ soa, double *__restrict__ aos, int64_t n, int K,
                                  int64_t ld) {
    __shared__ double tile[32][33];
    int64_t n_tiles_i = (n + 31) / 32;
    int k_tiles = (K + 31) / 32;
    int64_t total = n_tiles_i * k_tiles;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        int64_t ti = t / k_tiles;
        int tk = (int)(t - ti * k_tiles);
        int64_t i0 = ti * 32;
        int k0 = tk * 32;
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            int k = k0 + r;
            int64_t i = i0 + threadIdx.x;
            tile[r][threadIdx.x] = (k < K && i < n) ? soa[k * ld + i] : 0.0;
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            int64_t i = i0 + r;
            int k = k0 + threadIdx.x;
            if (i < n && k < K) aos[i * K + k] = tile[threadIdx.x][r];
        }
        __syncthreads();
    }
}

int launch_soa_to_aos(const double *soa, double *aos, int64_t n, int K, int64_t ld, cudaStream_t s) {
    if (n == 0) return PHMRF_OK;
    int64_t total = ((n + 31) / 32) * ((K + 31) / 32);
    int grid = (int)(total < 148 * 8 ? total : 148 * 8);
    soa_to_aos_kernel<<<grid, dim3(32, 8), 0, s>>>(soa, aos, n, K, ld);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

// ------------------------------------------------------------------------------------
// A1: emission.  One thread owns NPT consecutive nodes whose d features sit in registers.
// The packed per-state factors (a stream in the order the arithmetic consumes it:
// hc, then per row i: c_i, W_i0..W_ii) are staged once per block in shared memory and read
// back as warp-uniform (broadcast) 128-bit loads, so the inner loop is d(d+3)/2 DFMA per
// node-state and NPT*d(d+3)/2 DFMA per (PS+1)/2 shared loads.  (A first version indexed
// the constant bank with the state id: ncu showed the ADU pipe at 81 % and the FP64 pipe
// at 20 % -- indexed LDC is the wrong tool for per-state tables.)
// The FP64 pipe then saturates at ~75 %: a DFMA with three distinct vector-register
// operands is register-file limited to 27.4 TFLOP/s on B200 (csrc/probe.cu, probe 6), which
// is what this kernel reaches.  Passing the factors as kernel parameters (constant-bank /
// uniform-register multiplier, two vector operands) was tried and measured slower: 204
// registers -> 8 warps/SM, LDCU traffic on the ADU pipe, 63 % FP64 (profiles/r1_history.md).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long abs_bits(double v) {
    return (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
}

template <int D, bool SMEM_MODEL, int NPT>
__global__ void __launch_bounds__(256) emit_kernel(const double *__restrict__ Xs, int64_t n, int64_t ld, int K,
                                                    const double *__restrict__ model, double *__restrict__ logp,
                                                    unsigned long long *absmax_bits) {
    static_assert(NPT == 4, "the tiled log-likelihood layout keeps groups of 4 nodes contiguous");
    constexpr int PS = model_stride(D);
    constexpr int PSs = (PS + 1) & ~1;  // even stride: every state starts 16-byte aligned
    extern __shared__ __align__(16) double smodel[];
    if (SMEM_MODEL) {
        for (int e = threadIdx.x; e < K * PSs; e += blockDim.x) {
            const int k = e / PSs, q = e - k * PSs;
            smodel[e] = q < PS ? model[(int64_t)k * PS + q] : 0.0;
        }
        __syncthreads();
    }
    unsigned long long amax = 0ull;
    const int KP = logp_rows(K);
    const int64_t n_groups = ld / NPT;  // ld is a multiple of 64; the pad region holds zeros
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_groups;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = p * NPT;
        if (i0 >= n) break;
        double x[NPT][D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
#pragma unroll
            for (int u = 0; u < NPT; u += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(Xs + j * ld + i0 + u);
                x[u][j] = v.x;
                x[u + 1][j] = v.y;
            }
        }
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            double mv[PSs];
            if (SMEM_MODEL) {
                const double2 *m2 = reinterpret_cast<const double2 *>(smodel + k * PSs);
#pragma unroll
                for (int c = 0; c < PSs / 2; ++c) {
                    const double2 v = m2[c];
                    mv[2 * c] = v.x;
                    mv[2 * c + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < PS; ++c) mv[c] = __ldg(model + (int64_t)k * PS + c);
            }
            double acc[NPT];
#pragma unroll
            for (int u = 0; u < NPT; ++u) acc[u] = mv[0];
            int q = 1;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double z[NPT];
#pragma unroll
                for (int u = 0; u < NPT; ++u) z[u] = -mv[q];
                ++q;
#pragma unroll
                for (int j = 0; j <= i; ++j) {
#pragma unroll
                    for (int u = 0; u < NPT; ++u) z[u] = fma(mv[q], x[u][j], z[u]);
                    ++q;
                }
#pragma unroll
                for (int u = 0; u < NPT; ++u) acc[u] = fma(z[u], z[u], acc[u]);
            }
            double l[NPT];
#pragma unroll
            for (int u = 0; u < NPT; ++u) l[u] = -acc[u];
            // tiled layout (lp_index): NPT consecutive nodes stay contiguous under the swizzle
            double *dst = logp + lp_index(k, i0, KP);
#pragma unroll
            for (int u = 0; u < NPT; u += 2)
                *reinterpret_cast<double2 *>(dst + u) = make_double2(l[u], l[u + 1]);
#pragma unroll
            for (int u = 0; u < NPT; ++u) {
                const unsigned long long b = (i0 + u < n) ? abs_bits(l[u]) : 0ull;
                amax = b > amax ? b : amax;
            }
        }
    }
    // warp -> block -> device maximum; integer max of the |.| bit pattern is order
    // independent (deterministic) and lets a NaN win, like numpy's max.
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, amax, o);
        amax = other > amax ? other : amax;
    }
    __shared__ unsigned long long wmax[8];
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = wmax[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = wmax[w] > m ? wmax[w] : m;
        atomicMax(absmax_bits, m);
    }
}

template <int D>
static int launch_emit_d(const double *Xs, int64_t n, int64_t ld, int K, const double *model, double *logp,
                         unsigned long long *absmax_bits, int sm_count, cudaStream_t s) {
    constexpr int PSs = (model_stride(D) + 1) & ~1;
    constexpr int NPT = 4, TPB = 256;  // 8 nodes/thread measured identical (register-file bound either way)
    const int64_t n_groups = (n + NPT - 1) / NPT;
    const int64_t blocks = (n_groups + TPB - 1) / TPB;
    const int64_t cap = (int64_t)sm_count * 4;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (grid < 1) grid = 1;
    const size_t smem = (size_t)K * PSs * sizeof(double);
    if (smem <= 100 * 1024) {
        if (smem > 48 * 1024)
            PHMRF_CUDA(cudaFuncSetAttribute(emit_kernel<D, true, NPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        emit_kernel<D, true, NPT><<<grid, TPB, smem, s>>>(Xs, n, ld, K, model, logp, absmax_bits);
    } else {
        emit_kernel<D, false, NPT><<<grid, TPB, 0, s>>>(Xs, n, ld, K, model, logp, absmax_bits);
    }
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

int launch_emit(const double *Xs, int64_t n, int64_t ld, int D, int K, const double *model_global, double *logp,
                unsigned long long *absmax_bits, int sm_count, cudaStream_t s) {
    PHMRF_CUDA(cudaMemsetAsync(absmax_bits, 0, sizeof(unsigned long long), s));
    if (n == 0) return PHMRF_OK;
    switch (D) {
#define PHMRF_CASE(DD) \
    case DD:           \
        return launch_emit_d<DD>(Xs, n, ld, K, model_global, logp, absmax_bits, sm_count, s);
        PHMRF_CASE(1) PHMRF_CASE(2) PHMRF_CASE(3) PHMRF_CASE(4) PHMRF_CASE(5) PHMRF_CASE(6)
        PHMRF_CASE(7) PHMRF_CASE(8) PHMRF_CASE(9) PHMRF_CASE(10) PHMRF_CASE(11) PHMRF_CASE(12)
#undef PHMRF_CASE
    }
    set_error("n_features outside [1,12]");
    return PHMRF_E_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------
// dwf = max(max|unary|, max|w| * max V) + 1e-10   (pygco, down_weight_factor=None)
// ------------------------------------------------------------------------------------
__global__ void fill_kernel(double *p, double v, int64_t count) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

__global__ void dwf_kernel(const unsigned long long *absmax_bits, double wmax, double vmax, double dwf_in,
                           double *dwf_dev) {
    if (dwf_in > 0.0) {
        dwf_dev[0] = dwf_in;
    } else {
        double umax = __longlong_as_double((long long)absmax_bits[0]);
        double pw = __dmul_rn(wmax, vmax);
        // python's max(a, b) returns a unless b > a (so a NaN in `a` survives)
        double m = (pw > umax) ? pw : umax;
        dwf_dev[0] = __dadd_rn(m, 1e-10);
    }
    dwf_dev[1] = __longlong_as_double((long long)absmax_bits[0]);
}

__global__ void publish_kernel(unsigned long long *__restrict__ dst, const void *__restrict__ src, int count, bool src32) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        dst[i] = src32 ? (unsigned long long)(long long)reinterpret_cast<const int *>(src)[i]
                       : reinterpret_cast<const unsigned long long *>(src)[i];
    __threadfence_system();
}

int launch_publish(unsigned long long *dst_mapped, const void *src, int count, bool src32, cudaStream_t s) {
    if (count <= 0) return PHMRF_OK;
    const int blocks = (count + 255) / 256;
    publish_kernel<<<blocks < 32 ? blocks : 32, 256, 0, s>>>(dst_mapped, src, count, src32);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

int launch_fill(double *p, double v, int64_t count, cudaStream_t s) {
    if (count <= 0) return PHMRF_OK;
    const int64_t blocks = (count + 255) / 256;
    fill_kernel<<<(int)(blocks < 1184 ? blocks : 1184), 256, 0, s>>>(p, v, count);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

int launch_dwf(const unsigned long long *absmax_bits, double wmax, double vmax, double dwf_in, double *dwf_dev,
               cudaStream_t s) {
    dwf_kernel<<<1, 1, 0, s>>>(absmax_bits, wmax, vmax, dwf_in, dwf_dev);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

// ------------------------------------------------------------------------------------
// A2: integer unary.  Two separately rounded FP64 operations then truncation toward zero,
// exactly numpy's ((u / dwf) * 1e5).astype(intc).  The [K][ld] -> [n][K] transposition goes
// through shared memory so that both the FP64 reads and the int32 writes are coalesced.
// ------------------------------------------------------------------------------------
// Correctly rounded u / dwf from the correctly rounded reciprocal y = RN(1/dwf): two
// Markstein refinements (q1 is a faithful quotient, so q2 = RN(q1 + RN(u - dwf*q1)*y) is
// the IEEE quotient; Markstein 1990 / Handbook of Floating-Point Arithmetic, sec. 5.3).
// 5 FP64-pipe instructions instead of the ~15 + MUFU of the generic division routine.
__device__ __forceinline__ double div_by_dwf(double u, double dwf, double y) {
    const double q0 = __dmul_rn(u, y);
    const double r0 = __fma_rn(-q0, dwf, u);
    const double q1 = __fma_rn(r0, y, q0);
    const double r1 = __fma_rn(-q1, dwf, u);
    return __fma_rn(r1, y, q1);
}

// One warp owns 32 consecutive nodes (one tile of the log-likelihood layout): it reads the K
// state rows of the tile (256 B each, coalesced), converts, stages the 32 x K integers in its
// own shared-memory tile and streams them out as one contiguous 128*K-byte run.  No block-wide
// barrier: the warps of a CTA are independent, so loads of one overlap stores of another.
// On the way it writes the per-node maximum over the states (phase B's soft-max shift; the
// FP64 pipe idles in this HBM-bound kernel), and the same launch converts the edge weights
// (pygco converts them on every call because dwf changes with the model).
__global__ void __launch_bounds__(256, 4) quantise_kernel(const double *__restrict__ logp, int64_t n, int K,
                                                          const double *__restrict__ dwf_dev, double tol, double uprec,
                                                          int32_t *__restrict__ unary, double *__restrict__ rowmax,
                                                          long long *__restrict__ blist, long long bcap,
                                                          unsigned long long *bcount, const double *__restrict__ edge_w,
                                                          int64_t E, double wprec, int32_t *__restrict__ w_i32) {
    extern __shared__ __align__(16) int32_t tile_all[];  // [warps][32][K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t *tile = tile_all + (size_t)warp * 32 * K;
    const double dwf = dwf_dev[0];
    const double ydwf = 1.0 / dwf;  // IEEE division: the correctly rounded reciprocal
    // the refinement needs a finite, normal reciprocal; otherwise use the plain division
    const bool fast_div = (dwf > 1e-290) && (dwf < 1e290);
    constexpr int U = 6;  // states in flight per thread
    const int KP = logp_rows(K);
    const int64_t n_tiles = (n + 31) >> 5;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; t < n_tiles; t += wstride) {
        const int64_t base = t << 5;
        const int64_t i = base + lane;
        const int cnt = (int)((n - base) < 32 ? (n - base) : 32);
        if (i < n) {
            const double *src = logp + t * KP * 32;
            double rmax = -INFINITY;
            for (int k0 = 0; k0 < K; k0 += U) {
                double u[U];
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    const int k = k0 + q;
                    u[q] = k < K ? src[k * 32 + (lane ^ ((k & 7) << 2))] : -INFINITY;
                }
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    const int k = k0 + q;
                    if (k < K) {
                        rmax = fmax(rmax, u[q]);
                        const double nu = -u[q];
                        const double quo = fast_div ? div_by_dwf(nu, dwf, ydwf) : __ddiv_rn(nu, dwf);
                        const double tq = __dmul_rn(quo, uprec);
                        tile[lane * K + k] = __double2int_rz(tq);
                        // |tq| <= uprec < 2^51: nearest integer through the 2^52+2^51 constant
                        const double r = (tq + 6755399441055744.0) - 6755399441055744.0;
                        const double dist = fabs(tq - r);
                        if (dist <= tol || dist <= tol * fabs(tq)) {
                            unsigned long long pos = atomicAdd(bcount, 1ull);
                            if ((long long)pos < bcap) blist[pos] = (long long)(i * K + k);
                        }
                    }
                }
            }
            rowmax[i] = rmax;
        }
        __syncwarp();
        int32_t *dst = unary + base * K;
        const int total = cnt * K;
        if ((total & 3) == 0) {  // base*K*4 is a multiple of 128 bytes
            const int4 *src4 = reinterpret_cast<const int4 *>(tile);
            int4 *dst4 = reinterpret_cast<int4 *>(dst);
            for (int e = lane; e < (total >> 2); e += 32) dst4[e] = src4[e];
        } else {
            for (int e = lane; e < total; e += 32) dst[e] = tile[e];
        }
        __syncwarp();
    }
    // edge weights: w_i32[e] = trunc((w[e] / dwf) * wprec)
    if (edge_w != nullptr) {
        const int64_t tstride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += tstride) {
            const double w = edge_w[e];
            const double quo = fast_div ? div_by_dwf(w, dwf, ydwf) : __ddiv_rn(w, dwf);
            w_i32[e] = __double2int_rz(__dmul_rn(quo, wprec));
        }
    }
}

int launch_quantise(const double *logp, int64_t n, int K, const double *dwf_dev, double tol, double uprec,
                    int32_t *unary, double *rowmax, long long *blist, long long bcap, unsigned long long *bcount,
                    const double *edge_w, int64_t E, double wprec, int32_t *w_i32, int sm_count, cudaStream_t s) {
    PHMRF_CUDA(cudaMemsetAsync(bcount, 0, sizeof(unsigned long long), s));
    if (n == 0 && (edge_w == nullptr || E == 0)) return PHMRF_OK;
    int warps = 8;
    while (warps > 1 && (size_t)warps * 32 * K * 4 > 64 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * 32 * K * 4;
    if (smem > 200 * 1024) {
        set_error("n_states too large for the quantise tile");
        return PHMRF_E_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        PHMRF_CUDA(cudaFuncSetAttribute(quantise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_tiles = (n + 31) / 32;
    int64_t blocks = (n_tiles + warps - 1) / warps;
    const int64_t eblocks = edge_w ? (E + warps * 32 - 1) / (warps * 32) : 0;
    if (eblocks > blocks) blocks = eblocks;
    const int64_t cap = (int64_t)sm_count * 4;  // one resident wave (4 CTAs/SM at <= 64 registers)
    const int grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
    quantise_kernel<<<grid, warps * 32, smem, s>>>(logp, n, K, dwf_dev, tol, uprec, unary, rowmax, blist, bcap, bcount,
                                                    edge_w, edge_w ? E : 0, wprec, w_i32);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

// per-node maximum over the states for a log-likelihood the quantise kernel has not seen
// (phmrf_set_logprob followed directly by the E-step); optionally max|logp| as well
__global__ void rowmax_kernel(const double *__restrict__ logp, int64_t n, int K, double *__restrict__ rowmax,
                              unsigned long long *absmax_bits) {
    const int KP = logp_rows(K);
    unsigned long long amax = 0ull;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double m = -INFINITY;
        for (int k = 0; k < K; ++k) {
            const double v = logp[lp_index(k, i, KP)];
            m = fmax(m, v);
            const unsigned long long b = abs_bits(v);
            amax = b > amax ? b : amax;
        }
        rowmax[i] = m;
    }
    if (absmax_bits != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, amax, o);
            amax = other > amax ? other : amax;
        }
        if ((threadIdx.x & 31) == 0 && amax) atomicMax(absmax_bits, amax);
    }
}

int launch_rowmax(const double *logp, int64_t n, int K, double *rowmax, unsigned long long *absmax_bits,
                  cudaStream_t s) {
    if (n == 0) return PHMRF_OK;
    const int64_t blocks = (n + 255) / 256;
    rowmax_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, s>>>(logp, n, K, rowmax, absmax_bits);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

// labels must lie in [0, K): *first_bad = smallest offending index, -1 when all are valid
__global__ void check_labels_kernel(const int32_t *__restrict__ labels, int64_t n, int K, long long *first_bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = labels[i];
        if (l < 0 || l >= K) atomicMin((unsigned long long *)first_bad, (unsigned long long)i);
    }
}

int launch_check_labels(const int32_t *labels, int64_t n, int K, long long *first_bad, cudaStream_t s) {
    PHMRF_CUDA(cudaMemsetAsync(first_bad, 0xff, sizeof(long long), s));  // -1 == max as unsigned
    if (n == 0) return PHMRF_OK;
    const int64_t blocks = (n + 255) / 256;
    check_labels_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, s>>>(labels, n, K, first_bad);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

// zero the log-likelihood buffer; the padding row of an odd K gets kLogpPad (see common.cuh logp_rows)
__global__ void logp_init_kernel(double *__restrict__ logp, int64_t ld, int K, int KP, double pad) {
    const int64_t total = (ld >> 5) * KP * 32;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)((e >> 5) % KP);
        logp[e] = (k == K && (K & 1)) ? pad : 0.0;
    }
}

int launch_logp_init(double *logp, int64_t ld, int K, cudaStream_t s) {
    const int KP = logp_rows(K);
    const int64_t total = (ld >> 5) * KP * 32;
    const int64_t blocks = (total + 255) / 256;
    logp_init_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, s>>>(logp, ld, K, KP, kLogpPad);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

// Stand-in for the graph cut in benches/tests: first arg-min of the integer unary.
__global__ void argmin_unary_kernel(const int32_t *__restrict__ unary, int64_t n, int K, int32_t *__restrict__ labels) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t *row = unary + i * K;
        int best = 0;
        int32_t bv = row[0];
        for (int k = 1; k < K; ++k) {
            int32_t v = row[k];
            if (v < bv) {
                bv = v;
                best = k;
            }
        }
        labels[i] = best;
    }
}

int launch_argmin_unary(const int32_t *unary, int64_t n, int K, int32_t *labels, cudaStream_t s) {
    if (n == 0) return PHMRF_OK;
    int64_t blocks = (n + 255) / 256;
    int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    argmin_unary_kernel<<<grid, 256, 0, s>>>(unary, n, K, labels);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

}  // namespace phmrf
