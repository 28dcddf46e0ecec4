// Shared declarations between the host API (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/phmrf.h"

namespace phmrf {

constexpr int kMinFeatures = 1;
constexpr int kMaxFeatures = 12;      // D is a template parameter of the kernels

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
void count_launch(int n = 1);

#define PHMRF_CUDA(expr)                                                       \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return ::phmrf::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

// Per-state emission parameters as the kernels consume them (built on the host from the
// Cholesky factor): z = Wh*x - ch with Wh = sqrt(1/2)*L^-1 (lower triangular), ch = Wh*mu,
// and logp = -(hc + sum z^2) with hc = (d*ln(2pi) + logdet)/2.  Stored as the stream the
// emission kernel consumes: hc, then for each row i: ch_i, Wh_i0 .. Wh_ii.
__host__ __device__ constexpr int model_stride(int D) { return D * (D + 1) / 2 + D + 1; }
__host__ __device__ constexpr int n_stat_features(int D) { return 1 + D + D * (D + 1) / 2; }

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Log-likelihood layout in HBM: tile-major, 32 nodes per tile, KP = logp_rows(K) state rows of
// 32 doubles per tile, the node column XOR-swizzled by the state row:
//   (k, i)  ->  ((i / 32) * KP + k) * 32 + ((i % 32) ^ ((k % 8) * 4))
// One tile (KP * 256 bytes) is contiguous, so phase B fetches it with ONE bulk copy
// (cp.async.bulk) straight into shared memory, where the swizzle makes both access patterns
// conflict free: a lane per node along a row (per-node warps) and the FP64 mma operand
// fragments (8 states x 4 nodes; tools/swizzle_check.py).  The XOR acts on multiples of 4,
// so groups of 4 consecutive nodes stay contiguous (the emission kernel's 128-bit stores).
__host__ __device__ __forceinline__ int64_t lp_index(int k, int64_t i, int KP) {
    return ((i >> 5) * KP + k) * 32 + ((i & 31) ^ ((k & 7) << 2));
}

// ---- phase A (kernels_a.cu) -----------------------------------------------------------
// X_aos [n,D] row-major -> X_soa [D][ld]
int launch_aos_to_soa(const double *X_aos, double *X_soa, int64_t n, int D, int64_t ld, cudaStream_t s);
// [K][ld] -> out [n,K] row-major (posteriors, pairwise potential)
int launch_soa_to_aos(const double *soa, double *aos, int64_t n, int K, int64_t ld, cudaStream_t s);
// log-likelihood: host order [n,K] row-major <-> the tiled device layout (lp_index)
int launch_logp_from_aos(const double *aos, double *logp, int64_t n, int K, cudaStream_t s);
int launch_logp_to_aos(const double *logp, double *aos, int64_t n, int K, cudaStream_t s);
// emission: logp[k][i] for i<n, block maxima of |logp| folded into *absmax_bits (uint64 bit
// pattern of a non-negative double, atomicMax).
int launch_emit(const double *X_soa, int64_t n, int64_t ld, int D, int K, const double *model_global,
                double *logp, unsigned long long *absmax_bits, int sm_count, cudaStream_t s);
// per-node maximum over the states (phase B's soft-max shift); the quantise kernel writes it
// on its way, this is the stand-alone form for a caller-supplied log-likelihood
int launch_rowmax(const double *logp, int64_t n, int K, double *rowmax, unsigned long long *absmax_bits,
                  cudaStream_t s);
int launch_check_labels(const int32_t *labels, int64_t n, int K, long long *first_bad, cudaStream_t s);
int launch_fill(double *p, double v, int64_t count, cudaStream_t s);
// Small results to the host without the copy engines: the kernel stores `count` 64-bit words (or,
// with src32, sign-extended 32-bit words) into a host-mapped pinned buffer.  A DMA read-back of a few
// bytes queues behind whatever bulk download another region has in flight (tens of milliseconds);
// stores from an SM do not.
int launch_publish(unsigned long long *dst_mapped, const void *src, int count, bool src32, cudaStream_t s);
// rows of the log-likelihood matrix: K rounded up to the 8-state tiles of the pipeline kernel.  The
// padding rows keep their values for the lifetime of a region: row K of an odd K holds kLogpPad (the
// pipeline evaluates states in pairs: exp(-1e6 - shift) is a zero weight), the rows from the next even
// number on hold 0.0 -- the pipeline never evaluates them, and the bulk copy then drops a zero WEIGHT
// for them straight into the soft-max buffer the statistics product reads.
__host__ __device__ inline int logp_rows(int K) { return (K + 7) / 8 * 8; }
constexpr double kLogpPad = -1.0e6;
int launch_logp_init(double *logp, int64_t ld, int K, cudaStream_t s);
// dwf = max(absmax_u, wmax*vmax) + 1e-10 unless dwf_in > 0; written to *dwf_dev.
int launch_dwf(const unsigned long long *absmax_bits, double wmax, double vmax, double dwf_in, double *dwf_dev,
               cudaStream_t s);
// unary_i32[i*K+k] = trunc(((-logp[k][i])/dwf)*uprec); boundary entries appended to blist;
// rowmax[i] = max_k logp[k][i]; and, in the same launch, w_i32[e] = trunc((w[e]/dwf)*wprec)
// for the E edge weights (edge_w == nullptr skips it).
int launch_quantise(const double *logp, int64_t n, int K, const double *dwf_dev, double tol, double uprec,
                    int32_t *unary, double *rowmax, long long *blist, long long bcap, unsigned long long *bcount,
                    const double *edge_w, int64_t E, double wprec, int32_t *w_i32, int sm_count, cudaStream_t s);
int launch_argmin_unary(const int32_t *unary, int64_t n, int K, int32_t *labels, cudaStream_t s);

// ---- phase B (kernels_b.cu) -----------------------------------------------------------
struct EstepArgs {
    const double *X_soa;      // [D][ld]
    const double *logp;       // tiled [n/32][KP][32], see lp_index
    const double *rowmax;     // [ld] max_k logp (pipeline kernel only)
    const int32_t *labels;    // [n_window]
    const int32_t *nbr_id;    // [W][ld] window-local neighbour id, -1 = none
    const double *nbr_w;      // [W][ld] edge weight per slot
    const double *V;          // [K][K] (general path only)
    double beta;              // Potts strength (fast path)
    int potts;
    int64_t n, ld, own_offset;
    int D, K, W, estimate_type;
    double *post_soa;         // nullable [K][ld]
    double *pp_soa;           // nullable [K][ld] pairwise potential (signature-parity output)
    double *partials;         // [sm_count][K*F + 3] per-CTA partial sums
    double *stats_out;        // [K*(1+D+D*D) + 3]
    int *flags;               // [1] device flag: bit 0 = the pipeline met an overflow, rerun on the general path
    int force_general;        // skip the pipeline kernel
    double s_bound;           // upper bound of any neighbour weight sum: beta * W * max|w|
    const double *nbr_g;      // [W][ld] exp(beta * w) per slot (1 for an empty slot); pipeline kernel only
    double exp_beta;          // exp(beta)
    // implicit-grid form (regions built by phmrf_region_create_grid): neighbours are fixed offsets
    // from the row geometry, each edge weight is stored once, with its forward end
    const double2 *fwd_wg;    // [4][ldw] per window node: {w, exp(beta*w)} of right, lower-left, lower, lower-right
    int64_t ldw;
    int grid_kind;            // 1 diagonal region (upper triangle), 0 rectangle, -1: explicit neighbour slots
    int grid_nn;              // 8 or 4
    int64_t grid_n2, grid_rows;
    int64_t own_start_gid;    // region-global node id of the first owned node
};
// g[s][i] = exp(beta * (weighted ? w[s][i] : 1)) for occupied slots, 1 otherwise (kernels_b.cu)
int launch_nbr_g(const int32_t *nbr_id, const double *nbr_w, double *nbr_g, int64_t count, double beta, int weighted,
                 cudaStream_t s);
int launch_estep(const EstepArgs &a, int sm_count, cudaStream_t s);
// Warp-specialised bulk-copy pipeline (estep_bulk.cuh, kernels_b3*.cu), the fast path; *handled=false when the
// shape is outside its range.
int launch_estep_bulk(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled);
// Fold per-block partials into stats_out (defined in kernels_b.cu).
int launch_estep_finalize(const double *partials, int n_blocks, int K, int D, double *stats_out, cudaStream_t s);

// ---- region edge lists on the GPU (kernels_grid.cu) ----------------------
long long grid_edge_count(int kind, long long n1, long long n2, int nn);
int launch_grid_edges(const double *X_dev, int kind, long long n1, long long n2, int nn, int D, long long n,
                      long long n_edges, double *edge_list_dev, cudaStream_t s);

long long grid_row_start(int kind, long long n1, long long n2, long long row);
int launch_band_graph(const double *Xw_dev, int kind, long long n1, long long n2, int nn, int D, long long win_start,
                      long long own_start, long long own_end, long long n_window, double beta1, long long ld,
                      int32_t *nbr_id, double *nbr_w, long long *ids_dev, double *w_dev, long long *n_edges,
                      unsigned long long *wmax_bits, cudaStream_t s);

// forward-edge weights of the window nodes [0, n_fw) (implicit-grid form of phase B)
int launch_band_fwd(const double *Xw_dev, int kind, long long n1, long long n2, int nn, int D, long long win_start,
                    long long n_fw, double beta1, long long ldw, double2 *fwd, cudaStream_t s);
int launch_fwd_factor(double2 *fwd, long long count, double beta, int weighted, cudaStream_t s);

}  // namespace phmrf
