// the step immediately before the hot path -- the undirected 8-/4-neighbourhood
// edge list of a region with its edge distances, on the GPU.
//
// Reference: utility.py:1871-1973 edge_weightlist_grid3_undirected_unsym (diagonal region:
// row-major upper triangle incl. the diagonal, utility.py:2310-2317) and utility.py:1975-2053
// edge_weightlist_grid3_undirected (full n1 x n2 rectangle).  Forward directions right,
// lower-right, lower, lower-left (8) or right, lower (4) (utility.py:1898-1916), kept when
// the neighbour lies inside the region; d_ij = |xi-xj|^2 / (|xi||xj| + 1e-16), halved between
// two diagonal nodes of a diagonal region (utility.py:1919-1953); rows sorted by (id1,id2)
// (utility.py:1960), which for a dense region is: per node, neighbours in ascending id.
// Output format is the reference's: [E,3] float64 rows (id1, id2, d_ij).
#include <cub/cub.cuh>

#include "common.cuh"

namespace phmrf {

namespace {

struct Grid {
    int kind;  // 1 = diagonal region (upper triangle incl. diagonal), 0 = rectangle
    long long n1, n2;
    int nn;    // 8 or 4
};

__device__ __forceinline__ long long tri_start(long long B, long long x) { return x * B - (x * (x - 1)) / 2; }

// node id -> (x, y)
__device__ __forceinline__ void node_xy(const Grid &g, long long id, long long &x, long long &y) {
    if (g.kind == 0) {
        x = id / g.n2;
        y = id - x * g.n2;
    } else {
        const long long B = g.n2;
        // largest x with tri_start(x) <= id
        double bf = (double)B + 0.5;
        long long xx = (long long)(bf - sqrt(bf * bf - 2.0 * (double)id));
        if (xx < 0) xx = 0;
        if (xx > B - 1) xx = B - 1;
        while (xx > 0 && tri_start(B, xx) > id) --xx;
        while (xx < B - 1 && tri_start(B, xx + 1) <= id) ++xx;
        x = xx;
        y = xx + (id - tri_start(B, xx));
    }
}

__device__ __forceinline__ long long node_id(const Grid &g, long long x, long long y) {
    return g.kind == 0 ? x * g.n2 + y : tri_start(g.n2, x) + (y - x);
}

// the forward neighbours of (x,y) in ascending id order; returns how many are inside the region
__device__ __forceinline__ int forward_neighbours(const Grid &g, long long x, long long y, long long (&nid)[4],
                                                  bool (&both_diag)[4]) {
    int c = 0;
    const bool tri = g.kind != 0;
    const long long rows = tri ? g.n2 : g.n1, cols = g.n2;
    auto inside = [&](long long a, long long b) { return a >= 0 && a < rows && b >= 0 && b < cols && (!tri || a <= b); };
    // right
    if (inside(x, y + 1)) {
        nid[c] = node_id(g, x, y + 1);
        both_diag[c] = false;
        ++c;
    }
    if (g.nn == 8 && inside(x + 1, y - 1)) {  // lower left
        nid[c] = node_id(g, x + 1, y - 1);
        both_diag[c] = false;
        ++c;
    }
    if (inside(x + 1, y)) {  // lower
        nid[c] = node_id(g, x + 1, y);
        both_diag[c] = false;
        ++c;
    }
    if (g.nn == 8 && inside(x + 1, y + 1)) {  // lower right
        nid[c] = node_id(g, x + 1, y + 1);
        both_diag[c] = tri && x == y;
        ++c;
    }
    return c;
}

__global__ void grid_count_kernel(Grid g, long long n, int *__restrict__ counts) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long x, y, nid[4];
        bool bd[4];
        node_xy(g, i, x, y);
        counts[i] = forward_neighbours(g, x, y, nid, bd);
    }
}

__global__ void grid_fill_kernel(Grid g, long long n, int D, const double *__restrict__ X,
                                 const long long *__restrict__ offsets, double *__restrict__ edge_list) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long x, y, nid[4];
        bool bd[4];
        node_xy(g, i, x, y);
        const int c = forward_neighbours(g, x, y, nid, bd);
        const double *xi = X + i * D;
        double ni = 0.0;
        for (int j = 0; j < D; ++j) ni = fma(xi[j], xi[j], ni);
        ni = sqrt(ni);
        double *out = edge_list + offsets[i] * 3;
        for (int e = 0; e < c; ++e) {
            const double *xj = X + nid[e] * D;
            double nj = 0.0, dd = 0.0;
            for (int j = 0; j < D; ++j) {
                nj = fma(xj[j], xj[j], nj);
                const double t = xi[j] - xj[j];
                dd = fma(t, t, dd);
            }
            nj = sqrt(nj);
            double w = dd / (ni * nj + 1e-16);
            if (bd[e]) w *= 0.5;
            out[0] = (double)i;
            out[1] = (double)nid[e];
            out[2] = w;
            out += 3;
        }
    }
}

}  // namespace

// Closed-form edge count of a dense region.
long long grid_edge_count(int kind, long long n1, long long n2, int nn) {
    if (kind == 0) {
        long long e = n1 * (n2 - 1) + (n1 - 1) * n2;  // right + lower
        if (nn == 8) e += 2 * (n1 - 1) * (n2 - 1);     // both diagonals
        return e;
    }
    const long long B = n2;
    long long e = B * (B - 1) / 2 /*right*/ + B * (B - 1) / 2 /*lower: (x+1,y), y>=x+1*/;
    if (nn == 8) e += B * (B - 1) / 2 /*lower right*/ + (B - 1) * (B - 2) / 2 /*lower left: y-1>=x+1*/;
    return e;
}

int launch_grid_edges(const double *X_dev, int kind, long long n1, long long n2, int nn, int D, long long n,
                      long long n_edges, double *edge_list_dev, cudaStream_t s) {
    Grid g{kind, n1, n2, nn};
    int *counts = nullptr;
    long long *offsets = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    PHMRF_CUDA(cudaMalloc((void **)&counts, sizeof(int) * (n + 1)));
    PHMRF_CUDA(cudaMalloc((void **)&offsets, sizeof(long long) * (n + 1)));
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    grid_count_kernel<<<grid, 256, 0, s>>>(g, n, counts);
    count_launch();
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, offsets, n, s);
    PHMRF_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, offsets, n, s);
    count_launch();
    grid_fill_kernel<<<grid, 256, 0, s>>>(g, n, D, X_dev, offsets, edge_list_dev);
    count_launch();
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(counts);
    cudaFree(offsets);
    cudaFree(tmp);
    (void)n_edges;
    if (e != cudaSuccess) return cuda_fail(e, "grid edge kernels", __FILE__, __LINE__);
    PHMRF_CUDA(cudaGetLastError());
    return PHMRF_OK;
}

}  // namespace phmrf
