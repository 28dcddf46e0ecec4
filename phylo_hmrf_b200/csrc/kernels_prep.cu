// Preprocessing of one region: per-species rescale + log transform,
// contact image, 3x3 median hole fill, Perona-Malik diffusion, node-order flatten.
// Reference: utility.py:867-897, 362, 2192-2226, 2332-2365, 603-659, 1566-1573 (medpy call),
// 2295-2329, 2368-2400.  One-shot per data set and HBM-/latency-bound; everything runs on the device
// so that a 10 kb chr1 image (6.2e8 pixels per species) never exists on the host.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace phmrf {

namespace {

constexpr double kThresh = 1e-05;  // utility.py:47 THRESH1

// ---- normalise -------------------------------------------------------------------------------
// Column minima / maxima of max(x, 0): non-negative doubles order like their bit patterns.
__global__ void colminmax_kernel(const double *__restrict__ x, int64_t n, int d, unsigned long long *mn,
                                 unsigned long long *mx) {
    const int c = blockIdx.y;
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i * d + c];
        v = v < 0 ? 0.0 : v;  // (a NaN stays a NaN and is not ordered: left to the caller like numpy's min)
        const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
        lo = b < lo ? b : lo;
        hi = b > hi ? b : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mn + c, lo);
        atomicMax(mx + c, hi);
    }
}

// x = x_min + (x - m1) * 1.0 * (x_max - x_min) / (m2 - m1), then log(1 + x); numpy's operation
// order, no contraction.
__global__ void rescale_kernel(double *__restrict__ x, int64_t n, int d, const double *__restrict__ colmm, double x_min,
                               double x_max, int log1p) {
    const double span = __dsub_rn(x_max, x_min);
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n * d; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % d);
        const double m1 = colmm[2 * c], m2 = colmm[2 * c + 1];
        double v = x[e];
        v = v < 0 ? 0.0 : v;
        double y = __dmul_rn(__dmul_rn(__dsub_rn(v, m1), 1.0), span);
        y = __ddiv_rn(y, __dsub_rn(m2, m1));
        y = __dadd_rn(x_min, y);
        x[e] = log1p ? log(__dadd_rn(1.0, y)) : y;
    }
}

// ---- image -----------------------------------------------------------------------------------
__global__ void scatter_kernel(const double *__restrict__ value, const int64_t *__restrict__ pos, int64_t n, int d,
                               int c, int64_t start1, int64_t start2, int64_t n1, int64_t n2, int symmetric,
                               double *__restrict__ plane, int *bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = pos[2 * i] - start1, b = pos[2 * i + 1] - start2;
        if (a < 0 || b < 0 || a >= n1 || b >= n2 || (symmetric && (b >= n1 || a >= n2))) {
            *bad = 1;
            continue;
        }
        const double v = value[i * d + c];
        plane[a * n2 + b] = v;
        if (symmetric) plane[b * n2 + a] = v;
    }
}

// median of the 8 neighbours = mean of the 4th and 5th smallest (numpy.median of an even count)
__device__ __forceinline__ double median8(double (&w)[8]) {
#pragma unroll
    for (int a = 1; a < 8; ++a) {
#pragma unroll
        for (int b = a; b > 0; --b) {
            const double lo = fmin(w[b - 1], w[b]), hi = fmax(w[b - 1], w[b]);
            w[b - 1] = lo;
            w[b] = hi;
        }
    }
    return __ddiv_rn(__dadd_rn(w[3], w[4]), 2.0);
}

// 3x3 median hole fill in the reference's sequential raster order.  Cell (i,j) reads row i-1,
// (i,j-1) -- already updated -- and (i,j+1), row i+1 -- not yet updated.  All cells with the same
// t = 2i + j are independent and every updated neighbour has a smaller t, so sweeping t in order
// with a barrier in between reproduces the sequential result exactly.  One CTA per plane.
// symmetric != 0: only the upper triangle (j >= i) is scanned and a cell below the diagonal is
// read through its mirror, which is what the reference's mirrored writes amount to on a
// symmetric image (utility.py:603-630); the lower triangle is rewritten afterwards.
__global__ void __launch_bounds__(1024) holefill_kernel(double *__restrict__ planes, int64_t plane_stride, int64_t n1,
                                                        int64_t n2, int symmetric) {
    double *P = planes + (int64_t)blockIdx.x * plane_stride;
    const int64_t i_lo = 2, i_hi = n1 - 2, j_hi = n2 - 2;  // inclusive bounds of range(2, n-1)
    if (i_hi < i_lo || j_hi < 2) return;
    const int64_t j_lo0 = 2;
    const int64_t t_first = 2 * i_lo + (symmetric ? i_lo : j_lo0), t_last = 2 * i_hi + j_hi;
    for (int64_t t = t_first; t <= t_last; ++t) {
        // rows with a cell on this front: j = t - 2i within [jmin(i), j_hi]
        int64_t ia = (t - j_hi + 1) / 2;  // ceil((t - j_hi) / 2) for t - j_hi >= 0
        if (t - j_hi < 0) ia = i_lo;
        ia = ia < i_lo ? i_lo : ia;
        int64_t ib = symmetric ? t / 3 : (t - j_lo0) / 2;
        ib = ib > i_hi ? i_hi : ib;
        for (int64_t i = ia + threadIdx.x; i <= ib; i += blockDim.x) {
            const int64_t j = t - 2 * i;
            const double v = P[i * n2 + j];
            if (v < kThresh) {
                double w[8];
                int q = 0;
#pragma unroll
                for (int di = -1; di <= 1; ++di)
#pragma unroll
                    for (int dj = -1; dj <= 1; ++dj) {
                        if (di == 0 && dj == 0) continue;
                        int64_t a = i + di, b = j + dj;
                        if (symmetric && a > b) {
                            const int64_t s = a;
                            a = b;
                            b = s;
                        }
                        w[q++] = P[a * n2 + b];
                    }
                const double m1 = median8(w);
                if (m1 > kThresh) P[i * n2 + j] = m1;
            }
        }
        __syncthreads();
    }
}

__global__ void mirror_kernel(double *__restrict__ P, int64_t n) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n * n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / n, j = e % n;
        if (i > j) P[e] = P[j * n + i];
    }
}

__global__ void to_f32_kernel(const double *__restrict__ in, float *__restrict__ out, int64_t count) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = __double2float_rn(in[e]);
}
__global__ void to_f64_kernel(const float *__restrict__ in, double *__restrict__ out, int64_t count) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = (double)in[e];
}

// flux through the forward difference delta: exp(-(delta/kappa)^2) / 1 * delta, float32, numpy order
__device__ __forceinline__ float pm_flux(float delta, float kappa) {
    const float q = __fdiv_rn(delta, kappa);
    return __fmul_rn(expf(-__fmul_rn(q, q)), delta);
}

// One Perona-Malik step (medpy option 1): forward differences that vanish on the last row/column,
// flux divergence by backward differences (the first row/column keeps its own flux), out = in +
// gamma * (NS + EW).  4 B read + 4 B written per pixel (neighbours come from L1/L2).
__global__ void __launch_bounds__(256) diffuse_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t n1,
                                                      int64_t n2, float kappa, float gamma) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n2) return;
    for (int64_t i = blockIdx.y; i < n1; i += gridDim.y) {
        const float c = in[i * n2 + j];
        const float fs = i + 1 < n1 ? pm_flux(__fsub_rn(in[(i + 1) * n2 + j], c), kappa) : 0.0f;
        const float fe = j + 1 < n2 ? pm_flux(__fsub_rn(in[i * n2 + j + 1], c), kappa) : 0.0f;
        const float ns = i > 0 ? __fsub_rn(fs, pm_flux(__fsub_rn(c, in[(i - 1) * n2 + j]), kappa)) : fs;
        const float ew = j > 0 ? __fsub_rn(fe, pm_flux(__fsub_rn(c, in[i * n2 + j - 1]), kappa)) : fe;
        out[i * n2 + j] = __fadd_rn(c, __fmul_rn(gamma, __fadd_rn(ns, ew)));
    }
}

// One axis of scipy.ndimage.gaussian_filter (order 0, mode 'reflect'): symmetric correlation in
// scipy's summation order (NI_Correlate1D: centre tap first, then the pairs from the outermost in),
// separate multiply and add.
__global__ void __launch_bounds__(256) gauss1d_kernel(const double *__restrict__ in, double *__restrict__ out, int64_t n1,
                                                      int64_t n2, int axis, const double *__restrict__ w, int lw) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n2) return;
    const int64_t len = axis == 0 ? n1 : n2, stride = axis == 0 ? n2 : 1;
    for (int64_t i = blockIdx.y; i < n1; i += gridDim.y) {
        const int64_t pos = axis == 0 ? i : j;
        const double *line = in + (axis == 0 ? j : i * n2);
        auto at = [&](int64_t k) {  // 'reflect': d c b a | a b c d | d c b a
            const int64_t per = 2 * len;
            k %= per;
            if (k < 0) k += per;
            if (k >= len) k = per - 1 - k;
            return line[k * stride];
        };
        double tmp = __dmul_rn(at(pos), w[lw]);
        for (int jj = -lw; jj < 0; ++jj)
            tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(pos + jj), at(pos - jj)), w[jj + lw]));
        out[i * n2 + j] = tmp;
    }
}

// node order: upper triangle row by row (kind 1) or the whole block (kind 0)
__global__ void flatten_kernel(const double *__restrict__ plane, int64_t n1, int64_t n2, int kind, int d, int c,
                               double *__restrict__ data) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t i = blockIdx.y; i < n1; i += gridDim.y) {
        if (kind == 1) {
            if (j < i || j >= n2) continue;
            const int64_t row_start = i * n2 - i * (i - 1) / 2;  // nodes before row i
            data[(row_start + (j - i)) * d + c] = plane[i * n2 + j];
        } else {
            if (j >= n2) continue;
            data[(i * n2 + j) * d + c] = plane[i * n2 + j];
        }
    }
}

__global__ void interleave_kernel(const double *__restrict__ plane, int64_t count, int d, int c, double *__restrict__ img) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
        img[e * d + c] = plane[e];
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { cudaFree(p); }
    template <typename T>
    T *as() { return static_cast<T *>(p); }
};

inline int grid_for(int64_t count, int cap = 148 * 16) {
    const int64_t b = (count + 255) / 256;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace

}  // namespace phmrf

using namespace phmrf;

extern "C" int phmrf_prep_normalise(int device, double *x, int64_t n, int d, double *x_min, double *x_max,
                                    double *colminmax_out, int log1p) {
    if (!x || n <= 0 || d <= 0 || !x_min || !x_max) {
        set_error("phmrf_prep_normalise: bad argument");
        return PHMRF_E_INVALID;
    }
    PHMRF_CUDA(cudaSetDevice(device));
    DevBuf dx, dmn, dmx, dmm;
    PHMRF_CUDA(cudaMalloc(&dx.p, sizeof(double) * n * d));
    PHMRF_CUDA(cudaMalloc(&dmn.p, sizeof(unsigned long long) * d));
    PHMRF_CUDA(cudaMalloc(&dmx.p, sizeof(unsigned long long) * d));
    PHMRF_CUDA(cudaMalloc(&dmm.p, sizeof(double) * 2 * d));
    PHMRF_CUDA(cudaMemcpy(dx.p, x, sizeof(double) * n * d, cudaMemcpyHostToDevice));
    PHMRF_CUDA(cudaMemset(dmn.p, 0xff, sizeof(unsigned long long) * d));
    PHMRF_CUDA(cudaMemset(dmx.p, 0, sizeof(unsigned long long) * d));
    dim3 g1((unsigned)grid_for(n, 148 * 4), (unsigned)d);
    colminmax_kernel<<<g1, 256>>>(dx.as<double>(), n, d, dmn.as<unsigned long long>(), dmx.as<unsigned long long>());
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    std::vector<unsigned long long> hmn(d), hmx(d);
    PHMRF_CUDA(cudaMemcpy(hmn.data(), dmn.p, sizeof(unsigned long long) * d, cudaMemcpyDeviceToHost));
    PHMRF_CUDA(cudaMemcpy(hmx.data(), dmx.p, sizeof(unsigned long long) * d, cudaMemcpyDeviceToHost));
    std::vector<double> mm(2 * (size_t)d), lo(d), hi(d);
    for (int c = 0; c < d; ++c) {
        memcpy(&mm[2 * c], &hmn[c], 8);
        memcpy(&mm[2 * c + 1], &hmx[c], 8);
        lo[c] = mm[2 * c];
        hi[c] = mm[2 * c + 1];
    }
    auto median = [](std::vector<double> v) {  // numpy.median: mean of the two middle values
        std::sort(v.begin(), v.end());
        const size_t m = v.size();
        return (m & 1) ? v[m / 2] : (v[m / 2 - 1] + v[m / 2]) / 2.0;
    };
    if (*x_min < 0) *x_min = median(lo);
    if (*x_max < 0) *x_max = median(hi);
    if (colminmax_out) memcpy(colminmax_out, mm.data(), sizeof(double) * 2 * d);
    PHMRF_CUDA(cudaMemcpy(dmm.p, mm.data(), sizeof(double) * 2 * d, cudaMemcpyHostToDevice));
    rescale_kernel<<<grid_for(n * d), 256>>>(dx.as<double>(), n, d, dmm.as<double>(), *x_min, *x_max, log1p);
    count_launch();
    PHMRF_CUDA(cudaGetLastError());
    PHMRF_CUDA(cudaMemcpy(x, dx.p, sizeof(double) * n * d, cudaMemcpyDeviceToHost));
    return PHMRF_OK;
}

extern "C" int phmrf_prep_region_image(int device, const double *value, const int64_t *pos, int64_t n, int d, int kind,
                                       int64_t start1, int64_t start2, int64_t n1, int64_t n2, int filter_mode,
                                       int niter, double kappa, double gamma, double sigma, double *data_out,
                                       double *image_out) {
    if (!value || !pos || n <= 0 || d <= 0 || n1 <= 0 || n2 <= 0 || !data_out || (kind != 0 && kind != 1) ||
        (kind == 1 && (n1 != n2 || start1 != start2))) {
        set_error("phmrf_prep_region_image: bad argument");
        return PHMRF_E_INVALID;
    }
    PHMRF_CUDA(cudaSetDevice(device));
    const int64_t npix = n1 * n2;
    const int64_t n_nodes = kind == 1 ? n1 * (n1 + 1) / 2 : npix;
    DevBuf dval, dpos, dplane, df0, df1, ddata, dimg, dbad, dgw, dtmp;
    int lw = 0;
    if (filter_mode == 2 && sigma > 0) {
        // scipy.ndimage._filters._gaussian_kernel1d: radius int(4 sigma + 0.5), exp(-x^2 / (2 sigma^2)) normalised
        lw = (int)(4.0 * sigma + 0.5);
        std::vector<double> gw(2 * (size_t)lw + 1);
        const double s2 = sigma * sigma;
        double tot = 0.0;
        for (int x = -lw; x <= lw; ++x) {
            gw[x + lw] = std::exp(-0.5 / s2 * ((double)x * (double)x));
            tot += gw[x + lw];
        }
        for (auto &v : gw) v /= tot;
        PHMRF_CUDA(cudaMalloc(&dgw.p, sizeof(double) * gw.size()));
        PHMRF_CUDA(cudaMemcpy(dgw.p, gw.data(), sizeof(double) * gw.size(), cudaMemcpyHostToDevice));
        PHMRF_CUDA(cudaMalloc(&dtmp.p, sizeof(double) * npix));
    }
    // species are independent: the (latency-bound, one CTA per plane) hole fill runs for a batch of
    // planes at once, as many as fit a 48 GB budget
    int batch = (int)std::max<int64_t>(1, std::min<int64_t>(d, (int64_t)48e9 / (int64_t)(sizeof(double) * npix)));
    PHMRF_CUDA(cudaMalloc(&dval.p, sizeof(double) * n * d));
    PHMRF_CUDA(cudaMalloc(&dpos.p, sizeof(int64_t) * 2 * n));
    PHMRF_CUDA(cudaMalloc(&dplane.p, sizeof(double) * npix * batch));
    PHMRF_CUDA(cudaMalloc(&ddata.p, sizeof(double) * n_nodes * d));
    PHMRF_CUDA(cudaMalloc(&dbad.p, sizeof(int)));
    if (filter_mode == 0 && niter > 0) {
        PHMRF_CUDA(cudaMalloc(&df0.p, sizeof(float) * npix));
        PHMRF_CUDA(cudaMalloc(&df1.p, sizeof(float) * npix));
    }
    if (image_out) PHMRF_CUDA(cudaMalloc(&dimg.p, sizeof(double) * npix * d));
    PHMRF_CUDA(cudaMemcpy(dval.p, value, sizeof(double) * n * d, cudaMemcpyHostToDevice));
    PHMRF_CUDA(cudaMemcpy(dpos.p, pos, sizeof(int64_t) * 2 * n, cudaMemcpyHostToDevice));
    PHMRF_CUDA(cudaMemset(dbad.p, 0, sizeof(int)));
    const dim3 g2((unsigned)((n2 + 255) / 256), (unsigned)(n1 < 4096 ? n1 : 4096));
    for (int c0 = 0; c0 < d; c0 += batch) {
        const int nb = std::min(batch, d - c0);
        PHMRF_CUDA(cudaMemsetAsync(dplane.p, 0, sizeof(double) * npix * nb));
        for (int q = 0; q < nb; ++q)
            scatter_kernel<<<grid_for(n), 256>>>(dval.as<double>(), dpos.as<int64_t>(), n, d, c0 + q, start1, start2, n1,
                                                 n2, kind == 1, dplane.as<double>() + (int64_t)q * npix, dbad.as<int>());
        holefill_kernel<<<nb, 1024>>>(dplane.as<double>(), npix, n1, n2, kind == 1);
        count_launch(nb + 1);
        for (int q = 0; q < nb; ++q) {
            const int c = c0 + q;
            double *plane = dplane.as<double>() + (int64_t)q * npix;
            if (kind == 1) {
                mirror_kernel<<<grid_for(npix), 256>>>(plane, n1);
                count_launch();
            }
            if (filter_mode == 0 && niter > 0) {
                float *a = df0.as<float>(), *b = df1.as<float>();
                to_f32_kernel<<<grid_for(npix), 256>>>(plane, a, npix);
                for (int it = 0; it < niter; ++it) {
                    diffuse_kernel<<<g2, 256>>>(a, b, n1, n2, (float)kappa, (float)gamma);
                    std::swap(a, b);
                }
                to_f64_kernel<<<grid_for(npix), 256>>>(a, plane, npix);
                count_launch(niter + 2);
            }
            if (filter_mode == 2 && sigma > 0) {  // axis 0 then axis 1, like scipy
                gauss1d_kernel<<<g2, 256>>>(plane, dtmp.as<double>(), n1, n2, 0, dgw.as<double>(), lw);
                gauss1d_kernel<<<g2, 256>>>(dtmp.as<double>(), plane, n1, n2, 1, dgw.as<double>(), lw);
                count_launch(2);
            }
            flatten_kernel<<<g2, 256>>>(plane, n1, n2, kind, d, c, ddata.as<double>());
            count_launch();
            if (image_out) {
                interleave_kernel<<<grid_for(npix), 256>>>(plane, npix, d, c, dimg.as<double>());
                count_launch();
            }
        }
        PHMRF_CUDA(cudaGetLastError());
    }
    int bad = 0;
    PHMRF_CUDA(cudaMemcpy(&bad, dbad.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) {
        set_error("phmrf_prep_region_image: a bin pair lies outside the region window");
        return PHMRF_E_INVALID;
    }
    PHMRF_CUDA(cudaMemcpy(data_out, ddata.p, sizeof(double) * n_nodes * d, cudaMemcpyDeviceToHost));
    if (image_out) PHMRF_CUDA(cudaMemcpy(image_out, dimg.p, sizeof(double) * npix * d, cudaMemcpyDeviceToHost));
    return PHMRF_OK;
}
