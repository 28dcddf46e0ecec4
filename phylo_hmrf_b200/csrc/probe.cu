// Pipe probes: measure the roofline denominators the E-step kernels are judged against
// (FP64 FMA pipe; BASELINE.md section 2 leaves it "to be produced") plus a few mixes used
// while tuning.  A measurement tool, NOT part of the product: it builds into its own library,
// lib/libphmrf_probe.so (include/phmrf_probe.h), which only bench.py and tools/ load.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/phmrf_probe.h"

namespace phmrf {

enum { PHMRF_OK = 0, PHMRF_E_INVALID = -1, PHMRF_E_CUDA = -2 };
static thread_local std::string g_probe_error;
static void set_error(const std::string &msg) { g_probe_error = msg; }
static void count_launch(int = 1) {}
static int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    g_probe_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
                    std::to_string(line) + ")";
    cudaGetLastError();
    return PHMRF_E_CUDA;
}
#define PHMRF_CUDA(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) return ::phmrf::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

namespace {

__constant__ double c_probe[512];

template <int MODE>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
                a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
            }
        } else if (MODE == 1) {  // every second FMA takes an indexed constant-bank operand
            const int base = (it & 31) * 8;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double w = c_probe[base + u];
                a0 = fma(a0, w, c); a1 = fma(a1, w, c); a2 = fma(a2, w, c); a3 = fma(a3, w, c);
                a4 = fma(a4, w, c); a5 = fma(a5, w, c); a6 = fma(a6, w, c); a7 = fma(a7, w, c);
            }
        } else if (MODE == 3) {  // three distinct register operands per FMA (register-file bandwidth)
            double b0 = a0 * m, b1 = a1 * m, b2 = a2 * m, b3 = a3 * m;
            double c0 = a4 * m, c1 = a5 * m, c2 = a6 * m, c3 = a7 * m;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = fma(b0, c0, a0); a1 = fma(b1, c1, a1); a2 = fma(b2, c2, a2); a3 = fma(b3, c3, a3);
                a4 = fma(b0, c1, a4); a5 = fma(b1, c2, a5); a6 = fma(b2, c3, a6); a7 = fma(b3, c0, a7);
            }
        } else if (MODE == 4) {  // one multiplicand shared by four consecutive FMAs (the emission kernel's pattern)
            double b0 = a0 * m, b1 = a1 * m, c0 = a4 * m, c1 = a5 * m, c2 = a6 * m, c3 = a7 * m;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = fma(c0, b0, a0); a1 = fma(c1, b0, a1); a2 = fma(c2, b0, a2); a3 = fma(c3, b0, a3);
                a4 = fma(c0, b1, a4); a5 = fma(c1, b1, a5); a6 = fma(c2, b1, a6); a7 = fma(c3, b1, a7);
            }
        } else if (MODE == 5) {  // boustrophedon order: every FMA shares one multiplicand with its predecessor
            double b0 = a0 * m, b1 = a1 * m, c0 = a4 * m, c1 = a5 * m, c2 = a6 * m, c3 = a7 * m;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = fma(c0, b0, a0); a1 = fma(c1, b0, a1); a2 = fma(c2, b0, a2); a3 = fma(c3, b0, a3);
                a7 = fma(c3, b1, a7); a6 = fma(c2, b1, a6); a5 = fma(c1, b1, a5); a4 = fma(c0, b1, a4);
            }
        } else if (MODE == 2) {  // exp throughput: 8 independent exps per iteration
            a0 = exp(a0 * 1e-3 - 1.0); a1 = exp(a1 * 1e-3 - 1.0); a2 = exp(a2 * 1e-3 - 1.0); a3 = exp(a3 * 1e-3 - 1.0);
            a4 = exp(a4 * 1e-3 - 1.0); a5 = exp(a5 * 1e-3 - 1.0); a6 = exp(a6 * 1e-3 - 1.0); a7 = exp(a7 * 1e-3 - 1.0);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA m8n8k4: 256 FMA per warp instruction.
template <bool MIX>
__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-3, b = 1.0 - 1e-6;
    double c0 = 0, c1 = 0, d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    double x0 = seed, x1 = seed + 1, x2 = seed + 2, x3 = seed + 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(e0), "+d"(e1) : "d"(a), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(f0), "+d"(f1) : "d"(a), "d"(b));
            if (MIX) {
                x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
                x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
            }
        }
    }
    double s = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1 + x0 + x1 + x2 + x3;
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Few-warp issue probe: one CTA per SM, warps 0-3 (one per sub-partition) issue DMMA with NACC
// independent accumulators, the next `prod` warps issue DFMA in ILP independent chains.
// Every warp records its own cycle count; out[0] = mean cycles per DFMA instruction of a
// producer warp, out[1] = mean cycles per DMMA of a consumer warp (block 0).
template <int ILP, int NACC>
__global__ void __launch_bounds__(512, 1) mix_kernel(double *out, long long *cyc, int it_dfma, int it_dmma, int cons) {
    const int warp = threadIdx.x >> 5;
    const double m = 0.999999, c = 1e-9;
    double s = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < cons) {
        double acc[NACC][2];
#pragma unroll
        for (int u = 0; u < NACC; ++u) acc[u][0] = acc[u][1] = 0.0;
        const double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - 1e-6;
        for (int it = 0; it < it_dmma; ++it) {
#pragma unroll
            for (int u = 0; u < NACC; ++u)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[u][0]), "+d"(acc[u][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int u = 0; u < NACC; ++u) s += acc[u][0] + acc[u][1];
    } else {
        double x[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) x[u] = 1.0 + threadIdx.x + u;
        for (int it = 0; it < it_dfma; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int u = 0; u < ILP; ++u) x[u] = fma(x[u], m, c);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) s += x[u];
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
    if (s == 123.456) out[threadIdx.x] = s;
}

__global__ void copy_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}

template <typename Launch>
int time_best(Launch launch, int reps, float *best_ms) {
    cudaEvent_t e0, e1;
    PHMRF_CUDA(cudaEventCreate(&e0));
    PHMRF_CUDA(cudaEventCreate(&e1));
    *best_ms = 1e30f;
    for (int r = 0; r < reps + 2; ++r) {
        PHMRF_CUDA(cudaEventRecord(e0));
        launch();
        PHMRF_CUDA(cudaEventRecord(e1));
        PHMRF_CUDA(cudaEventSynchronize(e1));
        float ms;
        PHMRF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2 && ms < *best_ms) *best_ms = ms;
    }
    PHMRF_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return PHMRF_OK;
}

}  // namespace

// which: 0 DFMA TFLOP/s | 1 DFMA with indexed constant operand TFLOP/s | 2 exp Gexp/s
//        3 DMMA m8n8k4 TFLOP/s | 4 DMMA+DFMA interleaved, total TFLOP/s | 5 HBM copy GB/s
//        6 DFMA with three distinct register operands TFLOP/s
static int run_probe(int which, double *out) {
    int dev, sms;
    PHMRF_CUDA(cudaGetDevice(&dev));
    PHMRF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *buf = nullptr;
    const int grid = sms * 8, block = 256;
    float ms = 0;
    int rc = PHMRF_OK;
    if ((which >= 0 && which <= 4) || which == 6 || which == 15 || which == 16) {
        PHMRF_CUDA(cudaMalloc(&buf, sizeof(double) * grid * block));
        const int iters = which == 2 ? 2000 : 4000;
        if (which == 1) {
            double h[512];
            for (int i = 0; i < 512; ++i) h[i] = 1.0 - 1e-7 * (i + 1);
            PHMRF_CUDA(cudaMemcpyToSymbol(c_probe, h, sizeof(h)));
        }
        switch (which) {
            case 0: rc = time_best([&] { dfma_kernel<0><<<grid, block>>>(buf, iters, 1.0); }, 5, &ms); break;
            case 1: rc = time_best([&] { dfma_kernel<1><<<grid, block>>>(buf, iters, 1.0); }, 5, &ms); break;
            case 2: rc = time_best([&] { dfma_kernel<2><<<grid, block>>>(buf, iters, 1.0); }, 5, &ms); break;
            case 3: rc = time_best([&] { dmma_kernel<false><<<grid, block>>>(buf, iters, 1.0); }, 5, &ms); break;
            case 4: rc = time_best([&] { dmma_kernel<true><<<grid, block>>>(buf, iters, 1.0); }, 5, &ms); break;
            case 6: rc = time_best([&] { dfma_kernel<3><<<grid, block>>>(buf, iters, 1e-3); }, 5, &ms); break;
            case 15: rc = time_best([&] { dfma_kernel<4><<<grid, block>>>(buf, iters, 1e-3); }, 5, &ms); break;
            case 16: rc = time_best([&] { dfma_kernel<5><<<grid, block>>>(buf, iters, 1e-3); }, 5, &ms); break;
        }
        count_launch(7);
        cudaFree(buf);
        if (rc != PHMRF_OK) return rc;
        const double threads = (double)grid * block;
        if (which == 0 || which == 1 || which == 6 || which == 15 || which == 16) *out = threads * iters * 64.0 * 2.0 / (ms * 1e-3) / 1e12;
        if (which == 2) *out = threads * iters * 8.0 / (ms * 1e-3) / 1e9;
        if (which == 3) *out = (threads / 32.0) * iters * 16.0 * 256.0 * 2.0 / (ms * 1e-3) / 1e12;
        if (which == 4) *out = ((threads / 32.0) * iters * 16.0 * 256.0 * 2.0 + threads * iters * 32.0 * 2.0) / (ms * 1e-3) / 1e12;
        return PHMRF_OK;
    }
    if (which == 5) {
        const int64_t n = (int64_t)1 << 27;  // 2 GiB in, 2 GiB out
        double2 *a = nullptr, *b = nullptr;
        PHMRF_CUDA(cudaMalloc(&a, sizeof(double2) * n));
        PHMRF_CUDA(cudaMalloc(&b, sizeof(double2) * n));
        PHMRF_CUDA(cudaMemset(a, 0, sizeof(double2) * n));
        rc = time_best([&] { copy_kernel<<<sms * 16, 256>>>(a, b, n); }, 5, &ms);
        count_launch(7);
        cudaFree(a);
        cudaFree(b);
        if (rc != PHMRF_OK) return rc;
        *out = 2.0 * sizeof(double2) * (double)n / (ms * 1e-3) / 1e9;
        return PHMRF_OK;
    }
    if (which >= 7 && which <= 14) {
        // 7: 8 DFMA warps (ILP 8) alone   8: 4 DFMA warps alone   9: 4 DMMA warps alone (28 acc)
        // 10/11: 4 DMMA + 8 DFMA(ILP 8): cycles per DFMA / per DMMA
        // 12/13: 4 DMMA + 8 DFMA(ILP 4): cycles per DFMA / per DMMA     14: 12 DFMA warps alone (ILP 8)
        long long *cyc = nullptr;
        PHMRF_CUDA(cudaMalloc(&buf, sizeof(double) * 1024));
        PHMRF_CUDA(cudaMalloc(&cyc, sizeof(long long) * 16));
        PHMRF_CUDA(cudaMemset(cyc, 0, sizeof(long long) * 16));
        const int cons = (which == 7 || which == 8 || which == 14) ? 0 : 4;
        const int prod = which == 9 ? 0 : (which == 8 ? 4 : (which == 14 ? 12 : 8));
        const int it_dfma = 4000, it_dmma = 1200;
        const int threads = 32 * (cons + prod);
        for (int r = 0; r < 3; ++r) {
            if (which == 12 || which == 13)
                mix_kernel<4, 28><<<sms, threads>>>(buf, cyc, it_dfma * 2, it_dmma, cons);
            else
                mix_kernel<8, 28><<<sms, threads>>>(buf, cyc, it_dfma, it_dmma, cons);
        }
        PHMRF_CUDA(cudaDeviceSynchronize());
        count_launch(3);
        long long h[16];
        PHMRF_CUDA(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(buf);
        cudaFree(cyc);
        double cd = 0, cm = 0;
        for (int w = 0; w < cons; ++w) cm += (double)h[w] / cons;
        for (int w = cons; w < cons + prod; ++w) cd += (double)h[w] / prod;
        const bool want_dmma = which == 9 || which == 11 || which == 13;
        *out = want_dmma ? cm / (it_dmma * 28.0) : cd / (it_dfma * 64.0);
        return PHMRF_OK;
    }
    set_error("unknown probe");
    return PHMRF_E_INVALID;
}

}  // namespace phmrf

extern "C" {

const char *phmrf_probe_last_error(void) { return phmrf::g_probe_error.c_str(); }

int phmrf_probe(int device, int which, double *out) {
    if (!out) return phmrf::PHMRF_E_INVALID;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        phmrf::set_error("phmrf_probe: no such CUDA device");
        return phmrf::PHMRF_E_CUDA;
    }
    if (cudaSetDevice(device) != cudaSuccess) return phmrf::cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
    return phmrf::run_probe(which, out);
}

int phmrf_probe_fp64_tflops(int device, double *tflops_out) { return phmrf_probe(device, 0, tflops_out); }

}  // extern "C"
