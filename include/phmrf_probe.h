/*
 * phmrf_probe.h -- pipe probes (libphmrf_probe.so): the measured denominators of the rooflines bench.py reports
 * against (BASELINE.md section 2 asks for a DFMA-chain micro-benchmark: MEASURED_PEAKS.json carries no FP64 figure).
 * A measurement tool, not part of the product library; nothing under phylo_hmrf_b200/ needs it except
 * engine.probe(), which bench.py and tools/ call.
 */
#ifndef PHMRF_PROBE_H_
#define PHMRF_PROBE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Peak FP64 throughput of the DFMA pipe (chains of dependent FMAs, 8 per thread), TFLOP/s. */
int phmrf_probe_fp64_tflops(int device, double *tflops_out);
/* which: 0 DFMA TFLOP/s | 1 DFMA with an indexed constant operand | 2 exp Gexp/s | 3 DMMA m8n8k4 TFLOP/s |
 * 4 DMMA+DFMA interleaved | 5 HBM copy GB/s | 6 DFMA with three distinct register operands | 7-14 few-warp issue
 * probes (cycles per instruction) | 15, 16 operand-reuse orderings.  Returns 0, or a negative code. */
int phmrf_probe(int device, int which, double *out);
const char *phmrf_probe_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
