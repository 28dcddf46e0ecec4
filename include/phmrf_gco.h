/*
 * phmrf_gco.h -- C ABI of the host graph-cut step (libphmrf_gco.so).
 *
 * Replaces the C shim inside pygco (yujiali/pygco cgco.cpp; third-party, not in the
 * reference tree) that phylo_hmrf.py:496-498 reaches through pygco.cut_general_graph: it
 * drives the GCO v3.0 library the reference vendors (gco_source/GCoptimization.h:559-597)
 * with the same call sequence -- GCoptimizationGeneralGraph(n_sites, n_labels),
 * setDataCost(int*), setNeighbors(s1, s2, w) per edge in edge order, setSmoothCost(int*),
 * setLabel per site, swap(n_iter) / expansion(n_iter), whatLabel.
 * GCO stays on the host (north star); it consumes the integer arrays produced on the GPU
 * by phmrf_quantise.  One GCO instance per call, no global state, GCException is caught
 * and reported through the return code / phmrf_gco_last_error().
 */
#ifndef PHMRF_GCO_H_
#define PHMRF_GCO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PHMRF_GCO_SWAP = 0, PHMRF_GCO_EXPANSION = 1 };

/* unary [n_sites*n_labels] (site-major, GCoptimization.h:337), edge_ids [n_edges*2] with
 * id1<id2, edge_w [n_edges], smooth [n_labels*n_labels], init_labels [n_sites] or NULL,
 * labels_out [n_sites].  energy_out (nullable) receives the final total energy and
 * energy_before_out (nullable) the energy of the initial labelling.
 * Returns 0, or -1 on invalid arguments / a GCO exception. */
int phmrf_gco_cut_general_graph(int64_t n_sites, int32_t n_labels, const int32_t *unary, const int64_t *edge_ids,
                                const int32_t *edge_w, int64_t n_edges, const int32_t *smooth,
                                const int32_t *init_labels, int32_t n_iter, int32_t algorithm,
                                int32_t *labels_out, long long *energy_out, long long *energy_before_out);

const char *phmrf_gco_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PHMRF_GCO_H_ */
