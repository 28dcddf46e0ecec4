// Thin C wrapper around the GCO v3.0 library (compiled from the reference's vendored
// gco_source/ at build time; see build.py).  Mirrors the call sequence of pygco's general
// graph path behind phylo_hmrf.py:496-498.  See include/phmrf_gco.h.
#include <cstdint>
#include <string>

#include "GCoptimization.h"
#include "../../include/phmrf_gco.h"

static thread_local std::string g_gco_error;

extern "C" const char *phmrf_gco_last_error(void) { return g_gco_error.c_str(); }

extern "C" int phmrf_gco_cut_general_graph(int64_t n_sites, int32_t n_labels, const int32_t *unary,
                                           const int64_t *edge_ids, const int32_t *edge_w, int64_t n_edges,
                                           const int32_t *smooth, const int32_t *init_labels, int32_t n_iter,
                                           int32_t algorithm, int32_t *labels_out, long long *energy_out,
                                           long long *energy_before_out) {
    if (n_sites <= 0 || n_labels <= 0 || !unary || !smooth || !labels_out || n_edges < 0 ||
        (n_edges > 0 && (!edge_ids || !edge_w))) {
        g_gco_error = "phmrf_gco_cut_general_graph: invalid arguments";
        return -1;
    }
    // GCO indexes the data cost as site*n_labels+label with a 32-bit int (GCoptimization.h:337)
    if (n_sites * (int64_t)n_labels >= ((int64_t)1 << 31)) {
        g_gco_error = "phmrf_gco_cut_general_graph: n_sites*n_labels overflows GCO's 32-bit indexing";
        return -1;
    }
    for (int64_t e = 0; e < n_edges; ++e) {
        const int64_t a = edge_ids[2 * e], b = edge_ids[2 * e + 1];
        if (a < 0 || b < 0 || a >= n_sites || b >= n_sites || a >= b) {
            g_gco_error = "phmrf_gco_cut_general_graph: edges must satisfy 0 <= id1 < id2 < n_sites";
            return -1;
        }
    }
    try {
        GCoptimizationGeneralGraph gc((GCoptimization::SiteID)n_sites, (GCoptimization::LabelID)n_labels);
        gc.setDataCost(const_cast<GCoptimization::EnergyTermType *>(unary));
        for (int64_t e = 0; e < n_edges; ++e)
            gc.setNeighbors((GCoptimization::SiteID)edge_ids[2 * e], (GCoptimization::SiteID)edge_ids[2 * e + 1],
                            (GCoptimization::EnergyTermType)edge_w[e]);
        gc.setSmoothCost(const_cast<GCoptimization::EnergyTermType *>(smooth));
        if (init_labels)
            for (int64_t i = 0; i < n_sites; ++i) {
                if (init_labels[i] < 0 || init_labels[i] >= n_labels) {
                    g_gco_error = "phmrf_gco_cut_general_graph: init label out of range";
                    return -1;
                }
                gc.setLabel((GCoptimization::SiteID)i, (GCoptimization::LabelID)init_labels[i]);
            }
        if (energy_before_out) *energy_before_out = (long long)gc.compute_energy();
        long long en;
        if (algorithm == PHMRF_GCO_EXPANSION)
            en = (long long)gc.expansion(n_iter);
        else
            en = (long long)gc.swap(n_iter);
        if (energy_out) *energy_out = en;
        gc.whatLabel(0, (GCoptimization::SiteID)n_sites, labels_out);
    } catch (GCException &ex) {
        g_gco_error = std::string("GCO: ") + (ex.message ? ex.message : "unknown error");
        return -1;
    } catch (std::exception &ex) {
        g_gco_error = std::string("GCO: ") + ex.what();
        return -1;
    } catch (...) {
        g_gco_error = "GCO: unknown exception";
        return -1;
    }
    return 0;
}
