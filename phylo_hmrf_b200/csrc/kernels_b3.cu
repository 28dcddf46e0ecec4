// Phase B, bulk-copy pipeline: dispatch (the kernel lives in estep_bulk.cuh) and the instantiations
// for 1..4 features.
#define PHMRF_B3_ENTRY launch_estep_bulk_d14
#define PHMRF_B3_D0 1
#include "estep_bulk.cuh"

namespace phmrf {

int launch_estep_bulk_d58(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled);
int launch_estep_bulk_d9c(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled);

int launch_estep_bulk(const EstepArgs &a, int sm_count, cudaStream_t s, bool *handled) {
    *handled = false;
    if (!a.potts || a.W > estep::kFastSlots || a.pp_soa != nullptr || a.n == 0 || a.rowmax == nullptr) return PHMRF_OK;
    if (a.nbr_g == nullptr && (a.fwd_wg == nullptr || a.grid_kind < 0)) return PHMRF_OK;
    if (!(fabs(a.s_bound) < 100.0)) return PHMRF_OK;  // exp(S) * exp(600) must stay finite without range checks
    if (a.D >= 1 && a.D <= 4) return launch_estep_bulk_d14(a, sm_count, s, handled);
    if (a.D >= 5 && a.D <= 8) return launch_estep_bulk_d58(a, sm_count, s, handled);
    if (a.D >= 9 && a.D <= 12) return launch_estep_bulk_d9c(a, sm_count, s, handled);
    return PHMRF_OK;
}

}  // namespace phmrf
