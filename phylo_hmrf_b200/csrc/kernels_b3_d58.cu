// Phase B, bulk-copy pipeline: instantiations for 5..8 features (see estep_bulk.cuh).
#define PHMRF_B3_ENTRY launch_estep_bulk_d58
#define PHMRF_B3_D0 5
#include "estep_bulk.cuh"
