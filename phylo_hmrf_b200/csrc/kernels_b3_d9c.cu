// Phase B, bulk-copy pipeline: instantiations for 9..12 features (see estep_bulk.cuh).
#define PHMRF_B3_ENTRY launch_estep_bulk_d9c
#define PHMRF_B3_D0 9
#include "estep_bulk.cuh"
