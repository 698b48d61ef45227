// Translation unit of the persistent recurrence kernels (rec_fwd.cuh, rec_bwd.cuh).
#include "rec_fwd.cuh"
#include "rec_bwd.cuh"
