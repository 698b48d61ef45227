// Translation unit of the persistent recurrence kernels (rec_fwd.cuh, rec_bwd.cuh) and of the fused DR reduction (dr_pass.cuh).
#include "rec_fwd.cuh"
#include "rec_bwd.cuh"
#include "dr_pass.cuh"
