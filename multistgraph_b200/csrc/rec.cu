// Translation unit of the persistent recurrence kernels (rec_fwd.cuh).
#include "rec_fwd.cuh"
