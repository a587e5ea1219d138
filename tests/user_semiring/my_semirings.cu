// The device half of tests/user_semiring/my_semirings.h: one nvcc translation unit of the "application" that instantiates the
// accumulation engine for its own semiring structs and exports one id function per semiring.
// mpi.h here is the single-rank stand-in of the oracle directory (MPI_Op is only named by the host-side members).
#include "combblas_b200/device_semiring.cuh"
#include "my_semirings.h"

CBGPU_DEFINE_SEMIRING(ktips_or_and_id, KTipsOrAnd, bool, bool, bool)
CBGPU_DEFINE_SEMIRING(max_times_f64_id, MaxTimesF64, double, double, double)
CBGPU_DEFINE_SEMIRING(min_plus_i32_id, MinPlusI32, int32_t, int32_t, int32_t)
