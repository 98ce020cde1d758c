// One instantiation group of the fused per-walker kernel (qe_walker_kernel.cuh): GFMC_n / V elements / local energy,
// orbital padding NMO = 4, spherical basis.  One group per translation unit: each goes through a single-threaded
// (deterministic) ptxas, and build() compiles the files in parallel.
#define QE_EXP_ESTRIN 1  // qexp_s of this translation unit: see qe_device.cuh
#include "qe_walker_kernel.cuh"

template int launch_walker_one<false, 4, false>(qe_engine*, WalkerArgs&, cudaStream_t, int);
