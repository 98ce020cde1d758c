// One instantiation group of the fused per-walker kernel (qe_walker_kernel.cuh): GFMC_n / V elements / local energy,
// orbital padding NMO = 4, Cartesian basis.  One group per translation unit: each goes through a single-threaded
// (deterministic) ptxas, and build() compiles the files in parallel.
#define QE_EXP_ESTRIN 1  // qexp_s of this translation unit: see qe_device.cuh
#include "qe_walker_kernel.cuh"

#ifndef QE_DEV_MINIMAL  // (development builds instantiate the benchmark shape only)
template int launch_walker_one<false, 4, true>(qe_engine*, WalkerArgs&, cudaStream_t, int);
#endif
