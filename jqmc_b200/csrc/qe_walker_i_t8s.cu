// One instantiation group of the fused per-walker kernel (qe_walker_kernel.cuh): GFMC_t,
// orbital padding NMO = 8, spherical basis.  One group per translation unit: each goes through a single-threaded
// (deterministic) ptxas, and build() compiles the files in parallel.
#include "qe_walker_kernel.cuh"

#ifndef QE_DEV_MINIMAL  // (development builds instantiate the benchmark shape only)
template int launch_walker_one<true, 8, false>(qe_engine*, WalkerArgs&, cudaStream_t, int);
#endif
