// One instantiation group of the fused per-walker kernel (qe_walker_kernel.cuh): GFMC_t,
// orbital padding NMO = 4, spherical basis.  One group per translation unit: each goes through a single-threaded
// (deterministic) ptxas, and build() compiles the files in parallel.
#include "qe_walker_kernel.cuh"

template int launch_walker_one<true, 4, false>(qe_engine*, WalkerArgs&, cudaStream_t, int);
