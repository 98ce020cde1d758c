// Metropolis kernels for orbital padding NMO = 16 (spherical and Cartesian): see qe_mcmc_kernel.cuh
#include "qe_mcmc_kernel.cuh"

#ifndef QE_DEV_MINIMAL  // (development builds instantiate the benchmark shape only)
template int mcmc_launch_one<16, false>(qe_engine*, McmcArgs&, int, double, cudaStream_t);
template int mcmc_launch_one<16, true>(qe_engine*, McmcArgs&, int, double, cudaStream_t);
#endif
