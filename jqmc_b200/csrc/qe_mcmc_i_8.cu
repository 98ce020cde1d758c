// Metropolis kernels for orbital padding NMO = 8 (spherical and Cartesian): see qe_mcmc_kernel.cuh
#include "qe_mcmc_kernel.cuh"

#ifndef QE_DEV_MINIMAL  // (development builds instantiate the benchmark shape only)
template int mcmc_launch_one<8, false>(qe_engine*, McmcArgs&, int, double, cudaStream_t);
template int mcmc_launch_one<8, true>(qe_engine*, McmcArgs&, int, double, cudaStream_t);
#endif
