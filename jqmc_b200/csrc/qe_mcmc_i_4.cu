// Metropolis kernels for orbital padding NMO = 4 (spherical and Cartesian): see qe_mcmc_kernel.cuh
#include "qe_mcmc_kernel.cuh"

template int mcmc_launch_one<4, false>(qe_engine*, McmcArgs&, int, double, cudaStream_t);
#ifndef QE_DEV_MINIMAL
template int mcmc_launch_one<4, true>(qe_engine*, McmcArgs&, int, double, cudaStream_t);
#endif
