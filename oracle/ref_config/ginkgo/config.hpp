/* Build configuration used ONLY to compile the unmodified Ginkgo 1.5.0 sources
 * under /root/reference into oracle/_ref/ (test infrastructure, see
 * oracle/Makefile.ref).  It plays the role of the header CMake would generate
 * from include/ginkgo/config.hpp.in: CPU-only (reference + omp executors),
 * no MPI, no hwloc, no PAPI, no mixed precision. */
#ifndef GKO_INCLUDE_CONFIG_H
#define GKO_INCLUDE_CONFIG_H

#define GKO_VERSION_MAJOR 1
#define GKO_VERSION_MINOR 5
#define GKO_VERSION_PATCH 0
#define GKO_VERSION_TAG "master"
#define GKO_VERSION_STR 1, 5, 0

#define GKO_VERBOSE_LEVEL 1
#define GKO_HAVE_CXXABI_H
/* #undef GINKGO_JACOBI_FULL_OPTIMIZATIONS */
/* #undef GINKGO_BENCHMARK_ENABLE_TUNING */
/* #undef GINKGO_MIXED_PRECISION */
#define GINKGO_HIP_PLATFORM_HCC 0
#define GINKGO_HIP_PLATFORM_NVCC 0
#define GKO_HAVE_PAPI_SDE 0
#define GINKGO_BUILD_MPI 0
#define GINKGO_HAVE_GPU_AWARE_MPI 0
#define GKO_HAVE_HWLOC 0
/* #undef GINKGO_FORCE_SPMV_BLOCKING_COMM */

#endif  // GKO_INCLUDE_CONFIG_H
