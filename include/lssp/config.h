/* config.h -- build configuration of the B200-native LSSP facade.  Every third-party
 * adapter of the reference (include/config.h.in:18-34) is off: they are CPU libraries outside
 * the accelerated path.  USE_GPU marks this build.  USE_BLAS / USE_LAPACK only switch on what the
 * reference guards with them -- LSSP_PC_BILUK -- which this build implements itself. */
#ifndef LSSP_CONFIG_H
#define LSSP_CONFIG_H
#define LSSP_VER_MAJOR 1
#define LSSP_VER_MINOR 0
#define USE_GPU     1
#define USE_BLAS    1   /* block ILU(k) (pc-biluk.h) is built in: the dense block kernels are the library's own, */
#define USE_LAPACK  1   /* no BLAS / LAPACK is linked (reference include/type-defs.h:70-74, src/pc-biluk.cxx:3-4) */
#define USE_LASPACK 0
#define USE_SSPARSE 0
#define USE_MUMPS   0
#define USE_PETSC   0
#define USE_ITSOL   0
#define USE_LIS     0
#define USE_QR_MUMPS 0
#define USE_SUPERLU 0
#define USE_PARDISO 0
#define USE_FASP    0
#define USE_HSL_MI20 0
#define USE_SXAMG   1   /* the B200 build ships its own SX-AMG-style AMG (sxamg.h stand-in) */
#endif
