#pragma once
#define CXX_COMP_VENDOR "gnu"
#define GXX_VERSION "13"
#define GRID_TRACING_NONE 1
#define Config_Nc 3
#define Sp2n_config 0
#define AVX2 1
#define GRID_DEFAULT_PRECISION_DOUBLE 1
#define GRID_ALLOC_ALIGN (2*1024*1024)
#define ALLOCATION_CACHE 1
#define GRID_MPI3_SHM_NONE 1
#define GRID_SHM_PATH "/var/lib/hugetlbfs/global/pagesize-2MB/"
#define GRID_COMMS_NONE 1
#define RNG_SITMO 1
#define TIMERS_ON 1
#define HAVE_ZLIB 1
#define GRID_OMP 1
#define VERSION "0.7.0"
#define PACKAGE_STRING "Grid 0.7.0"
#define HAVE_EXECINFO_H 1
#define HAVE_DECL_BE64TOH 1
#define HAVE_DECL_NTOHLL 0
#define HAVE_ENDIAN_H 1
#define HAVE_MM_MALLOC_H 1
#define HAVE_MALLOC_H 1
