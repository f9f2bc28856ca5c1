/* Hand-written stand-in for the autoconf-generated fplll_config.h (autotools are absent here).
 * Test infrastructure only: used to compile the UNMODIFIED reference from /root/reference into oracle/_ref/. */
#ifndef FPLLL_CONFIG__H
#define FPLLL_CONFIG__H
#define FPLLL_MAJOR_VERSION 5
#define FPLLL_MINOR_VERSION 5
#define FPLLL_MICRO_VERSION 0
#define FPLLL_VERSION 5.5.0
#define FPLLL_VERSION_INFO
#define FPLLL_MAX_ENUM_DIM 256
#define FPLLL_WITH_RECURSIVE_ENUM 1
#define FPLLL_MAX_PARALLEL_ENUM_DIM 120
#define HAVE_LIBGMP 1
#endif
