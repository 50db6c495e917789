/* TEST INFRASTRUCTURE (oracle build only): force-included when compiling the reference
 * sources into oracle/_ref/.  The reference's CPU_GEMM path scales an uninitialised
 * output buffer by BETA=0 (Executable/gemm.c:71-72), which yields NaN whenever malloc
 * hands back dirty pages; zero-filled allocations make that path well defined. */
#ifndef SRT_ORACLE_ZERO_MALLOC_H
#define SRT_ORACLE_ZERO_MALLOC_H
#include <stdlib.h>
#define malloc(n) calloc(1, (n))
#endif
