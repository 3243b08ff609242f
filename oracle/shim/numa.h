/* oracle/shim/numa.h -- TEST INFRASTRUCTURE ONLY.
 *
 * libnuma is not installed in this image.  The reference's TPSM thread pool
 * (STMMQR/include/tpsm/tpsm_base.h:30 includes <numa.h>; the calls are in
 * STMMQR/src/base/tpsm_base.c:119-287) only needs a handful of entry points.
 * This header provides single-node stand-ins on top of calloc/realloc/free so
 * that the UNMODIFIED reference sources compile into oracle/_ref/.
 * It is never seen by the product (stmqr_b200/).
 */
#ifndef ORACLE_SHIM_NUMA_H
#define ORACLE_SHIM_NUMA_H
#include <stdlib.h>
#include <string.h>

static inline int numa_available (void) { return 0 ; }
static inline int numa_max_node (void) { return 0 ; }
static inline int numa_node_of_cpu (int cpu) { (void) cpu ; return 0 ; }
static inline int numa_distance (int a, int b) { return (a == b) ? 10 : 20 ; }
static inline void *numa_alloc_onnode (size_t bytes, int node)
{
    (void) node ;
    return calloc (1, bytes ? bytes : 1) ;
}
static inline void *numa_alloc_local (size_t bytes)
{
    return calloc (1, bytes ? bytes : 1) ;
}
static inline void *numa_alloc_interleaved (size_t bytes)
{
    return calloc (1, bytes ? bytes : 1) ;
}
static inline void *numa_realloc (void *old, size_t old_size, size_t new_size)
{
    (void) old_size ;
    return realloc (old, new_size ? new_size : 1) ;
}
static inline void numa_free (void *p, size_t bytes)
{
    (void) bytes ;
    free (p) ;
}
#endif
