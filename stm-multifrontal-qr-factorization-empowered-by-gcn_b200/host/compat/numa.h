/* host/compat/numa.h -- declarations only, searched AFTER the system include directories
 * (-idirafter): the reference's headers include <numa.h> (STMMQR/include/tpsm/tpsm_base.h:30) and a
 * machine that builds the reference has libnuma's own header, which then wins.  This image has no
 * libnuma; the drop-in itself never calls any of these (it does not use the TPSM pool), it only needs
 * the reference's headers to parse. */
#ifndef STMQR_B200_COMPAT_NUMA_H
#define STMQR_B200_COMPAT_NUMA_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
int   numa_available (void) ;
int   numa_max_node (void) ;
int   numa_node_of_cpu (int cpu) ;
void *numa_alloc_onnode (size_t size, int node) ;
void *numa_alloc_local (size_t size) ;
void *numa_alloc_interleaved (size_t size) ;
void *numa_realloc (void *old_addr, size_t old_size, size_t new_size) ;
void  numa_free (void *start, size_t size) ;
int   numa_distance (int node1, int node2) ;
#ifdef __cplusplus
}
#endif
#endif
