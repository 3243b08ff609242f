/* stmqr_b200.h -- C ABI of the B200-native numeric multifrontal-QR engine.
 *
 * This is the drop-in boundary for the reference's numeric phase
 *     qr_numeric *qr_factorize (sparse_csc **Ahandle, Long freeA, double tol,
 *                               Long ntol, qr_symbolic *QRsym, sparse_common *cc)
 * (reference: STMMQR/include/SparseQR.h:127-135, definition
 *  STMMQR/src/qr/SparseQR_factorize.c:222-749).  A host-side qr_factorize
 * written against the reference's own headers (stmqr_b200/host/qr_factorize_b200.c)
 * fills the plain views below from qr_symbolic / sparse_csc and calls these
 * entry points; no reference header and no torch type crosses this boundary.
 *
 * Everything here is extern "C", plain pointers and sizes.  All index arrays
 * are 64-bit signed (the reference's `Long` = Sparse_long), all values IEEE
 * double.  Every function returns STMQR_OK (0) or a negative status that the
 * host side maps onto the reference's cc->status codes
 * (STMMQR/include/SparseCore.h:48-54).
 *
 * There is NO CPU fallback: if no sm_100 device is present the calls fail
 * with STMQR_ERR_NO_DEVICE.
 */
#ifndef STMQR_B200_H
#define STMQR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STMQR_OK                 0
#define STMQR_ERR_NO_DEVICE    (-1)
#define STMQR_ERR_OUT_OF_MEMORY (-2)   /* -> SPARSE_OUT_OF_MEMORY (SparseCore.h:50) */
#define STMQR_ERR_TOO_LARGE    (-3)    /* -> SPARSE_TOO_LARGE     (SparseCore.h:51) */
#define STMQR_ERR_INVALID      (-4)    /* -> SPARSE_INVALID       (SparseCore.h:52) */
#define STMQR_ERR_CUDA         (-5)    /* any other CUDA runtime error -> SPARSE_INVALID */

typedef struct stmqr_handle_s *stmqr_handle ;

/* Read-only view of the reference's qr_symbolic (STMMQR/include/SparseQR_struct.h:26-137).
 * Field names are the reference's.  Only the fields the numeric phase reads are here
 * (the task-tree fields TaskChild/TaskFront/TaskStack/On_stack/Stack_maxstack belong to
 * the CPU scheduler, SparseQR_multithreads.c, which this engine replaces). */
typedef struct
{
    int64_t m, n, anz ;              /* S = A(P,Q) is m-by-n with anz entries              */
    int64_t nf ;                     /* number of fronts                                   */
    int64_t maxfn ;                  /* max # columns of any front                         */
    int64_t rjsize ;                 /* size of Rj, HStair, HTau                           */
    int64_t hisize ;                 /* size of Hii (= Hip[nf])                            */
    int64_t do_rank_detection ;      /* 0: Fm/Cm are exact, 1: they are upper bounds       */
    int64_t keepH ;                  /* always 1 in the reference (SparseQR_analyze.c:205) */
    const int64_t *Sp ;              /* [m+1]  row pointers of S                           */
    const int64_t *Sj ;              /* [anz]  column indices of S, ascending in each row  */
    const int64_t *Qfill ;           /* [n] or NULL: column k of S is column Qfill[k] of A */
    const int64_t *PLinv ;           /* [m]    row i of A is row PLinv[i] of S             */
    const int64_t *Sleft ;           /* [n+2]  rows Sleft[j]..Sleft[j+1]-1 have leftmost col j */
    const int64_t *Parent ;          /* [nf+1]                                             */
    const int64_t *Child ;           /* [nf+1]                                             */
    const int64_t *Childp ;          /* [nf+2]                                             */
    const int64_t *Super ;           /* [nf+1] pivot columns of front f: Super[f]..Super[f+1]-1 */
    const int64_t *Rp ;              /* [nf+1]                                             */
    const int64_t *Rj ;              /* [rjsize] columns of front f: Rj[Rp[f]..Rp[f+1]-1]  */
    const int64_t *Post ;            /* [nf+1] postordering of the fronts                  */
    const int64_t *Hip ;             /* [nf+1] Hii of front f starts at Hip[f]             */
    const int64_t *Fm ;              /* [nf+1] (bound on) # rows of each front             */
    const int64_t *Cm ;              /* [nf+1] (bound on) # rows of each contribution block*/
} stmqr_symbolic_view ;

/* The input matrix: the reference's sparse_csc, packed, Long indices, real double
 * (STMMQR/include/SparseCore.h:514).  Rows need not be sorted. */
typedef struct
{
    int64_t nrow, ncol, nzmax ;
    const int64_t *p ;               /* [ncol+1] */
    const int64_t *i ;               /* [p[ncol]] */
    const double  *x ;               /* [p[ncol]] */
} stmqr_csc_view ;

/* Scalars produced by the numeric phase (members of qr_numeric, SparseQR_struct.h:145-209). */
typedef struct
{
    int64_t rank ;                   /* # live pivot columns                         */
    int64_t rank1 ;                  /* # live pivot columns among the first ntol    */
    int64_t maxfrank ;               /* max # rows of any R block (>= 1)             */
    int64_t maxfm ;                  /* max (Hm [0..nf-1])                           */
    int64_t rh_size ;                /* total # doubles of all packed R+H blocks     */
    double  flops ;                  /* reference flop count, SparseQR_factorize.c:1571 */
} stmqr_numeric_info ;

/* Destination of the numeric factorization on the host.  All arrays are allocated by the
 * caller (the host side uses SparseCore_malloc so that qr_freenum, SparseQR.c:1245-1270,
 * can free them).  Layouts are the reference's qr_numeric. */
typedef struct
{
    double  *stack ;                 /* [rh_size] all packed R+H blocks, one stack (ns = 1);
                                        NULL: skip (already streamed)                       */
    int64_t *Roff ;                  /* [nf] Rblock[f] = stack + Roff[f]                     */
    char    *Rdead ;                 /* [n]                                                  */
    int64_t *HStair ;                /* [rjsize]                                             */
    double  *HTau ;                  /* [rjsize]                                             */
    int64_t *Hii ;                   /* [hisize] (already permuted as qr_hpinv leaves it)    */
    int64_t *Hm ;                    /* [nf]                                                 */
    int64_t *Hr ;                    /* [nf]                                                 */
    int64_t *HPinv ;                 /* [m]                                                  */
} stmqr_numeric_view ;

/* Timing / traffic counters of the last factorization (device times from CUDA events on the
 * engine's stream, milliseconds). */
typedef struct
{
    double ms_plan ;                 /* stmqr_b200_analyze: level sets, maps, arenas (host+device) */
    double ms_h2d ;                  /* upload of A                                              */
    double ms_numeric ;              /* all numeric kernels (build S .. hpinv), device time      */
    double ms_d2h ;                  /* download of the qr_numeric arrays                        */
    double ms_assemble ;             /* device time in setup+assemble+pack kernels (if profiled) */
    double ms_front ;                /* device time in front-QR kernels (if profiled)            */
    double bytes_assemble ;          /* algorithmic bytes of assembly+pack, SURVEY.md 8(d)       */
    double flops ;                   /* same as stmqr_numeric_info.flops                         */
    int64_t launches ;               /* kernels launched by the last factorization               */
    int64_t nlevels ;
    int64_t nf_small, nf_big ;       /* fronts that took the fused shared-memory / tiled path    */
    int64_t device_bytes ;           /* device memory held by the handle                         */
    /* per kernel class (only filled when options.profile_phases = 1: CUDA events around every
     * launch on the engine's stream).  Classes: 0 build_S, 1 front_setup, 2 assemble, 3 panel,
     * 4 update (WY trailing update), 5 finish+alloc, 6 pack, 7 hpinv+misc */
    double  ms_class [8] ;
    int64_t launches_class [8] ;
    double  update_flops ;           /* algorithmic flops of the trailing updates: 4*mr*nv*ncols  */
} stmqr_stats ;

/* Blocking parameters: the reference keeps them in mutable globals set by
 * chunk_getSettings (SparseQR_factorize.c:44-96, SparseQR.h:16-19).  They only affect
 * performance, never the result's structure. */
typedef struct
{
    int32_t panel ;                  /* Householder panel width of the tiled path (default 32) */
    int32_t small_elems ;            /* fronts with at most this many doubles take the fused
                                        shared-memory path (0 = engine default)               */
    int32_t profile_phases ;         /* 1: time assemble/front phases separately (adds syncs)  */
    int32_t reserved ;               /* A/B switches for tests and tuning (0 = production path): bit 0 no
                                        look-ahead, bit 1 no two-level (128-column) blocking, bit 2 no
                                        k_panel_grid, bit 3 no small-front kernel (read by analyze),
                                        bit 4 non-persistent K = 128 apply, bit 5 one column per panel
                                        exchange, bit 6 no recycling of the contribution-block arena (read
                                        by analyze); bits 8-15 ring stages of k_update_dmma (2/4),
                                        bits 16-23 max warps per panel CTA                            */
} stmqr_options ;

int  stmqr_b200_device_count (void) ;
int  stmqr_b200_create  (int device, stmqr_handle *out) ;
void stmqr_b200_destroy (stmqr_handle h) ;
int  stmqr_b200_set_options (stmqr_handle h, const stmqr_options *opt) ;

/* Build the device-side plan for a symbolic analysis: etree level sets, per-front static
 * bounds, child->parent column maps, S entry -> front column maps, memory arenas.
 * Replaces the per-stack workspace set-up of qr_factorize (SparseQR_factorize.c:295-468). */
int  stmqr_b200_analyze (stmqr_handle h, const stmqr_symbolic_view *sym) ;

/* Upload A (pattern + values) and build S = A(P,Q) values on the device
 * (replaces qr_stranspose2, SparseQR_factorize.c:755-785). */
int  stmqr_b200_upload_matrix (stmqr_handle h, const stmqr_csc_view *A) ;

/* Values-only refactorization: new values on the pattern given to the last
 * stmqr_b200_upload_matrix.  The first factorization of a pattern leaves the map "entry p of A -> slot of S"
 * on the device (the search of qr_stranspose2, SparseQR_factorize.c:755-785, is purely symbolic), so only
 * 8 bytes per entry cross the bus and S is built by one scatter.  The caller vouches that Ap/Ai are
 * unchanged.  stmqr_b200_factorize / _factorize_streamed do the same on their own whenever the handle holds
 * a pattern of the same shape: values first, and the caller's Ap/Ai are uploaded beside the numeric phase and
 * compared with the resident pattern on the device (a mismatch repeats the call through the full path;
 * STMQR_B200_SPECULATIVE=0 turns this off). */
int  stmqr_b200_upload_values (stmqr_handle h, const double *Ax, int64_t nnz) ;
int  stmqr_b200_refactorize_values (stmqr_handle h, const double *Ax, int64_t nnz, double tol, int64_t ntol,
                                    stmqr_numeric_info *info) ;

/* Numeric factorization of the resident matrix: per etree level, front set-up (qr_fsize
 * :1066), assembly (qr_assemble :1151), front QR (qr_front :1383, qr_larftb :1851) and
 * packing (qr_cpack :1639, qr_rhpack :1691), then the row permutation (qr_hpinv :991).
 * tol < 0 or !do_rank_detection disables rank detection (:285-289).  Results stay on the
 * device until stmqr_b200_download. */
int  stmqr_b200_factorize_resident (stmqr_handle h, double tol, int64_t ntol,
                                    stmqr_numeric_info *info) ;

/* Copy the factorization into caller-allocated host arrays (layout of qr_numeric). */
int  stmqr_b200_download (stmqr_handle h, const stmqr_numeric_view *out) ;

/* Convenience: upload_matrix + factorize_resident (the whole of qr_factorize's device work
 * with host buffers on both sides; download is separate because the caller must size
 * `stack` from info->rh_size first). */
int  stmqr_b200_factorize (stmqr_handle h, const stmqr_csc_view *A, double tol, int64_t ntol,
                           stmqr_numeric_info *info) ;

/* The same with the download of the packed R+H stack overlapped with the factorization: the blocks
 * of an etree level are final as soon as the level is packed and are copied into `stack` while the next
 * levels run (the reference's qr_factorize also allocates its stacks by the symbolic bound and shrinks
 * them afterwards, SparseQR_factorize.c:405-410,:560-660).  `stack` must hold stmqr_b200_rh_bound()
 * doubles; on return the first info->rh_size of them are valid.  Afterwards call stmqr_b200_download
 * with out->stack = NULL for the remaining arrays. */
int  stmqr_b200_rh_bound (stmqr_handle h, int64_t *doubles) ;
/* The same overlap for the numeric phase in pieces (several GPUs: every handle streams the blocks of its
 * own fronts): stream_begin before factorize_begin, stream_end after factorize_hpinv_b. */
int  stmqr_b200_stream_begin (stmqr_handle h, double *stack, int64_t capacity) ;
int  stmqr_b200_stream_end (stmqr_handle h) ;
int  stmqr_b200_factorize_streamed (stmqr_handle h, const stmqr_csc_view *A, double tol, int64_t ntol,
                                    double *stack, int64_t capacity, stmqr_numeric_info *info) ;

/* ---- the numeric phase in pieces, and the etree partitioned over several GPUs ------------------
 * One handle per GPU (one process per GPU under torchrun, or several handles in one process).
 * Replaces the reference's task tree + TPSM scheduler (SparseQR_analyze.c:705-1161,
 * SparseQR_multithreads.c:14-115): independent etree subtrees are factorized on different GPUs,
 * the contribution blocks of the subtree roots move to the GPU that owns the top of the tree
 * (NCCL send/recv or peer copies, done by the host between factorize_levels (1) and (2)), and the
 * global row permutation (qr_hpinv) is finished after an element-wise max-merge of Hm, Hr, Cm, Rdead
 * and W over the GPUs.  factorize_resident == begin, levels (0), hpinv_a, hpinv_b on one GPU. */
int  stmqr_b200_factorize_begin (stmqr_handle h, double tol, int64_t ntol) ;
int  stmqr_b200_factorize_levels (stmqr_handle h, int part /* 0 all, 1 my subtrees, 2 top of tree */) ;
int  stmqr_b200_factorize_hpinv_a (stmqr_handle h) ;
int  stmqr_b200_factorize_hpinv_b (stmqr_handle h, stmqr_numeric_info *info) ;
int  stmqr_b200_sync (stmqr_handle h) ;

/* Host only (no device needed), deterministic: owner[f] in [0,nparts) and is_top[f] for every front.
 * The top of the tree (fronts above the cut, closed upwards) belongs to part 0. */
int  stmqr_b200_partition_fronts (const stmqr_symbolic_view *sym, int nparts, int32_t *owner,
                                  int32_t *is_top) ;
int  stmqr_b200_set_partition (stmqr_handle h, int nparts, int mypart, const int32_t *owner,
                               const int32_t *is_top) ;

/* Host-only planner (no device needed): a handle on which stmqr_b200_analyze / stmqr_b200_set_partition
 * only compute the plan -- etree level schedule, arena sizes, the recycled contribution-block offsets --
 * and account for the device memory the plan would take.  Every other call fails with STMQR_ERR_NO_DEVICE /
 * STMQR_ERR_INVALID.  Replaces the reference's stack sizing (Stack_maxstack, SparseQR_analyze.c:1061-1161)
 * as the place where a caller learns whether a problem fits. */
typedef struct
{
    int64_t nlevels ;
    int64_t F_doubles ;              /* front arena: the widest etree level (bound sizes)                */
    int64_t C_doubles ;              /* contribution-block arena with recycling (high-water mark)        */
    int64_t C_doubles_unrecycled ;   /* sum of the bounds of all blocks (what it would be without)       */
    int64_t R_doubles ;              /* packed R+H arena (symbolic bound)                                */
    int64_t device_bytes ;           /* everything the handle allocates                                  */
    int64_t nparts, mypart ;
} stmqr_plan_info ;
int  stmqr_b200_create_planner (stmqr_handle *out) ;
/* Coff, Csize [nf] (offset and bound size of every contribution block in the arena, doubles), level [nf]
 * (etree level of every front); any of them may be NULL. */
int  stmqr_b200_plan_info (stmqr_handle h, stmqr_plan_info *out, int64_t *Coff, int64_t *Csize, int32_t *level) ;

/* ---- general ownership + the C data plane (csrc/multigpu.cuh) ---------------------------------------------
 * Every front has an owner GPU (stmqr_b200_map_fronts: subtrees below the cut as partition_fronts deals them,
 * a front above the cut on the GPU of its heaviest child, so the upper levels spread over the GPUs too).  Each
 * GPU walks the etree levels over its own fronts; after level l the contribution blocks (+ row ids + Cm/Hr/Hm)
 * of the level-l fronts whose parent lives elsewhere move there, in symbolic-bound sizes, ordered on the
 * engine's stream: no handshake, no host synchronisation.  At the end Hm|Hr|Cm, Rdead and the row permutation
 * are merged with element-wise max all-reduces.  Transports: NCCL (one process per GPU; libnccl is dlopen'ed)
 * and peer copies between handles of one process (one host thread per handle). */
int  stmqr_b200_map_fronts (const stmqr_symbolic_view *sym, int nparts, int32_t *owner) ;
int  stmqr_b200_set_ownership (stmqr_handle h, int nparts, int mypart, const int32_t *owner) ;
int  stmqr_b200_nccl_unique_id (void *id128) ;                 /* rank 0; ship the 128 bytes to the others */
int  stmqr_b200_comm_init (stmqr_handle h, int nranks, int rank, const void *id128) ;
int  stmqr_b200_peer_group_create (stmqr_handle *handles, int n, void **group) ;
void stmqr_b200_peer_group_destroy (void *group) ;
/* one GPU's share; all GPUs of the group call it at the same time */
int  stmqr_b200_factorize_dist (stmqr_handle h, double tol, int64_t ntol, stmqr_numeric_info *info) ;
/* every handle of a peer group, one host thread each; infos [n] (may be NULL) */
int  stmqr_b200_factorize_multi (void *group, double tol, int64_t ntol, stmqr_numeric_info *infos) ;
/* collective, after factorize_dist: HStair, HTau, the permuted Hii and the offset of every packed block in its
 * owner's stack become global on every GPU (the host then needs the integer side from ONE of them);
 * factorize_multi_ex (..., gather = 1) calls it in every thread */
int  stmqr_b200_gather_outputs (stmqr_handle h) ;
int  stmqr_b200_factorize_multi_ex (void *group, double tol, int64_t ntol, stmqr_numeric_info *infos, int gather) ;

/* Cooperative fronts (NCCL transport, 3 GPUs or more; STMQR_B200_COOP=0 off, =2 also on 2 GPUs).  An etree level
 * that consists of ONE large front -- the top chain of a 3-D problem, where subtree ownership leaves most of the
 * flops on one GPU (the reference's task tree has the same serial top, SparseQR_analyze.c:705-859)
 * -- is factorized by all GPUs: the front's home GPU keeps assembly, every panel and the pack; the K = 128
 * trailing update is spread over the GPUs by chunks of 512 columns.  Per outer block home broadcasts the block
 * reflector; the owner of the block after next applies it to that block first and sends the block home, where it
 * arrives while the block in between is being factorized.  factorize_dist decides per level (same test on every
 * GPU).  stmqr_b200_coop_chunks returns the chunk -> GPU map it uses for a front of fn columns (host only:
 * chunk_owner [ceil (fn / 512)]; chunk 0 always stays on home, home's share shrinks as GPUs are added because
 * it also runs the panels). */
int  stmqr_b200_coop_chunks (int nparts, int home, int64_t fn, int32_t *chunk_owner, int64_t *nchunks) ;

#define STMQR_ARRAY_HM    0   /* int32 [nf] */
#define STMQR_ARRAY_HR    1   /* int32 [nf] */
#define STMQR_ARRAY_CM    2   /* int32 [nf] */
#define STMQR_ARRAY_RDEAD 3   /* int8  [n]  */
#define STMQR_ARRAY_W     4   /* int32 [m]  row permutation workspace of qr_hpinv */
/* Device address of one of the arrays that are merged (element-wise max) over the GPUs. */
int  stmqr_b200_device_array (stmqr_handle h, int which, void **ptr, int64_t *count,
                              int32_t *elem_bytes) ;

/* Device regions of front f that its parent on another GPU needs (the packed contribution block,
 * qr_cpack layout SparseQR_factorize.c:1639-1685, and the row ids of its rows, Hii[Hip[f]+Hr[f]..)).
 * Owner side: cm = hr = hm = -1 (read from the device).  Receiver side: pass the owner's values. */
typedef struct
{
    void   *C ;   int64_t C_doubles ;
    void   *Hii ; int64_t Hii_ints ;        /* int32 */
    int64_t cm, hr, hm ;
} stmqr_front_regions ;
int  stmqr_b200_front_regions (stmqr_handle h, int64_t f, int64_t cm, int64_t hr, int64_t hm,
                               stmqr_front_regions *out) ;

/* ---- the consumers next to the path, on the resident factorization -----------------
 * Q-apply and R-solve that read the packed R+H blocks where the factorization left them in HBM, so a
 * least-squares solve does not download the factor at all.  They replace, for factorizations without
 * column singletons (the multifrontal part), the reference's
 *     QR_qmult (QR_QTX / QR_QX)           STMMQR/src/qr/SparseQR.c:1838-2110 (qr_private_Happly :1706,
 *                                          qr_private_get_H_vectors :1455, row permutation HPinv)
 *     qr_rsolve                           STMMQR/src/qr/SparseQR.c:2218-2465 (QR_solve systems
 *                                          QR_RX_EQUALS_B: use_Qfill = 0, QR_RETX_EQUALS_B: use_Qfill = 1)
 * All matrices are host arrays, column major: X, Y m-by-nx (ld = m); B m-by-nrhs (ld = m, rows >= rank
 * are ignored); the solution n-by-nrhs (ld = n).  Dead columns get the basic solution x_j = 0.
 * The etree is walked level by level (all fronts of a level in one launch). */
int  stmqr_b200_qmult (stmqr_handle h, int method /* 0: Y = Q'X, 1: Y = QX */, int64_t nx,
                       const double *X, double *Y) ;
int  stmqr_b200_rsolve (stmqr_handle h, int use_Qfill, int64_t nrhs, const double *B, double *X) ;
/* x = E * (R \ (Q'b)): min ||Ax - b|| as qrtest.c:11-53 does it with QR_qmult + QR_solve; device_ms
 * (optional) receives the device time including the two host<->device copies of b and x. */
int  stmqr_b200_solve_ls (stmqr_handle h, int64_t nrhs, const double *B, double *X, double *device_ms) ;

/* ---- the explicit factor ------------------------------------------------------------
 * R of the resident factorization as a compressed-column matrix, extracted on the device: replaces
 * qr_rcount / qr_rconvert (STMMQR/src/qr/SparseLQ.c:102-297, :299-520) in the configuration SparseQR()'s
 * getR path uses for the multifrontal part (n1rows = 0, n2 = n, getT = 0; Ra/Rap branch).  Columns are in the
 * order of S (apply Qfill for A's order), inside a column the entries are ordered as the reference appends
 * them (by front, then by row), exact zeros are dropped, only rows < econ are kept.
 *   rcount:   Rp [n+1] = column pointers (may be NULL), *nnzR = Rp [n]
 *   rconvert: Rp [n+1], Ri [nnzR], Rx [nnzR]   (call rcount first to size Ri/Rx) */
int  stmqr_b200_rcount (stmqr_handle h, int64_t econ, int64_t *Rp, int64_t *nnzR) ;
int  stmqr_b200_rconvert (stmqr_handle h, int64_t econ, int64_t *Rp, int64_t *Ri, double *Rx) ;

int  stmqr_b200_get_stats (stmqr_handle h, stmqr_stats *out) ;

/* FP64 peak microbenchmarks on the handle's device, used as roofline denominators (the driver's
 * MEASURED_PEAKS.json has no FP64 figure): register-resident DMMA (mma.sync f64) and DFMA loops,
 * one CTA set per SM, timed with CUDA events.  TFLOP/s. */
int  stmqr_b200_measure_fp64_peak (stmqr_handle h, double *dmma_tflops, double *dfma_tflops) ;
const char *stmqr_b200_last_error (stmqr_handle h) ;

/* Debug / parity taps used by the tests: copy one front's assembled F (before the QR) or
 * its factorized F to the host.  Only valid when the handle was created with
 * stmqr_b200_set_debug_capture (h, 1) before factorize. */
int  stmqr_b200_set_debug_capture (stmqr_handle h, int on) ;
int  stmqr_b200_get_front (stmqr_handle h, int64_t f, int which /*0 assembled, 1 factorized*/,
                           double *F, int64_t capacity, int64_t *fm, int64_t *fn) ;

#ifdef __cplusplus
}
#endif
#endif /* STMQR_B200_H */
