/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat, ctypes-friendly wrapper around the UNMODIFIED reference library
 * (oracle/_ref/libstmmqr_ref.so) so that tests/, bench.py's reference arm and
 * tests/golden/make_golden.py can drive the reference's own API
 * (SparseCore_read_matrix, SparseQR, QR_qmult, QR_solve ...) and look inside its
 * qr_symbolic / qr_numeric objects through the plain views of include/stmqr_b200.h.
 *
 * It also interposes `qr_factorize` (STMMQR/include/SparseQR.h:127): the reference's
 * SparseQR() (STMMQR/src/qr/SparseQR.c:349,371) then calls the definition below, which
 *   - optionally keeps a copy of the matrix actually factorized (A, or the singleton-pruned Y),
 *   - forwards either to the reference's own qr_factorize (RTLD_NEXT) or to the B200 drop-in
 *     (symbol stmqr_b200_qr_factorize of libstmqr_dropin.so) -- selected by rh_set_backend().
 * This is exactly the dynamic-link-level drop-in a maintainer gets with LD_PRELOAD.
 *
 * Compiled against the reference headers in place (-I/root/reference/STMMQR/include) into
 * oracle/_ref/libref_harness.so by __graft_entry__.build().
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <float.h>
#include <stdio.h>
#include <string.h>
#include <sys/time.h>
#include "SparseQR.h"
#include "tpsm.h"
#include "../include/stmqr_b200.h"

typedef qr_numeric *(*qr_factorize_fn) (sparse_csc **, Long, double, Long, qr_symbolic *,
    sparse_common *) ;

static int g_backend = 0 ;                 /* 0 reference CPU, 1 B200 drop-in */
static qr_factorize_fn g_dropin = NULL ;
static int g_tap = 0 ;
static sparse_csc *g_tap_A = NULL ;        /* copy of the matrix given to qr_factorize */
static double g_tap_tol = 0 ;
static Long g_tap_ntol = 0 ;
static double g_last_fac_seconds = 0 ;
typedef void (*rh_probe_fn) (void *QRsym) ;
static rh_probe_fn g_probe = NULL ;
extern void openblas_set_num_threads (int) ;

static double now_s (void)
{
    struct timeval tv ;
    gettimeofday (&tv, NULL) ;
    return tv.tv_sec + tv.tv_usec * 1e-6 ;
}

/* ---- the interposed entry point -------------------------------------------------------- */
qr_numeric *qr_factorize (sparse_csc **Ahandle, Long freeA, double tol, Long ntol,
    qr_symbolic *QRsym, sparse_common *cc)
{
    if (g_tap && Ahandle && *Ahandle)
    {
        if (g_tap_A) SparseCore_free_sparse (&g_tap_A, cc) ;
        g_tap_A = SparseCore_copy_sparse (*Ahandle, cc) ;
        g_tap_tol = tol ;
        g_tap_ntol = ntol ;
    }
    if (g_backend == 2)
    {
        /* symbolic-only probe: hand the symbolic object to the registered callback and stop.  SparseQR then
         * frees everything and returns NULL (used by tools/ to plan large configurations without running them). */
        if (g_probe) g_probe ((void *) QRsym) ;
        if (freeA) SparseCore_free_sparse (Ahandle, cc) ;
        cc->status = SPARSE_INVALID ;
        return NULL ;
    }
    qr_factorize_fn fn = NULL ;
    if (g_backend == 1) fn = g_dropin ;
    else fn = (qr_factorize_fn) dlsym (RTLD_NEXT, "qr_factorize") ;
    if (!fn)
    {
        fprintf (stderr, "ref_harness: no qr_factorize backend (%d)\n", g_backend) ;
        cc->status = SPARSE_INVALID ;
        return NULL ;
    }
    double t0 = now_s () ;
    qr_numeric *r = fn (Ahandle, freeA, tol, ntol, QRsym, cc) ;
    g_last_fac_seconds = now_s () - t0 ;
    return r ;
}

int rh_set_backend (int backend, const char *dropin_path)
{
    if (backend == 1)
    {
        void *h = dlopen (dropin_path, RTLD_NOW | RTLD_GLOBAL) ;
        if (!h) { fprintf (stderr, "ref_harness: %s\n", dlerror ()) ; return -1 ; }
        g_dropin = (qr_factorize_fn) dlsym (h, "stmqr_b200_qr_factorize") ;
        if (!g_dropin) { fprintf (stderr, "ref_harness: %s\n", dlerror ()) ; return -2 ; }
    }
    g_backend = backend ;
    return 0 ;
}
int rh_get_backend (void) { return g_backend ; }
void rh_set_tap (int on) { g_tap = on ; }
double rh_last_fac_seconds (void) { return g_last_fac_seconds ; }
void rh_set_blas_threads (int n) { openblas_set_num_threads (n) ; }

/* ---- sparse_common / matrices ------------------------------------------------------------ */
void *rh_start (void)
{
    sparse_common *cc = (sparse_common *) calloc (1, sizeof (sparse_common)) ;
    SparseCore_start (cc) ;
    return cc ;
}
void rh_finish (void *ccv)
{
    sparse_common *cc = (sparse_common *) ccv ;
    if (g_tap_A) SparseCore_free_sparse (&g_tap_A, cc) ;
    SparseCore_finish (cc) ;
    free (cc) ;
}
int rh_status (void *ccv) { return ((sparse_common *) ccv)->status ; }
void rh_clear_status (void *ccv) { ((sparse_common *) ccv)->status = SPARSE_OK ; }
long rh_memory_inuse (void *ccv) { return (long) ((sparse_common *) ccv)->memory_inuse ; }
long rh_malloc_count (void *ccv) { return (long) ((sparse_common *) ccv)->malloc_count ; }
double rh_flopcount (void *ccv) { return ((sparse_common *) ccv)->SPQR_flopcount ; }
double rh_flopcount_bound (void *ccv) { return ((sparse_common *) ccv)->SPQR_flopcount_bound ; }

/* qrtest.c:105-121: SparseCore_read_matrix (fp, prefer=1, ...) with the two graph-dump files */
void *rh_read_mtx (void *ccv, const char *path)
{
    sparse_common *cc = (sparse_common *) ccv ;
    FILE *fp = fopen (path, "r") ;
    if (!fp) return NULL ;
    FILE *n1 = fopen ("/dev/null", "a+"), *n2 = fopen ("/dev/null", "a+") ;
    int mtype = 0 ;
    sparse_csc *A = (sparse_csc *) SparseCore_read_matrix (fp, 1, &mtype, cc, n1, n2, 0) ;
    fclose (fp) ; fclose (n1) ; fclose (n2) ;
    if (A && mtype != SPARSE_CSC) { return NULL ; }
    return A ;
}

void *rh_csc_from_arrays (void *ccv, long m, long n, long nnz, const int64_t *Ap,
    const int64_t *Ai, const double *Ax)
{
    sparse_common *cc = (sparse_common *) ccv ;
    sparse_csc *A = SparseCore_allocate_sparse (m, n, nnz, TRUE, TRUE, 0, SPARSE_REAL, cc) ;
    if (!A) return NULL ;
    memcpy (A->p, Ap, (n+1) * sizeof (Long)) ;
    memcpy (A->i, Ai, nnz * sizeof (Long)) ;
    memcpy (A->x, Ax, nnz * sizeof (double)) ;
    return A ;
}

void rh_csc_view (void *Av, stmqr_csc_view *out)
{
    sparse_csc *A = (sparse_csc *) Av ;
    out->nrow = A->nrow ; out->ncol = A->ncol ; out->nzmax = A->nzmax ;
    out->p = (const int64_t *) A->p ; out->i = (const int64_t *) A->i ;
    out->x = (const double *) A->x ;
}
void *rh_tap_matrix (void) { return g_tap_A ; }
double rh_tap_tol (void) { return g_tap_tol ; }
long rh_tap_ntol (void) { return g_tap_ntol ; }

void rh_free_sparse (void *ccv, void *Av)
{
    sparse_csc *A = (sparse_csc *) Av ;
    SparseCore_free_sparse (&A, (sparse_common *) ccv) ;
}

/* qrtest.c:133-142 */
double rh_default_tol (void *ccv, void *Av)
{
    sparse_csc *A = (sparse_csc *) Av ;
    double mx = qr_maxcolnorm (A, (sparse_common *) ccv) ;
    if (mx == 0) mx = 1 ;
    return 20 * ((double) A->nrow + (double) A->ncol) * DBL_EPSILON * mx ;
}

/* The reference's TPSM pool cannot be re-initialised after TPSM_destroy (its distribution tables
 * are not reset: initializeThreadsDistribution, tpsm_distribution.c:37-38, fails the second
 * time), so the harness creates it once per process, as the reference driver does
 * (qrtest.c:150), and leaves it alive; idle workers sleep on a condition variable. */
static int g_pool_size = 0 ;
void rh_pool_ensure (int pool)
{
    if (g_pool_size == 0)
    {
        TPSM_init (pool, 2000, 3000, TPSM_NODE_AFFINITY) ;
        g_pool_size = pool ;
    }
}
int rh_pool_size (void) { return g_pool_size ; }

/* ---- the reference driver preamble + SparseQR (qrtest.c:144-180) ------------------------- */
/* ordering_arg: qrtest's third argument (0 AMD, 1 COLAMD, 2 METIS, 3 NESDIS, else DEFAULT).
 * grain <= 1: serial tree (no TPSM tasks).  grain > 1: TPSM pool of `pool` workers. */
void *rh_sparseqr (void *ccv, void *Av, int ordering_arg, double tol, double grain, int pool)
{
    sparse_common *cc = (sparse_common *) ccv ;
    sparse_csc *A = (sparse_csc *) Av ;
    long ordering ;
    switch (ordering_arg)
    {
        case 0: ordering = QR_ORDERING_AMD ; break ;
        case 1: ordering = QR_ORDERING_COLAMD ; break ;
        case 2: ordering = QR_ORDERING_ONLYMETIS ; break ;
        case 3: ordering = QR_ORDERING_NESDIS ; break ;
        default: ordering = QR_ORDERING_DEFAULT ;
    }
    cc->SPQR_grain = grain ;
    cc->status = SPARSE_OK ;
    int pooled = (grain > 1 && pool > 0) ;
    if (pooled) rh_pool_ensure (pool) ;
    chunk_getSettings (32, 5000, 4, 4) ;
    cc->QR_CHUNK_FLAG = 0 ;
    Relaxfactor_setting (A->ncol, SparseCore_nnz (A, cc), RELAX_FOR_QR, cc) ;
    char name [8] = "rh" ;
    SparseQR_factorization *QR = SparseQR (ordering, tol, A, cc, name) ;
    return QR ;
}

void rh_free_qr (void *ccv, void *QRv)
{
    SparseQR_factorization *QR = (SparseQR_factorization *) QRv ;
    SparseQR_free (&QR, (sparse_common *) ccv) ;
}

void rh_qr_info (void *QRv, double *out /* [12] */)
{
    SparseQR_factorization *QR = (SparseQR_factorization *) QRv ;
    out [0] = QR->Ana_time ; out [1] = QR->Fac_time ; out [2] = QR->tol ;
    out [3] = QR->n1rows ; out [4] = QR->n1cols ; out [5] = QR->rank ;
    out [6] = QR->QRnum->rank ; out [7] = QR->QRnum->rank1 ; out [8] = QR->QRnum->maxfrank ;
    out [9] = QR->QRnum->maxfm ; out [10] = QR->QRnum->ns ; out [11] = QR->QRnum->ntasks ;
}

/* R of the multifrontal part as CSC through the reference's own qr_rcount / qr_rconvert (SparseLQ.c:102,299;
 * n1rows = 0, n2 = n, getT = 0).  Pass Ri = Rx = NULL to get the column pointers / nnz only.  Returns nnz(R). */
long rh_rconvert (void *QRv, long econ, int64_t *Rp_out, int64_t *Ri, double *Rx)
{
    SparseQR_factorization *QR = (SparseQR_factorization *) QRv ;
    qr_symbolic *S = QR->QRsym ;
    Long n = S->n ;
    Long *Ra = (Long *) calloc ((size_t) n + 1, sizeof (Long)) ;
    Long nh = 0 ;
    qr_rcount (S, QR->QRnum, 0, econ, n, 0, Ra, NULL, NULL, &nh) ;
    Long tot = 0 ;
    for (Long j = 0 ; j < n ; j++) { Long c = Ra [j] ; Ra [j] = tot ; tot += c ; }
    Ra [n] = tot ;
    if (Rp_out) for (Long j = 0 ; j <= n ; j++) Rp_out [j] = Ra [j] ;
    if (Ri && Rx)
    {
        /* qr_rconvert advances the column pointers while it fills (Rap [j]++) */
        qr_rconvert (S, QR->QRnum, 0, econ, n, 0, Ra, (Long *) Ri, Rx, NULL, NULL, NULL, NULL, NULL, NULL, NULL) ;
    }
    free (Ra) ;
    return (long) tot ;
}

void rh_set_probe (rh_probe_fn fn) { g_probe = fn ; }

static void sym_view_of (qr_symbolic *S, stmqr_symbolic_view *v) ;
void rh_sym_view (void *QRv, stmqr_symbolic_view *v) { sym_view_of (((SparseQR_factorization *) QRv)->QRsym, v) ; }
void rh_symbolic_view (void *QRsymv, stmqr_symbolic_view *v) { sym_view_of ((qr_symbolic *) QRsymv, v) ; }

static void sym_view_of (qr_symbolic *S, stmqr_symbolic_view *v)
{
    v->m = S->m ; v->n = S->n ; v->anz = S->anz ; v->nf = S->nf ; v->maxfn = S->maxfn ;
    v->rjsize = S->rjsize ; v->hisize = S->hisize ;
    v->do_rank_detection = S->do_rank_detection ; v->keepH = S->keepH ;
    v->Sp = S->Sp ; v->Sj = S->Sj ; v->Qfill = S->Qfill ; v->PLinv = S->PLinv ;
    v->Sleft = S->Sleft ; v->Parent = S->Parent ; v->Child = S->Child ; v->Childp = S->Childp ;
    v->Super = S->Super ; v->Rp = S->Rp ; v->Rj = S->Rj ; v->Post = S->Post ; v->Hip = S->Hip ;
    v->Fm = S->Fm ; v->Cm = S->Cm ;
}

/* pointers into the reference's qr_numeric; Roff_out[f] = offset of Rblock[f] in a virtual
 * concatenation of the ns stacks (stack s starts at sum of Stack_size[0..s-1]). */
void rh_num_view (void *QRv, stmqr_numeric_view *v, int64_t *Roff_out, int64_t *total)
{
    qr_symbolic *S = ((SparseQR_factorization *) QRv)->QRsym ;
    qr_numeric *N = ((SparseQR_factorization *) QRv)->QRnum ;
    v->stack = NULL ; v->Roff = NULL ;
    v->Rdead = N->Rdead ; v->HStair = N->HStair ; v->HTau = N->HTau ; v->Hii = N->Hii ;
    v->Hm = N->Hm ; v->Hr = N->Hr ; v->HPinv = N->HPinv ;
    int64_t tot = 0 ;
    for (Long s = 0 ; s < N->ns ; s++) tot += N->Stack_size [s] ;
    *total = tot ;
    if (Roff_out)
    {
        for (Long f = 0 ; f < S->nf ; f++)
        {
            int64_t base = 0 ; int found = 0 ;
            for (Long s = 0 ; s < N->ns && !found ; s++)
            {
                double *b = N->Stacks [s] ;
                if (N->Rblock [f] >= b && N->Rblock [f] <= b + N->Stack_size [s])
                {
                    Roff_out [f] = base + (N->Rblock [f] - b) ; found = 1 ;
                }
                base += N->Stack_size [s] ;
            }
            if (!found) Roff_out [f] = -1 ;
        }
    }
}
/* copy the ns stacks back to back into dst[total] */
void rh_num_copy_stacks (void *QRv, double *dst)
{
    qr_numeric *N = ((SparseQR_factorization *) QRv)->QRnum ;
    for (Long s = 0 ; s < N->ns ; s++)
    {
        memcpy (dst, N->Stacks [s], N->Stack_size [s] * sizeof (double)) ;
        dst += N->Stack_size [s] ;
    }
}

/* ---- consumers of qr_numeric (reference code, untouched): Q apply and solve -------------- */
/* Y = op(Q) X, method as SparseQR_definitions.h:24-27; X is nrow-by-ncol column-major. */
int rh_qmult (void *ccv, void *QRv, int method, long nrow, long ncol, const double *X, double *Y)
{
    sparse_common *cc = (sparse_common *) ccv ;
    dense_array *Xd = SparseCore_allocate_dense (nrow, ncol, nrow, SPARSE_REAL, cc) ;
    memcpy (Xd->x, X, nrow*ncol*sizeof (double)) ;
    dense_array *Yd = QR_qmult (method, (SparseQR_factorization *) QRv, Xd, cc) ;
    SparseCore_free_dense (&Xd, cc) ;
    if (!Yd) return -1 ;
    memcpy (Y, Yd->x, Yd->nrow*Yd->ncol*sizeof (double)) ;
    SparseCore_free_dense (&Yd, cc) ;
    return 0 ;
}
/* X = solve (system, B); B is brow-by-ncol; X is xrow-by-ncol */
int rh_solve (void *ccv, void *QRv, int system, long brow, long ncol, const double *B,
    long xrow, double *X)
{
    sparse_common *cc = (sparse_common *) ccv ;
    dense_array *Bd = SparseCore_allocate_dense (brow, ncol, brow, SPARSE_REAL, cc) ;
    memcpy (Bd->x, B, brow*ncol*sizeof (double)) ;
    dense_array *Xd = QR_solve (system, (SparseQR_factorization *) QRv, Bd, cc) ;
    SparseCore_free_dense (&Bd, cc) ;
    if (!Xd) return -1 ;
    if ((long) Xd->nrow != xrow) { SparseCore_free_dense (&Xd, cc) ; return -2 ; }
    memcpy (X, Xd->x, Xd->nrow*Xd->ncol*sizeof (double)) ;
    SparseCore_free_dense (&Xd, cc) ;
    return 0 ;
}
/* Y = A*X (transpose=0) or A'*X (transpose=1): SparseCore_sdmult, SparseCore.h:1111 */
int rh_sdmult (void *ccv, void *Av, int transpose, long xrow, long ncol, const double *X,
    long yrow, double *Y)
{
    sparse_common *cc = (sparse_common *) ccv ;
    double one [2] = {1,0}, zero [2] = {0,0} ;
    dense_array *Xd = SparseCore_allocate_dense (xrow, ncol, xrow, SPARSE_REAL, cc) ;
    dense_array *Yd = SparseCore_zeros (yrow, ncol, SPARSE_REAL, cc) ;
    memcpy (Xd->x, X, xrow*ncol*sizeof (double)) ;
    int ok = SparseCore_sdmult ((sparse_csc *) Av, transpose, one, zero, Xd, Yd, cc) ;
    memcpy (Y, Yd->x, yrow*ncol*sizeof (double)) ;
    SparseCore_free_dense (&Xd, cc) ;
    SparseCore_free_dense (&Yd, cc) ;
    return ok ? 0 : -1 ;
}

/* qrtest.c:11-53 check_error restated (the driver's own residual: x = 0..n-1, b = A x,
 * x_sol = E (R \ (Q' b)), res = ||x_sol - x||_2 / n).  Square A only, as in the reference. */
double rh_check_error (void *ccv, void *Av, void *QRv)
{
    sparse_csc *A = (sparse_csc *) Av ;
    long n = A->ncol, m = A->nrow ;
    double *x = (double *) malloc (n * sizeof (double)) ;
    double *b = (double *) calloc (m, sizeof (double)) ;
    double *y = (double *) calloc (m, sizeof (double)) ;
    double *xs = (double *) calloc (n, sizeof (double)) ;
    for (long i = 0 ; i < n ; i++) x [i] = (double) i ;
    double res = -1 ;
    if (rh_sdmult (ccv, Av, 0, n, 1, x, m, b) == 0 &&
        rh_qmult (ccv, QRv, QR_QTX, m, 1, b, y) == 0 &&
        rh_solve (ccv, QRv, QR_RETX_EQUALS_B, m, 1, y, n, xs) == 0)
    {
        double s = 0 ;
        for (long i = 0 ; i < n ; i++) { double d = xs [i] - (double) i ; s += d*d ; }
        res = sqrt (s) / (double) n ;
    }
    free (x) ; free (b) ; free (y) ; free (xs) ;
    return res ;
}

/* ---- numeric refactorization on an existing symbolic object ------------------------------- */
/* Calls qr_factorize (the interposed one above -> selected backend) exactly as SparseQR.c:349
 * does (&A, FALSE, tol, n, QRsym, cc), replaces QR->QRnum by the result and returns the seconds
 * spent inside the call (the reference's Fac_time interval).  Only for factorizations without
 * singletons (QR->n1cols == 0).  pool > 0: a TPSM pool of that many workers is alive around the
 * call, as the reference driver arranges (qrtest.c:150,193); its creation is not timed. */
double rh_refactorize (void *ccv, void *Av, void *QRv, int pool)
{
    sparse_common *cc = (sparse_common *) ccv ;
    sparse_csc *A = (sparse_csc *) Av ;
    SparseQR_factorization *QR = (SparseQR_factorization *) QRv ;
    if (QR->n1cols != 0) return -1 ;
    if (QR->QRnum) qr_freenum (&QR->QRnum, cc) ;
    cc->status = SPARSE_OK ;
    if (pool > 0) rh_pool_ensure (pool) ;
    chunk_getSettings (32, 5000, 4, 4) ;
    if (cc->QR_CHUNK_FLAG) { FCHUNK = 80 ; SMALL = 8000 ; }     /* as qr_analyze left it (:666-670) */
    double t0 = now_s () ;
    QR->QRnum = qr_factorize (&A, FALSE, QR->tol, A->ncol, QR->QRsym, cc) ;
    double t = now_s () - t0 ;
    if (!QR->QRnum) return -2 ;
    QR->rank = QR->n1rows + QR->QRnum->rank1 ;
    return t ;
}
