/* qr_factorize_b200.c -- the drop-in replacement of the reference's numeric phase.
 *
 * Exports
 *     qr_numeric *qr_factorize (sparse_csc **Ahandle, Long freeA, double tol, Long ntol,
 *                               qr_symbolic *QRsym, sparse_common *cc)
 * with the prototype of STMMQR/include/SparseQR.h:127-135 and the contract of the reference
 * definition (STMMQR/src/qr/SparseQR_factorize.c:222-749):
 *   - QRsym == NULL            -> free A if freeA, return NULL                       (:247-254)
 *   - freeA                    -> SparseCore_free_sparse (Ahandle) always            (:324-327)
 *   - tol < 0 / !do_rank_detection -> no rank detection                              (:285-289)
 *   - every member of the returned qr_numeric is allocated with SparseCore_malloc/calloc with
 *     exactly the sizes qr_freenum (STMMQR/src/qr/SparseQR.c:1245-1270) frees
 *   - failure: return NULL with cc->status < SPARSE_OK                               (:329-333)
 * It is compiled against the reference's own headers (it is host code of the reference's
 * library: link it instead of -- or LD_PRELOAD it in front of -- SparseQR_factorize.o's
 * qr_factorize).  All numeric work goes through the C ABI of include/stmqr_b200.h to the
 * CUDA engine; there is no CPU path here.
 *
 * ns = 1: one exactly-sized host stack holds all packed R+H blocks (any front->stack
 * placement is legal for the consumers, which only use Rblock[f]).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/time.h>
#include "SparseQR.h"
#include "stmqr_b200.h"

/* One engine handle per process, re-planned when the symbolic object changes. */
static stmqr_handle g_handle = NULL ;
static int g_have_plan = 0 ;
static uint64_t g_planned_key = 0 ;

/* Key of the device plan: a 64-bit hash of the CONTENT of the symbolic object, not of its address (a
 * new qr_symbolic can be malloc'ed where a freed one was).  The arrays of size O(nf) are hashed in
 * full; of the O(n), O(m), O(anz) arrays every stride-th word plus the last 512 words are hashed so
 * that the key costs well under a millisecond on the 1M-unknown configs.  Two analyses of the same
 * pattern under another ordering / relaxation / tolerance mode differ in Super/Rp/Fm/Cm (hashed in
 * full) long before they agree on all sampled words.  STMQR_B200_CACHE_PLAN=0 re-plans on every call;
 * stmqr_b200_dropin_invalidate_plan() drops the cached plan explicitly. */
static uint64_t mix64 (uint64_t h, uint64_t v)
{
    h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2) ;
    h *= 0xff51afd7ed558ccdULL ;
    return h ^ (h >> 32) ;
}
static uint64_t hash_words (uint64_t h, const Long *a, Long count, Long max_samples)
{
    if (a == NULL || count <= 0) return mix64 (h, (uint64_t) count) ;
    Long stride = (max_samples > 0 && count > max_samples) ? (count + max_samples - 1) / max_samples : 1 ;
    if (stride == 1)
    {
        /* four independent chains (the multiply-xor chain of one accumulator is latency bound), folded in a
         * fixed order */
        uint64_t h0 = h, h1 = h ^ 0x9e3779b97f4a7c15ULL, h2 = h + 0x632be59bd9b4e019ULL, h3 = ~h ;
        Long i = 0 ;
        for ( ; i + 3 < count ; i += 4)
        {
            h0 = mix64 (h0, (uint64_t) a [i]) ;   h1 = mix64 (h1, (uint64_t) a [i+1]) ;
            h2 = mix64 (h2, (uint64_t) a [i+2]) ; h3 = mix64 (h3, (uint64_t) a [i+3]) ;
        }
        for ( ; i < count ; i++) h0 = mix64 (h0, (uint64_t) a [i]) ;
        h = mix64 (mix64 (mix64 (h0, h1), h2), h3) ;
        return mix64 (h, (uint64_t) count) ;
    }
    for (Long i = 0 ; i < count ; i += stride) h = mix64 (h, (uint64_t) a [i]) ;
    if (stride > 1) for (Long i = (count > 512) ? count - 512 : 0 ; i < count ; i++) h = mix64 (h, (uint64_t) a [i]) ;
    return mix64 (h, (uint64_t) count) ;
}
static uint64_t symbolic_key (const qr_symbolic *Q)
{
    const Long S = 65536 ;
    uint64_t h = 0x53544d5152423230ULL ;
    const Long sc [12] = { Q->m, Q->n, Q->anz, Q->nf, Q->maxfn, Q->rjsize, Q->hisize, Q->do_rank_detection,
        Q->keepH, Q->maxstack, Q->ntasks, Q->ns } ;
    for (int i = 0 ; i < 12 ; i++) h = mix64 (h, (uint64_t) sc [i]) ;
    h = hash_words (h, Q->Super, Q->nf + 1, 0) ;  h = hash_words (h, Q->Rp, Q->nf + 1, 0) ;
    h = hash_words (h, Q->Childp, Q->nf + 2, 0) ; h = hash_words (h, Q->Child, Q->nf + 1, 0) ;
    h = hash_words (h, Q->Hip, Q->nf + 1, 0) ;    h = hash_words (h, Q->Fm, Q->nf, 0) ;
    h = hash_words (h, Q->Cm, Q->nf, 0) ;
    h = hash_words (h, Q->Rj, Q->rjsize, S) ;     h = hash_words (h, Q->Sp, Q->m + 1, S) ;
    h = hash_words (h, Q->Sj, Q->anz, S) ;        h = hash_words (h, Q->PLinv, Q->m, S) ;
    h = hash_words (h, Q->Sleft, Q->n + 2, S) ;   h = hash_words (h, Q->Qfill, Q->Qfill ? Q->n : 0, S) ;
    return h ;
}

static double now_ms (void)
{
    struct timeval tv ;
    gettimeofday (&tv, NULL) ;
    return tv.tv_sec * 1e3 + tv.tv_usec * 1e-3 ;
}

static int map_status (int s)
{
    switch (s)
    {
        case STMQR_OK:                return SPARSE_OK ;
        case STMQR_ERR_OUT_OF_MEMORY: return SPARSE_OUT_OF_MEMORY ;
        case STMQR_ERR_TOO_LARGE:     return SPARSE_TOO_LARGE ;
        default:                      return SPARSE_INVALID ;
    }
}

static void report (sparse_common *cc, int s, const char *where)
{
    int st = map_status (s) ;
    fprintf (stderr, "stmqr_b200 %s: %s\n", where, g_handle ? stmqr_b200_last_error (g_handle) : "no device") ;
    SparseCore_error (st, __FILE__, __LINE__, "B200 numeric factorization failed", cc) ;
    if (cc->status >= SPARSE_OK) cc->status = st ;
}

/* ---- several GPUs behind the same entry point: STMQR_B200_DEVICES=0,1,2,3 ------------------------------------
 * One engine handle per listed device (a device may be listed twice: the handles then share it, which is how the
 * one-GPU test exercises this path).  Every front gets an owner GPU (stmqr_b200_map_fronts), the handles form a
 * peer group (contribution blocks move by cudaMemcpyPeerAsync between the etree levels, one host thread per
 * handle, csrc/multigpu.cuh), every GPU streams the packed R+H blocks of ITS fronts into ITS OWN host stack
 * while it factorizes: the returned qr_numeric has ns = #GPUs stacks, Rblock [f] points into the stack of
 * front f's owner -- any front-to-stack placement is legal for the consumers; the reference's
 * own task scheduler also produces several stacks (SparseQR_factorize.c:405-410).  The integer side is merged
 * on the devices and downloaded once. */
#define MAXDEV 16
static int g_ndev = -1 ;                    /* -1: STMQR_B200_DEVICES not parsed yet */
static int g_devs [MAXDEV] ;
static stmqr_handle g_hs [MAXDEV] ;
static void *g_group = NULL ;
static int32_t *g_owner = NULL ;
static Long g_owner_nf = 0 ;
static int g_multi_have_plan = 0 ;
static uint64_t g_multi_key = 0 ;

static void parse_devices (void)
{
    g_ndev = 0 ;
    const char *e = getenv ("STMQR_B200_DEVICES") ;
    if (!e) return ;
    while (*e && g_ndev < MAXDEV)
    {
        char *end ;
        long d = strtol (e, &end, 10) ;
        if (end == e) break ;
        g_devs [g_ndev++] = (int) d ;
        e = (*end == ',') ? end + 1 : end ;
    }
    if (g_ndev < 2) g_ndev = 0 ;            /* one device: the ordinary path (STMQR_B200_DEVICE) */
}

static void multi_shutdown (void)
{
    if (g_group) stmqr_b200_peer_group_destroy (g_group) ;
    g_group = NULL ;
    for (int i = 0 ; i < MAXDEV ; i++) { if (g_hs [i]) stmqr_b200_destroy (g_hs [i]) ; g_hs [i] = NULL ; }
    free (g_owner) ; g_owner = NULL ; g_owner_nf = 0 ;
    g_multi_have_plan = 0 ;
}

static qr_numeric *multi_fail (sparse_csc **Ahandle, Long freeA, qr_numeric **QRnum, Long *Roff, Long nf,
    sparse_common *cc, int s, const char *where, stmqr_handle h)
{
    if (s != STMQR_OK)
    {
        int st = map_status (s) ;
        fprintf (stderr, "stmqr_b200 (multi-GPU) %s: %s\n", where, h ? stmqr_b200_last_error (h) : "") ;
        SparseCore_error (st, __FILE__, __LINE__, "B200 numeric factorization failed", cc) ;
        if (cc->status >= SPARSE_OK) cc->status = st ;
    }
    if (Roff) SparseCore_free (nf, sizeof (Long), Roff, cc) ;
    if (QRnum && *QRnum) qr_freenum (QRnum, cc) ;
    if (freeA) SparseCore_free_sparse (Ahandle, cc) ;
    return (NULL) ;
}

static qr_numeric *qr_factorize_multi_gpu (sparse_csc **Ahandle, Long freeA, double tol, Long ntol,
    qr_symbolic *QRsym, sparse_common *cc)
{
    sparse_csc *A = *Ahandle ;
    const int verbose = getenv ("STMQR_B200_VERBOSE") != NULL ;
    const double t_start = now_ms () ;
    Long nf = QRsym->nf, m = QRsym->m, n = QRsym->n, rjsize = QRsym->rjsize, hisize = QRsym->hisize ;
    const int nd = g_ndev ;
    int s ;
    for (int i = 0 ; i < nd ; i++)
        if (g_hs [i] == NULL && (s = stmqr_b200_create (g_devs [i], &g_hs [i])) != STMQR_OK)
        {
            g_hs [i] = NULL ;
            return multi_fail (Ahandle, freeA, NULL, NULL, nf, cc, s, "create", NULL) ;
        }
    /* plan: the same symbolic plan on every GPU, the ownership map, the peer group */
    const uint64_t key = symbolic_key (QRsym) ;
    if (!(g_multi_have_plan && key == g_multi_key))
    {
        stmqr_symbolic_view v ;
        v.m = m ; v.n = n ; v.anz = QRsym->anz ; v.nf = nf ; v.maxfn = QRsym->maxfn ;
        v.rjsize = rjsize ; v.hisize = hisize ;
        v.do_rank_detection = QRsym->do_rank_detection ; v.keepH = QRsym->keepH ;
        v.Sp = QRsym->Sp ; v.Sj = QRsym->Sj ; v.Qfill = QRsym->Qfill ; v.PLinv = QRsym->PLinv ;
        v.Sleft = QRsym->Sleft ; v.Parent = QRsym->Parent ; v.Child = QRsym->Child ;
        v.Childp = QRsym->Childp ; v.Super = QRsym->Super ; v.Rp = QRsym->Rp ; v.Rj = QRsym->Rj ;
        v.Post = QRsym->Post ; v.Hip = QRsym->Hip ; v.Fm = QRsym->Fm ; v.Cm = QRsym->Cm ;
        g_multi_have_plan = 0 ;
        if (g_group) { stmqr_b200_peer_group_destroy (g_group) ; g_group = NULL ; }
        free (g_owner) ;
        g_owner = (int32_t *) malloc ((size_t) (nf > 0 ? nf : 1) * sizeof (int32_t)) ;
        g_owner_nf = nf ;
        if (!g_owner) return multi_fail (Ahandle, freeA, NULL, NULL, nf, cc, STMQR_ERR_OUT_OF_MEMORY, "owner map", NULL) ;
        if ((s = stmqr_b200_map_fronts (&v, nd, g_owner)) != STMQR_OK)
            return multi_fail (Ahandle, freeA, NULL, NULL, nf, cc, s, "map_fronts", NULL) ;
        for (int i = 0 ; i < nd ; i++)
        {
            if ((s = stmqr_b200_analyze (g_hs [i], &v)) != STMQR_OK) return multi_fail (Ahandle, freeA, NULL, NULL, nf, cc, s, "analyze", g_hs [i]) ;
            if ((s = stmqr_b200_set_ownership (g_hs [i], nd, i, g_owner)) != STMQR_OK) return multi_fail (Ahandle, freeA, NULL, NULL, nf, cc, s, "set_ownership", g_hs [i]) ;
        }
        if ((s = stmqr_b200_peer_group_create (g_hs, nd, &g_group)) != STMQR_OK)
            return multi_fail (Ahandle, freeA, NULL, NULL, nf, cc, s, "peer_group_create", g_hs [0]) ;
        g_multi_have_plan = 1 ;
        g_multi_key = key ;
    }
    const double t_plan = now_ms () ;
    SparseCore_allocate_work (0, (m > nf) ? m : nf, 0, cc) ;

    stmqr_csc_view Av ;
    Av.nrow = A->nrow ; Av.ncol = A->ncol ; Av.nzmax = A->nzmax ;
    Av.p = (const int64_t *) A->p ; Av.i = (const int64_t *) A->i ; Av.x = (const double *) A->x ;

    /* the numeric object: ns = #GPUs stacks, each allocated by the bound of its GPU's fronts and shrunk afterwards */
    qr_numeric *QRnum = (qr_numeric *) SparseCore_malloc (1, sizeof (qr_numeric), cc) ;
    if (cc->status < SPARSE_OK) return multi_fail (Ahandle, freeA, NULL, NULL, nf, cc, STMQR_OK, "", NULL) ;
    Long ns = nd ;
    QRnum->Rblock     = (double **) SparseCore_malloc (nf, sizeof (double *), cc) ;
    QRnum->Rdead      = (char *)    SparseCore_calloc (n,  sizeof (char), cc) ;
    QRnum->Stacks     = (double **) SparseCore_calloc (ns, sizeof (double *), cc) ;
    QRnum->Stack_size = (Long *)    SparseCore_calloc (ns, sizeof (Long), cc) ;
    QRnum->HStair = (Long *)   SparseCore_malloc (rjsize, sizeof (Long), cc) ;
    QRnum->HTau   = (double *) SparseCore_malloc (rjsize, sizeof (double), cc) ;
    QRnum->Hii    = (Long *)   SparseCore_malloc (hisize, sizeof (Long), cc) ;
    QRnum->Hm     = (Long *)   SparseCore_malloc (nf, sizeof (Long), cc) ;
    QRnum->Hr     = (Long *)   SparseCore_malloc (nf, sizeof (Long), cc) ;
    QRnum->HPinv  = (Long *)   SparseCore_malloc (m, sizeof (Long), cc) ;
    QRnum->n = n ; QRnum->m = m ; QRnum->nf = nf ;
    QRnum->rjsize = rjsize ; QRnum->hisize = hisize ; QRnum->keepH = QRsym->keepH ;
    QRnum->maxstack = QRsym->maxstack ;
    QRnum->ns = ns ; QRnum->ntasks = nd ;
    QRnum->maxfm = EMPTY ;
    QRnum->norm_E_fro = 0 ;
    Long *Roff = (Long *) SparseCore_malloc (nf, sizeof (Long), cc) ;
    int64_t cap [MAXDEV] ;
    for (int i = 0 ; i < nd && cc->status == SPARSE_OK ; i++)
    {
        cap [i] = 1 ;
        stmqr_b200_rh_bound (g_hs [i], &cap [i]) ;
        QRnum->Stack_size [i] = cap [i] ;
        QRnum->Stacks [i] = (double *) SparseCore_malloc (cap [i], sizeof (double), cc) ;
    }
    if (cc->status < SPARSE_OK) return multi_fail (Ahandle, freeA, &QRnum, Roff, nf, cc, STMQR_OK, "", NULL) ;

    /* upload A to every GPU (each builds the rows of S its fronts need from the whole matrix), start the
     * downloaders, factorize with one host thread per GPU, merge the integer side on the devices */
    for (int i = 0 ; i < nd ; i++)
        if ((s = stmqr_b200_upload_matrix (g_hs [i], &Av)) != STMQR_OK) return multi_fail (Ahandle, freeA, &QRnum, Roff, nf, cc, s, "upload_matrix", g_hs [i]) ;
    for (int i = 0 ; i < nd ; i++)
        if ((s = stmqr_b200_stream_begin (g_hs [i], QRnum->Stacks [i], cap [i])) != STMQR_OK)
        {
            for (int j = 0 ; j < i ; j++) stmqr_b200_stream_end (g_hs [j]) ;
            return multi_fail (Ahandle, freeA, &QRnum, Roff, nf, cc, s, "stream_begin", g_hs [i]) ;
        }
    stmqr_numeric_info infos [MAXDEV] ;
    s = stmqr_b200_factorize_multi_ex (g_group, tol, ntol, infos, 1) ;
    int s_end = STMQR_OK ;
    for (int i = 0 ; i < nd ; i++) { int e2 = stmqr_b200_stream_end (g_hs [i]) ; if (e2 != STMQR_OK && s_end == STMQR_OK) s_end = e2 ; }
    const double t_fact = now_ms () ;
    if (s != STMQR_OK || s_end != STMQR_OK)
    {
        stmqr_handle bad = g_hs [0] ;
        for (int i = 0 ; i < nd ; i++) if (stmqr_b200_last_error (g_hs [i]) [0]) { bad = g_hs [i] ; break ; }
        return multi_fail (Ahandle, freeA, &QRnum, Roff, nf, cc, (s != STMQR_OK) ? s : s_end, "factorize", bad) ;
    }
    if (freeA) SparseCore_free_sparse (Ahandle, cc) ;
    for (int i = 0 ; i < nd ; i++)
    {
        Long stacksize = (infos [i].rh_size > 0) ? infos [i].rh_size : 1 ;
        size_t cur = (size_t) cap [i] ;
        QRnum->Stacks [i] = (double *) SparseCore_realloc (stacksize, sizeof (double), QRnum->Stacks [i], &cur, cc) ;
        QRnum->Stack_size [i] = (Long) cur ;
    }
    /* the integer side is global on every GPU after the gather: take it from the first one */
    stmqr_numeric_view out ;
    out.stack = NULL ;
    out.Roff = (int64_t *) Roff ; out.Rdead = QRnum->Rdead ;
    out.HStair = (int64_t *) QRnum->HStair ; out.HTau = QRnum->HTau ;
    out.Hii = (int64_t *) QRnum->Hii ; out.Hm = (int64_t *) QRnum->Hm ;
    out.Hr = (int64_t *) QRnum->Hr ; out.HPinv = (int64_t *) QRnum->HPinv ;
    s = stmqr_b200_download (g_hs [0], &out) ;
    if (s != STMQR_OK || cc->status < SPARSE_OK)
        return multi_fail (Ahandle, 0, &QRnum, Roff, nf, cc, s, "download", g_hs [0]) ;
    for (Long f = 0 ; f < nf ; f++) QRnum->Rblock [f] = QRnum->Stacks [g_owner [f]] + Roff [f] ;
    SparseCore_free (nf, sizeof (Long), Roff, cc) ;
    if (verbose)
        fprintf (stderr, "stmqr_b200 qr_factorize on %d GPUs: plan %.1f ms, upload+numeric %.1f ms, download of the integer "
            "side %.1f ms\n", nd, t_plan - t_start, t_fact - t_plan, now_ms () - t_fact) ;
    QRnum->rank = infos [0].rank ;
    QRnum->rank1 = infos [0].rank1 ;
    QRnum->maxfrank = infos [0].maxfrank ;
    QRnum->maxfm = infos [0].maxfm ;
    cc->SPQR_flopcount = infos [0].flops ;
    return (QRnum) ;
}

qr_numeric *stmqr_b200_qr_factorize (sparse_csc **Ahandle, Long freeA, double tol, Long ntol,
    qr_symbolic *QRsym, sparse_common *cc)
{
    if (QRsym == NULL)
    {
        if (freeA) SparseCore_free_sparse (Ahandle, cc) ;
        return (NULL) ;
    }
    if (g_ndev < 0) parse_devices () ;
    if (g_ndev >= 2) return qr_factorize_multi_gpu (Ahandle, freeA, tol, ntol, QRsym, cc) ;
    sparse_csc *A = *Ahandle ;
    const int verbose = getenv ("STMQR_B200_VERBOSE") != NULL ;
    double t_start = now_ms (), t_plan, t_fact, t_alloc ;
    Long nf = QRsym->nf, m = QRsym->m, n = QRsym->n, rjsize = QRsym->rjsize,
        hisize = QRsym->hisize ;
    int s ;

    if (g_handle == NULL)
    {
        int dev = 0 ;
        const char *e = getenv ("STMQR_B200_DEVICE") ;
        if (e) dev = atoi (e) ;
        s = stmqr_b200_create (dev, &g_handle) ;
        if (s != STMQR_OK)
        {
            g_handle = NULL ;
            if (freeA) SparseCore_free_sparse (Ahandle, cc) ;
            report (cc, s, "create") ;
            return (NULL) ;
        }
    }

    /* plan (level sets, maps, arenas): depends only on the symbolic object.  It is kept across calls
     * and reused when the content key of the symbolic object is unchanged (a refactorization loop over
     * same-pattern matrices, SparseQR.c:349,371, pays for it once). */
    const char *cache_env = getenv ("STMQR_B200_CACHE_PLAN") ;
    const int use_cache = !(cache_env && cache_env [0] == '0') ;
    const uint64_t key = use_cache ? symbolic_key (QRsym) : 0 ;
    if (!(use_cache && g_have_plan && key == g_planned_key))
    {
        stmqr_symbolic_view v ;
        v.m = m ; v.n = n ; v.anz = QRsym->anz ; v.nf = nf ; v.maxfn = QRsym->maxfn ;
        v.rjsize = rjsize ; v.hisize = hisize ;
        v.do_rank_detection = QRsym->do_rank_detection ; v.keepH = QRsym->keepH ;
        v.Sp = QRsym->Sp ; v.Sj = QRsym->Sj ; v.Qfill = QRsym->Qfill ; v.PLinv = QRsym->PLinv ;
        v.Sleft = QRsym->Sleft ; v.Parent = QRsym->Parent ; v.Child = QRsym->Child ;
        v.Childp = QRsym->Childp ; v.Super = QRsym->Super ; v.Rp = QRsym->Rp ; v.Rj = QRsym->Rj ;
        v.Post = QRsym->Post ; v.Hip = QRsym->Hip ; v.Fm = QRsym->Fm ; v.Cm = QRsym->Cm ;
        g_have_plan = 0 ;
        s = stmqr_b200_analyze (g_handle, &v) ;
        if (s != STMQR_OK)
        {
            if (freeA) SparseCore_free_sparse (Ahandle, cc) ;
            report (cc, s, "analyze") ;
            return (NULL) ;
        }
        g_have_plan = 1 ;
        g_planned_key = key ;
    }

    t_plan = now_ms () ;
    /* the reference uses cc->Iwork (size max(m,nf)) as scratch and callers rely on it existing */
    SparseCore_allocate_work (0, (m > nf) ? m : nf, 0, cc) ;

    stmqr_csc_view Av ;
    Av.nrow = A->nrow ; Av.ncol = A->ncol ; Av.nzmax = A->nzmax ;
    Av.p = (const int64_t *) A->p ; Av.i = (const int64_t *) A->i ; Av.x = (const double *) A->x ;
    stmqr_numeric_info info ;

    /* ---- allocate the numeric object exactly as qr_freenum expects (:339-375) ----
     * The R+H stack is allocated by its symbolic bound and shrunk afterwards, as the reference does
     * with its own stacks (:405-410, :560-660): the engine can then copy the blocks of every finished
     * etree level into it while the next levels are still being factorized. */
    const int streamed = (getenv ("STMQR_B200_NO_STREAM") == NULL) ;
    int64_t cap = 0 ;
    if (streamed && stmqr_b200_rh_bound (g_handle, &cap) != STMQR_OK) cap = 0 ;
    qr_numeric *QRnum = (qr_numeric *) SparseCore_malloc (1, sizeof (qr_numeric), cc) ;
    if (cc->status < SPARSE_OK)
    {
        if (freeA) SparseCore_free_sparse (Ahandle, cc) ;
        return (NULL) ;
    }
    Long ns = 1 ;
    QRnum->Rblock     = (double **) SparseCore_malloc (nf, sizeof (double *), cc) ;
    QRnum->Rdead      = (char *)    SparseCore_calloc (n,  sizeof (char), cc) ;
    QRnum->Stacks     = (double **) SparseCore_calloc (ns, sizeof (double *), cc) ;
    QRnum->Stack_size = (Long *)    SparseCore_calloc (ns, sizeof (Long), cc) ;
    QRnum->HStair = (Long *)   SparseCore_malloc (rjsize, sizeof (Long), cc) ;
    QRnum->HTau   = (double *) SparseCore_malloc (rjsize, sizeof (double), cc) ;
    QRnum->Hii    = (Long *)   SparseCore_malloc (hisize, sizeof (Long), cc) ;
    QRnum->Hm     = (Long *)   SparseCore_malloc (nf, sizeof (Long), cc) ;
    QRnum->Hr     = (Long *)   SparseCore_malloc (nf, sizeof (Long), cc) ;
    QRnum->HPinv  = (Long *)   SparseCore_malloc (m, sizeof (Long), cc) ;
    QRnum->n = n ; QRnum->m = m ; QRnum->nf = nf ;
    QRnum->rjsize = rjsize ; QRnum->hisize = hisize ; QRnum->keepH = QRsym->keepH ;
    QRnum->maxstack = QRsym->maxstack ;
    QRnum->ns = ns ; QRnum->ntasks = 1 ;
    QRnum->maxfm = EMPTY ;
    QRnum->norm_E_fro = 0 ;
    Long *Roff = (Long *) SparseCore_malloc (nf, sizeof (Long), cc) ;
    if (cc->status == SPARSE_OK && streamed && cap > 0)
    {
        QRnum->Stack_size [0] = cap ;
        QRnum->Stacks [0] = (double *) SparseCore_malloc (cap, sizeof (double), cc) ;
    }
    if (cc->status < SPARSE_OK)
    {
        if (Roff) SparseCore_free (nf, sizeof (Long), Roff, cc) ;
        qr_freenum (&QRnum, cc) ;
        if (freeA) SparseCore_free_sparse (Ahandle, cc) ;
        return (NULL) ;
    }
#ifdef MADV_HUGEPAGE
#define STACK_HUGEPAGES(ptr, doubles) do { \
        /* the stack is written exactly once, front to back, by the download: ask for huge pages on its \
         * 2 MB-aligned interior so that the first touch takes ~500x fewer page faults */ \
        if ((ptr) && (size_t) (doubles) * sizeof (double) >= ((size_t) 8 << 20)) \
        { \
            size_t a_ = ((size_t) (ptr) + ((size_t) 2 << 20) - 1) & ~(((size_t) 2 << 20) - 1) ; \
            size_t e_ = ((size_t) (ptr) + (size_t) (doubles) * sizeof (double)) & ~(((size_t) 2 << 20) - 1) ; \
            if (e_ > a_) madvise ((void *) a_, e_ - a_, MADV_HUGEPAGE) ; \
        } } while (0)
#else
#define STACK_HUGEPAGES(ptr, doubles) do { } while (0)
#endif

    /* the big integer / tau arrays are also written once, front to back, by the download */
    STACK_HUGEPAGES (QRnum->HStair, rjsize) ;
    STACK_HUGEPAGES (QRnum->HTau, rjsize) ;
    STACK_HUGEPAGES (QRnum->Hii, hisize) ;
    STACK_HUGEPAGES (QRnum->HPinv, m) ;
    if (streamed && cap > 0)
    {
        STACK_HUGEPAGES (QRnum->Stacks [0], cap) ;
        s = stmqr_b200_factorize_streamed (g_handle, &Av, tol, ntol, QRnum->Stacks [0], cap, &info) ;
    }
    else
    {
        s = stmqr_b200_factorize (g_handle, &Av, tol, ntol, &info) ;
    }
    t_fact = now_ms () ;
    if (freeA) SparseCore_free_sparse (Ahandle, cc) ;           /* A is no longer needed (:324) */
    if (s != STMQR_OK || cc->status < SPARSE_OK)
    {
        if (s != STMQR_OK) report (cc, s, "factorize") ;
        SparseCore_free (nf, sizeof (Long), Roff, cc) ;
        qr_freenum (&QRnum, cc) ;
        return (NULL) ;
    }
    {
        Long stacksize = (info.rh_size > 0) ? info.rh_size : 1 ;
        if (streamed && cap > 0)
        {
            /* shrink the stack to what the factorization produced (:640-657) */
            size_t cur = (size_t) cap ;
            QRnum->Stacks [0] = (double *) SparseCore_realloc (stacksize, sizeof (double), QRnum->Stacks [0], &cur, cc) ;
            QRnum->Stack_size [0] = (Long) cur ;
        }
        else
        {
            QRnum->Stack_size [0] = stacksize ;
            QRnum->Stacks [0] = (double *) SparseCore_malloc (stacksize, sizeof (double), cc) ;
            STACK_HUGEPAGES (QRnum->Stacks [0], stacksize) ;
        }
    }
    t_alloc = now_ms () ;
    if (cc->status < SPARSE_OK)
    {
        SparseCore_free (nf, sizeof (Long), Roff, cc) ;
        qr_freenum (&QRnum, cc) ;
        return (NULL) ;
    }

    stmqr_numeric_view out ;
    out.stack = (streamed && cap > 0) ? NULL : QRnum->Stacks [0] ;
    out.Roff = (int64_t *) Roff ; out.Rdead = QRnum->Rdead ;
    out.HStair = (int64_t *) QRnum->HStair ; out.HTau = QRnum->HTau ;
    out.Hii = (int64_t *) QRnum->Hii ; out.Hm = (int64_t *) QRnum->Hm ;
    out.Hr = (int64_t *) QRnum->Hr ; out.HPinv = (int64_t *) QRnum->HPinv ;
    s = stmqr_b200_download (g_handle, &out) ;
    if (s != STMQR_OK)
    {
        SparseCore_free (nf, sizeof (Long), Roff, cc) ;
        qr_freenum (&QRnum, cc) ;
        report (cc, s, "download") ;
        return (NULL) ;
    }
    for (Long f = 0 ; f < nf ; f++) QRnum->Rblock [f] = QRnum->Stacks [0] + Roff [f] ;
    SparseCore_free (nf, sizeof (Long), Roff, cc) ;

    if (verbose)
        fprintf (stderr, "stmqr_b200 qr_factorize: plan %.1f ms, upload+numeric %.1f ms, host alloc %.1f ms, "
            "download %.1f ms (R+H %.1f MB)\n", t_plan - t_start, t_fact - t_plan, t_alloc - t_fact,
            now_ms () - t_alloc, info.rh_size * 8e-6) ;
    QRnum->rank = info.rank ;
    QRnum->rank1 = info.rank1 ;
    QRnum->maxfrank = info.maxfrank ;
    QRnum->maxfm = info.maxfm ;
    cc->SPQR_flopcount = info.flops ;           /* the reference's count (:504,:1571) */
    return (QRnum) ;
}

/* The reference's symbol.  When this library is linked (or preloaded) in front of the reference's
 * SparseQR_factorize.o, SparseQR() (SparseQR.c:349,371) and SparseLQ() call the B200 engine. */
qr_numeric *qr_factorize (sparse_csc **Ahandle, Long freeA, double tol, Long ntol,
    qr_symbolic *QRsym, sparse_common *cc)
{
    return stmqr_b200_qr_factorize (Ahandle, freeA, tol, ntol, QRsym, cc) ;
}

/* Release the device (optional; for drivers that want a clean exit). */
void stmqr_b200_dropin_shutdown (void)
{
    if (g_handle) stmqr_b200_destroy (g_handle) ;
    g_handle = NULL ;
    g_have_plan = 0 ;
    multi_shutdown () ;
    g_ndev = -1 ;                           /* STMQR_B200_DEVICES is read again by the next call */
}

/* Forget the cached device plan: the next qr_factorize re-plans whatever its symbolic object is. */
void stmqr_b200_dropin_invalidate_plan (void)
{
    g_have_plan = 0 ;
}
