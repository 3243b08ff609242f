/* oracle/stmqr_oracle.c -- TEST INFRASTRUCTURE ONLY (see stmqr_oracle.h).
 *
 * Single-threaded CPU restatement of the reference's numeric multifrontal QR.
 * Every routine cites the reference code it follows
 * (STMMQR/src/qr/SparseQR_factorize.c unless noted).  Storage is deliberately
 * naive (one heap block per front) -- this is a checker, not a fast path.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include "stmqr_oracle.h"

typedef int64_t Int ;
#define IMAX(a,b) (((a) > (b)) ? (a) : (b))
#define IMIN(a,b) (((a) < (b)) ? (a) : (b))

/* ------------------------------------------------------------------------- */
/* S = A(P,Q) values in row form: qr_stranspose2, SparseQR_factorize.c:755-785 */
/* ------------------------------------------------------------------------- */
void stmqr_oracle_stranspose2 (const stmqr_csc_view *A, const Int *Qfill,
    const Int *Sp, const Int *PLinv, double *Sx, Int *W)
{
    Int m = A->nrow, n = A->ncol ;
    for (Int r = 0 ; r < m ; r++) W [r] = Sp [r] ;
    for (Int k = 0 ; k < n ; k++)
    {
        Int j = Qfill ? Qfill [k] : k ;
        for (Int p = A->p [j] ; p < A->p [j+1] ; p++)
        {
            Int r = PLinv [A->i [p]] ;
            Sx [W [r]++] = A->x [p] ;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* LAPACK dlarfg (call site :1320): returns tau, overwrites alpha with beta   */
/* and x with v(1:).  SURVEY.md Appendix B.                                   */
/* ------------------------------------------------------------------------- */
static double nrm2 (Int n, const double *x)
{
    /* scaled 2-norm, as reference BLAS dnrm2 */
    double scale = 0, ssq = 1 ;
    for (Int i = 0 ; i < n ; i++)
    {
        if (x [i] != 0)
        {
            double a = fabs (x [i]) ;
            if (scale < a) { double r = scale / a ; ssq = 1 + ssq * r * r ; scale = a ; }
            else { double r = a / scale ; ssq += r * r ; }
        }
    }
    return scale * sqrt (ssq) ;
}

double stmqr_oracle_larfg (Int n, double *alpha, double *x)
{
    if (n <= 1) return 0 ;
    double xnorm = nrm2 (n-1, x) ;
    if (xnorm == 0) return 0 ;
    double a = *alpha ;
    double beta = -copysign (hypot (a, xnorm), a) ;
    const double safmin = DBL_MIN / (DBL_EPSILON * 0.5) ;
    const double rsafmn = 1.0 / safmin ;
    int knt = 0 ;
    if (fabs (beta) < safmin)
    {
        do
        {
            knt++ ;
            for (Int i = 0 ; i < n-1 ; i++) x [i] *= rsafmn ;
            beta *= rsafmn ;
            a *= rsafmn ;
        } while (fabs (beta) < safmin && knt < 20) ;
        xnorm = nrm2 (n-1, x) ;
        beta = -copysign (hypot (a, xnorm), a) ;
    }
    double tau = (beta - a) / beta ;
    double s = 1.0 / (a - beta) ;
    for (Int i = 0 ; i < n-1 ; i++) x [i] *= s ;
    for (int j = 0 ; j < knt ; j++) beta *= safmin ;
    *alpha = beta ;
    return tau ;
}

/* dlarf('L') with v[0] treated as 1 (qr_private_apply1 :1359-1381) */
static void apply1 (Int m, Int n, Int ldc, const double *V, double tau, double *C)
{
    if (m <= 0 || n <= 0 || tau == 0) return ;
    for (Int j = 0 ; j < n ; j++)
    {
        double *c = C + j*ldc ;
        double w = c [0] ;
        for (Int i = 1 ; i < m ; i++) w += V [i] * c [i] ;
        w *= tau ;
        c [0] -= w ;
        for (Int i = 1 ; i < m ; i++) c [i] -= V [i] * w ;
    }
}

/* dlarft('F','C') + dlarfb('L','T','F','C'): qr_larftb method QR_QTX :1876-1882.
 * V is m-by-k unit lower trapezoidal (diagonal / upper part never read), C is m-by-n.
 * W: workspace k*k + k*n. */
static void larftb (Int m, Int n, Int k, Int ldc, Int ldv, const double *V, const double *Tau,
    double *C, double *W)
{
    if (m <= 0 || n <= 0 || k <= 0) return ;
    double *T = W ;             /* k-by-k, ld k (:1869) */
    double *Wk = W + k*k ;      /* k-by-n, ld k */
    for (Int i = 0 ; i < k ; i++)
    {
        for (Int j = 0 ; j < k ; j++) T [j + i*k] = 0 ;
        if (Tau [i] == 0) continue ;
        /* t = -tau_i * V(i:m-1,0:i-1)' * V(i:m-1,i) */
        for (Int j = 0 ; j < i ; j++)
        {
            double s = V [i + j*ldv] ;              /* V(i,i) = 1 */
            for (Int r = i+1 ; r < m ; r++) s += V [r + j*ldv] * V [r + i*ldv] ;
            T [j + i*k] = -Tau [i] * s ;
        }
        /* t = T(0:i-1,0:i-1) * t  (upper triangular) */
        for (Int j = 0 ; j < i ; j++)
        {
            double s = 0 ;
            for (Int l = j ; l < i ; l++) s += T [j + l*k] * T [l + i*k] ;
            T [j + i*k] = s ;
        }
        T [i + i*k] = Tau [i] ;
    }
    for (Int c = 0 ; c < n ; c++)
    {
        double *cc = C + c*ldc ;
        double *w = Wk + c*k ;
        /* w = V' * C(:,c) */
        for (Int j = 0 ; j < k ; j++)
        {
            double s = cc [j] ;
            for (Int r = j+1 ; r < m ; r++) s += V [r + j*ldv] * cc [r] ;
            w [j] = s ;
        }
        /* w = T' * w   (row vector times upper-triangular T) */
        for (Int j = k-1 ; j >= 0 ; j--)
        {
            double s = 0 ;
            for (Int l = 0 ; l <= j ; l++) s += T [l + j*k] * w [l] ;
            w [j] = s ;
        }
        /* C(:,c) -= V * w */
        for (Int j = 0 ; j < k ; j++)
        {
            double wj = w [j] ;
            if (wj == 0) continue ;
            cc [j] -= wj ;
            for (Int r = j+1 ; r < m ; r++) cc [r] -= V [r + j*ldv] * wj ;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* qr_front :1383-1618 -- staircase blocked Householder QR with rank detection */
/* ------------------------------------------------------------------------- */
Int stmqr_oracle_front (Int m, Int n, Int npiv, double tol, Int ntol,
    Int fchunk, Int small, Int minchunk_g, Int minchunk_ratio,
    double *F, Int *Stair, char *Rdead, double *Tau, double *W,
    double *flops, double *min_margin)
{
    npiv = IMIN (n, IMAX (0, npiv)) ;
    fchunk = IMAX (fchunk, 1) ;
    Int minchunk = IMAX (minchunk_g, minchunk_ratio > 0 ? fchunk / minchunk_ratio : fchunk) ;
    Int rank = IMIN (m, npiv) ;
    ntol = IMIN (ntol, npiv) ;
    Int g = 0, g1 = 0, k1 = 0, k2 = 0, nv = 0, vzeros = 0, t = 0 ;
    const double *V = F ;

    for (Int k = 0 ; k < n ; k++)
    {
        Int t0 = t ;
        t = Stair [k] ;
        if (g >= m)
        {
            /* no rows left (:1444-1458) */
            for ( ; k < npiv ; k++) { Rdead [k] = 1 ; Stair [k] = 0 ; Tau [k] = 0 ; }
            for ( ; k < n ; k++) { Stair [k] = m ; Tau [k] = 0 ; }
            return rank ;
        }
        t = IMAX (g+1, t) ;
        Stair [k] = t ;

        /* staircase grew a lot: flush the pending block (:1467-1483) */
        vzeros += nv * (t - t0) ;
        if (nv >= minchunk)
        {
            Int vsize = (nv*(nv+1))/2 + nv*(t-g1-nv) ;
            if (vzeros > IMAX (16, vsize/2))
            {
                larftb (t0-g1, n-k2, nv, m, m, V, Tau + k1, F + g1 + k2*m, W) ;
                nv = 0 ; vzeros = 0 ;
            }
        }

        double *x = F + g + k*m ;
        double tau = stmqr_oracle_larfg (t-g, x, x+1) ;
        double wk = fabs (x [0]) ;
        if (k < ntol)
        {
            if (tol > 0 && min_margin)
            {
                double mg = fabs (wk - tol) / tol ;
                if (mg < *min_margin) *min_margin = mg ;
            }
        }
        if (k < ntol && wk <= tol)
        {
            /* dead pivot column (:1495-1543) */
            for (Int i = g ; i < m ; i++) F [i + k*m] = 0 ;
            Stair [k] = 0 ; Tau [k] = 0 ; Rdead [k] = 1 ;
            if (nv > 0)
            {
                larftb (t0-g1, n-k2, nv, m, m, V, Tau + k1, F + g1 + k2*m, W) ;
                nv = 0 ; vzeros = 0 ;
            }
        }
        else
        {
            Tau [k] = tau ;
            if (nv == 0)
            {
                /* start a new panel (:1553-1568) */
                g1 = g ; k1 = k ; k2 = IMIN (n, k+fchunk) ;
                V = F + g1 + k1*m ;
                Int mleft = m-g1, nleft = n-k1 ;
                if (mleft * (nleft-(fchunk+4)) < small || mleft <= fchunk/2 || fchunk <= 1)
                {
                    k2 = n ;
                }
            }
            nv++ ;
            if (flops) *flops += (double) ((t-g) * (3 + 4*(n-k-1))) ;
            apply1 (t-g, k2-k-1, m, x, tau, x + m) ;
            g++ ;
            if (k == k2-1 || g == m)
            {
                larftb (t-g1, n-k2, nv, m, m, V, Tau + k1, F + g1 + k2*m, W) ;
                nv = 0 ; vzeros = 0 ;
            }
        }
        if (k == npiv-1) rank = g ;
    }
    return rank ;
}

/* ------------------------------------------------------------------------- */
/* the whole numeric phase: qr_factorize :222-749 with one task (qr_kernel :791) */
/* ------------------------------------------------------------------------- */
stmqr_oracle_result *stmqr_oracle_factorize (const stmqr_symbolic_view *sym,
    const stmqr_csc_view *A, double tol, Int ntol,
    Int fchunk, Int small, Int minchunk, Int minchunk_ratio, int capture)
{
    Int m = sym->m, n = sym->n, nf = sym->nf, anz = sym->anz ;
    const Int *Super = sym->Super, *Rp = sym->Rp, *Rj = sym->Rj, *Sleft = sym->Sleft,
        *Sp = sym->Sp, *Sj = sym->Sj, *Child = sym->Child, *Childp = sym->Childp,
        *Post = sym->Post, *Hip = sym->Hip ;
    if (!sym->do_rank_detection) tol = -1 ;                 /* :285-289 */
    fchunk = IMIN (m, fchunk) ;                             /* :304 */

    stmqr_oracle_result *R = (stmqr_oracle_result *) calloc (1, sizeof (*R)) ;
    R->nf = nf ; R->n = n ; R->m = m ;
    R->Rdead  = (char *)   calloc (IMAX (n,1), 1) ;
    R->HStair = (Int *)    calloc (IMAX (sym->rjsize,1), sizeof (Int)) ;
    R->HTau   = (double *) calloc (IMAX (sym->rjsize,1), sizeof (double)) ;
    R->Hii    = (Int *)    calloc (IMAX (sym->hisize,1), sizeof (Int)) ;
    R->Hii_raw= (Int *)    calloc (IMAX (sym->hisize,1), sizeof (Int)) ;
    R->Hm     = (Int *)    calloc (IMAX (nf,1), sizeof (Int)) ;
    R->Hr     = (Int *)    calloc (IMAX (nf,1), sizeof (Int)) ;
    R->HPinv  = (Int *)    calloc (IMAX (m,1), sizeof (Int)) ;
    R->Cm     = (Int *)    calloc (IMAX (nf,1), sizeof (Int)) ;
    R->Roff   = (Int *)    calloc (IMAX (nf,1), sizeof (Int)) ;
    R->Sx     = (double *) calloc (IMAX (anz,1), sizeof (double)) ;
    R->min_tol_margin = HUGE_VAL ;
    double **Cblock = (double **) calloc (IMAX (nf,1), sizeof (double *)) ;
    double **Rblock = (double **) calloc (IMAX (nf,1), sizeof (double *)) ;
    Int *Rsize = (Int *) calloc (IMAX (nf,1), sizeof (Int)) ;
    if (capture)
    {
        R->Fasm = (double **) calloc (IMAX (nf,1), sizeof (double *)) ;
        R->Ffac = (double **) calloc (IMAX (nf,1), sizeof (double *)) ;
        R->Cblk = (double **) calloc (IMAX (nf,1), sizeof (double *)) ;
    }
    Int *Wi = (Int *) calloc (IMAX (IMAX (m,nf),1), sizeof (Int)) ;
    Int *Fmap = (Int *) calloc (IMAX (n,1), sizeof (Int)) ;
    Int *Cmap = (Int *) calloc (IMAX (sym->maxfn,1), sizeof (Int)) ;
    double *W = (double *) calloc (IMAX ((fchunk + 1) * (sym->maxfn + fchunk) + 1, 1), sizeof (double)) ;

    stmqr_oracle_stranspose2 (A, sym->Qfill, Sp, sym->PLinv, R->Sx, Wi) ;
    const double *Sx = R->Sx ;
    Int *Hii = R->Hii ;
    Int sumfrank = 0, maxfrank = 1 ;

    for (Int kf = 0 ; kf < nf ; kf++)
    {
        Int f = Post [kf] ;
        Int col1 = Super [f], fp = Super [f+1] - col1 ;
        Int p1 = Rp [f], fn = Rp [f+1] - p1 ;
        Int *Stair = R->HStair + p1 ;
        double *Tau = R->HTau + p1 ;

        /* ---- qr_fsize :1066-1145 ---- */
        for (Int j = 0 ; j < fn ; j++) Fmap [Rj [p1+j]] = j ;
        for (Int j = 0 ; j < fn ; j++)
            Stair [j] = (j < fp) ? (Sleft [col1+j+1] - Sleft [col1+j]) : 0 ;
        for (Int p = Childp [f] ; p < Childp [f+1] ; p++)
        {
            Int c = Child [p] ;
            Int pc = Rp [c] + (Super [c+1] - Super [c]) ;
            for (Int ci = 0 ; ci < R->Cm [c] ; ci++) Stair [Fmap [Rj [pc+ci]]]++ ;
        }
        Int fm = 0 ;
        for (Int j = 0 ; j < fn ; j++) { Int t = fm ; fm += Stair [j] ; Stair [j] = t ; }
        R->Hm [f] = fm ;

        /* ---- qr_assemble :1151-1285 ---- */
        double *F = (double *) calloc (IMAX (fm*fn,1), sizeof (double)) ;
        Int *Hi = Hii + Hip [f] ;
        for (Int k = 0 ; k < fp ; k++)
        {
            for (Int row = Sleft [col1+k] ; row < Sleft [col1+k+1] ; row++)
            {
                Int i = Stair [k]++ ;
                for (Int p = Sp [row] ; p < Sp [row+1] ; p++)
                    F [i + Fmap [Sj [p]] * fm] = Sx [p] ;
                Hi [i] = row ;
            }
        }
        for (Int p = Childp [f] ; p < Childp [f+1] ; p++)
        {
            Int c = Child [p] ;
            Int fpc = Super [c+1] - Super [c] ;
            Int pc = Rp [c] + fpc ;
            Int cn = (Rp [c+1] - Rp [c]) - fpc ;
            Int cm = R->Cm [c] ;
            const double *C = Cblock [c] ;
            const Int *Hichild = Hii + Hip [c] + R->Hr [c] ;
            for (Int ci = 0 ; ci < cm ; ci++)
            {
                Int i = Stair [Fmap [Rj [pc+ci]]]++ ;
                Cmap [ci] = i ;
                Hi [i] = Hichild [ci] ;
            }
            for (Int cj = 0 ; cj < cn ; cj++)
            {
                double *Fj = F + fm * Fmap [Rj [pc+cj]] ;
                Int len = IMIN (cj+1, cm) ;
                for (Int ci = 0 ; ci < len ; ci++) Fj [Cmap [ci]] = *(C++) ;
            }
            free (Cblock [c]) ; Cblock [c] = NULL ;        /* released after assembly (:925-933) */
        }
        if (capture)
        {
            R->Fasm [f] = (double *) malloc (IMAX (fm*fn,1) * sizeof (double)) ;
            memcpy (R->Fasm [f], F, fm*fn*sizeof (double)) ;
        }

        /* ---- qr_front :1383 ---- */
        Int frank = stmqr_oracle_front (fm, fn, fp, tol, ntol - col1, fchunk, small,
            minchunk, minchunk_ratio, F, Stair, R->Rdead + col1, Tau, W,
            &R->flops, &R->min_tol_margin) ;
        sumfrank += frank ;
        maxfrank = IMAX (maxfrank, frank) ;
        if (capture)
        {
            R->Ffac [f] = (double *) malloc (IMAX (fm*fn,1) * sizeof (double)) ;
            memcpy (R->Ffac [f], F, fm*fn*sizeof (double)) ;
        }

        /* ---- qr_fcsize :1623, qr_cpack :1639-1685 ---- */
        Int cn = fn - fp ;
        Int cm = IMIN (fm - frank, cn) ;
        if (cm <= 0 || cn <= 0) cm = 0 ;
        Int csize = (cm*(cm+1))/2 + cm*(cn-cm) ;
        double *C = (double *) malloc (IMAX (csize,1) * sizeof (double)) ;
        {
            double *cp = C ;
            const double *Fc = F + frank + fp*fm ;
            for (Int k = 0 ; k < cn && cm > 0 ; k++)
            {
                Int len = IMIN (k+1, cm) ;
                for (Int i = 0 ; i < len ; i++) *(cp++) = Fc [i] ;
                Fc += fm ;
            }
        }
        Cblock [f] = C ;
        R->Cm [f] = cm ;
        if (capture)
        {
            R->Cblk [f] = (double *) malloc (IMAX (csize,1) * sizeof (double)) ;
            memcpy (R->Cblk [f], C, csize*sizeof (double)) ;
        }

        /* ---- qr_rhpack :1691-1784 (keepH) ---- */
        double *Rb = (double *) malloc (IMAX (fm*fn,1) * sizeof (double)) ;
        double *rp = Rb ;
        Int rm = 0 ;
        if (fm > 0 && fn > 0)
        {
            const double *Fk = F ;
            Int k ;
            for (k = 0 ; k < fp ; k++)
            {
                Int t = Stair [k] ;
                if (t == 0) t = rm ; else if (rm < fm) rm++ ;
                for (Int i = 0 ; i < t ; i++) *(rp++) = Fk [i] ;
                Fk += fm ;
            }
            Int h = rm ;
            for ( ; k < fn ; k++)
            {
                for (Int i = 0 ; i < rm ; i++) *(rp++) = Fk [i] ;
                Int t = Stair [k] ;
                h = IMIN (h+1, fm) ;
                for (Int i = h ; i < t ; i++) *(rp++) = Fk [i] ;
                Fk += fm ;
            }
        }
        R->Hr [f] = rm ;
        Rblock [f] = Rb ;
        Rsize [f] = rp - Rb ;
        free (F) ;
    }

    /* concatenate the R+H blocks in postorder into one stack (any placement is legal:
     * consumers only use Rblock[f]) */
    Int tot = 0 ;
    for (Int kf = 0 ; kf < nf ; kf++) { Int f = Post [kf] ; R->Roff [f] = tot ; tot += Rsize [f] ; }
    R->rh_size = tot ;
    R->stack = (double *) malloc (IMAX (tot,1) * sizeof (double)) ;
    for (Int f = 0 ; f < nf ; f++)
    {
        memcpy (R->stack + R->Roff [f], Rblock [f], Rsize [f] * sizeof (double)) ;
        free (Rblock [f]) ;
        free (Cblock [f]) ;
    }
    R->rank = sumfrank ;
    R->maxfrank = maxfrank ;
    memcpy (R->Hii_raw, Hii, sym->hisize * sizeof (Int)) ;

    /* ---- qr_hpinv :991-1060 ---- */
    {
        Int row1 = 0, row2 = m, maxfm = 0 ;
        for (Int i = Sleft [n] ; i < m ; i++) Wi [i] = --row2 ;
        for (Int f = 0 ; f < nf ; f++)
        {
            Int *Hi = Hii + Hip [f] ;
            Int rm = R->Hr [f], fm = R->Hm [f] ;
            for (Int i = 0 ; i < rm ; i++) Wi [Hi [i]] = row1++ ;
            Int cn = (Rp [f+1] - Rp [f]) - (Super [f+1] - Super [f]) ;
            Int cm = IMIN (fm - rm, cn) ;
            maxfm = IMAX (maxfm, fm) ;
            for (Int i = fm-1 ; i >= rm + cm ; i--) Wi [Hi [i]] = --row2 ;
        }
        R->maxfm = maxfm ;
        for (Int i = 0 ; i < m ; i++) R->HPinv [i] = Wi [sym->PLinv [i]] ;
        for (Int f = 0 ; f < nf ; f++)
        {
            Int *Hi = Hii + Hip [f] ;
            for (Int i = 0 ; i < R->Hm [f] ; i++) Hi [i] = Wi [Hi [i]] ;
        }
    }

    /* rank1 :727-742 */
    if (ntol >= n) R->rank1 = R->rank ;
    else
    {
        Int r1 = 0 ;
        for (Int j = 0 ; j < ntol ; j++) if (!R->Rdead [j]) r1++ ;
        R->rank1 = r1 ;
    }

    free (Cblock) ; free (Rblock) ; free (Rsize) ; free (Wi) ; free (Fmap) ; free (Cmap) ; free (W) ;
    return R ;
}

void stmqr_oracle_free (stmqr_oracle_result *r)
{
    if (!r) return ;
    free (r->stack) ; free (r->Roff) ; free (r->Rdead) ; free (r->HStair) ; free (r->HTau) ;
    free (r->Hii) ; free (r->Hii_raw) ; free (r->Hm) ; free (r->Hr) ; free (r->HPinv) ;
    free (r->Cm) ; free (r->Sx) ;
    double **caps [3] = { r->Fasm, r->Ffac, r->Cblk } ;
    for (int a = 0 ; a < 3 ; a++)
    {
        if (caps [a])
        {
            for (Int f = 0 ; f < r->nf ; f++) free (caps [a][f]) ;
            free (caps [a]) ;
        }
    }
    free (r) ;
}

/* ------------------------------------------------------------------------- */
/* Consumer of the numeric object (the row next to the  */
/* path): Y = Q'X and Y = QX from the packed R+H blocks, for factorizations   */
/* without singletons.                                                        */
/*   qr_private_get_H_vectors   SparseQR.c:1455-1546  where the Householder   */
/*                              vectors of a front live inside its R+H block   */
/*   qr_private_Happly          :1706-1832  front order (forward for Q', back- */
/*                              ward for Q), rows Hii[Hip[f]+h ..] of vector h */
/*   QR_qmult                   :1838-2110  row permutation HPinv around it    */
/* The reference applies the vectors in panels of 32 through dlarft/dlarfb;    */
/* here one reflector at a time (same product, different rounding).            */
/* ------------------------------------------------------------------------- */
static Int oracle_get_H_vectors (Int f, const stmqr_symbolic_view *sym, const int64_t *HStair,
    const double *HTau, const int64_t *Hm, double *H_Tau, Int *H_start, Int *H_end)
{
    Int col1 = sym->Super [f], fp = sym->Super [f+1] - col1 ;
    Int pr = sym->Rp [f], fn = sym->Rp [f+1] - pr ;
    const int64_t *Stair = HStair + pr ;
    const double *Tau = HTau + pr ;
    Int fm = Hm [f], h = 0, nh = 0, p = 0, rm = 0 ;
    for (Int k = 0 ; k < fn && nh < fm ; k++)
    {
        Int t ;
        if (k < fp)
        {
            t = Stair [k] ;
            if (t == 0) { p += rm ; continue ; }        /* dead column: R part only (:1509-1513) */
            else if (rm < fm) rm++ ;
            h = rm ;
        }
        else
        {
            t = Stair [k] ;
            h = IMIN (h+1, fm) ;
        }
        p += rm ;
        H_Tau [nh] = Tau [k] ;
        H_start [nh] = p ;
        p += (t-h) ;
        H_end [nh] = p ;
        nh++ ;
        if (h == fm) break ;
    }
    return nh ;
}

/* method 0: Y = Q'X, method 1: Y = QX (QR_QTX / QR_QX, SparseQR_definitions.h); X, Y are m-by-nx,
 * column major, ld = m.  Returns 0, or -1 on a bad method / allocation failure. */
int stmqr_oracle_qmult (int method, const stmqr_symbolic_view *sym, const stmqr_numeric_view *num,
    int64_t nx, const double *X, double *Y)
{
    if (method != 0 && method != 1) return -1 ;
    Int m = sym->m, nf = sym->nf, maxfn = sym->maxfn ;
    double *H_Tau = (double *) malloc ((size_t) IMAX (maxfn, 1) * sizeof (double)) ;
    Int *H_start = (Int *) malloc ((size_t) IMAX (maxfn, 1) * sizeof (Int)) ;
    Int *H_end = (Int *) malloc ((size_t) IMAX (maxfn, 1) * sizeof (Int)) ;
    double *Z = (double *) malloc ((size_t) IMAX (m * nx, 1) * sizeof (double)) ;
    if (!H_Tau || !H_start || !H_end || !Z) { free (H_Tau) ; free (H_start) ; free (H_end) ; free (Z) ; return -1 ; }
    /* Q'X works on Y(HPinv[i],:) = X(i,:) (:2004-2014); QX on a copy of X, permuted at the end (:2036-2048) */
    for (Int k = 0 ; k < nx ; k++)
        for (Int i = 0 ; i < m ; i++)
        {
            if (method == 0) Z [num->HPinv [i] + k*m] = X [i + k*m] ;
            else Z [i + k*m] = X [i + k*m] ;
        }
    for (Int ff = 0 ; ff < nf ; ff++)
    {
        Int f = (method == 0) ? ff : (nf - 1 - ff) ;
        Int nh = oracle_get_H_vectors (f, sym, num->HStair, num->HTau, num->Hm, H_Tau, H_start, H_end) ;
        const double *R = num->stack + num->Roff [f] ;
        const int64_t *Hi = num->Hii + sym->Hip [f] ;
        for (Int hh = 0 ; hh < nh ; hh++)
        {
            Int h = (method == 0) ? hh : (nh - 1 - hh) ;
            double tau = H_Tau [h] ;
            if (tau == 0) continue ;
            Int len = H_end [h] - H_start [h] ;            /* entries below the unit diagonal */
            const double *v = R + H_start [h] ;
            for (Int k = 0 ; k < nx ; k++)
            {
                double *z = Z + k*m ;
                double s = z [Hi [h]] ;
                for (Int i = 0 ; i < len ; i++) s += v [i] * z [Hi [h+1+i]] ;
                s *= tau ;
                z [Hi [h]] -= s ;
                for (Int i = 0 ; i < len ; i++) z [Hi [h+1+i]] -= s * v [i] ;
            }
        }
    }
    for (Int k = 0 ; k < nx ; k++)
        for (Int i = 0 ; i < m ; i++)
            Y [i + k*m] = (method == 0) ? Z [i + k*m] : Z [num->HPinv [i] + k*m] ;
    free (H_Tau) ; free (H_start) ; free (H_end) ; free (Z) ;
    return 0 ;
}

/* ------------------------------------------------------------------------- */
/* X = R\B (use_Qfill = 0) or X = E*(R\B) (use_Qfill = 1) from the packed R+H  */
/* blocks: qr_rsolve, SparseQR.c:2218-2465 (multifrontal rows only; no          */
/* singletons, keepH = 1).  B is m-by-nrhs (ld = m; only its first `rank` rows  */
/* are used), X is n-by-nrhs.  Dead columns get the basic solution x_j = 0.     */
/* ------------------------------------------------------------------------- */
int stmqr_oracle_rsolve (const stmqr_symbolic_view *sym, const stmqr_numeric_view *num, int64_t rank,
    int64_t maxfrank, int use_Qfill, int64_t nrhs, const double *B, double *X)
{
    Int m = sym->m, n = sym->n, nf = sym->nf ;
    const int64_t *Qfill = use_Qfill ? sym->Qfill : NULL ;
    Int mf = IMAX (maxfrank, 1) ;
    const double **Rcolp = (const double **) malloc ((size_t) mf * sizeof (double *)) ;
    Int *Rlive = (Int *) malloc ((size_t) mf * sizeof (Int)) ;
    double *W = (double *) malloc ((size_t) (mf * IMAX (nrhs, 1)) * sizeof (double)) ;
    if (!Rcolp || !Rlive || !W) { free (Rcolp) ; free (Rlive) ; free (W) ; return -1 ; }
    for (Int e = 0 ; e < n * nrhs ; e++) X [e] = 0 ;
    Int row2 = rank ;                                   /* last row of R + 1 (:2307) */
    for (Int f = nf-1 ; f >= 0 ; f--)
    {
        const double *R = num->stack + num->Roff [f] ;
        Int col1 = sym->Super [f], fp = sym->Super [f+1] - col1 ;
        Int pr = sym->Rp [f], fn = sym->Rp [f+1] - pr ;
        const int64_t *Stair = num->HStair + pr ;
        Int fm = num->Hm [f], h = 0, t = 0, rm = 0, k ;
        for (k = 0 ; k < fp ; k++)                      /* live pivot columns (:2331-2381) */
        {
            Int j = col1 + k, live ;
            t = Stair [k] ;
            if (t == 0) { live = 0 ; t = rm ; h = rm ; }
            else { live = (rm < fm) ; h = rm + 1 ; }
            if (live) { Rcolp [rm] = R ; Rlive [rm] = j ; rm++ ; }
            else
            {
                Int ii = Qfill ? Qfill [j] : j ;
                if (ii < n) for (Int kk = 0 ; kk < nrhs ; kk++) X [ii + kk*n] = 0 ;
            }
            R += rm + (t-h) ;
        }
        Int row1 = row2 - rm ;
        for (Int kk = 0 ; kk < nrhs ; kk++)             /* right-hand side of these rm equations */
            for (Int i = 0 ; i < rm ; i++)
            {
                Int ii = row1 + i ;
                W [i + kk*rm] = (ii < rank) ? B [ii + kk*m] : 0 ;
            }
        for ( ; k < fn ; k++)                           /* rectangular part: W -= R2 x2 (:2409-2444) */
        {
            Int j = sym->Rj [pr + k] ;
            Int ii = Qfill ? Qfill [j] : j ;
            if (ii >= n) break ;
            if (!num->Rdead [j])
                for (Int kk = 0 ; kk < nrhs ; kk++)
                {
                    double xi = X [ii + kk*n] ;
                    if (xi != 0) for (Int i = 0 ; i < rm ; i++) W [i + kk*rm] -= R [i] * xi ;
                }
            R += rm ;
            t = Stair [k] ;
            h = IMIN (h+1, fm) ;
            R += (t-h) ;
        }
        for (k = rm-1 ; k >= 0 ; k--)                   /* packed upper triangular part (:2450-2478) */
        {
            const double *Rk = Rcolp [k] ;
            Int j = Rlive [k] ;
            Int ii = Qfill ? Qfill [j] : j ;
            if (ii < n)
                for (Int kk = 0 ; kk < nrhs ; kk++)
                {
                    double xi = W [k + kk*rm] / Rk [k] ;
                    X [ii + kk*n] = xi ;
                    if (xi != 0) for (Int i = 0 ; i < k ; i++) W [i + kk*rm] -= Rk [i] * xi ;
                }
        }
        row2 = row1 ;
    }
    free (Rcolp) ; free (Rlive) ; free (W) ;
    return 0 ;
}
