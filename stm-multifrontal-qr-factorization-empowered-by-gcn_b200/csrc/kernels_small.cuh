// kernels_small.cuh -- the small-front batching path: ONE launch per etree level factorizes all the
// small fronts of the level, one warp per front, the whole front in shared memory.
//
// Fuses, for a front whose bound Fm x fn fits the shared-memory budget and fn <= 64:
//   qr_fsize     (SparseQR_factorize.c:1066-1145)  Stair, fm
//   qr_assemble  (:1151-1285)  rows of S, extend-add of the children's packed C blocks -> F in
//                              shared memory (the zero fill and the scatter never touch HBM)
//   qr_front     (:1383-1618)  staircase Householder QR, dlarfg / dlarf column by column, dead
//                              columns, early exit, rank (the reference also drops to one panel
//                              spanning all columns for fronts this small, :1563-1567)
//   qr_fcsize    (:1623-1634)  + the column lengths of qr_rhpack (:1726-1780), Hr, Cm
// F is written ONCE to the level's front arena (ld = fm); k_level_alloc / k_pack then treat small
// and large fronts alike.  The tens of thousands of leaf-level fronts of a 2-D / 3-D mesh problem
// (71 438 fronts of <= 36 x 61 on level 0 of config 2) cost one launch.
//
// lane = column (two columns per lane when fn > 32), rows are looped: a dot product of the pivot
// column with every other column is a private FMA chain per lane, no warp reductions at all.
#pragma once
#include "engine.cuh"
#include "kernels_panel.cuh"

namespace stmqr {

constexpr int SMALL_MAX_FN = 64 ;

// dynamic shared memory of one CTA (= one warp = one front): F[cap] | st[64] | live[64] | cmap[maxrows]
__host__ __device__ constexpr size_t small_smem_bytes (int cap_doubles, int maxrows)
{
    return sizeof (double) * (size_t) cap_doubles + sizeof (I32) * (size_t) (2 * SMALL_MAX_FN + maxrows + 4) ;
}

__device__ __forceinline__ I32 warp_excl_scan (I32 v, I32 &total)
{
    const int lane = threadIdx.x & 31 ;
    I32 inc = v ;
#pragma unroll
    for (int o = 1 ; o < 32 ; o <<= 1)
    {
        const I32 u = __shfl_up_sync (STMQR_FULL_MASK, inc, o) ;
        if (lane >= o) inc += u ;
    }
    total = __shfl_sync (STMQR_FULL_MASK, inc, 31) ;
    return inc - v ;
}
__device__ __forceinline__ I64 warp_excl_scan64 (I64 v, I64 &total)
{
    const int lane = threadIdx.x & 31 ;
    I64 inc = v ;
#pragma unroll
    for (int o = 1 ; o < 32 ; o <<= 1)
    {
        const I64 u = __shfl_up_sync (STMQR_FULL_MASK, inc, o) ;
        if (lane >= o) inc += u ;
    }
    total = __shfl_sync (STMQR_FULL_MASK, inc, 31) ;
    return inc - v ;
}

__global__ void __launch_bounds__ (32) k_front_small (const I32 *__restrict__ fronts, DSym S, DNum N, double tol,
    I64 ntol_all, I32 cap, I32 maxrows)
{
    extern __shared__ double smraw [] ;
    double *F = smraw ;
    I32 *st = (I32 *) (smraw + cap) ;           // [64] staircase
    I32 *cmap = st + 2 * SMALL_MAX_FN ;         // [maxrows] child C row -> front row, all children back to back
    const int lane = threadIdx.x ;
    const I32 f = fronts [blockIdx.x] ;
    const I32 col1 = S.Super [f], fp = S.Super [f+1] - col1 ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const I32 c1 = S.Childp [f], c2 = S.Childp [f+1] ;
    const I32 ntol = (I32) max ((I64) 0, min (ntol_all - (I64) col1, (I64) fp)) ;

    // ---- qr_fsize: # rows starting in every column, exclusive scan -> row start, fm ----------------
    for (I32 j = lane ; j < SMALL_MAX_FN ; j += 32)
        st [j] = (j < fp) ? (S.Sleft [col1+j+1] - S.Sleft [col1+j]) : 0 ;
    __syncwarp () ;
    for (I32 q = c1 ; q < c2 ; q++)
    {
        const I32 c = S.Child [q] ;
        const I32 pc = S.Rp [c] + (S.Super [c+1] - S.Super [c]) ;
        const I32 cm = N.Cm [c] ;
        for (I32 ci = lane ; ci < cm ; ci += 32) st [S.Cj [pc+ci]] += 1 ;     // distinct columns inside a child
        __syncwarp () ;
    }
    I32 fm ;
    {
        const I32 a = st [lane], b = st [lane + 32] ;
        I32 ta, tb ;
        const I32 ea = warp_excl_scan (a, ta) ;
        const I32 eb = warp_excl_scan (b, tb) ;
        st [lane] = ea ; st [lane + 32] = ta + eb ;
        fm = ta + tb ;
    }
    __syncwarp () ;
    const I32 ld = fm | 1 ;                     // odd: the lanes of a row hit distinct banks
    I32 *Hi = N.Hii + S.Hip [f] ;
    // zero F
    for (I32 e = lane ; e < ld * fn ; e += 32) F [e] = 0.0 ;
    __syncwarp () ;

    // ---- rows of S whose leftmost column is a pivot of this front (:1188-1208) ----------------------
    const I32 r1 = S.Sleft [col1], r2 = S.Sleft [col1+fp] ;
    for (I32 r = r1 + lane ; r < r2 ; r += 32)
    {
        const I32 pb = S.Sp [r], pe = S.Sp [r+1] ;
        const I32 k = S.Sj [pb] - col1 ;
        const I32 i = st [k] + (r - S.Sleft [col1+k]) ;
        Hi [i] = r ;
        for (I32 p = pb ; p < pe ; p++) F [i + S.Sjf [p] * ld] = N.Sx [p] ;
    }
    __syncwarp () ;
    for (I32 k = lane ; k < fp ; k += 32) st [k] += S.Sleft [col1+k+1] - S.Sleft [col1+k] ;
    __syncwarp () ;
    // ---- children: row map, row ids, extend-add of the packed C blocks (:1239-1281) -------------------
    I32 cbase = 0 ;
    for (I32 q = c1 ; q < c2 ; q++)
    {
        const I32 c = S.Child [q] ;
        const I32 fpc = S.Super [c+1] - S.Super [c] ;
        const I32 pc = S.Rp [c] + fpc ;
        const I32 cn = (S.Rp [c+1] - S.Rp [c]) - fpc ;
        const I32 cm = N.Cm [c] ;
        const I32 *Hichild = N.Hii + S.Hip [c] + N.Hr [c] ;
        for (I32 ci = lane ; ci < cm ; ci += 32)
        {
            const I32 j = S.Cj [pc+ci] ;
            const I32 i = st [j] ;
            st [j] = i + 1 ;
            cmap [cbase + ci] = i ;
            Hi [i] = Hichild [ci] ;
        }
        __syncwarp () ;
        if (cm > 0)
        {
            const double *C = N.C + S.Coff [c] ;
            for (I32 cj = 0 ; cj < cn ; cj++)
            {
                const I32 j = S.Cj [pc+cj] ;
                const I32 len = min (cj+1, cm) ;
                const double *src = C + cblock_col_offset (cj, cm) ;
                for (I32 ci = lane ; ci < len ; ci += 32) F [cmap [cbase + ci] + j * ld] = __ldcg (src + ci) ;
            }
        }
        cbase += cm ;
        __syncwarp () ;
    }
    // st[j] is now the row END of column j (the staircase)

    // ---- qr_front: unblocked staircase Householder QR, lane = column (lane, lane+32) -------------------
    double *Tau = N.HTau + p1 ;
    char *Rdead = N.Rdead + col1 ;
    const I32 ja = lane, jb = lane + 32 ;
    double *Fa = F + ja * ld, *Fb = F + ((jb < fn) ? jb : 0) * ld ;
    const bool hasa = (ja < fn), hasb = (jb < fn) ;
    I32 g = 0, rank = min (fm, fp) ;
    double flops = 0 ;
    I32 nlive = 0 ;                 // live pivots so far (for rm)
    for (I32 k = 0 ; k < fn ; k++)
    {
        if (g >= fm)
        {
            // no rows left (:1444-1458): remaining pivots are dead, remaining columns Stair = m
            for (I32 kk = k + lane ; kk < fn ; kk += 32)
            {
                if (kk < fp) { Rdead [kk] = 1 ; st [kk] = 0 ; }
                else st [kk] = fm ;
                Tau [kk] = 0 ;
            }
            break ;
        }
        const I32 t = max (g + 1, st [k]) ;
        const double *x = F + k * ld ;
        // dots of column k with my columns over rows (g, t); F(g, .) for the pivot row
        double sa = 0, sb = 0 ;
        const bool doa = hasa && ja >= k, dob = hasb && jb >= k ;
        for (I32 i = g + 1 ; i < t ; i++)
        {
            const double xv = x [i] ;
            if (doa) sa = fma (xv, Fa [i], sa) ;
            if (dob) sb = fma (xv, Fb [i], sb) ;
        }
        const double mine = (k < 32) ? sa : sb ;
        double ss = __shfl_sync (STMQR_FULL_MASK, mine, k & 31) ;
        const double alpha = x [g] ;
        double beta = alpha, tau = 0, scale = 0 ;
        if (t - g > 1)
        {
            double nrm ;
            if (ss > 1e-280 && ss < 1e280 && fabs (alpha) < 1e140) nrm = sqrt (fma (alpha, alpha, ss)) ;
            else if (ss <= 1e-280 && fabs (alpha) > 1e-120 && fabs (alpha) < 1e140) { nrm = 0 ; ss = 0 ; }   // noise: H = I (see kernels_panel.cuh)
            else
            {
                // rare: zero or badly scaled sub-column (dnrm2 semantics)
                double mx = 0 ;
                for (I32 i = g + 1 + lane ; i < t ; i += 32) mx = fmax (mx, fabs (x [i])) ;
                mx = warp_max (mx) ;
                if (mx > 0)
                {
                    const double inv = 1.0 / mx ;
                    double s2 = 0 ;
                    for (I32 i = g + 1 + lane ; i < t ; i += 32) { const double v = x [i] * inv ; s2 += v * v ; }
                    s2 = warp_sum (s2) ;
                    nrm = hypot (alpha, mx * sqrt (s2)) ;
                    ss = 1.0 ;
                    if (!(nrm >= 1e-290)) { nrm = 0 ; ss = 0 ; }    // underflowed noise: H = I (see kernels_panel.cuh)
                }
                else { nrm = 0 ; ss = 0 ; }
            }
            if (ss != 0)
            {
                beta = -copysign (nrm, alpha) ;
                tau = (beta - alpha) / beta ;
                scale = 1.0 / (alpha - beta) ;
                if (!(fabs (tau) <= 2.0) || !(fabs (scale) < 1e300)) { beta = alpha ; tau = 0 ; scale = 0 ; }
            }
        }
        const bool dead = (k < ntol) && (fabs (beta) <= tol) ;
        __syncwarp () ;
        if (dead)
        {
            for (I32 i = g + lane ; i < fm ; i += 32) F [i + k * ld] = 0.0 ;
            if (lane == 0) { st [k] = 0 ; Tau [k] = 0 ; Rdead [k] = 1 ; }
        }
        else
        {
            if (tau != 0)
            {
                // dlarf on my columns right of k
                if (hasa && ja > k)
                {
                    const double wv = tau * (Fa [g] + scale * sa), fct = scale * wv ;
                    Fa [g] -= wv ;
                    for (I32 i = g + 1 ; i < t ; i++) Fa [i] = fma (-x [i], fct, Fa [i]) ;
                }
                if (hasb && jb > k)
                {
                    const double wv = tau * (Fb [g] + scale * sb), fct = scale * wv ;
                    Fb [g] -= wv ;
                    for (I32 i = g + 1 ; i < t ; i++) Fb [i] = fma (-x [i], fct, Fb [i]) ;
                }
                __syncwarp () ;
                // v = x * scale below the diagonal
                for (I32 i = g + 1 + lane ; i < t ; i += 32) F [i + k * ld] *= scale ;
            }
            if (lane == 0) { F [g + k * ld] = beta ; Tau [k] = tau ; st [k] = t ; }
            flops += (double) (t - g) * (3.0 + 4.0 * (double) (fn - k - 1)) ;
            if (k < fp) nlive++ ;
            g++ ;
        }
        if (k == fp - 1) rank = g ;
        __syncwarp () ;
    }

    // ---- sizes of the packed blocks (k_front_finish), Hr, Cm, HStair ------------------------------------
    // rm before pivot column k = # live pivots among columns < k, capped at fm
    I64 *colp = N.colp + p1 ;
    I32 *stg = N.stair + p1 ;
    {
        const I32 la = (ja < fp && st [ja] != 0) ? 1 : 0, lb = (jb < fp && st [jb] != 0) ? 1 : 0 ;
        I32 ta, tb ;
        const I32 ea = warp_excl_scan (la, ta) ;
        const I32 eb = ta + warp_excl_scan (lb, tb) ;
        const I32 nl = ta + tb ;
        const I32 rm = min (fm, nl) ;
        auto collen = [&] (I32 k, I32 before) -> I64 {
            if (k >= fn) return 0 ;
            if (k < fp)
            {
                const I32 tt = st [k] ;
                return (tt == 0) ? (I64) min (fm, before) : (I64) tt ;
            }
            const I32 hh = min (rm + (k - fp + 1), fm) ;
            return (I64) rm + max (0, st [k] - hh) ;
        } ;
        const I64 lena = collen (ja, ea), lenb = collen (jb, eb) ;
        I64 sa64, sb64 ;
        const I64 offa = warp_excl_scan64 (lena, sa64) ;
        const I64 offb = sa64 + warp_excl_scan64 (lenb, sb64) ;
        const I64 rsize = (fm > 0) ? sa64 + sb64 : 0 ;
        if (hasa) { colp [ja] = offa ; stg [ja] = st [ja] ; }
        if (hasb) { colp [jb] = offb ; stg [jb] = st [jb] ; }
        (void) nlive ;
        if (lane == 0)
        {
            const I32 cn = fn - fp ;
            I32 cm = min (fm - rank, cn) ;
            if (cm < 0 || cn <= 0) cm = 0 ;
            N.Hm [f] = fm ;
            N.rank [f] = rank ;
            N.Cm [f] = cm ;
            N.Hr [f] = (fm > 0) ? rm : 0 ;
            N.rsize [f] = rsize ;
            atomicMax (N.maxfm, fm) ;
            atomicAdd (N.sumrank, rank) ;
            atomicMax (N.maxfrank, rank) ;
            if (flops != 0) atomicAdd (N.flops, flops) ;
            // algorithmic bytes of assembly + pack, same formula as k_front_finish
            double csz_children = 0, ids = fm ;
            for (I32 q = c1 ; q < c2 ; q++)
            {
                const I32 c = S.Child [q] ;
                const double cmc = N.Cm [c] ;
                const double cnc = (S.Rp [c+1] - S.Rp [c]) - (S.Super [c+1] - S.Super [c]) ;
                csz_children += cmc * (cmc + 1) / 2 + cmc * (cnc - cmc) ;
                ids += cmc + cnc ;
            }
            const double snz = (double) (S.Sp [r2] - S.Sp [r1]) ;
            const double csz = (double) cm * (cm + 1) / 2 + (double) cm * (cn - cm) ;
            atomicAdd (N.flops + 2, 8.0 * (2.0 * (double) fm * fn + csz_children + (double) rsize + csz) + 16.0 * snz + 8.0 * ids) ;
        }
    }
    // ---- F to the level's front arena (ld = fm) for k_pack ---------------------------------------------
    if (fm > 0)
    {
        double *Fg = N.F + S.Foff [f] ;
        for (I32 j = 0 ; j < fn ; j++)
            for (I32 i = lane ; i < fm ; i += 32) Fg [i + (I64) j * fm] = F [i + j * ld] ;
    }
}

} // namespace stmqr
