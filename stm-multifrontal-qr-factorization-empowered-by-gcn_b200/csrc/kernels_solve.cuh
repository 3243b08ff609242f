// kernels_solve.cuh -- the consumers next to the path, on the device: applying Q
// or Q' and solving with R straight from the packed R+H blocks that the factorization left in HBM, so a
// least-squares solve never downloads the factor (1.7 GB on BASELINE config 2, 29 GB on config 5).
//
//   QR_qmult / qr_private_Happly / qr_private_get_H_vectors   STMMQR/src/qr/SparseQR.c:1455-2110
//   qr_rsolve                                                   STMMQR/src/qr/SparseQR.c:2218-2465
//
// The reference walks the fronts one after the other (forward for Q', backward for Q and for R\b).  Fronts
// of one etree level touch disjoint rows of the right-hand side (a front's rows are its own rows of S plus
// the rows its children passed up) and write disjoint entries of x (their own pivot columns), so here all
// fronts of a level run in ONE launch: one CTA per large front, one warp per small front.
#pragma once
#include "engine.cuh"

namespace stmqr {

// ---------------------------------------------------------------------------------------------
// Householder table: hcol [Rp[f] + q] = column of the q-th Householder vector of front f, nh [f] =
// their number (qr_private_get_H_vectors, SparseQR.c:1455-1546: live pivot columns in order, then the
// non-pivot columns while rows are left).  Vector q has its unit entry on front row q and
// HStair[k] - (q+1) entries below it, stored in column k of the packed block behind the R part (q+1
// entries for a pivot column, Hr[f] for a non-pivot column).  One warp per front.
// ---------------------------------------------------------------------------------------------
// rlen [Rp[f] + k] (optional) = # entries of the R part of column k of the packed block (qr_rhpack: the rows
// above the Householder vector): live pivot column = # live pivots up to and including it, dead pivot column =
// # live pivots before it, non-pivot column = Hr[f].
__global__ void k_htable (DSym S, DNum N, I32 *__restrict__ hcol, I32 *__restrict__ nh, I32 *__restrict__ rlen)
{
    const int lane = threadIdx.x & 31 ;
    const I32 f = (I32) (((I64) blockIdx.x * blockDim.x + threadIdx.x) >> 5) ;
    if (f >= S.nf) return ;
    const I32 fp = S.Super [f+1] - S.Super [f] ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const I32 fm = N.Hm [f] ;
    const I32 *st = N.stair + p1 ;
    I32 q = 0 ;                     // vectors so far = rm while in the pivot columns, = h afterwards
    I32 base = 0 ;
    for ( ; base < fn && q < fm ; base += 32)
    {
        const I32 k = base + lane ;
        bool is = false ;
        if (k < fn)
        {
            if (k < fp) is = (st [k] != 0) ;
            else is = true ;        // every non-pivot column carries a vector until the rows run out
        }
        const unsigned mask = __ballot_sync (STMQR_FULL_MASK, is) ;
        const I32 mine = q + __popc (mask & ((1u << lane) - 1u)) ;
        // the loop of the reference stops once h (= vectors so far) reaches fm
        if (is && mine < fm) hcol [p1 + mine] = k ;
        if (rlen && k < fp) rlen [p1 + k] = min (fm, mine + (is ? 1 : 0)) ;
        q = min (fm, q + __popc (mask)) ;
    }
    if (lane == 0) nh [f] = q ;
    if (rlen)
    {
        // non-pivot columns: Hr entries; pivot columns the scan above did not reach (the rows ran out before
        // them, so they are dead): Hr entries as well
        const I32 rm = N.Hr [f] ;
        for (I32 k = lane ; k < fn ; k += 32)
            if (k >= fp || k >= base) rlen [p1 + k] = rm ;
    }
}

// group = the threads that work on one front: a whole CTA (GROUP = blockDim.x) or one warp (GROUP = 32)
template <int GROUP>
__device__ __forceinline__ void group_sync ()
{
    if (GROUP == 32) __syncwarp () ; else __syncthreads () ;
}

// sum of NX values over the group; every thread gets the totals.  sh: [NX * 32] doubles per CTA (unused by warps)
template <int GROUP, int NX>
__device__ __forceinline__ void group_sum (double (&s) [NX], double *sh)
{
#pragma unroll
    for (int c = 0 ; c < NX ; c++) s [c] = warp_sum (s [c]) ;
    if (GROUP == 32) return ;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5 ;
    constexpr int NWG = GROUP / 32 ;
    __syncthreads () ;                      // (sh may still be read from the previous vector)
    if (lane == 0)
    {
#pragma unroll
        for (int c = 0 ; c < NX ; c++) sh [c * 32 + w] = s [c] ;
    }
    __syncthreads () ;
#pragma unroll
    for (int c = 0 ; c < NX ; c++)
    {
        double t = 0 ;
#pragma unroll
        for (int ww = 0 ; ww < NWG ; ww++) t += sh [c * 32 + ww] ;      // fixed order: deterministic
        s [c] = t ;
    }
}

// ---------------------------------------------------------------------------------------------
// Z <- Q' Z (method 0: fronts of the level, vectors forward) or Z <- Q Z (method 1: vectors backward) for
// the fronts [first, first+count) of one etree level.  Z is m-by-nx (ld = m) in the row order of the
// factorization (HPinv applied by the caller's gather/scatter kernels); Hii holds the permuted row ids.
// ---------------------------------------------------------------------------------------------
template <int GROUP, int NX>
__global__ void __launch_bounds__ (GROUP == 32 ? 256 : GROUP) k_qapply (const I32 *__restrict__ fronts, I32 count, DSym S, DNum N,
    const I64 *__restrict__ Hii, const I32 *__restrict__ hcol, const I32 *__restrict__ nhv, int method, I32 nx0,
    double *__restrict__ Z)
{
    __shared__ double sh [NX * 32] ;
    const int gtid = (GROUP == 32) ? (threadIdx.x & 31) : threadIdx.x ;
    const I32 slot = (GROUP == 32) ? (I32) (((I64) blockIdx.x * blockDim.x + threadIdx.x) >> 5) : (I32) blockIdx.x ;
    if (slot >= count) return ;
    const I32 f = fronts [slot] ;
    const I32 fp = S.Super [f+1] - S.Super [f] ;
    const I32 p1 = S.Rp [f] ;
    const I32 nh = nhv [f] ;
    const I32 rm = N.Hr [f] ;
    const I64 m = S.m ;
    const double *R = N.R + N.Roff [f] ;
    const I64 *Hi = Hii + S.Hip [f] ;
    const I32 *st = N.stair + p1 ;
    const I64 *colp = N.colp + p1 ;
    double *Zc = Z + (I64) nx0 * m ;
    for (I32 qq = 0 ; qq < nh ; qq++)
    {
        const I32 q = (method == 0) ? qq : (nh - 1 - qq) ;
        const I32 k = hcol [p1 + q] ;
        const double tau = N.HTau [p1 + k] ;
        if (tau == 0) continue ;                        // (uniform over the group)
        const I32 len = st [k] - (q + 1) ;
        const double *v = R + colp [k] + ((k < fp) ? (q + 1) : rm) ;
        const I64 r0 = Hi [q] ;
        double s [NX] ;
#pragma unroll
        for (int c = 0 ; c < NX ; c++) s [c] = 0 ;
        for (I32 i = gtid ; i < len ; i += GROUP)
        {
            const double vi = v [i] ;
            const I64 r = Hi [q + 1 + i] ;
#pragma unroll
            for (int c = 0 ; c < NX ; c++) s [c] = fma (vi, Zc [r + c * m], s [c]) ;
        }
        group_sum<GROUP, NX> (s, sh) ;
#pragma unroll
        for (int c = 0 ; c < NX ; c++) s [c] = tau * (s [c] + Zc [r0 + c * m]) ;
        group_sync<GROUP> () ;                          // everybody has read Z (r0) before it changes
        for (I32 i = gtid ; i < len ; i += GROUP)
        {
            const double vi = v [i] ;
            const I64 r = Hi [q + 1 + i] ;
#pragma unroll
            for (int c = 0 ; c < NX ; c++) Zc [r + c * m] = fma (-s [c], vi, Zc [r + c * m]) ;
        }
        if (gtid == 0)
        {
#pragma unroll
            for (int c = 0 ; c < NX ; c++) Zc [r0 + c * m] -= s [c] ;
        }
        group_sync<GROUP> () ;
    }
}

// Z (HPinv [i], :) = X (i, :)  (scatter = 1) or Y (i, :) = Z (HPinv [i], :)  (scatter = 0): the row
// permutation around the Householder products (QR_qmult, SparseQR.c:2004-2014, :2036-2048)
__global__ void k_permute_rows (I64 m, I64 nx, const I64 *__restrict__ HPinv, const double *__restrict__ src,
    double *__restrict__ dst, int scatter)
{
    const I64 tot = m * nx ;
    for (I64 e = (I64) blockIdx.x * blockDim.x + threadIdx.x ; e < tot ; e += (I64) gridDim.x * blockDim.x)
    {
        const I64 i = e % m, c = e / m ;
        if (scatter) dst [HPinv [i] + c * m] = src [e] ;
        else dst [e] = src [HPinv [i] + c * m] ;
    }
}

// ---------------------------------------------------------------------------------------------
// Back substitution with the packed R blocks for the fronts of one etree level (levels from the root
// down): qr_rsolve, SparseQR.c:2307-2478.  B is m-by-nrhs (ld = m; rows >= rank are not used), X is
// n-by-nrhs and zero on entry, W is n-by-nrhs scratch indexed by the global row of R (row1 + i).
// Qfill == nullptr: X = R \ B in the column order of S, else X = E * (R \ B).
// ---------------------------------------------------------------------------------------------
template <int GROUP, int NX>
__global__ void __launch_bounds__ (GROUP == 32 ? 256 : GROUP) k_rsolve (const I32 *__restrict__ fronts, I32 count, DSym S, DNum N,
    const I32 *__restrict__ hcol, const I32 *__restrict__ Qfill, I64 rank, I32 c0, const double *__restrict__ B,
    double *__restrict__ X, double *__restrict__ W)
{
    const int gtid = (GROUP == 32) ? (threadIdx.x & 31) : threadIdx.x ;
    const I32 slot = (GROUP == 32) ? (I32) (((I64) blockIdx.x * blockDim.x + threadIdx.x) >> 5) : (I32) blockIdx.x ;
    if (slot >= count) return ;
    const I32 f = fronts [slot] ;
    const I32 col1 = S.Super [f], fp = S.Super [f+1] - col1 ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const I32 rm = N.Hr [f] ;
    if (rm <= 0) return ;
    const I64 m = S.m, n = S.n ;
    const I64 row1 = N.base1 [f] ;                     // first row of this front's block of R (exclusive scan of Hr)
    const double *R = N.R + N.Roff [f] ;
    const I64 *colp = N.colp + p1 ;
    const double *Bc = B + (I64) c0 * m ;
    double *Xc = X + (I64) c0 * n, *Wc = W + (I64) c0 * n ;
    // right-hand side of these rm equations minus the rectangular part times the known x (:2409-2444):
    // thread i owns row i and walks along the columns (coalesced along i)
    for (I32 i = gtid ; i < rm ; i += GROUP)
    {
        double w [NX] ;
#pragma unroll
        for (int c = 0 ; c < NX ; c++) w [c] = (row1 + i < rank) ? Bc [row1 + i + c * m] : 0.0 ;
        for (I32 k = fp ; k < fn ; k++)
        {
            const I32 j = S.Rj [p1 + k] ;
            const I64 ii = Qfill ? (I64) Qfill [j] : (I64) j ;
            if (ii >= n) break ;
            if (N.Rdead [j]) continue ;
            const double rik = R [colp [k] + i] ;
#pragma unroll
            for (int c = 0 ; c < NX ; c++) w [c] = fma (-rik, Xc [ii + c * n], w [c]) ;
        }
#pragma unroll
        for (int c = 0 ; c < NX ; c++) Wc [row1 + i + c * n] = w [c] ;
    }
    group_sync<GROUP> () ;
    // packed upper triangular part, last live pivot column first (:2450-2478); vector q of the Householder
    // table is the q-th live pivot column for q < rm
    for (I32 q = rm - 1 ; q >= 0 ; q--)
    {
        const I32 k = hcol [p1 + q] ;
        const I32 j = col1 + k ;
        const I64 ii = Qfill ? (I64) Qfill [j] : (I64) j ;
        const double *Rk = R + colp [k] ;
        double xq [NX] ;
        const double d = Rk [q] ;
#pragma unroll
        for (int c = 0 ; c < NX ; c++) xq [c] = Wc [row1 + q + c * n] / d ;
        if (ii < n)
        {
            for (I32 i = gtid ; i < q ; i += GROUP)
            {
                const double rik = Rk [i] ;
#pragma unroll
                for (int c = 0 ; c < NX ; c++) Wc [row1 + i + c * n] = fma (-rik, xq [c], Wc [row1 + i + c * n]) ;
            }
            if (gtid == 0)
            {
#pragma unroll
                for (int c = 0 ; c < NX ; c++) Xc [ii + c * n] = xq [c] ;
            }
        }
        group_sync<GROUP> () ;
    }
}

// ---------------------------------------------------------------------------------------------
// R as a compressed-column matrix from the packed blocks (qr_rcount / qr_rconvert, STMMQR/src/qr/
// SparseLQ.c:102-297, :299-520, the Ra/Rap branch: n1rows = 0, n2 = n, getT = 0).  The reference walks
// the fronts in order and appends to every column through a running cursor, so inside a column the
// entries are ordered by front, then by row; exact zeros are dropped (:235, :460).  Here:
//   k_rcount      one warp per (front, column): # non-zeros of its R part with row < econ
//   k_rcol_scan   one warp per column of R: exclusive running sum of those counts over the fronts that
//                 hold the column, in front order, through the transpose of Rj (built at analyze time:
//                 purely symbolic) -> where each front's piece starts inside the column, column totals
//   (scan of the column totals -> Rp)
//   k_rfill       one warp per (front, column): ballot-compacts the non-zeros into its piece
// ---------------------------------------------------------------------------------------------
__global__ void k_rcount (DSym S, DNum N, const I32 *__restrict__ posfront, const I32 *__restrict__ rlen,
    I64 rjsize, I64 econ, I32 *__restrict__ cnt)
{
    const int lane = threadIdx.x & 31 ;
    const I64 p = ((I64) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    if (p >= rjsize) return ;
    const I32 f = posfront [p] ;
    const I32 len = rlen [p] ;
    const I64 row1 = N.base1 [f] ;
    const double *Rk = N.R + N.Roff [f] + N.colp [p] ;
    I32 c = 0 ;
    for (I32 i = lane ; i < len ; i += 32) c += (Rk [i] != 0.0 && row1 + i < econ) ? 1 : 0 ;
    c = warp_sum_i (c) ;
    if (lane == 0) cnt [p] = c ;
}

__global__ void k_rcol_scan (I64 n, const I32 *__restrict__ RjTp, const I32 *__restrict__ RjTi,
    const I32 *__restrict__ cnt, I32 *__restrict__ off, I64 *__restrict__ colcount)
{
    const int lane = threadIdx.x & 31 ;
    const I64 j = ((I64) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    if (j >= n) return ;
    I32 run = 0 ;
    for (I32 b = RjTp [j] ; b < RjTp [j+1] ; b += 32)
    {
        const I32 e = b + lane ;
        const I32 p = (e < RjTp [j+1]) ? RjTi [e] : -1 ;
        const I32 c = (p >= 0) ? cnt [p] : 0 ;
        I32 inc = c ;
#pragma unroll
        for (int o = 1 ; o < 32 ; o <<= 1)
        {
            const I32 u = __shfl_up_sync (STMQR_FULL_MASK, inc, o) ;
            if (lane >= o) inc += u ;
        }
        if (p >= 0) off [p] = run + inc - c ;
        run += __shfl_sync (STMQR_FULL_MASK, inc, 31) ;
    }
    if (lane == 0) colcount [j] = run ;
}

__global__ void k_rfill (DSym S, DNum N, const I32 *__restrict__ posfront, const I32 *__restrict__ rlen,
    const I32 *__restrict__ off, const I64 *__restrict__ Rp_out, I64 rjsize, I64 econ,
    I64 *__restrict__ Ri, double *__restrict__ Rx)
{
    const int lane = threadIdx.x & 31 ;
    const I64 p = ((I64) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    if (p >= rjsize) return ;
    const I32 f = posfront [p] ;
    const I32 len = rlen [p] ;
    const I64 row1 = N.base1 [f] ;
    const I32 j = S.Rj [p] ;
    const double *Rk = N.R + N.Roff [f] + N.colp [p] ;
    I64 dst = Rp_out [j] + off [p] ;
    for (I32 b = 0 ; b < len ; b += 32)
    {
        const I32 i = b + lane ;
        const double v = (i < len) ? Rk [i] : 0.0 ;
        const bool nz = (i < len) && (v != 0.0) && (row1 + i < econ) ;
        const unsigned mask = __ballot_sync (STMQR_FULL_MASK, nz) ;
        if (nz)
        {
            const I64 q = dst + __popc (mask & ((1u << lane) - 1u)) ;
            Ri [q] = row1 + i ;
            Rx [q] = v ;
        }
        dst += __popc (mask) ;
    }
}

} // namespace stmqr
