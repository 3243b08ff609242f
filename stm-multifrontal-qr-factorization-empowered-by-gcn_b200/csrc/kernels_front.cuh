// kernels_front.cuh -- tiled staircase Householder QR of the fronts of one etree level
// (qr_front, SparseQR_factorize.c:1383-1618, with the dlarfg/dlarf/dlarft/dlarfb semantics of
// SURVEY.md Appendix B restated as device code).
//
// One "panel step" = k_panel (factorize columns [k1,k1+PB) of every active front of the level,
// one CTA per front, and build the T factor) followed by k_update (apply the block reflector
// to the trailing columns, one CTA per 32-column tile per front).
#pragma once
#include "engine.cuh"

namespace stmqr {

struct LevelArgs
{
    const I32 *fronts ;     // fronts of the level, sorted by # columns descending
    I32 count ;
    double tol ;
    I64 ntol ;
} ;

// ---------------------------------------------------------------------------------------------
// Panel factorization.  Semantics that must match the reference bit for bit (integer outputs):
//   t = max (g+1, Stair[k])                       :1460
//   dead pivot (k < ntol, |F(g,k)| <= tol): zero F(g:m-1,k), Stair=0, Tau=0, Rdead=1   :1495-1528
//   out of rows (g >= m): remaining pivots dead, remaining columns Stair = m          :1444-1458
//   rank = g sampled after pivot column npiv-1                                        :1604-1608
// Blocking (panel width, when T is applied) is performance-only.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (1024) k_panel (LevelArgs L, DSym S, DNum N, I32 k1, I32 PB)
{
    __shared__ double red [64] ;
    __shared__ double Tsh [PANEL_MAX * (PANEL_MAX + 1)] ;
    __shared__ double taus [PANEL_MAX] ;
    __shared__ I32 cols [PANEL_MAX], tq [PANEL_MAX] ;

    const I32 slot = blockIdx.x ;
    const I32 f = L.fronts [slot] ;
    const I32 col1 = S.Super [f], fp = S.Super [f+1] - col1 ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const int tid = threadIdx.x, nt = blockDim.x ;
    const int lane = tid & 31, w = tid >> 5, nw = nt >> 5 ;

    if (k1 >= fn || N.done [slot])
    {
        if (tid == 0) N.pnl_nv [slot] = 0 ;
        return ;
    }
    const I32 fm = N.Hm [f] ;
    const I64 ld = fm ;
    double *F = N.F + S.Foff [f] ;
    I32 *st = N.stair + p1 ;
    double *Tau = N.HTau + p1 ;
    char *Rdead = N.Rdead + col1 ;
    const I32 ntol = (I32) max ((I64) 0, min (L.ntol - (I64) col1, (I64) fp)) ;
    const double tol = L.tol ;

    I32 g = N.g [slot] ;
    const I32 g1 = g ;
    const I32 k2 = min (fn, k1 + PB) ;
    I32 nv = 0 ;
    double flops = 0 ;
    bool out_of_rows = false ;

    for (I32 k = k1 ; k < k2 ; k++)
    {
        if (g >= fm)
        {
            // no rows left: qr_front early exit (:1444-1458) for ALL remaining columns
            for (I32 kk = k + tid ; kk < fn ; kk += nt)
            {
                if (kk < fp) { Rdead [kk] = 1 ; st [kk] = 0 ; }
                else st [kk] = fm ;
                Tau [kk] = 0 ;
            }
            out_of_rows = true ;
            break ;
        }
        const I32 t = max (g + 1, st [k]) ;
        double *x = F + (I64) k * ld ;

        // ---- dlarfg on F(g:t-1,k) -----------------------------------------------------------
        double ss = 0, mx = 0 ;
        for (I32 i = g + 1 + tid ; i < t ; i += nt)
        {
            const double v = x [i] ;
            ss += v * v ;
            mx = fmax (mx, fabs (v)) ;
        }
        block_sum_max (ss, mx, red) ;
        if (mx > 0 && !(ss > 1e-280 && ss < 1e280))
        {
            // rare: rescale to avoid under/overflow of the sum of squares (dnrm2 semantics)
            double s2 = 0, dummy = 0 ;
            const double inv = 1.0 / mx ;
            for (I32 i = g + 1 + tid ; i < t ; i += nt) { const double v = x [i] * inv ; s2 += v * v ; }
            block_sum_max (s2, dummy, red) ;
            ss = s2 ; // xnorm = mx * sqrt (s2)
        }
        else mx = 1.0 ;
        const double xnorm = mx * sqrt (ss) ;
        const double alpha = x [g] ;
        double beta = alpha, tau = 0, scale = 0 ;
        if (t - g > 1 && xnorm != 0)
        {
            beta = -copysign (hypot (alpha, xnorm), alpha) ;
            tau = (beta - alpha) / beta ;
            scale = 1.0 / (alpha - beta) ;
        }
        const bool dead = (k < ntol) && (fabs (beta) <= tol) ;
        __syncthreads () ;      // everybody has read x[g] before it is overwritten

        if (dead)
        {
            for (I32 i = g + tid ; i < fm ; i += nt) x [i] = 0.0 ;
            if (tid == 0) { st [k] = 0 ; Tau [k] = 0 ; Rdead [k] = 1 ; }
        }
        else
        {
            if (tid == 0)
            {
                x [g] = beta ; Tau [k] = tau ; st [k] = t ;
                cols [nv] = k ; tq [nv] = t ; taus [nv] = tau ;
            }
            if (tau != 0)
            {
                for (I32 i = g + 1 + tid ; i < t ; i += nt) x [i] *= scale ;
            }
            flops += (double) (t - g) * (3.0 + 4.0 * (double) (fn - k - 1)) ;
            __syncthreads () ;
            // ---- dlarf: apply H_k to the remaining columns of the panel, one warp per column
            if (tau != 0)
            {
                for (I32 c = k + 1 + w ; c < k2 ; c += nw)
                {
                    double *y = F + (I64) c * ld ;
                    double s = 0 ;
                    for (I32 i = g + 1 + lane ; i < t ; i += 32) s += x [i] * y [i] ;
                    s = warp_sum (s) ;
                    const double wv = tau * (y [g] + s) ;
                    __syncwarp () ;
                    if (lane == 0) y [g] -= wv ;
                    for (I32 i = g + 1 + lane ; i < t ; i += 32) y [i] -= x [i] * wv ;
                }
            }
            nv++ ;
            g++ ;
        }
        if (k == fp - 1 && tid == 0) N.rank [f] = g ;
        __syncthreads () ;
    }

    // ---- dlarft: T of the nv live reflectors of this panel (forward, columnwise) ------------
    // V(:,q) lives in column cols[q] of F: unit diagonal at row g1+q, entries down to tq[q]-1.
    if (nv > 0 && k2 < fn)
    {
        const int npairs = nv * (nv - 1) / 2 ;
        for (int pidx = w ; pidx < npairs ; pidx += nw)
        {
            // pair (j,i), j < i, enumerated column by column
            int i = 1, rem = pidx ;
            while (rem >= i) { rem -= i ; i++ ; }
            const int j = rem ;
            const double *vj = F + (I64) cols [j] * ld ;
            const double *vi = F + (I64) cols [i] * ld ;
            const I32 r0 = g1 + i ;
            double s = 0 ;
            for (I32 r = r0 + 1 + lane ; r < tq [j] ; r += 32) s += vj [r] * vi [r] ;
            s = warp_sum (s) ;
            if (lane == 0)
            {
                if (r0 < tq [j]) s += vj [r0] ;          // V(g1+i,i) = 1
                Tsh [j + i * (PANEL_MAX + 1)] = -taus [i] * s ;
            }
        }
        __syncthreads () ;
        if (w == 0)
        {
            for (int i = 0 ; i < nv ; i++)
            {
                // T(0:i-1,i) = T(0:i-1,0:i-1) * T(0:i-1,i)
                double s = 0 ;
                if (lane < i)
                {
                    for (int l = lane ; l < i ; l++)
                        s += Tsh [lane + l * (PANEL_MAX + 1)] * Tsh [l + i * (PANEL_MAX + 1)] ;
                }
                __syncwarp () ;
                if (lane < i) Tsh [lane + i * (PANEL_MAX + 1)] = s ;
                if (lane == i) Tsh [i + i * (PANEL_MAX + 1)] = taus [i] ;
                __syncwarp () ;
            }
        }
        __syncthreads () ;
        double *Tg = N.Tws + (I64) slot * (PANEL_MAX * PANEL_MAX) ;
        for (int e = tid ; e < nv * nv ; e += nt)
        {
            const int j = e % nv, i = e / nv ;
            Tg [j + i * PANEL_MAX] = (j <= i) ? Tsh [j + i * (PANEL_MAX + 1)] : 0.0 ;
        }
        for (int q = tid ; q < nv ; q += nt) N.pnl_cols [slot * PANEL_MAX + q] = cols [q] ;
    }
    if (tid == 0)
    {
        N.g [slot] = g ;
        N.pnl_g1 [slot] = g1 ;
        N.pnl_nv [slot] = (k2 < fn) ? nv : 0 ;
        N.pnl_tend [slot] = (nv > 0) ? tq [nv-1] : g1 ;
        if (out_of_rows) N.done [slot] = 1 ;
        if (flops != 0) atomicAdd (N.flops, flops) ;
        if (nv > 0 && k2 < fn)
            atomicAdd (N.flops + 1, 4.0 * (double) (tq [nv-1] - g1) * (double) nv * (double) (fn - k2)) ;
    }
}

// ---------------------------------------------------------------------------------------------
// Trailing update C := (I - V T V')' C = C - V T' (V' C)   (dlarfb 'L','T','F','C'; the
// reference's qr_larftb method QR_QTX, SparseQR_factorize.c:1851-1882).
// grid = (active fronts, column tiles of 32).  FP64 FMA version (the DMMA tensor-core version is
// kernels_update_dmma.cuh); rows streamed through shared memory in chunks of RC.
// ---------------------------------------------------------------------------------------------
constexpr int UPD_TB = 32 ;     // columns per CTA
constexpr int UPD_RC = 32 ;     // rows per chunk

__global__ void __launch_bounds__ (256) k_update (LevelArgs L, DSym S, DNum N, I32 k2)
{
    __shared__ double Vs [UPD_RC][PANEL_MAX + 1] ;
    __shared__ double Cs [UPD_RC][UPD_TB + 1] ;
    __shared__ double Ws [PANEL_MAX][UPD_TB + 1] ;
    __shared__ double Ts [PANEL_MAX][PANEL_MAX + 1] ;
    __shared__ I32 cols [PANEL_MAX] ;

    const I32 slot = blockIdx.x ;
    const I32 nv = N.pnl_nv [slot] ;
    if (nv == 0) return ;
    const I32 f = L.fronts [slot] ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    const I32 c0 = k2 + blockIdx.y * UPD_TB ;
    if (c0 >= fn) return ;
    const I32 ncol = min (UPD_TB, fn - c0) ;
    const I64 ld = N.Hm [f] ;
    double *F = N.F + S.Foff [f] ;
    const I32 g1 = N.pnl_g1 [slot], tend = N.pnl_tend [slot] ;
    const I32 mr = tend - g1 ;
    const int tid = threadIdx.x ;
    const int tx = tid & 31, ty = tid >> 5 ;       // 32 x 8

    if (tid < PANEL_MAX) cols [tid] = (tid < nv) ? N.pnl_cols [slot * PANEL_MAX + tid] : 0 ;
    {
        const double *Tg = N.Tws + (I64) slot * (PANEL_MAX * PANEL_MAX) ;
        for (int e = tid ; e < PANEL_MAX * PANEL_MAX ; e += 256)
        {
            const int j = e % PANEL_MAX, i = e / PANEL_MAX ;
            Ts [j][i] = (j < nv && i < nv) ? Tg [j + i * PANEL_MAX] : 0.0 ;
        }
    }
    __syncthreads () ;

    // ---- W = V' C -------------------------------------------------------------------------
    double acc [4] = {0, 0, 0, 0} ;
    for (I32 r0 = 0 ; r0 < mr ; r0 += UPD_RC)
    {
        // load V chunk (masked to unit lower trapezoidal) and C chunk; lanes run along rows
#pragma unroll
        for (int a = 0 ; a < 4 ; a++)
        {
            const int q = ty + 8 * a ;
            const I32 r = r0 + tx ;
            double v = 0 ;
            if (q < nv && r < mr)
            {
                if (r == q) v = 1.0 ;
                else if (r > q) v = F [(g1 + r) + (I64) cols [q] * ld] ;
            }
            Vs [tx][q] = v ;
            const int c = ty + 8 * a ;
            double cv = 0 ;
            if (c < ncol && r < mr) cv = F [(g1 + r) + (I64) (c0 + c) * ld] ;
            Cs [tx][c] = cv ;
        }
        __syncthreads () ;
#pragma unroll 8
        for (int r = 0 ; r < UPD_RC ; r++)
        {
            const double cv = Cs [r][tx] ;
#pragma unroll
            for (int a = 0 ; a < 4 ; a++) acc [a] += Vs [r][ty + 8 * a] * cv ;
        }
        __syncthreads () ;
    }
#pragma unroll
    for (int a = 0 ; a < 4 ; a++) Ws [ty + 8 * a][tx] = acc [a] ;
    __syncthreads () ;
    // ---- W = T' W ---------------------------------------------------------------------------
    double w2 [4] ;
#pragma unroll
    for (int a = 0 ; a < 4 ; a++)
    {
        const int i = ty + 8 * a ;
        double s = 0 ;
        for (int j = 0 ; j <= i ; j++) s += Ts [j][i] * Ws [j][tx] ;
        w2 [a] = s ;
    }
    __syncthreads () ;
#pragma unroll
    for (int a = 0 ; a < 4 ; a++) Ws [ty + 8 * a][tx] = w2 [a] ;
    __syncthreads () ;
    // ---- C -= V W ---------------------------------------------------------------------------
    for (I32 r0 = 0 ; r0 < mr ; r0 += UPD_RC)
    {
#pragma unroll
        for (int a = 0 ; a < 4 ; a++)
        {
            const int q = ty + 8 * a ;
            const I32 r = r0 + tx ;
            double v = 0 ;
            if (q < nv && r < mr)
            {
                if (r == q) v = 1.0 ;
                else if (r > q) v = F [(g1 + r) + (I64) cols [q] * ld] ;
            }
            Vs [tx][q] = v ;
        }
        __syncthreads () ;
        const I32 r = r0 + tx ;
#pragma unroll
        for (int a = 0 ; a < 4 ; a++)
        {
            const int c = ty + 8 * a ;
            if (c < ncol && r < mr)
            {
                double s = 0 ;
#pragma unroll 8
                for (int q = 0 ; q < PANEL_MAX ; q++) s += Vs [tx][q] * Ws [q][c] ;
                F [(g1 + r) + (I64) (c0 + c) * ld] -= s ;
            }
        }
        __syncthreads () ;
    }
}

} // namespace stmqr
