// kernels_update.cuh -- trailing-matrix update of the blocked Householder QR on FP64 tensor
// cores:  C := (I - V T V')' C = C - V (T' (V' C))   (dlarfb 'L','T','F','C'; the reference's
// qr_larftb method QR_QTX, SparseQR_factorize.c:1851-1882).
//
// V (mr x nv, nv <= 32) is the unit lower trapezoidal block of Householder vectors that the
// panel kernel left in columns cols[0..nv) of the front, rows g1 .. tend-1; T is the nv x nv
// upper triangular factor from dlarft.  grid = (active fronts, column tiles of NC columns).
// One CTA owns NC columns of C over ALL mr rows, so C is read from HBM once for W = V'C, again
// (normally from L2) for the rank-nv update, and written once.
//
// Both contractions run as mma.sync.m8n8k4.f64 (DMMA; tcgen05.mma has no FP64 kind):
//   phase 1  W(q,c) = sum_r V(r,q) C(r,c)     M = 32 (q), N = NC (c), K = rows, K split over the
//            8 warps of the CTA (each warp owns 8 rows of every 64-row chunk and all 4 x NC/8
//            accumulator tiles), partial sums reduced in a fixed warp order (deterministic)
//   phase 2  W := T' W  (tiny, plain FMA)
//   phase 3  C(r,c) -= sum_q V(r,q) W(q,c)    M = rows, N = NC, K = 32; each warp owns 8 rows of the
//            chunk, keeps the whole W tile in registers as B fragments, and read-modify-writes C
//            straight from the accumulator layout.
// Shared-memory tiles are column-major with a column stride = 4 (mod 16) doubles, which makes every
// A/B fragment load (8 rows/cols x 4 k per warp) bank-conflict free.
#pragma once
#include "engine.cuh"
#include "kernels_panel.cuh"

namespace stmqr {

constexpr int UPD_RC = 64 ;             // rows per chunk
constexpr int UPD_LDS = UPD_RC + 4 ;    // column stride of the V and C tiles (68 = 4 mod 16)
constexpr int UPD_LDW = PANEL_MAX + 4 ; // column stride of the W tile (36 = 4 mod 16)

__device__ __forceinline__ void dmma_m8n8k4 (double &d0, double &d1, const double a, const double b)
{
    asm volatile ("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d" (d0), "+d" (d1) : "d" (a), "d" (b)) ;
}

// load rows [r0, r0+64) of the unit lower trapezoidal V into Vs (column-major, stride UPD_LDS)
__device__ __forceinline__ void load_V_chunk (double *Vs, const double *__restrict__ F, const I64 ld,
    const I32 g1, const I32 mr, const I32 nv, const I32 *cols, const I32 r0, const int tid)
{
#pragma unroll
    for (int a = 0 ; a < (PANEL_MAX * UPD_RC) / 256 ; a++)
    {
        const int e = tid + a * 256 ;
        const int q = e >> 6, rr = e & 63 ;
        const I32 r = r0 + rr ;
        double v = 0 ;
        if (q < nv && r < mr)
        {
            if (r == q) v = 1.0 ;
            else if (r > q) v = F [(g1 + r) + (I64) cols [q] * ld] ;
        }
        Vs [q * UPD_LDS + rr] = v ;
    }
}

template <int NC>
__global__ void __launch_bounds__ (256) k_update_dmma (LevelArgs L, DSym S, DNum N, I32 cbeg, I32 cend,
    I32 parity)
{
    constexpr int NT = NC / 8 ;                     // accumulator tiles along the columns
    extern __shared__ double sm [] ;
    double *Vs = sm ;                               // [32][UPD_LDS]
    double *Cs = Vs + PANEL_MAX * UPD_LDS ;         // [NC][UPD_LDS]
    double *Ws = Cs + NC * UPD_LDS ;                // [NC][UPD_LDW]   W(q,c) at Ws[c*LDW + q]
    double *Ts = Ws + NC * UPD_LDW ;                // [32][33]        T(j,i) at Ts[j*33 + i]
    __shared__ I32 cols [PANEL_MAX] ;

    const I32 slot = blockIdx.x ;
    const I32 slotp = parity * L.count + slot ;
    const I32 nv = N.pnl_nv [slotp] ;
    if (nv == 0) return ;
    const I32 f = L.fronts [slot] ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    const I32 c0 = cbeg + blockIdx.y * NC ;
    const I32 clim = min (cend, fn) ;
    if (c0 >= clim) return ;
    const I32 ncol = min (NC, clim - c0) ;
    const I64 ld = N.Hm [f] ;
    double *F = N.F + S.Foff [f] ;
    const I32 g1 = N.pnl_g1 [slotp], tend = N.pnl_tend [slotp] ;
    const I32 mr = tend - g1 ;
    const int tid = threadIdx.x ;
    const int lane = tid & 31, w = tid >> 5 ;
    const int grp = lane >> 2, tig = lane & 3 ;     // mma fragment coordinates

    if (tid < PANEL_MAX) cols [tid] = (tid < nv) ? N.pnl_cols [slotp * PANEL_MAX + tid] : 0 ;
    {
        const double *Tg = N.Tws + (I64) slotp * (PANEL_MAX * PANEL_MAX) ;
        for (int e = tid ; e < PANEL_MAX * PANEL_MAX ; e += 256)
        {
            const int j = e % PANEL_MAX, i = e / PANEL_MAX ;
            Ts [j * (PANEL_MAX + 1) + i] = (j < nv && i < nv) ? Tg [j + i * PANEL_MAX] : 0.0 ;
        }
    }
    __syncthreads () ;

    // ---- phase 1: W = V' C, K (rows) split over the warps ------------------------------------
    double acc [4][NT][2] ;
#pragma unroll
    for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
        for (int ni = 0 ; ni < NT ; ni++) { acc [mi][ni][0] = 0 ; acc [mi][ni][1] = 0 ; }

    for (I32 r0 = 0 ; r0 < mr ; r0 += UPD_RC)
    {
        load_V_chunk (Vs, F, ld, g1, mr, nv, cols, r0, tid) ;
#pragma unroll
        for (int a = 0 ; a < (NC * UPD_RC) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            const int c = e >> 6, rr = e & 63 ;
            const I32 r = r0 + rr ;
            double cv = 0 ;
            if (c < ncol && r < mr) cv = F [(g1 + r) + (I64) (c0 + c) * ld] ;
            Cs [c * UPD_LDS + rr] = cv ;
        }
        __syncthreads () ;
#pragma unroll
        for (int ks = 0 ; ks < 2 ; ks++)
        {
            const int rr = w * 8 + ks * 4 + tig ;
            double af [4], bf [NT] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++) af [mi] = Vs [(mi * 8 + grp) * UPD_LDS + rr] ;
#pragma unroll
            for (int ni = 0 ; ni < NT ; ni++) bf [ni] = Cs [(ni * 8 + grp) * UPD_LDS + rr] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
                for (int ni = 0 ; ni < NT ; ni++)
                    dmma_m8n8k4 (acc [mi][ni][0], acc [mi][ni][1], af [mi], bf [ni]) ;
        }
        __syncthreads () ;
    }
    // deterministic reduction of the 8 warps' partial W into Ws (fixed order 0..7)
    for (int ww = 0 ; ww < 8 ; ww++)
    {
        if (w == ww)
        {
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
                for (int ni = 0 ; ni < NT ; ni++)
#pragma unroll
                    for (int e = 0 ; e < 2 ; e++)
                    {
                        const int q = mi * 8 + grp, c = ni * 8 + tig * 2 + e ;
                        double *dst = Ws + c * UPD_LDW + q ;
                        *dst = (ww == 0) ? acc [mi][ni][e] : (*dst + acc [mi][ni][e]) ;
                    }
        }
        __syncthreads () ;
    }
    // ---- phase 2: W = T' W ---------------------------------------------------------------------
    {
        double w2 [(NC * PANEL_MAX) / 256] ;
#pragma unroll
        for (int a = 0 ; a < (NC * PANEL_MAX) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            const int i = e & 31, c = e >> 5 ;
            double s = 0 ;
            for (int j = 0 ; j <= i ; j++) s += Ts [j * (PANEL_MAX + 1) + i] * Ws [c * UPD_LDW + j] ;
            w2 [a] = s ;
        }
        __syncthreads () ;
#pragma unroll
        for (int a = 0 ; a < (NC * PANEL_MAX) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            Ws [(e >> 5) * UPD_LDW + (e & 31)] = w2 [a] ;
        }
        __syncthreads () ;
    }
    // ---- phase 3: C -= V W; W as B fragments in registers ------------------------------------------
    double wf [8][NT] ;
#pragma unroll
    for (int ks = 0 ; ks < 8 ; ks++)
#pragma unroll
        for (int ni = 0 ; ni < NT ; ni++) wf [ks][ni] = Ws [(ni * 8 + grp) * UPD_LDW + ks * 4 + tig] ;

    for (I32 r0 = 0 ; r0 < mr ; r0 += UPD_RC)
    {
        load_V_chunk (Vs, F, ld, g1, mr, nv, cols, r0, tid) ;
        __syncthreads () ;
        double d [NT][2] ;
#pragma unroll
        for (int ni = 0 ; ni < NT ; ni++) { d [ni][0] = 0 ; d [ni][1] = 0 ; }
#pragma unroll
        for (int ks = 0 ; ks < 8 ; ks++)
        {
            const double af = Vs [(ks * 4 + tig) * UPD_LDS + w * 8 + grp] ;
#pragma unroll
            for (int ni = 0 ; ni < NT ; ni++) dmma_m8n8k4 (d [ni][0], d [ni][1], af, wf [ks][ni]) ;
        }
        const I32 r = r0 + w * 8 + grp ;
        if (r < mr)
        {
#pragma unroll
            for (int ni = 0 ; ni < NT ; ni++)
#pragma unroll
                for (int e = 0 ; e < 2 ; e++)
                {
                    const int c = ni * 8 + tig * 2 + e ;
                    if (c < ncol) F [(g1 + r) + (I64) (c0 + c) * ld] -= d [ni][e] ;
                }
        }
        __syncthreads () ;
    }
}

template <int NC> constexpr size_t update_smem_bytes ()
{
    return sizeof (double) * (PANEL_MAX * UPD_LDS + NC * UPD_LDS + NC * UPD_LDW + PANEL_MAX * (PANEL_MAX + 1)) ;
}

} // namespace stmqr
