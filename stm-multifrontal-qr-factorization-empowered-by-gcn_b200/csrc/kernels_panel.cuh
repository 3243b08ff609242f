// kernels_panel.cuh -- panel factorization of the staircase Householder QR (qr_front,
// SparseQR_factorize.c:1383-1618: the dlarfg / dlarf column loop, plus dlarft for the block
// reflector of the panel) for all active fronts of one etree level.
//
// B200 design: one thread-block CLUSTER per front.  The rows of the panel that can still change
// (pivot row g .. staircase end of the last panel column) are split into CS contiguous slabs, one
// per CTA of the cluster, and each CTA keeps its slab (rows x <=32 columns) in shared memory for
// the whole panel, so the column loop never touches L2/HBM.  Per Householder column there is
// exactly ONE cluster-wide exchange: every CTA publishes the partial dot products of the current
// column with all panel columns over its rows (the norm, the dlarf inner products and the
// V'V entries that dlarft needs, all from the same pass) and the owner of the pivot row publishes
// that row; every CTA then reads the CS records through distributed shared memory, derives the
// same beta/tau/dead decision and updates its own rows.  Slabs that do not fit in shared memory
// (fronts with tens of thousands of rows) run the same code in place on the front (L2 resident).
//
// Semantics that must match the reference bit for bit (integer outputs):
//   t = max (g+1, Stair[k])                                                    :1460
//   dead pivot (k < ntol, |F(g,k)| <= tol): zero F(g:m-1,k), Stair=0, Tau=0, Rdead=1   :1495-1528
//   out of rows (g >= m): remaining pivots dead, remaining columns Stair = m   :1444-1458
//   rank = g sampled after pivot column npiv-1                                 :1604-1608
// Blocking (panel width, when T is applied) is performance-only.
#pragma once
#include <cooperative_groups.h>
#include "engine.cuh"

namespace stmqr {

namespace cg = cooperative_groups ;

#ifdef STMQR_PANEL_TIMING
#define PT_DECL long long pt_t0 = clock64 (), pt_acc [8] = {0,0,0,0,0,0,0,0}
#define PT_MARK(j) do { long long pt_t1 = clock64 () ; pt_acc [j] += pt_t1 - pt_t0 ; pt_t0 = pt_t1 ; } while (0)
#define PT_FLUSH(base) do { if (threadIdx.x == 0 && leader && L.count == 1) { for (int j_ = 0 ; j_ < 8 ; j_++) \
    atomicAdd (N.dbg + (base) + j_, (unsigned long long) pt_acc [j_]) ; } } while (0)
#else
#define PT_DECL
#define PT_MARK(j)
#define PT_FLUSH(base)
#endif

struct LevelArgs
{
    const I32 *fronts ;     // fronts of the level, sorted by # columns descending
    I32 count ;
    double tol ;
    I64 ntol ;
    I32 flags ;             // bit 0: one column per exchange in the panel (A/B switch of the look-ahead)
} ;

// what one CTA publishes to its cluster for one Householder column
struct PanelXch
{
    double dot [PANEL_MAX] ;    // sum over my rows in (g,t) of F(i,k) * F(i,c), every panel column c
    double rowg [PANEL_MAX] ;   // F(g,c), published by the CTA whose slab holds the pivot row g
    double mx ;                 // max over my rows in (g,t) of |F(i,k)|
    double ssq2 ;               // rescaled sum of squares (rare under/overflow path)
} ;

// dlarft for the panel's live reflectors from their V'V entries (Gs), and the panel's outputs for
// the trailing update (T, live columns, row window), written by the leader CTA of the cluster.
template <bool BUILD_T>
__device__ __forceinline__ void panel_epilogue (const LevelArgs &L, const DSym &S, const DNum &N,
    const I32 slot, const I32 f, const I32 k2, const I32 parity, const bool leader, const I32 nv,
    const I32 g, const I32 g1, const bool out_of_rows, const double flops, double *Gs, double *Tsh,
    double *taus, I32 *cols, I32 *tq)
{
    const int tid = threadIdx.x, nt = blockDim.x ;
    const int lane = tid & 31, w = tid >> 5 ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    // ---- dlarft: T of the nv live reflectors of this panel (forward, columnwise) --------------
    const I32 slotp = parity * L.count + slot ;
    if (BUILD_T && leader && nv > 0 && k2 < fn)
    {
        for (int e = tid ; e < nv * nv ; e += nt)
        {
            const int j = e % nv, i = e / nv ;
            if (j < i) Tsh [j + i * (PANEL_MAX + 1)] = -taus [i] * Gs [j + i * (PANEL_MAX + 1)] ;
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
        double *Tg = N.Tws + (I64) slotp * (PANEL_MAX * PANEL_MAX) ;
        for (int e = tid ; e < nv * nv ; e += nt)
        {
            const int j = e % nv, i = e / nv ;
            Tg [j + i * PANEL_MAX] = (j <= i) ? Tsh [j + i * (PANEL_MAX + 1)] : 0.0 ;
        }
        for (int q = tid ; q < nv ; q += nt) N.pnl_cols [slotp * PANEL_MAX + q] = cols [q] ;
    }
    if (leader && tid == 0)
    {
        N.g [slot] = g ;
        N.pnl_g1 [slotp] = g1 ;
        N.pnl_nv [slotp] = (k2 < fn) ? nv : 0 ;
        N.pnl_tend [slotp] = (nv > 0) ? tq [nv-1] : g1 ;
        if (out_of_rows) N.done [slot] = 1 ;
        if (flops != 0) atomicAdd (N.flops, flops) ;
        if (nv > 0 && k2 < fn)
            atomicAdd (N.flops + 1, 4.0 * (double) (tq [nv-1] - g1) * (double) nv * (double) (fn - k2)) ;
    }
}

// Row-per-lane mapping (one warp per panel column): used in place on the front (global memory,
// coalesced along the rows) when the row slabs do not fit in shared memory.
template <bool INSMEM>
__device__ __forceinline__ void panel_columns (cg::cluster_group &cluster, double *__restrict__ P,
    const I64 ldp, const LevelArgs &L, const DSym &S, const DNum &N, const I32 slot, const I32 f,
    const I32 k1, const I32 k2, const I32 parity, const I32 lrow0, const I32 nloc, const I32 rbeg,
    const I32 RL, const I32 rend, PanelXch *xch, double *tot, double *rg, double *Gs, double *Tsh,
    double *taus, I32 *cols, I32 *tq)
{
    const unsigned CS = cluster.num_blocks (), cr = cluster.block_rank () ;
    const int tid = threadIdx.x, nt = blockDim.x ;
    const int lane = tid & 31, w = tid >> 5, nw = nt >> 5 ;
    const I32 col1 = S.Super [f], fp = S.Super [f+1] - col1 ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const I32 fm = N.Hm [f] ;
    const I64 ld = fm ;
    double *F = N.F + S.Foff [f] ;
    I32 *st = N.stair + p1 ;
    double *Tau = N.HTau + p1 ;
    char *Rdead = N.Rdead + col1 ;
    const I32 ntol = (I32) max ((I64) 0, min (L.ntol - (I64) col1, (I64) fp)) ;
    const double tol = L.tol ;
    const I32 np = k2 - k1 ;
    const bool leader = (cr == 0) ;

    I32 g = rbeg ;                  // the panel starts at the front's current pivot row
    const I32 g1 = g ;
    I32 nv = 0 ;
    double flops = 0 ;
    bool out_of_rows = false ;
    int step = 0 ;

    for (I32 k = k1 ; k < k2 ; k++)
    {
        if (g >= fm)
        {
            // no rows left: qr_front early exit (:1444-1458) for ALL remaining columns
            if (leader)
            {
                for (I32 kk = k + tid ; kk < fn ; kk += nt)
                {
                    if (kk < fp) { Rdead [kk] = 1 ; st [kk] = 0 ; }
                    else st [kk] = fm ;
                    Tau [kk] = 0 ;
                }
            }
            out_of_rows = true ;
            break ;
        }
        const I32 c = k - k1 ;
        const I32 t = max (g + 1, st [k]) ;
        // my rows strictly below the pivot row and inside the staircase: local [i0,i1)
        const I32 i0 = max (g + 1, lrow0) - lrow0 ;
        const I32 i1 = min (min (t, rend), lrow0 + nloc) - lrow0 ;
        const unsigned owner = (unsigned) ((g - rbeg) / RL) ;
        const I32 gi = g - lrow0 ;
        PanelXch &X = xch [step & 1] ;
        step++ ;

        // ---- one pass: dot products of column k with every panel column over my rows ----------
        const double *xc = P + (I64) c * ldp ;
        for (I32 cc = w ; cc < np ; cc += nw)
        {
            const double *yc = P + (I64) cc * ldp ;
            double s = 0, mx = 0 ;
            for (I32 i = i0 + lane ; i < i1 ; i += 32)
            {
                const double xv = xc [i] ;
                s += xv * yc [i] ;
                mx = fmax (mx, fabs (xv)) ;
            }
            s = warp_sum (s) ;
            if (cc == c) mx = warp_max (mx) ;
            if (lane == 0)
            {
                X.dot [cc] = s ;
                X.rowg [cc] = (cr == owner) ? yc [gi] : 0.0 ;
                if (cc == c) X.mx = mx ;
            }
        }
        cluster.sync () ;
        // ---- gather the CS records (same order in every CTA: bitwise identical decisions) -----
        if (tid < np)
        {
            double s = 0 ;
            for (unsigned r = 0 ; r < CS ; r++) s += cluster.map_shared_rank (&X, r)->dot [tid] ;
            tot [tid] = s ;
            rg [tid] = cluster.map_shared_rank (&X, owner)->rowg [tid] ;
        }
        else if (tid == PANEL_MAX)
        {
            double mx = 0 ;
            for (unsigned r = 0 ; r < CS ; r++) mx = fmax (mx, cluster.map_shared_rank (&X, r)->mx) ;
            tot [PANEL_MAX] = mx ;
        }
        __syncthreads () ;
        double ss = tot [c], mx = tot [PANEL_MAX] ;
        const double alpha = rg [c] ;
        if (mx > 0 && !(ss > 1e-280 && ss < 1e280))
        {
            // rare: rescale to avoid under/overflow of the sum of squares (dnrm2 semantics);
            // the decision is uniform over the cluster, so one more exchange is safe
            const double inv = 1.0 / mx ;
            double s2 = 0 ;
            for (I32 i = i0 + tid ; i < i1 ; i += nt) { const double v = xc [i] * inv ; s2 += v * v ; }
            double dummy = 0 ;
            block_sum_max (s2, dummy, tot + PANEL_MAX + 2) ;
            if (tid == 0) X.ssq2 = s2 ;
            cluster.sync () ;
            s2 = 0 ;
            for (unsigned r = 0 ; r < CS ; r++) s2 += cluster.map_shared_rank (&X, r)->ssq2 ;
            cluster.sync () ;                // X is reused two steps later: keep the step parity
            ss = s2 ;
        }
        else mx = 1.0 ;
        const double xnorm = mx * sqrt (ss) ;
        double beta = alpha, tau = 0, scale = 0 ;
        if (t - g > 1 && xnorm != 0 && hypot (alpha, xnorm) >= 1e-290)     // (underflowed noise: H = I)
        {
            beta = -copysign (hypot (alpha, xnorm), alpha) ;
            tau = (beta - alpha) / beta ;
            scale = 1.0 / (alpha - beta) ;
            if (!(fabs (tau) <= 2.0) || !(fabs (scale) < 1e300)) { beta = alpha ; tau = 0 ; scale = 0 ; }
        }
        const bool dead = (k < ntol) && (fabs (beta) <= tol) ;

        if (dead)
        {
            // zero F(g:m-1,k): my rows, and (leader) whatever lies below the panel's row window
            double *xw = P + (I64) c * ldp ;
            for (I32 i = max (g, lrow0) - lrow0 + tid ; i < nloc ; i += nt) xw [i] = 0.0 ;
            if (leader)
            {
                double *xg = F + (I64) k * ld ;
                for (I32 i = rend + tid ; i < fm ; i += nt) xg [i] = 0.0 ;
                if (tid == 0) { st [k] = 0 ; Tau [k] = 0 ; Rdead [k] = 1 ; }
            }
        }
        else
        {
            // ---- dlarf: apply H_k to the remaining columns of the panel (my rows) -------------
            if (tau != 0)
            {
                for (I32 cc = c + 1 + w ; cc < np ; cc += nw)
                {
                    double *yc = P + (I64) cc * ldp ;
                    const double wv = tau * (rg [cc] + scale * tot [cc]) ;
                    const double fct = scale * wv ;
                    for (I32 i = i0 + lane ; i < i1 ; i += 32) yc [i] -= xc [i] * fct ;
                    if (cr == owner && lane == 0) yc [gi] -= wv ;
                }
            }
            // V'V entries for dlarft: v_j' v_k = scale * (v_j' x) + v_j (g), j an earlier reflector
            if (tid < nv) Gs [tid + nv * (PANEL_MAX + 1)] = scale * tot [cols [tid] - k1] + rg [cols [tid] - k1] ;
            __syncthreads () ;
            // ---- finish column k: v = x * scale below the diagonal, beta on it ------------------
            double *xw = P + (I64) c * ldp ;
            if (tau != 0) for (I32 i = i0 + tid ; i < i1 ; i += nt) xw [i] *= scale ;
            if (cr == owner && tid == 0) xw [gi] = beta ;
            if (tid == 0)
            {
                if (leader) { Tau [k] = tau ; st [k] = t ; }
                cols [nv] = k ; tq [nv] = t ; taus [nv] = tau ;
            }
            flops += (double) (t - g) * (3.0 + 4.0 * (double) (fn - k - 1)) ;
            nv++ ;
            g++ ;
        }
        if (k == fp - 1 && leader && tid == 0) N.rank [f] = g ;
        __syncthreads () ;
    }

    panel_epilogue<true> (L, S, N, slot, f, k2, parity, leader, nv, g, g1, out_of_rows, flops, Gs, Tsh, taus, cols, tq) ;
}

// Column-per-lane mapping for a slab in shared memory: lane = panel column, warp w owns the slab
// rows i = w (mod NW).  A row is only ever written by its owner warp, so the whole column step
// (partial dots -> one block barrier -> [cluster exchange] -> rank-1 update + scaling of the owned
// rows) needs ONE __syncthreads; the per-step records are double buffered by step parity.
// ldp is odd: the 32 lanes of a warp hit 32 distinct 8-byte banks.  Nothing in the column loop
// touches global memory: the staircase of the panel's columns is staged in shared memory and the
// per-column outputs (Stair, Tau, Rdead) are written once at the end.
constexpr int PANEL_NW_MAX = 16 ;       // 512 threads
constexpr int PANEL_XR_CTAS = 8 ;       // largest cluster of k_panel_cluster (portable size)
constexpr int PANEL_XR_DOUBLES = 2 * PANEL_XR_CTAS * 2 * PANEL_MAX + 2 * 2 * PANEL_MAX ;   // xrec [2][8][64] + xrows [2][64]

// per-CTA scratch carved from dynamic shared memory after the slab (sized by the # of warps)
__host__ __device__ constexpr int panel_scratch_doubles (int nw)
{
    return 4 * nw * PANEL_MAX           // part [2][nw][64]  per-warp partial dots of columns k and k+1
        + 4 * PANEL_MAX                 // prow [2][64]      rows g and g+1, published by their owner warps
        + 2 * nw                        // red  [2][nw]      block reductions of the rescale path
        + PANEL_MAX * (PANEL_MAX + 1)   // Gs   [32][33]     V'V, then T in place (dlarft)
        + 4 * PANEL_MAX ;               // taus, stair in, stair out, flags (as doubles/ints)
}

// Exchange between the CTAs of one front when they are NOT a cluster (k_panel_grid: fronts too tall
// for 8 shared-memory slabs): records in global memory (L2), a monotone arrival counter as barrier.
// All CTAs of the front are resident at the same time (<= 1 CTA per SM, grid <= # SMs).
struct GridComm
{
    double *rec ;           // [2][G][64]  per step parity and CTA: dot[32], rowg[32]
    double *red ;           // [2][G]      rare rescale path
    unsigned *ctr ;         // 4 arrival counters of this front, 32 words apart
    unsigned G ;            // CTAs taking part
    unsigned cr ;           // my rank
    unsigned epoch ;        // barriers passed
    int4 *ll ;              // [2][G][128] the records (a | b | row g | row g+1) as 16-byte lines {lo, tag, hi, tag}: data and
                            // flag travel together (8-byte stores are atomic), so a step needs no
                            // barrier: readers poll the lines until the tag of the step shows up
    unsigned tagbase ;      // launch sequence number * 64
} ;
__device__ __forceinline__ void ll_store (int4 *p, const double v, const unsigned tag)
{
    const unsigned lo = (unsigned) __double2loint (v), hi = (unsigned) __double2hiint (v) ;
    asm volatile ("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l" (p), "r" (lo), "r" (tag), "r" (hi), "r" (tag) : "memory") ;
}
__device__ __forceinline__ bool ll_load (const int4 *p, const unsigned tag, double &v)
{
    unsigned lo, f1, hi, f2 ;
    asm volatile ("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r" (lo), "=r" (f1), "=r" (hi), "=r" (f2) : "l" (p) : "memory") ;
    v = __hiloint2double ((int) hi, (int) lo) ;
    return (f1 == tag) && (f2 == tag) ;
}
__device__ __forceinline__ unsigned ld_relaxed_u32 (const unsigned *p)
{
    unsigned v ;
    asm volatile ("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r" (v) : "l" (p) : "memory") ;
    return v ;
}
__device__ __forceinline__ void red_add_u32 (unsigned *p, unsigned v)
{
    asm volatile ("red.relaxed.gpu.global.add.u32 [%0], %1;" :: "l" (p), "r" (v) : "memory") ;
}
__device__ __forceinline__ void grid_barrier (GridComm &gc)
{
    __syncthreads () ;
    gc.epoch++ ;
    if (threadIdx.x == 0)
    {
        // arrivals are spread over 4 counters in different L2 sectors (atomics on one address
        // serialise at ~27 cycles each); the waiter polls their sum
        // (relaxed polling, one fence on each side: acquire loads would invalidate L1 every time)
        __threadfence () ;
        red_add_u32 (gc.ctr + 32 * (gc.cr & 3), 1u) ;
        const unsigned target = gc.epoch * gc.G ;
        for ( ; ; )
        {
            const unsigned a = ld_relaxed_u32 (gc.ctr), b = ld_relaxed_u32 (gc.ctr + 32),
                c = ld_relaxed_u32 (gc.ctr + 64), d = ld_relaxed_u32 (gc.ctr + 96) ;
            if (a + b + c + d >= target) break ;
        }
        __threadfence () ;
    }
    __syncthreads () ;
}
constexpr int GRID_CTR_STRIDE = 512 ;      // unsigned per front: start barrier [0..128), step barrier [128..256), exit [256]

// all-reduce of one double over the CTA's warps and then over the cluster (rare rescale path)
template <int NW, bool ISMAX, bool GRID>
__device__ __forceinline__ double panel_allreduce (cg::cluster_group &cluster, GridComm &gc, const unsigned ECS, double v,
    double *red, PanelXch &X, const int par)
{
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5 ;
    if (lane == 0) red [w] = v ;
    __syncthreads () ;
    v = red [0] ;
#pragma unroll
    for (int ww = 1 ; ww < NW ; ww++) v = ISMAX ? fmax (v, red [ww]) : (v + red [ww]) ;
    __syncthreads () ;
    if (ECS > 1 && GRID)
    {
        if (tid == 0) __stcg (gc.red + par * gc.G + gc.cr, v) ;
        grid_barrier (gc) ;
        v = __ldcg (gc.red + par * gc.G) ;
        for (unsigned r = 1 ; r < ECS ; r++)
        {
            const double u = __ldcg (gc.red + par * gc.G + r) ;
            v = ISMAX ? fmax (v, u) : (v + u) ;
        }
        grid_barrier (gc) ;
    }
    else if (ECS > 1)
    {
        if (tid == 0) X.ssq2 = v ;
        cluster.sync () ;
        v = cluster.map_shared_rank (&X, 0)->ssq2 ;
        for (unsigned r = 1 ; r < ECS ; r++)
        {
            const double u = cluster.map_shared_rank (&X, r)->ssq2 ;
            v = ISMAX ? fmax (v, u) : (v + u) ;
        }
        cluster.sync () ;
    }
    return v ;
}

// Two columns per exchange (look-ahead inside the panel).  The per-column cost of this loop is latency:
// one block barrier, one cluster / L2 exchange and a sqrt + two divisions, all serial.  So the pass
// that computes the dots of column k with every panel column (a) also computes the dots of column k+1
// with every panel column (b), both over the rows from g+2 on, and the owners of rows g and g+1
// publish those rows.  After ONE exchange every thread can derive H_k exactly as before (the row g+1
// terms are added from the published row, no cancellation), and then H_{k+1} from the Gram entries:
//   x' = x_{k+1} - v_k w           (w = tau_k v_k' x_{k+1})
//   ||x'(g+2:)||^2 = b_{k+1} - 2 w s a_{k+1} + (w s)^2 a_k              (s = 1/(alpha_k - beta_k))
//   x'(g+2:)' y'_c = (b_c - s w_c a_{k+1}) - w s (a_c - s w_c a_k)      (y'_c = y_c - v_k w_c)
// The down-dated norm is only trusted when it keeps at least LA_THETA of the magnitude of its terms
// (cancellation bounded: relative error <= eps / LA_THETA, and the error of the second reflector's
// inner products stays a small multiple of eps ||y||, i.e. the update remains backward stable with a
// constant 1/sqrt(LA_THETA)); otherwise, or when column k+1 looks dead, or in any of dlarfg's rare
// scaling paths, the step handles column k alone and column k+1 gets its own direct pass -- exactly
// the one-column algorithm.  Both reflectors are applied in one sweep: y <- ya y - fa x_k - fb x_{k+1}.
constexpr double LA_THETA = 1.0 / 32.0 ;

template <int NW, bool GRID>
__device__ __forceinline__ void panel_columns_smem (cg::cluster_group &cluster, GridComm &gc, const unsigned ECS,
    const I32 slab_cap, const I32 ldp, const LevelArgs &L, const DSym &S, const DNum &N,
    const I32 slot, const I32 f, const I32 k1, const I32 k2, const I32 parity, const I32 lrow0,
    const I32 nloc, const I32 rbeg, const I32 RL, const I32 rend, PanelXch *xch, double *xrec, double *xrows,
    I32 *cols, I32 *tq)
{
    extern __shared__ double slab [] ;
    constexpr int PM2 = 2 * PANEL_MAX ;
    double *const P = slab ;
    double *const part = slab + slab_cap ;              // [2][NW][64]  per-warp partial dots: a (0..31) | b (32..63)
    double *const prow = part + 2 * NW * PM2 ;          // [2][64]      row g | row g+1, published by their owner warps
    double *const red = prow + 2 * PM2 ;
    double *const Gs = red + 2 * NW ;
    double *const taus = Gs + PANEL_MAX * (PANEL_MAX + 1) ;
    I32 *const stl = (I32 *) (taus + PANEL_MAX) ;       // [32] staircase of the panel's columns (in)
    I32 *const sto = stl + PANEL_MAX ;                  // [32] new Stair, -1 = column not processed
    double *const tauo = taus + 2 * PANEL_MAX ;         // [32] Tau out

    const unsigned cr = (ECS > 1) ? (GRID ? gc.cr : cluster.block_rank ()) : 0 ;
    const int tid = threadIdx.x ;
    constexpr int nt = NW * 32 ;
    const int lane = tid & 31, w = tid >> 5 ;
    const I32 col1 = S.Super [f], fp = S.Super [f+1] - col1 ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const I32 fm = N.Hm [f] ;
    const I64 ld = fm ;
    double *F = N.F + S.Foff [f] ;
    I32 *st = N.stair + p1 ;
    double *Tau = N.HTau + p1 ;
    char *Rdead = N.Rdead + col1 ;
    const I32 ntol = (I32) max ((I64) 0, min (L.ntol - (I64) col1, (I64) fp)) ;
    const double tol = L.tol ;
    const I32 np = k2 - k1 ;
    const bool leader = (cr == 0) ;
    const bool mycol = (lane < np) ;
    const bool la_on = (L.flags & 1) == 0 ;
    double *yc = P + (I64) (mycol ? lane : 0) * ldp ;       // my column of the slab

    if (tid < PANEL_MAX)
    {
        stl [tid] = (tid < np) ? st [k1 + tid] : 0 ;
        sto [tid] = -1 ;
    }
    __syncthreads () ;

    I32 g = rbeg ;
    const I32 g1 = g ;
    I32 nv = 0 ;
    I32 myq = -1 ;                  // index of my column among the live reflectors of this panel
    double flops = 0 ;
    bool out_of_rows = false ;
    int step = 0 ;
    I32 kstop = k2 ;
    PT_DECL ;

    for (I32 k = k1 ; k < k2 ; )
    {
        PT_MARK (7) ;
        if (g >= fm) { out_of_rows = true ; kstop = k ; break ; }
        const I32 c = k - k1, c1 = c + 1 ;
        const I32 t = max (g + 1, stl [c]) ;
        // may column k+1 ride on this exchange?  (it needs a pivot row g+1 inside the front)
        const bool la = la_on && (k + 1 < k2) && (g + 1 < fm) ;
        const I32 t1 = la ? max (g + 2, stl [c1]) : t ;
        // slab rows from row g+2 on inside the staircase of column k: local [j0,e0); of column k+1: [j0,e1)
        const I32 j0 = max (g + 2, lrow0) - lrow0 ;
        const I32 e0 = min (min (t, rend), lrow0 + nloc) - lrow0 ;
        const I32 e1 = min (min (t1, rend), lrow0 + nloc) - lrow0 ;
        const bool have1 = (g + 1 < rend) ;                 // row g+1 lies in the panel's row window
        const unsigned owner0 = (ECS > 1) ? (unsigned) ((g - rbeg) / RL) : 0u ;
        const unsigned owner1 = (ECS > 1 && have1) ? (unsigned) ((g + 1 - rbeg) / RL) : 0u ;
        const I32 gi0 = g - lrow0, gi1 = g + 1 - lrow0 ;
        const bool own_g0 = (cr == owner0) && ((gi0 & (NW - 1)) == w) ;             // my warp owns the pivot row
        const bool own_g1 = have1 && (cr == owner1) && ((gi1 & (NW - 1)) == w) ;    // ... the row below it
        const int par = step & 1 ;
        step++ ;
        const double *x0 = P + (I64) c * ldp ;
        const double *x1 = P + (I64) (la ? c1 : c) * ldp ;
        // first row of my warp at or after j0
        const I32 ifirst = j0 + ((w - j0) & (NW - 1)) ;

        // ---- partial dots of columns k and k+1 with my column over my warp's rows -----------------
        double sA, sB = 0 ;
        if (la)
        {
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0 ;
            I32 i = ifirst ;
            for ( ; i + 3 * NW < e0 ; i += 4 * NW)
            {
                const double y0 = yc [i], y1 = yc [i + NW], y2 = yc [i + 2*NW], y3 = yc [i + 3*NW] ;
                a0 = fma (x0 [i], y0, a0) ;             a1 = fma (x0 [i + NW], y1, a1) ;
                a2 = fma (x0 [i + 2*NW], y2, a2) ;      a3 = fma (x0 [i + 3*NW], y3, a3) ;
                b0 = fma (x1 [i], y0, b0) ;             b1 = fma (x1 [i + NW], y1, b1) ;
                b2 = fma (x1 [i + 2*NW], y2, b2) ;      b3 = fma (x1 [i + 3*NW], y3, b3) ;
            }
            for ( ; i < e0 ; i += NW)
            {
                const double y0 = yc [i] ;
                a0 = fma (x0 [i], y0, a0) ; b0 = fma (x1 [i], y0, b0) ;
            }
            // rows below the staircase of column k: column k is zero there
            for ( ; i + 3 * NW < e1 ; i += 4 * NW)
            {
                b0 = fma (x1 [i], yc [i], b0) ;                 b1 = fma (x1 [i + NW], yc [i + NW], b1) ;
                b2 = fma (x1 [i + 2*NW], yc [i + 2*NW], b2) ;   b3 = fma (x1 [i + 3*NW], yc [i + 3*NW], b3) ;
            }
            for ( ; i < e1 ; i += NW) b0 = fma (x1 [i], yc [i], b0) ;
            sA = (a0 + a1) + (a2 + a3) ;
            sB = (b0 + b1) + (b2 + b3) ;
        }
        else
        {
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0 ;
            I32 i = ifirst ;
            for ( ; i + 7 * NW < e0 ; i += 8 * NW)
            {
                a0 = fma (x0 [i], yc [i], a0) ;                 a1 = fma (x0 [i + NW], yc [i + NW], a1) ;
                a2 = fma (x0 [i + 2*NW], yc [i + 2*NW], a2) ;   a3 = fma (x0 [i + 3*NW], yc [i + 3*NW], a3) ;
                a4 = fma (x0 [i + 4*NW], yc [i + 4*NW], a4) ;   a5 = fma (x0 [i + 5*NW], yc [i + 5*NW], a5) ;
                a6 = fma (x0 [i + 6*NW], yc [i + 6*NW], a6) ;   a7 = fma (x0 [i + 7*NW], yc [i + 7*NW], a7) ;
            }
            if (i + 3 * NW < e0)
            {
                a0 = fma (x0 [i], yc [i], a0) ;                 a1 = fma (x0 [i + NW], yc [i + NW], a1) ;
                a2 = fma (x0 [i + 2*NW], yc [i + 2*NW], a2) ;   a3 = fma (x0 [i + 3*NW], yc [i + 3*NW], a3) ;
                i += 4 * NW ;
            }
            if (i < e0) a4 = fma (x0 [i], yc [i], a4) ;
            if (i + NW < e0) a5 = fma (x0 [i + NW], yc [i + NW], a5) ;
            if (i + 2 * NW < e0) a6 = fma (x0 [i + 2*NW], yc [i + 2*NW], a6) ;
            sA = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)) ;
        }
        if (lane == c && sA == 0.0)
        {
            // ||x||^2 = 0 over my rows: exactly zero entries, or squares that underflow?  Mark the
            // second case with a positive value below every regular sum of squares, so that the total
            // is 0 if and only if the sub-column is exactly zero (no extra exchange needed later)
            bool nz = false ;
            for (I32 i = ifirst ; i < e0 ; i += NW) nz |= (x0 [i] != 0.0) ;
            if (nz) sA = 1e-300 ;
        }
        PT_MARK (0) ;
        part [(par * NW + w) * PM2 + lane] = mycol ? sA : 0.0 ;
        if (la) part [(par * NW + w) * PM2 + PANEL_MAX + lane] = mycol ? sB : 0.0 ;
        if (own_g0) prow [par * PM2 + lane] = mycol ? yc [gi0] : 0.0 ;
        if (own_g1) prow [par * PM2 + PANEL_MAX + lane] = mycol ? yc [gi1] : 0.0 ;
        __syncthreads () ;
        PT_MARK (1) ;
        {
            // FP64 adds have a long dependent latency: sum the NW partials as a tree
            double pv [NW], pb [NW] ;
#pragma unroll
            for (int ww = 0 ; ww < NW ; ww++) pv [ww] = part [(par * NW + ww) * PM2 + lane] ;
            if (la)
            {
#pragma unroll
                for (int ww = 0 ; ww < NW ; ww++) pb [ww] = part [(par * NW + ww) * PM2 + PANEL_MAX + lane] ;
            }
            else
            {
#pragma unroll
                for (int ww = 0 ; ww < NW ; ww++) pb [ww] = 0.0 ;
            }
#pragma unroll
            for (int h = NW / 2 ; h > 0 ; h >>= 1)
#pragma unroll
                for (int ww = 0 ; ww < h ; ww++) { pv [ww] += pv [ww + h] ; pb [ww] += pb [ww + h] ; }
            sA = pv [0] ; sB = pb [0] ;
        }
        double r0 = (cr == owner0) ? prow [par * PM2 + lane] : 0.0 ;
        double r1 = (have1 && cr == owner1) ? prow [par * PM2 + PANEL_MAX + lane] : 0.0 ;
        if (ECS > 1 && GRID)
        {
            // one record per CTA in global memory (L2): a | b | row g | row g+1, summed in CTA order by
            // every CTA: bitwise identical decisions everywhere
            int4 *rec = gc.ll + (I64) par * gc.G * 128 ;
            const unsigned tag = gc.tagbase + (unsigned) step ;
            if (w == 0)
            {
                ll_store (rec + cr * 128 + lane, sA, tag) ;
                if (cr == owner0) ll_store (rec + cr * 128 + 64 + lane, r0, tag) ;
            }
            else if (w == 1)
            {
                if (la) ll_store (rec + cr * 128 + 32 + lane, sB, tag) ;
                if (have1 && cr == owner1) ll_store (rec + cr * 128 + 96 + lane, r1, tag) ;
            }
            __syncthreads () ;          // part[par] is overwritten below: everybody has read its partials
            {
                // warp w gathers the records w, w+NW, ...: all loads in flight at once, polled until the
                // tag of this step shows up; summed in a fixed order, the same in every CTA
                static_assert (!GRID || NW * 10 >= 148, "records per warp") ;
                double v [10], vb [10] ;
                bool ok ;
                do
                {
                    ok = true ;
#pragma unroll
                    for (int j = 0 ; j < 10 ; j++)
                    {
                        const unsigned r = (unsigned) (w + NW * j) ;
                        v [j] = 0.0 ; vb [j] = 0.0 ;
                        if (r < ECS)
                        {
                            ok &= ll_load (rec + r * 128 + lane, tag, v [j]) ;
                            if (la) ok &= ll_load (rec + r * 128 + 32 + lane, tag, vb [j]) ;
                        }
                    }
                } while (!__all_sync (STMQR_FULL_MASK, ok)) ;
                if (w == 0)
                {
                    double rv ;
                    while (!__all_sync (STMQR_FULL_MASK, ll_load (rec + owner0 * 128 + 64 + lane, tag, rv))) { }
                    prow [par * PM2 + lane] = rv ;
                }
                else if (w == 1)
                {
                    double rv = 0.0 ;
                    if (have1) while (!__all_sync (STMQR_FULL_MASK, ll_load (rec + owner1 * 128 + 96 + lane, tag, rv))) { }
                    prow [par * PM2 + PANEL_MAX + lane] = rv ;
                }
                part [(par * NW + w) * PM2 + lane] = (((v [0] + v [1]) + (v [2] + v [3])) + ((v [4] + v [5]) + (v [6] + v [7])))
                    + (v [8] + v [9]) ;
                part [(par * NW + w) * PM2 + PANEL_MAX + lane] = (((vb [0] + vb [1]) + (vb [2] + vb [3])) + ((vb [4] + vb [5]) + (vb [6] + vb [7])))
                    + (vb [8] + vb [9]) ;
            }
            __syncthreads () ;
            {
                double pv [NW], pb [NW] ;
#pragma unroll
                for (int ww = 0 ; ww < NW ; ww++)
                {
                    pv [ww] = part [(par * NW + ww) * PM2 + lane] ;
                    pb [ww] = part [(par * NW + ww) * PM2 + PANEL_MAX + lane] ;
                }
#pragma unroll
                for (int h = NW / 2 ; h > 0 ; h >>= 1)
#pragma unroll
                    for (int ww = 0 ; ww < h ; ww++) { pv [ww] += pv [ww + h] ; pb [ww] += pb [ww + h] ; }
                sA = pv [0] ; sB = pb [0] ;
            }
            r0 = prow [par * PM2 + lane] ; r1 = prow [par * PM2 + PANEL_MAX + lane] ;
        }
        else if (ECS > 1)
        {
            // PUSH through distributed shared memory: warp w stores this CTA's record (a | b, and the rows
            // g, g+1 if it holds them) straight into the shared memory of the CTAs w, w+NW, ... of the
            // cluster -- remote stores are fire-and-forget -- ONE cluster barrier makes all records
            // visible, and then every thread sums the ECS records from its OWN shared memory in rank
            // order: bitwise identical totals everywhere, no remote-load latency, no re-publishing.
            // (A record buffer is reused two steps later; by then every CTA has passed the barrier of
            // the step in between, which it reaches only after reading this step's records.)
            double *rec = xrec + par * (PANEL_XR_CTAS * PM2) ;      // [ECS][64]
            double *rws = xrows + par * PM2 ;                       // row g | row g+1
            for (unsigned r = (unsigned) w ; r < ECS ; r += NW)
            {
                double *dst = cluster.map_shared_rank (rec, r) + cr * PM2 ;
                dst [lane] = sA ;
                dst [PANEL_MAX + lane] = sB ;
                double *drw = cluster.map_shared_rank (rws, r) ;
                if (cr == owner0) drw [lane] = r0 ;
                if (have1 && cr == owner1) drw [PANEL_MAX + lane] = r1 ;
            }
            cluster.sync () ;
            {
                double dv [PANEL_XR_CTAS], db [PANEL_XR_CTAS] ;
#pragma unroll
                for (unsigned r = 0 ; r < PANEL_XR_CTAS ; r++)
                {
                    dv [r] = (r < ECS) ? rec [r * PM2 + lane] : 0.0 ;
                    db [r] = (la && r < ECS) ? rec [r * PM2 + PANEL_MAX + lane] : 0.0 ;
                }
                sA = dv [0] ; sB = db [0] ;
#pragma unroll
                for (unsigned r = 1 ; r < PANEL_XR_CTAS ; r++) if (r < ECS) { sA += dv [r] ; sB += db [r] ; }
            }
            r0 = rws [lane] ;
            r1 = have1 ? rws [PANEL_MAX + lane] : 0.0 ;
        }
        PT_MARK (2) ;
        // ---- H_k: dlarfg on F(g:t-1,k) -------------------------------------------------------------
        // F(g+1,k): inside the staircase it comes from the published row g+1, below it is structurally zero
        const double x0g1 = (g + 1 < t) ? __shfl_sync (STMQR_FULL_MASK, r1, c) : 0.0 ;
        const double s = fma (x0g1, r1, sA) ;       // dot of column k with my column over the rows (g, t)
        double ss = __shfl_sync (STMQR_FULL_MASK, s, c) ;
        if (ss == 0.0 && x0g1 != 0.0) ss = 1e-300 ;         // (underflowed square of a non-zero entry, see the marker above)
        const double alpha = __shfl_sync (STMQR_FULL_MASK, r0, c) ;
        double beta = alpha, tau = 0, scale = 0, rinv = 0 ;
        bool rare = false ;
        if (t - g > 1)
        {
            double nrm ;
            if (ss > 1e-280 && ss < 1e280 && fabs (alpha) < 1e140)
            {
                const double q = fma (alpha, alpha, ss) ;
                rinv = rsqrt (q) ;
                nrm = q * rinv ;                                // hypot (alpha, ||x||)
            }
            else if (ss == 0)
            {
                // exactly zero sub-column (see the marker above): dlarfg returns tau = 0, beta = alpha
                nrm = 0 ;
            }
            else if (ss <= 1e-280 && fabs (alpha) > 1e-120 && fabs (alpha) < 1e140)
            {
                // ||x|| <= 1e-140 (every x_i^2 <= ss) next to a pivot >= 1e-120: rounding noise of a
                // sub-column that is zero in exact arithmetic (frequent in the top fronts of mesh
                // problems).  H = I to far below double precision: same as dlarfg for x = 0 (tau = 0,
                // beta = alpha); for noise dlarfg would return tau = 2, which only flips the sign of R's
                // row.  No exchange needed (the rescale path below costs several cluster/grid barriers).
                nrm = 0 ; ss = 0 ;
            }
            else
            {
                // rare: zero or badly scaled sub-column: max |x| first, then the rescaled sum of
                // squares (dnrm2 semantics).  The decision is uniform over the cluster.
                rare = true ;
#ifdef STMQR_PANEL_TIMING
                if (tid == 0 && leader)
                {
                    atomicAdd (N.dbg + 63, 1ULL) ;
                    if (ss == 0) atomicAdd (N.dbg + 62, 1ULL) ;
                    if (ss > 0 && ss <= 1e-280) atomicAdd (N.dbg + 61, 1ULL) ;
                    if (ss >= 1e280) atomicAdd (N.dbg + 60, 1ULL) ;
                    if (!(ss == ss)) atomicAdd (N.dbg + 59, 1ULL) ;
                    if (fabs (alpha) <= 1e-120) atomicAdd (N.dbg + 58, 1ULL) ;
                    if (alpha == 0) atomicAdd (N.dbg + 57, 1ULL) ;
                    if (fabs (alpha) >= 1e140) atomicAdd (N.dbg + 56, 1ULL) ;
                }
#endif
                // my warp's rows in (g, t): local [i0, e0)
                const I32 i0 = max (g + 1, lrow0) - lrow0 ;
                const I32 if1 = i0 + ((w - i0) & (NW - 1)) ;
                double mx = 0 ;
                for (I32 i = if1 ; i < e0 ; i += NW) mx = fmax (mx, fabs (x0 [i])) ;
                mx = panel_allreduce<NW, true, GRID> (cluster, gc, ECS, mx, red, xch [par], par) ;
                if (mx > 0)
                {
                    const double inv = 1.0 / mx ;
                    double s2 = 0 ;
                    for (I32 i = if1 ; i < e0 ; i += NW) { const double v = x0 [i] * inv ; s2 += v * v ; }
                    s2 = panel_allreduce<NW, false, GRID> (cluster, gc, ECS, s2, red + NW, xch [par], par) ;
                    nrm = hypot (alpha, mx * sqrt (s2)) ;
                    ss = 1.0 ;
                    // pivot and sub-column both below ~safmin/eps (underflowed rounding noise deep in a
                    // large front): 1/(alpha-beta) would overflow.  dlarfg rescales by 1/safmin here; the
                    // column is numerically zero either way, so H = I (tau = 0, beta = alpha)
                    if (!(nrm >= 1e-290)) { nrm = 0 ; ss = 0 ; }
                }
                else { nrm = 0 ; ss = 0 ; }
            }
            if (ss != 0)
            {
                beta = -copysign (nrm, alpha) ;
                if (rinv != 0)
                {
                    // regular path: tau = (beta-alpha)/beta = 1 + |alpha|/nrm and 1/(alpha-beta) =
                    // sign(alpha) / (|alpha| + nrm) from the reciprocal square root: one reciprocal instead
                    // of two divisions behind the square root (this scalar chain is serial in every step)
                    tau = fma (fabs (alpha), rinv, 1.0) ;
                    scale = copysign (1.0 / (fabs (alpha) + nrm), alpha) ;
                }
                else
                {
                    tau = (beta - alpha) / beta ;
                    scale = 1.0 / (alpha - beta) ;
                }
                // safety net (overflow of the norm or of 1/(alpha-beta)): never emit a non-finite reflector
                if (!(fabs (tau) <= 2.0) || !(fabs (scale) < 1e300)) { beta = alpha ; tau = 0 ; scale = 0 ; }
            }
        }
        const bool dead = (k < ntol) && (fabs (beta) <= tol) ;
        const double wv = tau * (r0 + scale * s) ;          // w_c = tau v_k' y_c (meaningful on lanes > c)
        const double vg1 = scale * x0g1 ;                   // v_k (g+1)

        // ---- H_{k+1} from the same exchange (see the header of this function) ----------------------
        bool two = false ;
        double beta1 = 0, tau1 = 0, scale1 = 0, so = 0, A_c = 0, A_c1 = 0 ;
        if (la && !dead && !rare)
        {
            A_c = __shfl_sync (STMQR_FULL_MASK, sA, c) ;
            if (A_c == 1e-300) A_c = 0 ;                    // (the exact-zero marker is not a value)
            A_c1 = __shfl_sync (STMQR_FULL_MASK, sA, c1) ;
            const double B_c1 = __shfl_sync (STMQR_FULL_MASK, sB, c1) ;
            const double om = __shfl_sync (STMQR_FULL_MASK, wv, c1) ;
            const double alpha1 = __shfl_sync (STMQR_FULL_MASK, r1, c1) - vg1 * om ;
            so = scale * om ;
            bool ok1 = false ;
            if (t1 - (g + 1) <= 1)
            {
                // no rows below the new pivot: H = I (dlarfg with n = 1)
                beta1 = alpha1 ; ok1 = true ;
            }
            else
            {
                const double m1 = 2.0 * so * A_c1, m2 = so * so * A_c ;
                const double ss1 = (B_c1 - m1) + m2 ;
                const double mag = B_c1 + fabs (m1) + m2 ;
                if (ss1 >= LA_THETA * mag && ss1 > 1e-280 && mag < 1e280 && fabs (alpha1) < 1e140)
                {
                    const double q1 = fma (alpha1, alpha1, ss1) ;
                    const double rinv1 = rsqrt (q1) ;
                    const double nrm1 = q1 * rinv1 ;
                    beta1 = -copysign (nrm1, alpha1) ;
                    tau1 = fma (fabs (alpha1), rinv1, 1.0) ;
                    scale1 = copysign (1.0 / (fabs (alpha1) + nrm1), alpha1) ;
                    ok1 = (fabs (tau1) <= 2.0) && (fabs (scale1) < 1e300) ;
                }
            }
            // a column that looks dead gets the direct (one-column) pass: its decision is then made on
            // a norm without any down-date
            two = ok1 && !((k + 1 < ntol) && (fabs (beta1) <= tol)) ;
#ifdef STMQR_PANEL_TIMING
            if (tid == 0 && leader) atomicAdd (N.dbg + (two ? 54 : 55), 1ULL) ;
#endif
        }
        PT_MARK (3) ;

        if (dead)
        {
            // zero F(g:m-1,k): my warp's rows, and (leader) whatever lies below the panel's window
            if (lane == c)
            {
                const I32 z0 = max (g, lrow0) - lrow0 ;
                for (I32 i = z0 + ((w - z0) & (NW - 1)) ; i < nloc ; i += NW) yc [i] = 0.0 ;
            }
            if (leader)
            {
                double *xg = F + (I64) k * ld ;
                for (I32 i = rend + tid ; i < fm ; i += nt) xg [i] = 0.0 ;
                if (tid == 0) { sto [c] = 0 ; tauo [c] = 0 ; }
            }
            k++ ;
        }
        else
        {
            // ---- dlarf on my warp's rows from g+2 on: y <- ya y - fa x_k - fb x_{k+1} -----------------
            //   lane c    : v_k = scale x_k                                    (ya, fa, fb) = (scale, 0, 0)
            //   lane c+1  : v_{k+1} = scale1 (x_{k+1} - scale w x_k)           (scale1, scale1 scale w, 0)
            //   lanes > c+1: y - v_k w_c - v_{k+1} w'_c
            double r1p = r1, w1 = 0 ;                       // row g+1 of my column after H_k; w'_c
            if (lane > c) r1p = fma (-vg1, wv, r1) ;
            else if (lane == c) r1p = vg1 ;
            double d1 = 0 ;                                 // x'(g+2:)' y'_c  (v_c for the lanes <= c)
            if (two)
            {
                const double swv = (lane > c) ? scale * wv : 0.0 ;
                d1 = (lane == c) ? scale * (A_c1 - so * A_c) : ((sB - swv * A_c1) - so * (sA - swv * A_c)) ;
                w1 = tau1 * (r1p + scale1 * d1) ;
            }
            double ya = 1.0, fa = 0.0, fb = 0.0 ;
            if (lane == c) { if (tau != 0) ya = scale ; }
            else if (two && lane == c1) { if (tau1 != 0) ya = scale1 ; fa = ya * so ; }
            else if (lane > c) { fb = scale1 * w1 ; fa = fma (-fb, so, scale * wv) ; }
            if ((tau != 0 || two) && lane >= c && mycol)
            {
                const I32 ee = two ? e1 : e0 ;
                I32 i = ifirst ;
                if (two)
                {
                    for ( ; i + 3 * NW < ee ; i += 4 * NW)
                    {
                        const double u0 = x0 [i], u1 = x0 [i + NW], u2 = x0 [i + 2*NW], u3 = x0 [i + 3*NW] ;
                        const double z0 = x1 [i], z1 = x1 [i + NW], z2 = x1 [i + 2*NW], z3 = x1 [i + 3*NW] ;
                        const double y0 = yc [i], y1 = yc [i + NW], y2 = yc [i + 2*NW], y3 = yc [i + 3*NW] ;
                        __syncwarp (__activemask ()) ;
                        yc [i] = fma (-z0, fb, fma (-u0, fa, y0 * ya)) ;
                        yc [i + NW] = fma (-z1, fb, fma (-u1, fa, y1 * ya)) ;
                        yc [i + 2*NW] = fma (-z2, fb, fma (-u2, fa, y2 * ya)) ;
                        yc [i + 3*NW] = fma (-z3, fb, fma (-u3, fa, y3 * ya)) ;
                    }
                    for ( ; i < ee ; i += NW)
                    {
                        const double u0 = x0 [i], z0 = x1 [i], y0 = yc [i] ;
                        __syncwarp (__activemask ()) ;
                        yc [i] = fma (-z0, fb, fma (-u0, fa, y0 * ya)) ;
                    }
                }
                else
                {
                    for ( ; i + 3 * NW < ee ; i += 4 * NW)
                    {
                        const double u0 = x0 [i], u1 = x0 [i + NW], u2 = x0 [i + 2*NW], u3 = x0 [i + 3*NW] ;
                        const double y0 = yc [i], y1 = yc [i + NW], y2 = yc [i + 2*NW], y3 = yc [i + 3*NW] ;
                        __syncwarp (__activemask ()) ;
                        yc [i] = fma (-u0, fa, y0 * ya) ; yc [i + NW] = fma (-u1, fa, y1 * ya) ;
                        yc [i + 2*NW] = fma (-u2, fa, y2 * ya) ; yc [i + 3*NW] = fma (-u3, fa, y3 * ya) ;
                    }
                    for ( ; i < ee ; i += NW)
                    {
                        const double u0 = x0 [i], y0 = yc [i] ;
                        __syncwarp (__activemask ()) ;
                        yc [i] = fma (-u0, fa, y0 * ya) ;
                    }
                }
            }
            // the two pivot rows, by their owner warps
            if (own_g0 && mycol)
            {
                if (lane == c) yc [gi0] = beta ;
                else if (lane > c && tau != 0) yc [gi0] -= wv ;
            }
            if (own_g1 && mycol && lane >= c)
            {
                if (two && lane == c1) yc [gi1] = beta1 ;
                else if (two && lane > c1) yc [gi1] = r1p - w1 ;
                else if (tau != 0 && g + 1 < t) yc [gi1] = r1p ;   // (lane c: v_k (g+1); lanes > c: row g+1 after H_k)
            }
            // V'V entries for dlarft: v_j' v_k = scale * (v_j' x) + v_j (g), j an earlier reflector
            if (leader && w == 0 && myq >= 0) Gs [myq + nv * (PANEL_MAX + 1)] = scale * s + r0 ;
            if (leader && tid == 0)
            {
                sto [c] = t ; tauo [c] = tau ;
                cols [nv] = k ; tq [nv] = t ; taus [nv] = tau ;
            }
            if (lane == c) myq = nv ;
            flops += (double) (t - g) * (3.0 + 4.0 * (double) (fn - k - 1)) ;
            nv++ ;
            g++ ;
            if (k == fp - 1 && leader && tid == 0) N.rank [f] = g ;
            k++ ;
            if (two)
            {
                // v_j' v_{k+1} = v_j (g+1) + scale1 * (v_j' x'),  j <= k
                if (leader && w == 0 && myq >= 0) Gs [myq + nv * (PANEL_MAX + 1)] = r1p + scale1 * d1 ;
                if (leader && tid == 0)
                {
                    sto [c1] = t1 ; tauo [c1] = tau1 ;
                    cols [nv] = k ; tq [nv] = t1 ; taus [nv] = tau1 ;
                }
                if (lane == c1) myq = nv ;
                flops += (double) (t1 - g) * (3.0 + 4.0 * (double) (fn - k - 1)) ;
                nv++ ;
                g++ ;
                if (k == fp - 1 && leader && tid == 0) N.rank [f] = g ;
                k++ ;
            }
        }
        if (dead && k - 1 == fp - 1 && leader && tid == 0) N.rank [f] = g ;
        // the next step's dots read, on this warp's rows, columns that other lanes of the warp just wrote
        __syncwarp () ;
        PT_MARK (4) ;
    }
    __syncthreads () ;
    PT_MARK (5) ;
    if (leader)
    {
        // per-column outputs of the columns this panel processed
        if (tid < np && sto [tid] >= 0)
        {
            const I32 k = k1 + tid ;
            st [k] = sto [tid] ; Tau [k] = tauo [tid] ;
            if (sto [tid] == 0 && k < fp) Rdead [k] = 1 ;
        }
        if (out_of_rows)
        {
            // no rows left: qr_front early exit (:1444-1458) for ALL remaining columns
            for (I32 kk = kstop + tid ; kk < fn ; kk += nt)
            {
                if (kk < fp) { Rdead [kk] = 1 ; st [kk] = 0 ; }
                else st [kk] = fm ;
                Tau [kk] = 0 ;
            }
        }
    }
    // the slab goes back to the front before T is built: the leader reuses it as scratch
    for (I32 c = w ; c < np ; c += NW)
    {
        double *dst = F + (I64) (k1 + c) * ld + lrow0 ;
        const double *src = P + (I64) c * ldp ;
#pragma unroll 4
        for (I32 i = lane ; i < nloc ; i += 32) dst [i] = src [i] ;
    }
    if (leader && nv > 1 && k2 < fn)
    {
        // dlarft by recursive doubling instead of its 496-step dependent chain: for a block reflector
        // split as [V1 V2], T = [T11 T12 ; 0 T22] with T12 = -T11 (V1'V2) T22.  Start from the 1 x 1
        // blocks T(i,i) = tau_i and merge neighbouring blocks of size B = 1, 2, 4, 8, 16; every merge is
        // two small triangular products, one output entry per thread, dependent chains of length B.
        // Column i of T comes out zero when tau_i = 0, as dlarft leaves it.
        __syncthreads () ;                               // the slab was read by the write-back above
        constexpr int LT = PANEL_MAX + 1 ;
        double *Tm = P, *Wm = P + PANEL_MAX * LT ;      // host guarantees slab_cap >= 2*32*33
        for (int e = tid ; e < PANEL_MAX * PANEL_MAX ; e += nt)
        {
            const int j = e & 31, i = e >> 5 ;
            Tm [j + i * LT] = (j == i && i < nv) ? taus [i] : 0.0 ;
        }
        __syncthreads () ;
        for (int B = 1 ; B < nv ; B *= 2)
        {
            // W = (V1'V2) T22 :  W(r,c) = sum_{l = s+B .. c} G(r,l) T(l,c),   r in [s,s+B), c in [s+B,s+2B)
            for (int e = tid ; e < 16 * B ; e += nt)
            {
                const int blk = e / (B * B), rr = (e / B) % B, cc = e % B ;
                const int s0 = blk * 2 * B, r = s0 + rr, c = s0 + B + cc ;
                if (c < nv)
                {
                    double a0 = 0, a1 = 0 ;
                    int l = s0 + B ;
                    for ( ; l + 1 <= c ; l += 2)
                    {
                        a0 = fma (Gs [r + l * LT], Tm [l + c * LT], a0) ;
                        a1 = fma (Gs [r + (l+1) * LT], Tm [(l+1) + c * LT], a1) ;
                    }
                    if (l <= c) a0 = fma (Gs [r + l * LT], Tm [l + c * LT], a0) ;
                    Wm [r + c * LT] = a0 + a1 ;
                }
            }
            __syncthreads () ;
            // T12 = -T11 W :  T(r,c) = -sum_{l = r .. s+B-1} T(r,l) W(l,c)
            for (int e = tid ; e < 16 * B ; e += nt)
            {
                const int blk = e / (B * B), rr = (e / B) % B, cc = e % B ;
                const int s0 = blk * 2 * B, r = s0 + rr, c = s0 + B + cc ;
                if (c < nv)
                {
                    double a0 = 0, a1 = 0 ;
                    int l = r ;
                    for ( ; l + 1 < s0 + B ; l += 2)
                    {
                        a0 = fma (Tm [r + l * LT], Wm [l + c * LT], a0) ;
                        a1 = fma (Tm [r + (l+1) * LT], Wm [(l+1) + c * LT], a1) ;
                    }
                    if (l < s0 + B) a0 = fma (Tm [r + l * LT], Wm [l + c * LT], a0) ;
                    Tm [r + c * LT] = -(a0 + a1) ;
                }
            }
            __syncthreads () ;
        }
        const I32 slotp = parity * L.count + slot ;
        double *Tg = N.Tws + (I64) slotp * (PANEL_MAX * PANEL_MAX) ;
        for (int e = tid ; e < nv * nv ; e += nt)
        {
            const int j = e % nv, i = e / nv ;
            Tg [j + i * PANEL_MAX] = (j <= i) ? Tm [j + i * LT] : 0.0 ;
        }
        for (int q = tid ; q < nv ; q += nt) N.pnl_cols [slotp * PANEL_MAX + q] = cols [q] ;
        panel_epilogue<false> (L, S, N, slot, f, k2, parity, leader, nv, g, g1, out_of_rows, flops, Gs, Gs, taus, cols, tq) ;
    }
    else
    {
        panel_epilogue<true> (L, S, N, slot, f, k2, parity, leader, nv, g, g1, out_of_rows, flops, Gs, Gs, taus, cols, tq) ;
    }
    PT_MARK (6) ;
    PT_FLUSH ((NW == 16 ? 0 : (NW == 8 ? 8 : 16)) + (ECS > 1 ? 24 : 0)) ;
}

// grid = (# active fronts) * CS CTAs, cluster dimension CS (launch attribute); dynamic shared
// memory = slab_cap doubles (the row slab of one CTA: RL rows x np columns, column-major, odd ld)
// followed by panel_scratch_doubles (NT/32) doubles of scratch.  NT = threads per CTA: small
// fronts use small CTAs so that many of them are resident per SM.
template <int NT, int MINB>
__global__ void __launch_bounds__ (NT, MINB) k_panel_cluster (LevelArgs L, DSym S, DNum N, I32 k1, I32 PB,
    I32 parity, I32 slab_cap)
{
    extern __shared__ double slab [] ;
    __shared__ PanelXch xch [2] ;
    __shared__ double tot [PANEL_MAX + 2 + 64] ;
    __shared__ double rg [PANEL_MAX] ;
    __shared__ I32 cols [PANEL_MAX], tq [PANEL_MAX] ;
    constexpr int NW = NT / 32 ;
    // records pushed by the cluster's CTAs (panel_columns_smem): behind the scratch, only allocated by the
    // host when the launch has clusters of more than one CTA
    double *xrec = slab + slab_cap + panel_scratch_doubles (NW) ;
    double *xrows = xrec + 2 * PANEL_XR_CTAS * 2 * PANEL_MAX ;
    // global-mode scratch (V'V / T, taus) lives behind the slab like the shared-memory mode's
    double *Gs = slab + slab_cap + 4 * NW * PANEL_MAX + 4 * PANEL_MAX + 2 * NW ;
    double *Tsh = Gs ;
    double *taus = Gs + PANEL_MAX * (PANEL_MAX + 1) ;

    cg::cluster_group cluster = cg::this_cluster () ;
    const unsigned CS = cluster.num_blocks (), cr = cluster.block_rank () ;
    const I32 slot = blockIdx.x / CS ;
    const I32 f = L.fronts [slot] ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const int tid = threadIdx.x, nt = blockDim.x ;
    const I32 slotp = parity * L.count + slot ;

    // every CTA of the cluster reads the front's state before any of them may change it, and
    // then takes the same branch (same inputs), so no CTA is left waiting in a cluster barrier
    const I32 done0 = N.done [slot] ;
    const I32 g1 = N.g [slot] ;
    const I32 stlast = (k1 < fn) ? N.stair [p1 + min (fn, k1 + PB) - 1] : 0 ;
    cluster.sync () ;
    if (k1 >= fn || done0)
    {
        if (cr == 0 && tid == 0) N.pnl_nv [slotp] = 0 ;
        return ;
    }
    const I32 fm = N.Hm [f] ;
    const I32 k2 = min (fn, k1 + PB) ;
    const I32 np = k2 - k1 ;
    const I32 *st = N.stair + p1 ;
    // rows that this panel can touch: [g1, rend)
    const I32 rend = min (fm, max (stlast, g1 + np)) ;
    const I32 nrows = max (rend - g1, 0) ;
    // effective cluster size: a window that fits one CTA's slab is done by the leader alone (no
    // cluster barriers at all); the other CTAs of the cluster leave
    const unsigned ECS = ((I64) ((nrows + 3) | 1) * np <= (I64) slab_cap) ? 1u : CS ;
    if (cr >= ECS) return ;
    const I32 RL = max (4, (((nrows + (I32) ECS - 1) / (I32) ECS) + 3) & ~3) ;
    const I32 lrow0 = g1 + (I32) cr * RL ;
    const I32 nloc = max (0, min (rend - lrow0, RL)) ;
    const I32 ldp = RL | 1 ;
    const bool insmem = ((I64) ldp * np <= (I64) slab_cap) ;
    double *F = N.F + S.Foff [f] ;
    const I64 ld = fm ;

    if (insmem)
    {
        // one warp per column, lanes along the rows (coalesced); the loads of a thread are
        // independent, so the whole slab costs about one memory latency
        {
            const int lane = tid & 31, w = tid >> 5 ;
            for (I32 c = w ; c < np ; c += NW)
            {
                const double *src = F + (I64) (k1 + c) * ld + lrow0 ;
                double *dst = slab + (I64) c * ldp ;
#pragma unroll 4
                for (I32 i = lane ; i < nloc ; i += 32) dst [i] = __ldcg (src + i) ;
            }
        }
        __syncthreads () ;
        GridComm nogrid {} ;
        panel_columns_smem<NW, false> (cluster, nogrid, ECS, slab_cap, ldp, L, S, N, slot, f, k1, k2, parity, lrow0, nloc, g1,
            RL, rend, xch, xrec, xrows, cols, tq) ;
    }
    else
    {
        panel_columns<false> (cluster, F + (I64) k1 * ld + lrow0, ld, L, S, N, slot, f, k1, k2, parity, lrow0,
            nloc, g1, RL, rend, xch, tot, rg, Gs, Tsh, taus, cols, tq) ;
    }
    if (ECS == 1) return ;
    // nobody may leave while a peer can still read its exchange records
    cluster.sync () ;
}

// ---------------------------------------------------------------------------------------------
// The same panel factorization for fronts whose row window does not fit the 8 shared-memory slabs
// of a cluster: G CTAs per front (G x fronts <= # SMs, one CTA per SM: all resident), each with its
// slab in shared memory, exchanging through global memory with an arrival-counter barrier.
// grid = G x nfronts, slot = slot0 + blockIdx.x / G.  ctr: 2 counters per slot, zero on entry (the
// last CTA to leave resets them).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (512, 1) k_panel_grid (LevelArgs L, DSym S, DNum N, I32 k1, I32 PB, I32 parity,
    I32 slab_cap, I32 G, I32 slot0, unsigned seq)
{
    extern __shared__ double slab [] ;
    __shared__ PanelXch xch [2] ;
    __shared__ I32 cols [PANEL_MAX], tq [PANEL_MAX] ;
    constexpr int NW = 16 ;
    cg::cluster_group cluster = cg::this_cluster () ;
    const I32 slot = slot0 + blockIdx.x / G ;
    const unsigned cr = blockIdx.x % G ;
    const I32 f = L.fronts [slot] ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const int tid = threadIdx.x ;
    const I32 slotp = parity * L.count + slot ;
    unsigned *ctr = N.gridctr + (I64) GRID_CTR_STRIDE * slot ;

    // every CTA of the front reads the front's state before any of them may change it
    const I32 done0 = N.done [slot] ;
    const I32 g1 = N.g [slot] ;
    const I32 stlast = (k1 < fn) ? N.stair [p1 + min (fn, k1 + PB) - 1] : 0 ;
    const I32 fm = N.Hm [f] ;
    GridComm g0 ; g0.rec = nullptr ; g0.red = nullptr ; g0.ctr = ctr ; g0.G = (unsigned) G ; g0.cr = cr ; g0.epoch = 0 ;
    g0.ll = nullptr ; g0.tagbase = 0 ;
    const I32 k2 = min (fn, k1 + PB) ;
    const I32 np = k2 - k1 ;
    const I32 rend = min (fm, max (stlast, g1 + np)) ;
    const I32 nrows = max (rend - g1, 0) ;
    // CTAs that take part: a step costs ~5.7 cycles per slab row (shared-memory sweeps) plus
    // ~27/4 cycles per arriving CTA (L2 atomics): ECS ~ sqrt (0.84 nrows), at least what fits
    const I32 rlcap = max (4, (slab_cap / max (np, 1) - 4) & ~3) ;
    const I32 ecs_fit = (nrows + rlcap - 1) / rlcap ;
    const I32 ecs_opt = (I32) sqrtf (0.84f * (float) nrows) ;
    unsigned ECS = (unsigned) max (1, min ((I32) G, max (ecs_fit, ecs_opt))) ;
    const bool idle = (k1 >= fn || done0) ;
    const I32 RL = max (4, (((nrows + (I32) ECS - 1) / (I32) ECS) + 3) & ~3) ;
    const I32 ldp = RL | 1 ;
    const bool fits = ((I64) ldp * np <= (I64) slab_cap) ;
    // the staircase of the panel's columns is read inside panel_columns_smem: stage it before the barrier
    grid_barrier (g0) ;
    if (!idle && fits && cr < ECS)
    {
        const I32 lrow0 = g1 + (I32) cr * RL ;
        const I32 nloc = max (0, min (rend - lrow0, RL)) ;
        double *F = N.F + S.Foff [f] ;
        const I64 ld = fm ;
        {
            const int lane = tid & 31, w = tid >> 5 ;
            for (I32 c = w ; c < np ; c += NW)
            {
                const double *src = F + (I64) (k1 + c) * ld + lrow0 ;
                double *dst = slab + (I64) c * ldp ;
#pragma unroll 4
                for (I32 i = lane ; i < nloc ; i += 32) dst [i] = __ldcg (src + i) ;
            }
        }
        __syncthreads () ;
        GridComm gc ;
        gc.rec = N.gridrec + (I64) slot * (2 * 148 * 64) ;
        gc.red = N.gridred + (I64) slot * (2 * 148) ;
        gc.ctr = ctr + 128 ; gc.G = ECS ; gc.cr = cr ; gc.epoch = 0 ;
        gc.ll = N.gridll + (I64) slot * (2 * 148 * 128) ; gc.tagbase = seq * 64u ;
        panel_columns_smem<NW, true> (cluster, gc, ECS, slab_cap, ldp, L, S, N, slot, f, k1, k2, parity, lrow0, nloc, g1,
            RL, rend, xch, nullptr, nullptr, cols, tq) ;
    }
    else if (cr == 0 && tid == 0)
    {
        N.pnl_nv [slotp] = 0 ;
        if (!idle && !fits) atomicExch (N.griderr, 1) ;      // host sized G too small: reported as an error
    }
    // the last CTA of the front to leave resets the counters for the next launch
    __syncthreads () ;
    if (tid == 0)
    {
        __threadfence () ;
        if (atomicAdd (ctr + 256, 1u) == (unsigned) G - 1)
        {
            for (int j = 0 ; j < 8 ; j++) ctr [32 * j] = 0 ;
            ctr [256] = 0 ;
            __threadfence () ;
        }
    }
}

} // namespace stmqr
