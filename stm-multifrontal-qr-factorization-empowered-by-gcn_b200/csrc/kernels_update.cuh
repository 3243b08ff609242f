// kernels_update.cuh -- trailing-matrix update of the blocked Householder QR on FP64 tensor
// cores:  C := (I - V T V')' C = C - V (T' (V' C))   (dlarfb 'L','T','F','C'; the reference's
// qr_larftb method QR_QTX, SparseQR_factorize.c:1851-1882).
//
// V (mr x nv, nv <= 32) is the unit lower trapezoidal block of Householder vectors that the
// panel kernel left in columns cols[0..nv) of the front, rows g1 .. tend-1; T is the nv x nv
// upper triangular factor from dlarft.  grid = (active fronts, column tiles of 32 columns).
// One CTA owns 32 columns of C over ALL mr rows, so C is read from HBM once for W = V'C, again
// (normally from L2) for the rank-nv update, and written once.
//
// Both contractions run as mma.sync.m8n8k4.f64 (DMMA; tcgen05.mma has no FP64 kind):
//   phase 1  W(q,c) = sum_r V(r,q) C(r,c)     M = 32 (q), N = 32 (c), K = rows, K split over the
//            8 warps of the CTA (each warp owns 8 rows of every 64-row chunk and all 4 x 4
//            accumulator tiles); the 8 partial W are summed as a tree (deterministic)
//   phase 2  W := -T' W  (tiny, plain FMA)
//   phase 3  C(r,c) += sum_q V(r,q) W(q,c)    M = rows, N = 32, K = 32; each warp owns 8 rows of the
//            chunk, keeps the whole W tile in registers as B fragments, starts its accumulators
//            from the C tile and stores them straight to the front.
// The rows stream through an NSTAGE-deep ring of shared-memory tiles filled by cp.async (LDGSTS,
// 8-byte granules because fronts are only 8-byte aligned; out-of-range granules are zero-filled
// by the copy itself), so the DMMA pipe works on chunk i while chunks i+1.. are in flight.
// Shared-memory tiles are column-major with a column stride = 4 (mod 16) doubles, which makes every
// A/B fragment load (8 rows/cols x 4 k per warp) bank-conflict free.
#pragma once
#include "engine.cuh"
#include "kernels_panel.cuh"

namespace stmqr {

constexpr int UPD_NC = 32 ;             // columns of C per CTA
constexpr int UPD_RC = 64 ;             // rows per chunk
constexpr int UPD_LDS = UPD_RC + 4 ;    // column stride of the V and C tiles (68 = 4 mod 16)
constexpr int UPD_LDW = PANEL_MAX + 4 ; // column stride of the W tile (36 = 4 mod 16)
constexpr int UPD_STAGE = (PANEL_MAX + UPD_NC) * UPD_LDS ;      // doubles per ring stage (V tile | C tile)

__device__ __forceinline__ void dmma_m8n8k4 (double &d0, double &d1, const double a, const double b)
{
    asm volatile ("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d" (d0), "+d" (d1) : "d" (a), "d" (b)) ;
}
// 8-byte asynchronous global->shared copy; valid = false writes zeros instead
__device__ __forceinline__ void cp_async8 (double *smem, const double *gmem, const bool valid)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared (smem) ;
    const int sz = valid ? 8 : 0 ;
    asm volatile ("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r" (s), "l" (gmem), "r" (sz) : "memory") ;
}
__device__ __forceinline__ void cp_async_commit () { asm volatile ("cp.async.commit_group;" ::: "memory") ; }
template <int N> __device__ __forceinline__ void cp_async_wait () { asm volatile ("cp.async.wait_group %0;" :: "n" (N) : "memory") ; }

// enqueue rows [r0, r0+64) of V (strictly lower part; diagonal and above arrive as zeros) and of the
// CTA's C tile into one ring stage
__device__ __forceinline__ void issue_chunk (double *stage, const double *__restrict__ F, const I64 ld,
    const I32 g1, const I32 mr, const I32 nv, const I32 *cols, const I32 c0, const I32 ncol, const I32 r0,
    const int tid)
{
    double *Vs = stage, *Cs = stage + PANEL_MAX * UPD_LDS ;
#pragma unroll
    for (int a = 0 ; a < (PANEL_MAX * UPD_RC) / 256 ; a++)
    {
        const int e = tid + a * 256 ;
        const int q = e >> 6, rr = e & 63 ;
        const I32 r = r0 + rr ;
        const bool ok = (q < nv) && (r < mr) && (r > q) ;
        cp_async8 (Vs + q * UPD_LDS + rr, ok ? (F + (g1 + r) + (I64) cols [q] * ld) : F, ok) ;
    }
#pragma unroll
    for (int a = 0 ; a < (UPD_NC * UPD_RC) / 256 ; a++)
    {
        const int e = tid + a * 256 ;
        const int c = e >> 6, rr = e & 63 ;
        const I32 r = r0 + rr ;
        const bool ok = (c < ncol) && (r < mr) ;
        cp_async8 (Cs + c * UPD_LDS + rr, ok ? (F + (g1 + r) + (I64) (c0 + c) * ld) : F, ok) ;
    }
}

// rsf > 1: the rows are split over a cluster of rsf CTAs (cluster dimension z): every CTA computes
// the partial W of its rows, the partials are summed in rank order through distributed shared memory,
// and every CTA updates its own rows -- mid-size fronts (a few column tiles, thousands of rows)
// then occupy rsf times as many SMs and every CTA streams 1/rsf of the rows.
template <int NSTAGE>
__global__ void __launch_bounds__ (256) k_update_dmma (LevelArgs L, DSym S, DNum N, I32 cbeg, I32 cend,
    I32 parity, I32 rsf)
{
    constexpr int NT = UPD_NC / 8 ;                 // accumulator tiles along the columns
    extern __shared__ double sm [] ;
    double *ring = sm ;                             // [NSTAGE][UPD_STAGE]
    double *Ws = ring + NSTAGE * UPD_STAGE ;        // [32][UPD_LDW]   W(q,c) at Ws[c*LDW + q]
    double *Ts = Ws + UPD_NC * UPD_LDW ;            // [32][33]        T(j,i) at Ts[j*33 + i]
    __shared__ I32 cols [PANEL_MAX] ;

    const I32 slot = blockIdx.x ;
    const I32 slotp = parity * L.count + slot ;
    const I32 nv = N.pnl_nv [slotp] ;
    if (nv == 0) return ;
    const I32 f = L.fronts [slot] ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    const I32 c0 = cbeg + blockIdx.y * UPD_NC ;
    const I32 clim = min (cend, fn) ;
    if (c0 >= clim) return ;
    const I32 ncol = min (UPD_NC, clim - c0) ;
    const I64 ld = N.Hm [f] ;
    double *F = N.F + S.Foff [f] ;
    const I32 g1 = N.pnl_g1 [slotp], tend = N.pnl_tend [slotp] ;
    const I32 mr = tend - g1 ;
    const I32 nch = (mr + UPD_RC - 1) / UPD_RC ;
    const I32 cpr = (nch + rsf - 1) / rsf ;                         // chunks per rank of the cluster
    const I32 ch0 = min (nch, (I32) blockIdx.z * cpr), ch1 = min (nch, ch0 + cpr) ;
    const I32 nmy = ch1 - ch0 ;
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

#pragma unroll
    for (int p = 0 ; p < NSTAGE - 1 ; p++)
    {
        if (p < nmy) issue_chunk (ring + p * UPD_STAGE, F, ld, g1, mr, nv, cols, c0, ncol, (ch0 + p) * UPD_RC, tid) ;
        cp_async_commit () ;
    }
    for (I32 ch = 0 ; ch < nmy ; ch++)
    {
        // keep NSTAGE-1 chunks in flight; the stage refilled here was consumed in iteration ch-1
        if (ch + NSTAGE - 1 < nmy)
            issue_chunk (ring + ((ch + NSTAGE - 1) % NSTAGE) * UPD_STAGE, F, ld, g1, mr, nv, cols, c0, ncol,
                (ch0 + ch + NSTAGE - 1) * UPD_RC, tid) ;
        cp_async_commit () ;
        cp_async_wait<NSTAGE - 1> () ;
        __syncthreads () ;
        double *Vs = ring + (ch % NSTAGE) * UPD_STAGE, *Cs = Vs + PANEL_MAX * UPD_LDS ;
        if (ch0 + ch == 0)
        {
            if (tid < nv) Vs [tid * UPD_LDS + tid] = 1.0 ;          // unit diagonal of V
            __syncthreads () ;
        }
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
    cp_async_wait<0> () ;
    // the 8 warps' partial W: staged in the (now idle) ring, summed as a tree in a fixed order
    {
        double *part = ring + w * (PANEL_MAX * UPD_NC) ;        // 8 x 1024 doubles <= NSTAGE x 4352
#pragma unroll
        for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
            for (int ni = 0 ; ni < NT ; ni++)
#pragma unroll
                for (int e = 0 ; e < 2 ; e++)
                    part [(ni * 8 + tig * 2 + e) * PANEL_MAX + mi * 8 + grp] = acc [mi][ni][e] ;
    }
    __syncthreads () ;
    // ---- phase 2: W = -T' W -------------------------------------------------------------------
    {
        double wsum [(UPD_NC * PANEL_MAX) / 256] ;
#pragma unroll
        for (int a = 0 ; a < (UPD_NC * PANEL_MAX) / 256 ; a++)
        {
            const int e = tid + a * 256 ;               // e = c*32 + q
            double p [8] ;
#pragma unroll
            for (int ww = 0 ; ww < 8 ; ww++) p [ww] = ring [ww * (PANEL_MAX * UPD_NC) + e] ;
            wsum [a] = ((p [0] + p [1]) + (p [2] + p [3])) + ((p [4] + p [5]) + (p [6] + p [7])) ;
        }
#pragma unroll
        for (int a = 0 ; a < (UPD_NC * PANEL_MAX) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            Ws [(e >> 5) * UPD_LDW + (e & 31)] = wsum [a] ;
        }
        if (rsf > 1)
        {
            // W = sum over the ranks of the cluster (fixed order: every rank gets the same bits)
            cg::cluster_group cluster = cg::this_cluster () ;
            cluster.sync () ;
#pragma unroll
            for (int a = 0 ; a < (UPD_NC * PANEL_MAX) / 256 ; a++)
            {
                const int e = tid + a * 256 ;
                double *mine = Ws + (e >> 5) * UPD_LDW + (e & 31) ;
                double t = 0 ;
                for (int r = 0 ; r < rsf ; r++) t += *cluster.map_shared_rank (mine, r) ;
                wsum [a] = t ;
            }
            cluster.sync () ;           // all remote reads done: Ws may be overwritten, peers may leave
#pragma unroll
            for (int a = 0 ; a < (UPD_NC * PANEL_MAX) / 256 ; a++)
            {
                const int e = tid + a * 256 ;
                Ws [(e >> 5) * UPD_LDW + (e & 31)] = wsum [a] ;
            }
        }
        __syncthreads () ;
        double w2 [(UPD_NC * PANEL_MAX) / 256] ;
#pragma unroll
        for (int a = 0 ; a < (UPD_NC * PANEL_MAX) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            const int i = e & 31, c = e >> 5 ;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0 ;
            // T is upper triangular: rows j <= i (the rest of the column is stored as zeros)
#pragma unroll
            for (int j = 0 ; j < PANEL_MAX ; j += 4)
            {
                s0 = fma (Ts [j * (PANEL_MAX + 1) + i], Ws [c * UPD_LDW + j], s0) ;
                s1 = fma (Ts [(j+1) * (PANEL_MAX + 1) + i], Ws [c * UPD_LDW + j + 1], s1) ;
                s2 = fma (Ts [(j+2) * (PANEL_MAX + 1) + i], Ws [c * UPD_LDW + j + 2], s2) ;
                s3 = fma (Ts [(j+3) * (PANEL_MAX + 1) + i], Ws [c * UPD_LDW + j + 3], s3) ;
            }
            w2 [a] = -((s0 + s1) + (s2 + s3)) ;
        }
        __syncthreads () ;
#pragma unroll
        for (int a = 0 ; a < (UPD_NC * PANEL_MAX) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            Ws [(e >> 5) * UPD_LDW + (e & 31)] = w2 [a] ;
        }
        __syncthreads () ;
    }
    // ---- phase 3: C += V W (W already negated); W as B fragments in registers -------------------
    double wf [8][NT] ;
#pragma unroll
    for (int ks = 0 ; ks < 8 ; ks++)
#pragma unroll
        for (int ni = 0 ; ni < NT ; ni++) wf [ks][ni] = Ws [(ni * 8 + grp) * UPD_LDW + ks * 4 + tig] ;

#pragma unroll
    for (int p = 0 ; p < NSTAGE - 1 ; p++)
    {
        if (p < nmy) issue_chunk (ring + p * UPD_STAGE, F, ld, g1, mr, nv, cols, c0, ncol, (ch0 + p) * UPD_RC, tid) ;
        cp_async_commit () ;
    }
    for (I32 ch = 0 ; ch < nmy ; ch++)
    {
        if (ch + NSTAGE - 1 < nmy)
            issue_chunk (ring + ((ch + NSTAGE - 1) % NSTAGE) * UPD_STAGE, F, ld, g1, mr, nv, cols, c0, ncol,
                (ch0 + ch + NSTAGE - 1) * UPD_RC, tid) ;
        cp_async_commit () ;
        cp_async_wait<NSTAGE - 1> () ;
        __syncthreads () ;
        double *Vs = ring + (ch % NSTAGE) * UPD_STAGE, *Cs = Vs + PANEL_MAX * UPD_LDS ;
        if (ch0 + ch == 0)
        {
            if (tid < nv) Vs [tid * UPD_LDS + tid] = 1.0 ;
            __syncthreads () ;
        }
        double d [NT][2] ;
#pragma unroll
        for (int ni = 0 ; ni < NT ; ni++)
#pragma unroll
            for (int e = 0 ; e < 2 ; e++) d [ni][e] = Cs [(ni * 8 + tig * 2 + e) * UPD_LDS + w * 8 + grp] ;
#pragma unroll
        for (int ks = 0 ; ks < 8 ; ks++)
        {
            const double af = Vs [(ks * 4 + tig) * UPD_LDS + w * 8 + grp] ;
#pragma unroll
            for (int ni = 0 ; ni < NT ; ni++) dmma_m8n8k4 (d [ni][0], d [ni][1], af, wf [ks][ni]) ;
        }
        const I32 r = (ch0 + ch) * UPD_RC + w * 8 + grp ;
        if (r < mr)
        {
#pragma unroll
            for (int ni = 0 ; ni < NT ; ni++)
#pragma unroll
                for (int e = 0 ; e < 2 ; e++)
                {
                    const int c = ni * 8 + tig * 2 + e ;
                    if (c < ncol) F [(g1 + r) + (I64) (c0 + c) * ld] = d [ni][e] ;
                }
        }
        __syncthreads () ;
    }
    cp_async_wait<0> () ;
}

template <int NSTAGE> constexpr size_t update_smem_bytes ()
{
    return sizeof (double) * (NSTAGE * UPD_STAGE + UPD_NC * UPD_LDW + PANEL_MAX * (PANEL_MAX + 1)) ;
}

} // namespace stmqr
