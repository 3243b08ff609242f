// kernels_wide.cuh -- two-level blocked Householder QR for the LARGE fronts: the trailing-matrix
// update with a 128-column block reflector on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64).
//
// Reference semantics: qr_front / qr_larftb (SparseQR_factorize.c:1383-1618, :1851-1904) apply the
// block reflector of every `fchunk` (32) columns to ALL remaining columns of the front, a K = 32
// contraction: ~4 flop per byte of C streamed, i.e. memory bound on a B200 (ridge ~ 6 flop/B).
// Blocking is performance-only, so here four 32-column panels form one
// OUTER block of 128 columns:
//     for each panel p of the block:  panel factorization (k_panel_cluster)
//                                     V_p -> clean buffer Vb (k_wide_vextract)
//                                     update of the REST OF THE BLOCK only (<= 96 columns, K = 32)
//     Gram matrix Vb'Vb -> T of the whole block (k_wide_tmerge; the blocked dlarft recurrence
//         T(0:j,j) = -T(0:j,0:j) (V(:,0:j)' V(:,j)) T(j,j))
//     update of all columns right of the block with K = 128:
//         W  = Vb' C          k_wide_vtc    (split over row ranges, deterministic partials)
//         W2 = -T' sum(W)     k_wide_wt
//         C += Vb W2          k_wide_apply
// Every kernel is a grid over (tiles, fronts of the level): the same launch serves all active fronts
// of an etree level.  The contractions run as 4x4 register tiles of m8n8k4 DMMAs per warp (8 fragment
// loads per 16 DMMAs), operands staged in shared memory by cp.async rings.
//
// Vb is a "clean" copy of the block's Householder vectors relative to the block's first pivot row
// g0: column 32p+q is reflector q of panel p (zero above its diagonal, explicit 1 on it, zero below
// the staircase, zero for q >= nv_p), rows padded with zeros, leading dimension even: 16-byte
// cp.async granules and no masking in the GEMM loops.
#pragma once
#include "engine.cuh"
#include "kernels_panel.cuh"
#include "kernels_update.cuh"

namespace stmqr {

constexpr int WB = 128 ;                // columns of an outer block
constexpr int WB_PANELS = WB / PANEL_MAX ;
constexpr int W_NC = 64 ;               // columns of C per CTA
constexpr int W_RC = 32 ;               // rows per chunk of k_wide_vtc
constexpr int W_LDC = W_RC + 4 ;        // column stride of its tiles (36 = 4 mod 16: conflict-free fragments)
constexpr int W_RT = 128 ;              // rows per CTA of k_wide_apply
constexpr int W_LDV = W_RT + 4 ;        // column stride of its V slab (132 = 4 mod 16)
constexpr int W_NST = 3 ;               // ring stages

enum { WIDE_INNER = 0, WIDE_OUTER = 1, WIDE_GRAM = 2 } ;

struct WideArgs
{
    const I32 *fronts ;     // fronts of the level (sorted by # columns descending)
    I32 count ;             // # fronts of the level = slots
    I32 buf ;               // parity of the outer block: selects Vb / Tbt / wblk
    I32 ldv ;               // leading dimension of Vb (multiple of 128, >= max rows + 128)
    I32 rs ;                // rows per split of k_wide_vtc (multiple of 32)
    I32 nsplit ;            // max # splits
    I32 ncmax ;             // column capacity of Wp / W2 per front
} ;

__device__ __forceinline__ void cp_async16 (double *smem, const double *gmem)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared (smem) ;
    asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r" (s), "l" (gmem) : "memory") ;
}
__device__ __forceinline__ void cp_async16z (double *smem, const double *gmem, const bool valid)
{
    const unsigned s = (unsigned) __cvta_generic_to_shared (smem) ;
    const int sz = valid ? 16 : 0 ;
    asm volatile ("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r" (s), "l" (gmem), "r" (sz) : "memory") ;
}

__device__ __forceinline__ double *wide_vb (const WideArgs &A, const DNum &N, const I32 slot)
{
    return N.wVb + ((I64) A.buf * A.count + slot) * ((I64) A.ldv * WB) ;
}
__device__ __forceinline__ I32 *wide_blk (const WideArgs &A, const DNum &N, const I32 slot)
{
    return N.wblk + ((I64) A.buf * A.count + slot) * 4 ;    // g0, mr, nvtot, -
}

// ---------------------------------------------------------------------------------------------
// V of panel p -> Vb columns [32p, 32p+32).  grid = (row chunks of 256, fronts); warp per column.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (256) k_wide_vextract (WideArgs A, DSym S, DNum N, I32 p)
{
    const I32 slot = blockIdx.y ;
    const I32 slotp = p * A.count + slot ;
    const I32 nv = N.pnl_nv [slotp] ;
    I32 *blk = wide_blk (A, N, slot) ;
    const I32 g1 = N.pnl_g1 [slotp] ;
    const I32 g0 = (p == 0) ? g1 : blk [0] ;
    const I32 nvtot0 = (p == 0) ? 0 : blk [2] ;
    const I32 f = A.fronts [slot] ;
    const I32 fm = N.Hm [f] ;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5 ;
    if (nv == 0 && nvtot0 == 0)
    {
        // nothing live in this block so far: nobody reads Vb
        if (blockIdx.x == 0 && tid == 0 && p == 0) { blk [0] = g1 ; blk [1] = 0 ; blk [2] = 0 ; }
        return ;
    }
    const I32 rows = min (A.ldv, ((max (fm - g0, 0) + W_RT - 1) / W_RT) * W_RT + W_RT) ;   // zero padded
    const I32 rb = blockIdx.x * 256 ;
    if (rb < rows)
    {
        const double *F = N.F + S.Foff [f] ;
        double *Vb = wide_vb (A, N, slot) ;
        for (I32 q = w ; q < PANEL_MAX ; q += 8)
        {
            const I32 col = (q < nv) ? N.pnl_cols [slotp * PANEL_MAX + q] : -1 ;
            const I32 d = (g1 - g0) + q ;                   // row of the unit diagonal
            const double *Fc = F + (I64) max (col, 0) * fm + g0 ;
            double *Vc = Vb + (I64) (p * PANEL_MAX + q) * A.ldv ;
#pragma unroll 4
            for (I32 r = rb + lane ; r < min (rows, rb + 256) ; r += 32)
            {
                double v = 0.0 ;
                if (col >= 0 && r >= d && g0 + r < fm) v = (r == d) ? 1.0 : __ldcg (Fc + r) ;
                Vc [r] = v ;
            }
        }
    }
    if (blockIdx.x == 0 && tid == 0)
    {
        if (p == 0) { blk [0] = g1 ; blk [1] = (nv > 0) ? N.pnl_tend [slotp] - g1 : 0 ; blk [2] = nv ; }
        else if (nv > 0) { blk [1] = N.pnl_tend [slotp] - g0 ; blk [2] = nvtot0 + nv ; }
    }
}

// ---------------------------------------------------------------------------------------------
// W_partial(split) = V' B over the rows of one split.   KV = 32 KVT reflector columns.
//   WIDE_INNER  V = Vb(:, 32p..32p+32),  B = F(:, cbeg..cend)      -> wWpi [slot][split][c-cbeg][32]
//   WIDE_OUTER  V = Vb(:, 0..128),       B = F(:, cbeg..cend)      -> wWp  [slot][split][c][128]
//   WIDE_GRAM   V = Vb(:, 0..128),       B = Vb(:, 0..128)         -> wGp  [slot][split][c][128]
// grid = (column tiles of 64 x splits, fronts).
// ---------------------------------------------------------------------------------------------
template <int KVT>
__global__ void __launch_bounds__ (256) k_wide_vtc (WideArgs A, DSym S, DNum N, I32 mode, I32 p, I32 cbeg,
    I32 cend, I32 nct_in)
{
    constexpr int KV = PANEL_MAX * KVT ;
    constexpr int STAGE = (KV + W_NC) * W_LDC ;
    extern __shared__ double sm [] ;
    const I32 slot = blockIdx.y ;
    const I32 *blk = wide_blk (A, N, slot) ;
    if (mode == WIDE_INNER ? (N.pnl_nv [p * A.count + slot] == 0) : (blk [2] == 0)) return ;
    const I32 g0 = blk [0], mr = blk [1] ;
    const I32 f = A.fronts [slot] ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    const I32 nct = (KVT == 1) ? nct_in + (p * PANEL_MAX + W_NC - 1) / W_NC : nct_in ;     // column tiles per split
    I32 ct = blockIdx.x % nct ;
    const I32 split = blockIdx.x / nct ;
    const I32 rbeg = split * A.rs ;
    if (rbeg >= mr) return ;
    const I32 rend = min (mr, rbeg + A.rs) ;
    // WIDE_INNER launches carry extra column tiles (ct >= nct_in) that multiply V_p with the Householder
    // vectors of the EARLIER panels of the block: the Gram blocks k_wide_tmerge needs, for free
    if (mode == WIDE_INNER && ct >= nct_in) { mode = WIDE_GRAM ; ct -= nct_in ; cbeg = 0 ; }
    const I32 clim = (mode == WIDE_GRAM) ? ((KVT == 1) ? p * PANEL_MAX : WB) : min (cend, fn) ;
    const I32 c0 = cbeg + ct * W_NC ;
    if (c0 >= clim) return ;
    const I32 ncol = min (W_NC, clim - c0) ;
    const I64 fm = N.Hm [f] ;
    const double *Vb = wide_vb (A, N, slot) ;
    const double *Va = Vb + (I64) ((KVT == 1) ? p * PANEL_MAX : 0) * A.ldv ;
    const double *Bp = (mode == WIDE_GRAM) ? Vb : (N.F + S.Foff [f] + g0) ;
    const I64 ldb = (mode == WIDE_GRAM) ? (I64) A.ldv : fm ;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5 ;
    const int grp = lane >> 2, tig = lane & 3 ;
    // warp tile: 32 reflectors x 32 columns; KVT = 4: 4 x 2 warps over (q, c), every warp all rows of
    // the chunk; KVT = 1: 2 warps over c, 4 over the rows of the chunk (partials summed at the end)
    const int wq = (KVT == 4) ? (w & 3) : 0 ;
    const int wc = (KVT == 4) ? (w >> 2) : (w & 1) ;
    const int wr = (KVT == 4) ? 0 : (w >> 1) ;

    auto issue = [&] (double *stage, const I32 r0)
    {
        double *Vs = stage, *Cs = stage + KV * W_LDC ;
#pragma unroll
        for (int a = 0 ; a < (KV * (W_RC / 2)) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            const int q = e >> 4, gq = e & 15 ;
            cp_async16 (Vs + q * W_LDC + 2 * gq, Va + (I64) q * A.ldv + r0 + 2 * gq) ;
        }
#pragma unroll
        for (int a = 0 ; a < (W_NC * W_RC) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            const int c = e >> 5, rr = e & 31 ;
            const bool ok = (c < ncol) && (r0 + rr < rend) ;
            cp_async8 (Cs + c * W_LDC + rr, ok ? (Bp + (I64) (c0 + c) * ldb + r0 + rr) : Bp, ok) ;
        }
    } ;

    double acc [4][4][2] ;
#pragma unroll
    for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
        for (int ni = 0 ; ni < 4 ; ni++) { acc [mi][ni][0] = 0 ; acc [mi][ni][1] = 0 ; }

    const I32 nch = (rend - rbeg + W_RC - 1) / W_RC ;
#pragma unroll
    for (int s = 0 ; s < W_NST - 1 ; s++)
    {
        if (s < nch) issue (sm + s * STAGE, rbeg + s * W_RC) ;
        cp_async_commit () ;
    }
    for (I32 ch = 0 ; ch < nch ; ch++)
    {
        if (ch + W_NST - 1 < nch) issue (sm + ((ch + W_NST - 1) % W_NST) * STAGE, rbeg + (ch + W_NST - 1) * W_RC) ;
        cp_async_commit () ;
        cp_async_wait<W_NST - 1> () ;
        __syncthreads () ;
        const double *Vs = sm + (ch % W_NST) * STAGE, *Cs = Vs + KV * W_LDC ;
        constexpr int KS = (KVT == 4) ? (W_RC / 4) : (W_RC / 16) ;
#pragma unroll
        for (int ks = 0 ; ks < KS ; ks++)
        {
            const int rr = (wr * KS + ks) * 4 + tig ;
            double af [4], bf [4] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++) af [mi] = Vs [(wq * 32 + mi * 8 + grp) * W_LDC + rr] ;
#pragma unroll
            for (int ni = 0 ; ni < 4 ; ni++) bf [ni] = Cs [(wc * 32 + ni * 8 + grp) * W_LDC + rr] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
                for (int ni = 0 ; ni < 4 ; ni++) dmma_m8n8k4 (acc [mi][ni][0], acc [mi][ni][1], af [mi], bf [ni]) ;
        }
        __syncthreads () ;
    }
    cp_async_wait<0> () ;

    double *out ;
    I64 ldo ;
    if (mode == WIDE_OUTER) { out = N.wWp + ((I64) slot * A.nsplit + split) * ((I64) A.ncmax * WB) + (I64) c0 * WB ; ldo = WB ; }
    else if (mode == WIDE_GRAM && KVT == 1)
    {
        // (V_p' V_prev)(q, c), c < 32 p:  [slot][split][p-1][c][q]
        out = N.wGp + (((I64) slot * A.nsplit + split) * (WB_PANELS - 1) + (p - 1)) * (96 * PANEL_MAX) + (I64) c0 * PANEL_MAX ;
        ldo = PANEL_MAX ;
    }
    else if (mode == WIDE_GRAM) { out = N.wGp + ((I64) slot * A.nsplit + split) * (WB * WB) + (I64) c0 * WB ; ldo = WB ; }
    else { out = N.wWpi + (((I64) slot * A.nsplit + split) * WB + (c0 - cbeg)) * PANEL_MAX ; ldo = PANEL_MAX ; }
    if (KVT == 1)
    {
        // the 4 row groups of the chunk: partial tiles through shared memory, summed in a fixed order
        double *part = sm ;                                 // [4][64][33]
#pragma unroll
        for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
            for (int ni = 0 ; ni < 4 ; ni++)
#pragma unroll
                for (int e = 0 ; e < 2 ; e++)
                    part [(wr * W_NC + wc * 32 + ni * 8 + tig * 2 + e) * 33 + mi * 8 + grp] = acc [mi][ni][e] ;
        __syncthreads () ;
        for (int e = tid ; e < W_NC * PANEL_MAX ; e += 256)
        {
            const int c = e >> 5, q = e & 31 ;
            if (c < ncol)
                out [(I64) c * ldo + q] = (part [c * 33 + q] + part [(W_NC + c) * 33 + q])
                    + (part [(2 * W_NC + c) * 33 + q] + part [(3 * W_NC + c) * 33 + q]) ;
        }
    }
    else
    {
#pragma unroll
        for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
            for (int ni = 0 ; ni < 4 ; ni++)
#pragma unroll
                for (int e = 0 ; e < 2 ; e++)
                {
                    const int c = wc * 32 + ni * 8 + tig * 2 + e ;
                    if (c < ncol) out [(I64) c * ldo + wq * 32 + mi * 8 + grp] = acc [mi][ni][e] ;
                }
    }
}
template <int KVT> constexpr size_t wide_vtc_smem_bytes ()
{
    return sizeof (double) * (size_t) ((W_NST * (PANEL_MAX * KVT + W_NC) * W_LDC > 4 * W_NC * 33)
        ? W_NST * (PANEL_MAX * KVT + W_NC) * W_LDC : 4 * W_NC * 33) ;
}

// ---------------------------------------------------------------------------------------------
// W2 = -T' (sum over the splits of W_partial).  grid = (column tiles of 16, fronts).
// ---------------------------------------------------------------------------------------------
template <int KVT>
__global__ void __launch_bounds__ (256) k_wide_wt (WideArgs A, DSym S, DNum N, I32 mode, I32 p, I32 cbeg, I32 cend)
{
    constexpr int KV = PANEL_MAX * KVT ;
    constexpr int NCW = 16 ;
    __shared__ double Ws [NCW * (KV + 1)] ;
    const I32 slot = blockIdx.y ;
    const I32 *blk = wide_blk (A, N, slot) ;
    const I32 nvp = (mode == WIDE_INNER) ? N.pnl_nv [p * A.count + slot] : blk [2] ;
    if (nvp == 0) return ;
    const I32 mr = blk [1] ;
    const I32 f = A.fronts [slot] ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    const I32 clim = min (cend, fn) ;
    const I32 c0 = cbeg + blockIdx.x * NCW ;
    if (c0 >= clim) return ;
    const I32 ncol = min (NCW, clim - c0) ;
    const I32 nsp = min (A.nsplit, (mr + A.rs - 1) / A.rs) ;
    const int tid = threadIdx.x ;
    const double *Wp ;
    I64 sstride ;
    if (mode == WIDE_OUTER) { Wp = N.wWp + (I64) slot * A.nsplit * ((I64) A.ncmax * WB) + (I64) c0 * WB ; sstride = (I64) A.ncmax * WB ; }
    else { Wp = N.wWpi + ((I64) slot * A.nsplit * WB + (c0 - cbeg)) * PANEL_MAX ; sstride = (I64) WB * PANEL_MAX ; }
    for (int e = tid ; e < NCW * KV ; e += 256)
    {
        const int cc = e / KV, q = e % KV ;
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0 ;
        if (cc < ncol)
        {
            const double *src = Wp + (I64) cc * KV + q ;
            I32 sp = 0 ;
            for ( ; sp + 3 < nsp ; sp += 4)
            {
                const double a = __ldcg (src + sp * sstride), b = __ldcg (src + (sp+1) * sstride),
                    c = __ldcg (src + (sp+2) * sstride), d = __ldcg (src + (sp+3) * sstride) ;
                s0 += a ; s1 += b ; s2 += c ; s3 += d ;
            }
            for ( ; sp < nsp ; sp++) s0 += __ldcg (src + sp * sstride) ;
        }
        Ws [cc * (KV + 1) + q] = (s0 + s1) + (s2 + s3) ;
    }
    __syncthreads () ;
    // thread (q', column group): out(q', c) = -sum_q T(q, q') Wsum(q, c)
    constexpr int NCG = 256 / KV, CPT = NCW / NCG ;
    const int qp = tid % KV, cg = tid / KV ;
    double acc [CPT] ;
#pragma unroll
    for (int j = 0 ; j < CPT ; j++) acc [j] = 0 ;
    if (mode == WIDE_OUTER)
    {
        // Tt[q' + q*128] = T(q,q'): 16 rows q at a time through shared memory (coalesced, conflict-free)
        __shared__ double Tsm [16 * WB] ;
        const double *Tt = N.wTbt + ((I64) A.buf * A.count + slot) * (WB * WB) ;
        for (int q0 = 0 ; q0 < KV ; q0 += 16)
        {
            for (int e = tid ; e < 16 * WB ; e += 256) Tsm [e] = __ldcg (Tt + q0 * WB + e) ;
            __syncthreads () ;
#pragma unroll 8
            for (int q = 0 ; q < 16 ; q++)
            {
                const double t = Tsm [q * WB + (qp & (WB - 1))] ;
#pragma unroll
                for (int j = 0 ; j < CPT ; j++) acc [j] = fma (t, Ws [(cg * CPT + j) * (KV + 1) + q0 + q], acc [j]) ;
            }
            __syncthreads () ;
        }
    }
    else
    {
        const double *Tg = N.Tws + (I64) (p * A.count + slot) * (PANEL_MAX * PANEL_MAX) ;  // T(j,i) at Tg[j + 32 i]
        if (qp < nvp)
            for (int q = 0 ; q <= qp ; q++)
            {
                const double t = __ldg (Tg + q + qp * PANEL_MAX) ;
#pragma unroll
                for (int j = 0 ; j < CPT ; j++) acc [j] = fma (t, Ws [(cg * CPT + j) * (KV + 1) + q], acc [j]) ;
            }
    }
    double *W2 = (mode == WIDE_OUTER) ? (N.wW2 + (I64) slot * ((I64) A.ncmax * WB) + (I64) c0 * WB)
        : (N.wW2i + ((I64) slot * WB + (c0 - cbeg)) * PANEL_MAX) ;
#pragma unroll
    for (int j = 0 ; j < CPT ; j++)
    {
        const int cc = cg * CPT + j ;
        if (cc < ncol) W2 [(I64) cc * KV + qp] = -acc [j] ;
    }
}

// The same for the update inside the block (32 reflectors of panel p, <= 96 columns): 8 columns per
// CTA, one thread per entry; the partial sums over the row splits are independent loads.
__global__ void __launch_bounds__ (256) k_wide_wt_inner (WideArgs A, DSym S, DNum N, I32 p, I32 cbeg, I32 cend)
{
    __shared__ double Ws [8][PANEL_MAX + 1] ;
    __shared__ double Ts [PANEL_MAX][PANEL_MAX + 1] ;
    const I32 slot = blockIdx.y ;
    const I32 nvp = N.pnl_nv [p * A.count + slot] ;
    if (nvp == 0) return ;
    const I32 *blk = wide_blk (A, N, slot) ;
    const I32 mr = blk [1] ;
    const I32 f = A.fronts [slot] ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    const I32 clim = min (cend, fn) ;
    const I32 c0 = cbeg + blockIdx.x * 8 ;
    if (c0 >= clim) return ;
    const I32 nsp = min (A.nsplit, (mr + A.rs - 1) / A.rs) ;
    const int tid = threadIdx.x, cc = tid >> 5, q = tid & 31 ;
    const double *Tg = N.Tws + (I64) (p * A.count + slot) * (PANEL_MAX * PANEL_MAX) ;   // T(j,i) at Tg[j + 32 i]
    for (int e = tid ; e < PANEL_MAX * PANEL_MAX ; e += 256)
    {
        const int j = e & 31, i = e >> 5 ;
        Ts [j][i] = (j <= i && i < nvp) ? Tg [e] : 0.0 ;
    }
    const bool live = (c0 + cc < clim) ;
    const double *Wp = N.wWpi + ((I64) slot * A.nsplit * WB + (c0 - cbeg + cc)) * PANEL_MAX + q ;
    const I64 sstride = (I64) WB * PANEL_MAX ;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0 ;
    if (live)
    {
        I32 sp = 0 ;
        for ( ; sp + 3 < nsp ; sp += 4)
        {
            const double a = __ldcg (Wp + sp * sstride), b = __ldcg (Wp + (sp+1) * sstride),
                c = __ldcg (Wp + (sp+2) * sstride), d = __ldcg (Wp + (sp+3) * sstride) ;
            s0 += a ; s1 += b ; s2 += c ; s3 += d ;
        }
        for ( ; sp < nsp ; sp++) s0 += __ldcg (Wp + sp * sstride) ;
    }
    Ws [cc][q] = (s0 + s1) + (s2 + s3) ;
    __syncthreads () ;
    double a0 = 0, a1 = 0 ;
#pragma unroll
    for (int j = 0 ; j < PANEL_MAX ; j += 2)
    {
        a0 = fma (Ts [j][q], Ws [cc][j], a0) ;          // T(j, q), zero for j > q
        a1 = fma (Ts [j+1][q], Ws [cc][j+1], a1) ;
    }
    if (live) N.wW2i [((I64) slot * WB + (c0 - cbeg + cc)) * PANEL_MAX + q] = -(a0 + a1) ;
}

// ---------------------------------------------------------------------------------------------
// C(rows of the block, cbeg..cend) += V W2.   grid = (column tiles of 64 x row tiles of 128, fronts).
// ---------------------------------------------------------------------------------------------
template <int KVT>
__global__ void __launch_bounds__ (256) k_wide_apply (WideArgs A, DSym S, DNum N, I32 mode, I32 p, I32 cbeg,
    I32 cend, I32 nct)
{
    constexpr int KV = PANEL_MAX * KVT ;
    constexpr int STAGE = PANEL_MAX * W_LDV + W_NC * W_LDC ;
    constexpr int NST = (KVT == 1) ? 1 : 2 ;        // 2 stages = 104 KB: two CTAs per SM overlap fill and drain
    extern __shared__ double sm [] ;
    const I32 slot = blockIdx.y ;
    const I32 *blk = wide_blk (A, N, slot) ;
    if (mode == WIDE_INNER ? (N.pnl_nv [p * A.count + slot] == 0) : (blk [2] == 0)) return ;
    const I32 g0 = blk [0], mr = blk [1] ;
    const I32 f = A.fronts [slot] ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    const I32 ct = blockIdx.x % nct, rt = blockIdx.x / nct ;
    const I32 r0 = rt * W_RT ;
    if (r0 >= mr) return ;
    const I32 clim = min (cend, fn) ;
    const I32 c0 = cbeg + ct * W_NC ;
    if (c0 >= clim) return ;
    const I32 ncol = min (W_NC, clim - c0) ;
    const I64 fm = N.Hm [f] ;
    double *C = N.F + S.Foff [f] + g0 + r0 + (I64) c0 * fm ;
    const double *Va = wide_vb (A, N, slot) + (I64) ((mode == WIDE_INNER) ? p * PANEL_MAX : 0) * A.ldv + r0 ;
    const double *W2 = (mode == WIDE_OUTER) ? (N.wW2 + (I64) slot * ((I64) A.ncmax * WB) + (I64) c0 * WB)
        : (N.wW2i + ((I64) slot * WB + (c0 - cbeg)) * PANEL_MAX) ;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5 ;
    const int grp = lane >> 2, tig = lane & 3 ;
    const int wr = w & 3, wc = w >> 2 ;             // 4 x 2 warps over (rows, columns), 32 x 32 each

    auto issue = [&] (double *stage, const int kk)
    {
        double *Vs = stage, *Ws = stage + PANEL_MAX * W_LDV ;
#pragma unroll
        for (int a = 0 ; a < (PANEL_MAX * (W_RT / 2)) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            const int k = e >> 6, g = e & 63 ;
            cp_async16 (Vs + k * W_LDV + 2 * g, Va + (I64) (kk * PANEL_MAX + k) * A.ldv + 2 * g) ;
        }
#pragma unroll
        for (int a = 0 ; a < (W_NC * (PANEL_MAX / 2)) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            const int c = e >> 4, g = e & 15 ;
            const bool ok = (c < ncol) ;
            cp_async16z (Ws + c * W_LDC + 2 * g, ok ? (W2 + (I64) c * KV + kk * PANEL_MAX + 2 * g) : W2, ok) ;
        }
    } ;

#pragma unroll
    for (int s = 0 ; s < NST - 1 ; s++)
    {
        if (s < KVT) issue (sm + s * STAGE, s) ;
        cp_async_commit () ;
    }
    if (NST == 1) { issue (sm, 0) ; cp_async_commit () ; }

    double acc [4][4][2], cold [4][4][2] ;
#pragma unroll
    for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
        for (int ni = 0 ; ni < 4 ; ni++)
#pragma unroll
            for (int e = 0 ; e < 2 ; e++)
            {
                const int r = wr * 32 + mi * 8 + grp, c = wc * 32 + ni * 8 + tig * 2 + e ;
                acc [mi][ni][e] = 0.0 ;
                cold [mi][ni][e] = (r0 + r < mr && c < ncol) ? __ldcg (C + r + (I64) c * fm) : 0.0 ;
            }

    for (int kk = 0 ; kk < KVT ; kk++)
    {
        if (NST > 1)
        {
            if (kk + NST - 1 < KVT) issue (sm + ((kk + NST - 1) % NST) * STAGE, kk + NST - 1) ;
            cp_async_commit () ;
            cp_async_wait<NST - 1> () ;
        }
        else cp_async_wait<0> () ;
        __syncthreads () ;
        const double *Vs = sm + (kk % NST) * STAGE, *Ws = Vs + PANEL_MAX * W_LDV ;
#pragma unroll
        for (int ks = 0 ; ks < PANEL_MAX / 4 ; ks++)
        {
            double af [4], bf [4] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++) af [mi] = Vs [(ks * 4 + tig) * W_LDV + wr * 32 + mi * 8 + grp] ;
#pragma unroll
            for (int ni = 0 ; ni < 4 ; ni++) bf [ni] = Ws [(wc * 32 + ni * 8 + grp) * W_LDC + ks * 4 + tig] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
                for (int ni = 0 ; ni < 4 ; ni++) dmma_m8n8k4 (acc [mi][ni][0], acc [mi][ni][1], af [mi], bf [ni]) ;
        }
        __syncthreads () ;
    }
    cp_async_wait<0> () ;
#pragma unroll
    for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
        for (int ni = 0 ; ni < 4 ; ni++)
#pragma unroll
            for (int e = 0 ; e < 2 ; e++)
            {
                const int r = wr * 32 + mi * 8 + grp, c = wc * 32 + ni * 8 + tig * 2 + e ;
                if (r0 + r < mr && c < ncol) C [r + (I64) c * fm] = cold [mi][ni][e] + acc [mi][ni][e] ;
            }
}
template <int KVT> constexpr size_t wide_apply_smem_bytes ()
{
    return sizeof (double) * (size_t) (((KVT == 1) ? 1 : 2) * (PANEL_MAX * W_LDV + W_NC * W_LDC)) ;
}

// ---------------------------------------------------------------------------------------------
// The K = 128 apply, persistent over row tiles: a CTA owns 64 columns, keeps its 128 x 64 tile of W2
// in shared memory and walks down the rows (tiles rg, rg + nrg, ...) with ONE continuous cp.async ring
// over the (row tile, 32-reflector slab) pairs of V: no pipeline fill / drain per 128 rows, W2 loaded
// once.  The old values of C are loaded into registers when a row tile starts and added at its end.
// grid = (column tiles x nrg, fronts).
// ---------------------------------------------------------------------------------------------
constexpr int WP_NST = 3 ;
__global__ void __launch_bounds__ (256) k_wide_apply_rows (WideArgs A, DSym S, DNum N, I32 cbeg, I32 cend, I32 nct,
    I32 nrg)
{
    constexpr int KVT = WB_PANELS ;
    constexpr int VSLAB = PANEL_MAX * W_LDV ;           // one 32-reflector slab of V: [32][132]
    constexpr int WSLAB = W_NC * W_LDC ;                // one 32-reflector slab of W2: [64][36]
    extern __shared__ double sm [] ;
    double *Wsm = sm ;                                  // [4][WSLAB]
    double *ring = sm + KVT * WSLAB ;                   // [WP_NST][VSLAB]
    const I32 slot = blockIdx.y ;
    const I32 *blk = wide_blk (A, N, slot) ;
    if (blk [2] == 0) return ;
    const I32 g0 = blk [0], mr = blk [1] ;
    const I32 f = A.fronts [slot] ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    const I32 ct = blockIdx.x % nct, rg = blockIdx.x / nct ;
    const I32 nrt = (mr + W_RT - 1) / W_RT ;
    if (rg >= nrt) return ;
    const I32 clim = min (cend, fn) ;
    const I32 c0 = cbeg + ct * W_NC ;
    if (c0 >= clim) return ;
    const I32 ncol = min (W_NC, clim - c0) ;
    const I64 fm = N.Hm [f] ;
    double *Cbase = N.F + S.Foff [f] + g0 + (I64) c0 * fm ;
    const double *Vb = wide_vb (A, N, slot) ;
    const double *W2 = N.wW2 + (I64) slot * ((I64) A.ncmax * WB) + (I64) c0 * WB ;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5 ;
    const int grp = lane >> 2, tig = lane & 3 ;
    const int wr = w & 3, wc = w >> 2 ;

    const I32 ntile = (nrt - rg + nrg - 1) / nrg ;      // my row tiles: rg, rg + nrg, ...
    const I32 nq = ntile * KVT ;                        // (tile, slab) pairs

    // W2 tile, all four slabs (first cp.async group)
#pragma unroll
    for (int a = 0 ; a < (KVT * W_NC * (PANEL_MAX / 2)) / 256 ; a++)
    {
        const int e = tid + a * 256 ;
        const int kk = e >> 10, c = (e >> 4) & 63, g = e & 15 ;
        const bool ok = (c < ncol) ;
        cp_async16z (Wsm + kk * WSLAB + c * W_LDC + 2 * g, ok ? (W2 + (I64) c * WB + kk * PANEL_MAX + 2 * g) : W2, ok) ;
    }
    auto issue = [&] (const I32 q)
    {
        const I32 r0 = (rg + (q / KVT) * nrg) * W_RT ;
        const int kk = q % KVT ;
        double *Vs = ring + (q % WP_NST) * VSLAB ;
        const double *Va = Vb + (I64) (kk * PANEL_MAX) * A.ldv + r0 ;
#pragma unroll
        for (int a = 0 ; a < (PANEL_MAX * (W_RT / 2)) / 256 ; a++)
        {
            const int e = tid + a * 256 ;
            const int k = e >> 6, g = e & 63 ;
            cp_async16 (Vs + k * W_LDV + 2 * g, Va + (I64) k * A.ldv + 2 * g) ;
        }
    } ;
#pragma unroll
    for (int s = 0 ; s < WP_NST - 1 ; s++)
    {
        if (s < nq) issue (s) ;
        cp_async_commit () ;            // (the W2 tile rides in the first group)
    }

    double acc [4][4][2], cold [4][4][2] ;
    for (I32 q = 0 ; q < nq ; q++)
    {
        const int kk = q % KVT ;
        const I32 r0 = (rg + (q / KVT) * nrg) * W_RT ;
        if (kk == 0)
        {
            double *C = Cbase + r0 ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
                for (int ni = 0 ; ni < 4 ; ni++)
#pragma unroll
                    for (int e = 0 ; e < 2 ; e++)
                    {
                        const int r = wr * 32 + mi * 8 + grp, c = wc * 32 + ni * 8 + tig * 2 + e ;
                        acc [mi][ni][e] = 0.0 ;
                        cold [mi][ni][e] = (r0 + r < mr && c < ncol) ? __ldcg (C + r + (I64) c * fm) : 0.0 ;
                    }
        }
        if (q + WP_NST - 1 < nq) issue (q + WP_NST - 1) ;
        cp_async_commit () ;
        cp_async_wait<WP_NST - 1> () ;
        __syncthreads () ;
        const double *Vs = ring + (q % WP_NST) * VSLAB, *Ws = Wsm + kk * WSLAB ;
#pragma unroll
        for (int ks = 0 ; ks < PANEL_MAX / 4 ; ks++)
        {
            double af [4], bf [4] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++) af [mi] = Vs [(ks * 4 + tig) * W_LDV + wr * 32 + mi * 8 + grp] ;
#pragma unroll
            for (int ni = 0 ; ni < 4 ; ni++) bf [ni] = Ws [(wc * 32 + ni * 8 + grp) * W_LDC + ks * 4 + tig] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
                for (int ni = 0 ; ni < 4 ; ni++) dmma_m8n8k4 (acc [mi][ni][0], acc [mi][ni][1], af [mi], bf [ni]) ;
        }
        __syncthreads () ;
        if (kk == KVT - 1)
        {
            double *C = Cbase + r0 ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
                for (int ni = 0 ; ni < 4 ; ni++)
#pragma unroll
                    for (int e = 0 ; e < 2 ; e++)
                    {
                        const int r = wr * 32 + mi * 8 + grp, c = wc * 32 + ni * 8 + tig * 2 + e ;
                        if (r0 + r < mr && c < ncol) C [r + (I64) c * fm] = cold [mi][ni][e] + acc [mi][ni][e] ;
                    }
        }
    }
    cp_async_wait<0> () ;
}
constexpr size_t wide_apply_rows_smem_bytes ()
{
    return sizeof (double) * (size_t) (WB_PANELS * W_NC * W_LDC + WP_NST * PANEL_MAX * W_LDV) ;
}

// ---------------------------------------------------------------------------------------------
// T of the outer block from its panels' T (dlarft of each panel, k_panel_cluster) and the Gram
// matrix G = Vb'Vb:  T(0:32p, p) = -T(0:32p,0:32p) G(0:32p, p) T(p,p).  One CTA per front.  Output
// transposed (k_wide_wt reads it coalesced): Tbt[q' + 128 q] = T(q,q').
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (1024) k_wide_tmerge (WideArgs A, DSym S, DNum N)
{
    extern __shared__ double sm [] ;
    constexpr int LT = WB + 1 ;
    constexpr int NTH = 1024 ;
    double *T = sm ;                        // [128][129]
    double *X = T + WB * LT ;               // [96][33]
    double *G = X + 96 * 33 ;               // [96][33]
    const I32 slot = blockIdx.x ;
    const I32 *blk = wide_blk (A, N, slot) ;
    if (blk [2] == 0) return ;
    const I32 g0 = blk [0] ;
    const int tid = threadIdx.x ;
    for (int e = tid ; e < WB * LT ; e += NTH) T [e] = 0.0 ;
    __syncthreads () ;
    for (int p = 0 ; p < WB_PANELS ; p++)
    {
        const I32 nv = N.pnl_nv [p * A.count + slot] ;
        const double *Tg = N.Tws + (I64) (p * A.count + slot) * (PANEL_MAX * PANEL_MAX) ;
        {
            const int e = tid ;
            const int j = e & 31, i = e >> 5 ;          // T_pp(j,i), j <= i
            if (j <= i && i < nv) T [(p * 32 + j) * LT + p * 32 + i] = Tg [j + i * PANEL_MAX] ;
        }
    }
    __syncthreads () ;
    // Gram blocks (V_p' V_prev), partial per row split, written by the inner k_wide_vtc launches
    const double *Gp = N.wGp + (I64) slot * A.nsplit * (WB_PANELS - 1) * (96 * PANEL_MAX) ;
    const I64 sstride = (I64) (WB_PANELS - 1) * (96 * PANEL_MAX) ;
    for (int p = 1 ; p < WB_PANELS ; p++)
    {
        const int np = 32 * p ;
        const I32 nvp = N.pnl_nv [p * A.count + slot] ;
        if (nvp == 0) continue ;                        // uniform over the CTA: block column stays zero
        const I32 mrp = N.pnl_tend [p * A.count + slot] - g0 ;
        const I32 nsp = min (A.nsplit, (mrp + A.rs - 1) / A.rs) ;
        for (int e = tid ; e < np * 32 ; e += NTH)
        {
            const int l = e & 31, i = e >> 5 ;          // G(i, 32p + l) = (V_p' V_prev)(l, i)
            const double *src = Gp + (I64) (p - 1) * (96 * PANEL_MAX) + i * PANEL_MAX + l ;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0 ;
            I32 sp = 0 ;
            for ( ; sp + 3 < nsp ; sp += 4)
            {
                const double a = __ldcg (src + sp * sstride), b = __ldcg (src + (sp+1) * sstride),
                    c = __ldcg (src + (sp+2) * sstride), d = __ldcg (src + (sp+3) * sstride) ;
                s0 += a ; s1 += b ; s2 += c ; s3 += d ;
            }
            for ( ; sp < nsp ; sp++) s0 += __ldcg (src + sp * sstride) ;
            G [i * 33 + l] = (s0 + s1) + (s2 + s3) ;
        }
        __syncthreads () ;
        for (int e = tid ; e < np * 32 ; e += NTH)
        {
            const int j = e & 31, i = e >> 5 ;          // X(i,j) = sum_l G(i,l) T_pp(l,j)
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0 ;
#pragma unroll
            for (int l = 0 ; l < 32 ; l += 4)
            {
                s0 = fma (G [i * 33 + l], T [(np + l) * LT + np + j], s0) ;
                s1 = fma (G [i * 33 + l + 1], T [(np + l + 1) * LT + np + j], s1) ;
                s2 = fma (G [i * 33 + l + 2], T [(np + l + 2) * LT + np + j], s2) ;
                s3 = fma (G [i * 33 + l + 3], T [(np + l + 3) * LT + np + j], s3) ;
            }
            X [i * 33 + j] = (s0 + s1) + (s2 + s3) ;
        }
        __syncthreads () ;
        for (int e = tid ; e < np * 32 ; e += NTH)
        {
            const int j = e & 31, i = e >> 5 ;          // T(i, 32p+j) = -sum_{l>=i} T(i,l) X(l,j)
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0 ;
            int l = i ;
            for ( ; l + 3 < np ; l += 4)
            {
                s0 = fma (T [i * LT + l], X [l * 33 + j], s0) ;
                s1 = fma (T [i * LT + l + 1], X [(l + 1) * 33 + j], s1) ;
                s2 = fma (T [i * LT + l + 2], X [(l + 2) * 33 + j], s2) ;
                s3 = fma (T [i * LT + l + 3], X [(l + 3) * 33 + j], s3) ;
            }
            for ( ; l < np ; l++) s0 = fma (T [i * LT + l], X [l * 33 + j], s0) ;
            T [i * LT + np + j] = -((s0 + s1) + (s2 + s3)) ;
        }
        __syncthreads () ;
    }
    double *Tt = N.wTbt + ((I64) A.buf * A.count + slot) * (WB * WB) ;
    for (int e = tid ; e < WB * WB ; e += NTH)
    {
        const int qp = e & (WB - 1), q = e >> 7 ;
        Tt [qp + q * WB] = T [q * LT + qp] ;
    }
}
constexpr size_t wide_tmerge_smem_bytes () { return sizeof (double) * (size_t) (WB * (WB + 1) + 2 * 96 * 33) ; }

} // namespace stmqr
