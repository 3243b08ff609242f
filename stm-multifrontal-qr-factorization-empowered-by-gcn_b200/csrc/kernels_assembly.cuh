// kernels_assembly.cuh -- S construction, front set-up (qr_fsize), assembly (qr_assemble),
// packing (qr_cpack, qr_rhpack) and the row permutation (qr_hpinv).  All HBM-bound
// integer/copy work: coalesced along front columns (column-major F, ld = fm).
#pragma once
#include "engine.cuh"

namespace stmqr {

// ---------------------------------------------------------------------------------------------
// S = A(P,Q) values in row form.  Reference: qr_stranspose2, SparseQR_factorize.c:755-785, a
// serial scatter through a running row cursor.  Here one thread per entry of A finds its slot by
// binary search of the permuted column inside the (ascending) row of S: no atomics, deterministic.
// Duplicate entries of A (legal in a sparse_csc) land in consecutive slots in storage order, exactly
// as the reference's cursor places them (the later one then wins in qr_assemble, :1199-1203).  (Through
// SparseQR() the case cannot arise: the reference's own analysis refuses a matrix with duplicate entries
// with SPARSE_INVALID -- measured with oracle/_ref, "HNUCHOL error: all methods failed",
// SparseChol_analyze.c:640 -- so this only matters to callers of the C ABI with their own symbolic object.)
// ---------------------------------------------------------------------------------------------
// slot [p] (optional): the slot of S that entry p of A goes to -- the whole search is symbolic, so a
// refactorization with new values on the same pattern is a plain scatter (k_scatter_values).
__global__ void k_build_S (I32 n, const I64 *__restrict__ Ap, const I64 *__restrict__ Ai,
    const double *__restrict__ Ax, DSym S, double *__restrict__ Sx, I32 *__restrict__ slot, I32 *err)
{
    // one warp per column of A keeps the reads of Ai/Ax coalesced
    const int lane = threadIdx.x & 31 ;
    const I64 warp = ((I64) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    const I64 nwarps = ((I64) gridDim.x * blockDim.x) >> 5 ;
    for (I64 j = warp ; j < n ; j += nwarps)
    {
        const I32 col = S.Qinv [j] ;
        const I64 p1 = Ap [j], p2 = Ap [j+1] ;
        for (I64 p = p1 + lane ; p < p2 ; p += 32)
        {
            const I32 row = S.PLinv [Ai [p]] ;
            I32 lo = S.Sp [row], hi = S.Sp [row+1] - 1 ;
            while (lo < hi)
            {
                I32 mid = (lo + hi) >> 1 ;
                if (S.Sj [mid] < col) lo = mid + 1 ; else hi = mid ;
            }
            if (lo < S.Sp [row+1] && S.Sj [lo] == col)
            {
                // duplicate (i,j) entries of A occupy consecutive slots of S in the order they are
                // stored in the column (qr_stranspose2's running cursor W[row]++, :779-781): the d-th
                // duplicate of this entry goes to slot lo + d.  Rare: only looked at when the next slot
                // holds the same column.
                if (lo + 1 < S.Sp [row+1] && S.Sj [lo+1] == col)
                {
                    const I64 ai = Ai [p] ;
                    I32 d = 0 ;
                    for (I64 q = p1 ; q < p ; q++) d += (Ai [q] == ai) ;
                    lo += d ;
                    if (!(lo < S.Sp [row+1] && S.Sj [lo] == col)) { atomicExch (err, 1) ; if (slot) slot [p] = -1 ; continue ; }
                }
                Sx [lo] = Ax [p] ;
                if (slot) slot [p] = lo ;
            }
            else { atomicExch (err, 1) ; if (slot) slot [p] = -1 ; }
        }
    }
}

// Values-only refactorization (same pattern, new values): Sx [slot [p]] = Ax [p].  Coalesced reads of Ax
// and slot, 8-byte scattered writes (every slot of S is written exactly once when A has no duplicates;
// with duplicates the later entry wins, as in the reference, because slots of duplicates are distinct).
__global__ void k_scatter_values (I64 nnz, const double *__restrict__ Ax, const I32 *__restrict__ slot,
    double *__restrict__ Sx)
{
    I64 p = (I64) blockIdx.x * blockDim.x + threadIdx.x ;
    const I64 stride = (I64) gridDim.x * blockDim.x ;
    for ( ; p < nnz ; p += stride)
    {
        const I32 sl = slot [p] ;
        if (sl >= 0) Sx [sl] = Ax [p] ;
    }
}

// The pattern of the matrix handed to a speculative values-only refactorization, compared on the device
// with the resident pattern while the numeric phase runs: *differs != 0 if any column pointer or row
// index changed (the host then repeats the factorization through the full path).
__global__ void k_compare_pattern (I64 ncol1, I64 nnz, const I64 *__restrict__ Ap, const I64 *__restrict__ Ai,
    const I64 *__restrict__ Bp, const I64 *__restrict__ Bi, I32 *differs)
{
    I64 p = (I64) blockIdx.x * blockDim.x + threadIdx.x ;
    const I64 stride = (I64) gridDim.x * blockDim.x ;
    bool bad = false ;
    for (I64 q = p ; q < ncol1 ; q += stride) bad |= (Ap [q] != Bp [q]) ;
    for (I64 q = p ; q < nnz ; q += stride) bad |= (Ai [q] != Bi [q]) ;
    if (bad) atomicExch (differs, 1) ;
}

// ---------------------------------------------------------------------------------------------
// Front set-up for every front of one etree level: qr_fsize (SparseQR_factorize.c:1066-1145)
// plus the integer half of qr_assemble (:1188-1248): row start of every column (Stair), # rows
// fm, the row of every original row of S (rowpos), the row of every child C row (Cmap) and the
// row ids Hii.  One CTA per front.  After the kernel stair[] holds the row END of each column.
// ---------------------------------------------------------------------------------------------
__global__ void k_front_setup (const I32 *__restrict__ fronts, DSym S, DNum N)
{
    __shared__ I32 sh [34] ;
    const I32 slot = blockIdx.x ;
    const I32 f = fronts [slot] ;
    const I32 col1 = S.Super [f], fp = S.Super [f+1] - col1 ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    I32 *st = N.stair + p1 ;
    const int tid = threadIdx.x, nt = blockDim.x ;

    for (I32 j = tid ; j < fn ; j += nt)
        st [j] = (j < fp) ? (S.Sleft [col1+j+1] - S.Sleft [col1+j]) : 0 ;
    __syncthreads () ;
    const I32 c1 = S.Childp [f], c2 = S.Childp [f+1] ;
    for (I32 q = c1 ; q < c2 ; q++)
    {
        const I32 c = S.Child [q] ;
        const I32 pc = S.Rp [c] + (S.Super [c+1] - S.Super [c]) ;
        const I32 cm = N.Cm [c] ;
        // the cm columns of one child are distinct: no atomics needed inside a child
        for (I32 ci = tid ; ci < cm ; ci += nt) st [S.Cj [pc+ci]] += 1 ;
        __syncthreads () ;
    }
    const I32 fm = block_exclusive_scan<I32> (st, fn, sh) ;
    __syncthreads () ;

    I32 *Hi = N.Hii + S.Hip [f] ;
    // original rows of S whose leftmost column is a pivot of this front (:1188-1208)
    const I32 r1 = S.Sleft [col1], r2 = S.Sleft [col1+fp] ;
    for (I32 r = r1 + tid ; r < r2 ; r += nt)
    {
        const I32 k = S.Sj [S.Sp [r]] - col1 ;          // leftmost column of the row
        const I32 i = st [k] + (r - S.Sleft [col1+k]) ;
        N.rowpos [r] = i ;
        Hi [i] = r ;
    }
    __syncthreads () ;
    for (I32 k = tid ; k < fp ; k += nt) st [k] += S.Sleft [col1+k+1] - S.Sleft [col1+k] ;
    __syncthreads () ;
    // child rows, children in order (:1239-1248)
    for (I32 q = c1 ; q < c2 ; q++)
    {
        const I32 c = S.Child [q] ;
        const I32 pc = S.Rp [c] + (S.Super [c+1] - S.Super [c]) ;
        const I32 cm = N.Cm [c] ;
        const I32 *Hichild = N.Hii + S.Hip [c] + N.Hr [c] ;
        for (I32 ci = tid ; ci < cm ; ci += nt)
        {
            const I32 j = S.Cj [pc+ci] ;
            const I32 i = st [j] ;
            st [j] = i + 1 ;
            N.Cmap [pc+ci] = i ;
            Hi [i] = Hichild [ci] ;
        }
        __syncthreads () ;
    }
    if (tid == 0)
    {
        N.Hm [f] = fm ;
        N.rank [f] = min (fm, fp) ;
        N.g [slot] = 0 ;
        N.done [slot] = 0 ;
        atomicMax (N.maxfm, fm) ;
    }
}

// debug (STMQR_B200_CHECK=1): first front of the level whose F holds a non-finite value
__global__ void k_check_finite (const I32 *__restrict__ fronts, I32 count, DSym S, DNum N, I32 *out)
{
    const I32 f = fronts [blockIdx.x] ;
    const I32 fn = S.Rp [f+1] - S.Rp [f] ;
    const I64 fm = N.Hm [f] ;
    const double *F = N.F + S.Foff [f] ;
    bool bad = false ;
    I64 where = -1 ;
    for (I64 e = threadIdx.x ; e < fm * fn ; e += blockDim.x)
        if (!(fabs (F [e]) < 1e100)) { if (!bad) where = e ; bad = true ; }
    if (bad)
    {
        const I32 old = atomicCAS (out, -1, f) ;
        if (old == -1) { out [1] = (I32) (where % fm) ; out [2] = (I32) (where / fm) ; }
    }
}

// max actual # rows over the fronts of one level (read back by the host for the large-front levels)
__global__ void k_level_maxfm (const I32 *__restrict__ fronts, I32 count, DNum N)
{
    __shared__ I32 sh [32] ;
    I32 m = 0 ;
    for (I32 i = threadIdx.x ; i < count ; i += blockDim.x) m = max (m, N.Hm [fronts [i]]) ;
    for (int o = 16 ; o > 0 ; o >>= 1) m = max (m, __shfl_xor_sync (STMQR_FULL_MASK, m, o)) ;
    if ((threadIdx.x & 31) == 0) sh [threadIdx.x >> 5] = m ;
    __syncthreads () ;
    if (threadIdx.x < 32)
    {
        m = (threadIdx.x < (blockDim.x >> 5)) ? sh [threadIdx.x] : 0 ;
        for (int o = 16 ; o > 0 ; o >>= 1) m = max (m, __shfl_xor_sync (STMQR_FULL_MASK, m, o)) ;
        if (threadIdx.x == 0) N.lvlstat [0] = m ;
    }
}

// ---------------------------------------------------------------------------------------------
// Numeric assembly (qr_assemble, SparseQR_factorize.c:1176-1281) in two launches per level:
//   k_zero_fronts   F <- 0 for every front of the level: one contiguous 16-byte store stream per front
//                   (front offsets are 16-byte aligned), grid = (fronts, chunks)
//   k_assemble      scatter the rows of S and stack the children's packed C blocks.  In multifrontal QR
//                   every row of F has exactly one source (an original row of S or one row of one child's
//                   C block), so the writes are disjoint and need no order among themselves.  grid =
//                   (fronts, slices); the slices deal out the SOURCES (rows of S, columns of each child C
//                   block) round-robin, so nothing is read twice and nothing is read to be discarded.
// ---------------------------------------------------------------------------------------------
__global__ void k_zero_fronts (const I32 *__restrict__ fronts, DSym S, DNum N)
{
    const I32 f = fronts [blockIdx.x] ;
    const I64 fn = S.Rp [f+1] - S.Rp [f] ;
    const I64 cnt = (I64) N.Hm [f] * fn ;
    if (cnt == 0) return ;
    double *F = N.F + S.Foff [f] ;
    double2 *F2 = reinterpret_cast<double2 *> (F) ;
    const I64 n2 = cnt >> 1 ;
    const I64 stride = (I64) gridDim.y * blockDim.x ;
    const double2 z = make_double2 (0.0, 0.0) ;
    I64 i = (I64) blockIdx.y * blockDim.x + threadIdx.x ;
    for ( ; i + 3 * stride < n2 ; i += 4 * stride)
    {
        F2 [i] = z ; F2 [i + stride] = z ; F2 [i + 2 * stride] = z ; F2 [i + 3 * stride] = z ;
    }
    for ( ; i < n2 ; i += stride) F2 [i] = z ;
    if ((cnt & 1) && blockIdx.y == 0 && threadIdx.x == 0) F [cnt - 1] = 0.0 ;
}

__global__ void k_assemble (const I32 *__restrict__ fronts, DSym S, DNum N)
{
    const I32 f = fronts [blockIdx.x] ;
    const I32 col1 = S.Super [f], fp = S.Super [f+1] - col1 ;
    const I32 fm = N.Hm [f] ;
    if (fm == 0) return ;
    double *F = N.F + S.Foff [f] ;
    const int lane = threadIdx.x & 31 ;
    const I32 gw = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5) ;        // my warp among the front's warps
    const I32 ngw = gridDim.y * (blockDim.x >> 5) ;

    // rows of S: one warp per row, lanes over its entries
    const I32 r1 = S.Sleft [col1], r2 = S.Sleft [col1+fp] ;
    for (I32 r = r1 + gw ; r < r2 ; r += ngw)
    {
        const I32 i = N.rowpos [r] ;
        const I32 e2 = S.Sp [r+1] ;
        for (I32 p = S.Sp [r] + lane ; p < e2 ; p += 32) F [i + (I64) S.Sjf [p] * fm] = N.Sx [p] ;
    }

    // children: one warp per column of the child's C block (contiguous source, one column of F)
    for (I32 q = S.Childp [f] ; q < S.Childp [f+1] ; q++)
    {
        const I32 c = S.Child [q] ;
        const I32 fpc = S.Super [c+1] - S.Super [c] ;
        const I32 pc = S.Rp [c] + fpc ;
        const I32 cn = (S.Rp [c+1] - S.Rp [c]) - fpc ;
        const I32 cm = N.Cm [c] ;
        if (cm <= 0) continue ;
        const double *C = N.C + S.Coff [c] ;
        const I32 *Cmap = N.Cmap + pc ;
        // (start at a different warp for every child: short children do not all land on the first slices)
        for (I32 cj = (gw + ngw - (q % ngw)) % ngw ; cj < cn ; cj += ngw)
        {
            const I32 len = min (cj+1, cm) ;
            const double *src = C + cblock_col_offset (cj, cm) ;
            double *Fj = F + (I64) S.Cj [pc+cj] * fm ;
            I32 ci = lane ;
            for ( ; ci + 96 < len ; ci += 128)
            {
                const double v0 = src [ci], v1 = src [ci+32], v2 = src [ci+64], v3 = src [ci+96] ;
                const I32 i0 = Cmap [ci], i1 = Cmap [ci+32], i2 = Cmap [ci+64], i3 = Cmap [ci+96] ;
                Fj [i0] = v0 ; Fj [i1] = v1 ; Fj [i2] = v2 ; Fj [i3] = v3 ;
            }
            for ( ; ci < len ; ci += 32) Fj [Cmap [ci]] = src [ci] ;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// After the front QR: sizes of the packed blocks.  qr_fcsize (:1623), the column lengths of
// qr_rhpack (:1726-1780) as two block scans, Hr, Cm, HStair.  One CTA per front.
// ---------------------------------------------------------------------------------------------
__global__ void k_front_finish (const I32 *__restrict__ fronts, DSym S, DNum N)
{
    __shared__ I64 sh64 [34] ;
    __shared__ I32 sh32 [34] ;
    const I32 f = fronts [blockIdx.x] ;
    const I32 fp = S.Super [f+1] - S.Super [f] ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const I32 fm = N.Hm [f] ;
    const I32 *st = N.stair + p1 ;
    I64 *colp = N.colp + p1 ;
    const int tid = threadIdx.x, nt = blockDim.x ;

    // rm before/after each pivot column = # live pivots so far (capped at fm)
    // use colp as int64 scratch for the live flags
    for (I32 k = tid ; k < fp ; k += nt) colp [k] = (st [k] != 0) ? 1 : 0 ;
    __syncthreads () ;
    const I64 nlive = block_exclusive_scan<I64> (colp, fp, sh64) ;
    __syncthreads () ;
    const I32 rm = (I32) min ((I64) fm, nlive) ;
    // column lengths
    for (I32 k = tid ; k < fn ; k += nt)
    {
        I64 len ;
        if (k < fp)
        {
            const I32 t = st [k] ;
            const I64 before = min ((I64) fm, colp [k]) ;           // rm before this column
            len = (t == 0) ? before : (I64) t ;
        }
        else
        {
            const I32 h = min (rm + (k - fp + 1), fm) ;
            len = (I64) rm + max (0, st [k] - h) ;
        }
        colp [k] = len ;
    }
    __syncthreads () ;
    const I64 rsize = (fm > 0) ? block_exclusive_scan<I64> (colp, fn, sh64) : 0 ;
    (void) sh32 ;
    if (tid == 0)
    {
        const I32 cn = fn - fp ;
        const I32 rank = N.rank [f] ;
        I32 cm = min (fm - rank, cn) ;
        if (cm < 0 || cn <= 0) cm = 0 ;
        N.Cm [f] = cm ;
        N.Hr [f] = (fm > 0) ? rm : 0 ;
        N.rsize [f] = rsize ;
        {
            // algorithmic bytes of assembly + pack for this front
            const I32 col1 = S.Super [f] ;
            double csz_children = 0, ids = fm ;
            for (I32 q = S.Childp [f] ; q < S.Childp [f+1] ; q++)
            {
                const I32 c = S.Child [q] ;
                const double cmc = N.Cm [c] ;
                const double cnc = (S.Rp [c+1] - S.Rp [c]) - (S.Super [c+1] - S.Super [c]) ;
                csz_children += cmc * (cmc + 1) / 2 + cmc * (cnc - cmc) ;
                ids += cmc + cnc ;
            }
            const double snz = (double) (S.Sp [S.Sleft [col1+fp]] - S.Sp [S.Sleft [col1]]) ;
            const double csz = (double) cm * (cm + 1) / 2 + (double) cm * (cn - cm) ;
            const double b = 8.0 * (2.0 * (double) fm * fn + csz_children + (double) rsize + csz)
                + 16.0 * snz + 8.0 * ids ;
            atomicAdd (N.flops + 2, b) ;
        }
        atomicAdd (N.sumrank, rank) ;
        atomicMax (N.maxfrank, rank) ;
    }
}

// Deterministic bump allocation of the level's R+H blocks: exclusive scan over the level's
// fronts (in list order) on one CTA.
// cap = doubles of the R+H arena: if the level does not fit (the symbolic bound was violated -- never seen,
// the bound is the reference's own) the overflow flag N.griderr [1] is raised and k_pack writes no R+H block any
// more, so nothing is stored outside the arena and the host reports the failure.
__global__ void k_level_alloc (const I32 *__restrict__ fronts, I32 count, DNum N, I64 cap)
{
    __shared__ I64 sh [34] ;
    __shared__ I64 carry_s ;
    const int tid = threadIdx.x, nt = blockDim.x ;
    if (tid == 0) carry_s = (I64) *N.rcursor ;
    __syncthreads () ;
    const int lane = tid & 31, w = tid >> 5, nw = nt >> 5 ;
    for (I32 base = 0 ; base < count ; base += nt)
    {
        const I32 i = base + tid ;
        const I64 v = (i < count) ? N.rsize [fronts [i]] : 0 ;
        I64 inc = v ;
        for (int o = 1 ; o < 32 ; o <<= 1)
        {
            I64 u = __shfl_up_sync (STMQR_FULL_MASK, inc, o) ;
            if (lane >= o) inc += u ;
        }
        if (lane == 31) sh [w] = inc ;
        __syncthreads () ;
        if (w == 0)
        {
            I64 t = (lane < nw) ? sh [lane] : 0 ;
            I64 ti = t ;
            for (int o = 1 ; o < 32 ; o <<= 1)
            {
                I64 u = __shfl_up_sync (STMQR_FULL_MASK, ti, o) ;
                if (lane >= o) ti += u ;
            }
            sh [lane] = ti - t ;
            if (lane == 31) sh [32] = ti ;
        }
        __syncthreads () ;
        if (i < count) N.Roff [fronts [i]] = carry_s + sh [w] + inc - v ;
        __syncthreads () ;
        if (tid == 0) carry_s += sh [32] ;
        __syncthreads () ;
    }
    if (tid == 0)
    {
        *N.rcursor = (unsigned long long) carry_s ;
        if (carry_s > cap) N.griderr [1] = 1 ;
    }
}

// ---------------------------------------------------------------------------------------------
// Pack: copy the upper-trapezoidal C block (qr_cpack :1639-1685) and the R+H staircase
// (qr_rhpack :1691-1784) out of F.  grid = (fronts, column slices); reads and writes are
// contiguous per column.
// ---------------------------------------------------------------------------------------------
__global__ void k_pack (const I32 *__restrict__ fronts, DSym S, DNum N)
{
    const I32 f = fronts [blockIdx.x] ;
    const I32 fp = S.Super [f+1] - S.Super [f] ;
    const I32 p1 = S.Rp [f], fn = S.Rp [f+1] - p1 ;
    const I32 fm = N.Hm [f] ;
    if (fm == 0) return ;
    const I32 nsl = gridDim.y ;
    const I32 cb = (fn + nsl - 1) / nsl ;
    const I32 j1 = blockIdx.y * cb, j2 = min (fn, j1 + cb) ;
    if (j1 >= j2) return ;
    const double *F = N.F + S.Foff [f] ;
    const I64 *colp = N.colp + p1 ;
    double *R = N.R + N.Roff [f] ;
    double *C = N.C + S.Coff [f] ;
    const I32 rm = N.Hr [f], cm = N.Cm [f], rank = N.rank [f] ;
    const I64 rsize = N.rsize [f] ;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5 ;
    const bool keepR = (N.griderr [1] == 0) ;           // (R+H arena overflow: pack the C blocks only)

    // every column is a contiguous read and a contiguous write; four independent loads are in flight per lane
    // before the first store (source and destination never overlap)
    for (I32 k = j1 + w ; k < j2 ; k += nw)
    {
        const double *__restrict__ Fk = F + (I64) k * fm ;
        double *__restrict__ Rk = R + colp [k] ;
        const I64 len = ((k+1 < fn) ? colp [k+1] : rsize) - colp [k] ;
        if (!keepR) { if (k < fp) continue ; }
        if (k < fp)
        {
            I64 i = lane ;
            for ( ; i + 96 < len ; i += 128)
            {
                const double v0 = Fk [i], v1 = Fk [i+32], v2 = Fk [i+64], v3 = Fk [i+96] ;
                Rk [i] = v0 ; Rk [i+32] = v1 ; Rk [i+64] = v2 ; Rk [i+96] = v3 ;
            }
            for ( ; i < len ; i += 32) Rk [i] = Fk [i] ;
        }
        else
        {
            // rows 0..rm-1, then rows h..t-1 with h = min (rm + (k-fp+1), fm)
            const I32 h = min (rm + (k - fp + 1), fm) ;
            const I64 sh = (I64) h - rm ;
            I64 i = keepR ? (I64) lane : len ;
            for ( ; i + 96 < len ; i += 128)
            {
                const double v0 = Fk [i + ((i < rm) ? 0 : sh)], v1 = Fk [i + 32 + ((i + 32 < rm) ? 0 : sh)],
                    v2 = Fk [i + 64 + ((i + 64 < rm) ? 0 : sh)], v3 = Fk [i + 96 + ((i + 96 < rm) ? 0 : sh)] ;
                Rk [i] = v0 ; Rk [i+32] = v1 ; Rk [i+64] = v2 ; Rk [i+96] = v3 ;
            }
            for ( ; i < len ; i += 32) Rk [i] = Fk [i + ((i < rm) ? 0 : sh)] ;
            // contribution block column cj = k - fp: rows rank .. rank+min(cj+1,cm)-1
            const I32 cj = k - fp ;
            const I32 clen = min (cj+1, cm) ;
            double *__restrict__ Ck = C + cblock_col_offset (cj, cm) ;
            const double *__restrict__ Fr = Fk + rank ;
            I32 ci = lane ;
            for ( ; ci + 96 < clen ; ci += 128)
            {
                const double v0 = Fr [ci], v1 = Fr [ci+32], v2 = Fr [ci+64], v3 = Fr [ci+96] ;
                Ck [ci] = v0 ; Ck [ci+32] = v1 ; Ck [ci+64] = v2 ; Ck [ci+96] = v3 ;
            }
            for ( ; ci < clen ; ci += 32) Ck [ci] = Fr [ci] ;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// qr_hpinv (SparseQR_factorize.c:991-1060) in parallel: the serial row counters row1/row2
// become two exclusive scans over the fronts.
// ---------------------------------------------------------------------------------------------
__global__ void k_hpinv_counts (DSym S, DNum N)
{
    const I32 f = blockIdx.x * blockDim.x + threadIdx.x ;
    if (f >= S.nf) return ;
    const I32 fm = N.Hm [f], rm = N.Hr [f] ;
    const I32 cn = (S.Rp [f+1] - S.Rp [f]) - (S.Super [f+1] - S.Super [f]) ;
    const I32 cm = min (fm - rm, cn) ;
    N.base1 [f] = rm ;
    N.base2 [f] = max (0, fm - rm - cm) ;
}
__global__ void k_scan_i64 (I64 *a, I64 *b, I32 n)
{
    __shared__ I64 sh [34] ;
    block_exclusive_scan<I64> (a, n, sh) ;
    __syncthreads () ;
    block_exclusive_scan<I64> (b, n, sh) ;
}
__global__ void k_scan1_i64 (I64 *a, I32 n)
{
    __shared__ I64 sh [34] ;
    block_exclusive_scan<I64> (a, n, sh) ;
}
__global__ void k_hpinv_rows (DSym S, DNum N)
{
    // one warp per front
    const int lane = threadIdx.x & 31 ;
    const I32 f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    if (f >= S.nf) return ;
    if (N.owned && !N.owned [f]) return ;       // another GPU holds this front's row ids
    const I32 fm = N.Hm [f], rm = N.Hr [f] ;
    const I32 cn = (S.Rp [f+1] - S.Rp [f]) - (S.Super [f+1] - S.Super [f]) ;
    const I32 cm = min (fm - rm, cn) ;
    const I32 *Hi = N.Hii + S.Hip [f] ;
    const I64 b1 = N.base1 [f] ;
    const I64 nempty = S.m - S.Sleft [S.n] ;
    const I64 b2 = (I64) S.m - nempty - N.base2 [f] ;    // value of row2 when front f is reached
    for (I32 i = lane ; i < rm ; i += 32) N.W [Hi [i]] = (I32) (b1 + i) ;
    for (I32 i = rm + cm + lane ; i < fm ; i += 32) N.W [Hi [i]] = (I32) (b2 - (fm - i)) ;
}
__global__ void k_hpinv_empty (DSym S, DNum N)
{
    const I32 i = S.Sleft [S.n] + blockIdx.x * blockDim.x + threadIdx.x ;
    if (i < S.m) N.W [i] = S.m - 1 - (i - S.Sleft [S.n]) ;
}
__global__ void k_hpinv_apply (DSym S, DNum N, I64 *HPinv64, I64 *Hii64)
{
    const int lane = threadIdx.x & 31 ;
    const I64 gw = ((I64) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    const I64 nwarp = ((I64) gridDim.x * blockDim.x) >> 5 ;
    const I64 gt = (I64) blockIdx.x * blockDim.x + threadIdx.x ;
    const I64 ntot = (I64) gridDim.x * blockDim.x ;
    for (I64 i = gt ; i < S.m ; i += ntot) HPinv64 [i] = N.W [S.PLinv [i]] ;
    for (I64 f = gw ; f < S.nf ; f += nwarp)
    {
        if (N.owned && !N.owned [f]) continue ;
        const I32 fm = N.Hm [f] ;
        const I32 *Hi = N.Hii + S.Hip [f] ;
        I64 *Ho = Hii64 + S.Hip [f] ;
        for (I32 i = lane ; i < fm ; i += 32) Ho [i] = N.W [Hi [i]] ;
    }
}

__global__ void k_rank1 (const char *Rdead, I64 ntol, I32 *rank1)
{
    I64 i = (I64) blockIdx.x * blockDim.x + threadIdx.x ;
    int live = (i < ntol) ? (Rdead [i] == 0) : 0 ;
    live = warp_sum_i (live) ;
    if ((threadIdx.x & 31) == 0 && live) atomicAdd (rank1, live) ;
}

__global__ void k_widen (const I32 *__restrict__ a, I64 *__restrict__ b, I64 n)
{
    I64 i = (I64) blockIdx.x * blockDim.x + threadIdx.x ;
    const I64 s = (I64) gridDim.x * blockDim.x ;
    for ( ; i < n ; i += s) b [i] = a [i] ;
}

} // namespace stmqr
