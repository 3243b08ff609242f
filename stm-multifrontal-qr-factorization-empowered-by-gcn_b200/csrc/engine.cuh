// engine.cuh -- device-side data structures of the B200 multifrontal-QR engine.
//
// Naming follows the reference's domain (fronts, staircase, contribution blocks, stacks):
// qr_symbolic / qr_numeric of STMMQR/include/SparseQR_struct.h.
#pragma once
#include "common.cuh"

namespace stmqr {

constexpr int PANEL_MAX = 32 ;      // max Householder panel width of the tiled path

// Immutable per-analysis data (device pointers).  int32 on the device: the host side rejects
// problems whose index ranges do not fit (STMQR_ERR_TOO_LARGE).
struct DSym
{
    I32 m, n, nf ;
    const I32 *Super ;      // [nf+1]
    const I32 *Rp ;         // [nf+1]
    const I32 *Rj ;         // [rjsize]
    const I32 *Sleft ;      // [n+2]
    const I32 *Sp ;         // [m+1]
    const I32 *Sj ;         // [anz]
    const I32 *Child ;      // [nf+1]
    const I32 *Childp ;     // [nf+2]
    const I32 *Hip ;        // [nf+1]
    const I32 *PLinv ;      // [m]
    const I32 *Qinv ;       // [n]   column j of A is column Qinv[j] of S
    const I32 *Qfill ;      // [n]   column k of S is column Qfill[k] of A (identity if the analysis has none)
    const I32 *Cj ;         // [rjsize] for entry p of Rj at local position >= fp of front c:
                            //          index of that column inside the PARENT front (Fmap of
                            //          qr_fsize/qr_assemble, precomputed: it is purely symbolic)
    const I32 *Sjf ;        // [anz]  front-local column of every entry of S (Fmap[Sj[p]])
    const I64 *Foff ;       // [nf]   offset of the front's F inside the per-level scratch arena
    const I64 *Coff ;       // [nf]   offset of the front's contribution block in the C arena
} ;

// Mutable per-factorization state.
struct DNum
{
    double *Sx ;            // [anz]
    double *F ;             // front scratch arena (fronts of the level being processed)
    double *C ;             // contribution-block arena
    double *R ;             // packed R+H arena (the final "stack")
    double *HTau ;          // [rjsize]
    double *Tws ;           // [maxLevelWidth * 32*32] T of the current panel of each front
    I32 *stair ;            // [rjsize] counts -> row start -> row end (staircase) -> HStair
    I32 *Cmap ;             // [rjsize] child C row ci -> row of the parent front
    I32 *rowpos ;           // [m] row of S -> row inside its front
    I32 *Hii ;              // [hisize] S row ids per front row (before qr_hpinv)
    I32 *Hm, *Hr, *Cm ;     // [nf] actual # rows of F, of R, of C
    I32 *rank ;             // [nf]
    I64 *colp ;             // [rjsize] offset of each column inside the front's packed R+H
    I64 *rsize ;            // [nf] # doubles of the packed R+H block
    I64 *Roff ;             // [nf] its offset in R
    char *Rdead ;           // [n]
    // per level-slot state of the tiled Householder path
    I32 *g ;                // current pivot row of the front
    I32 *done ;             // front ran out of rows (qr_front early exit)
    I32 *pnl_g1, *pnl_nv, *pnl_tend ;
    I32 *pnl_cols ;         // [slot*32] live columns of the current panel
    // global scalars
    unsigned long long *rcursor ;   // bump pointer of the R arena
    I32 *sumrank, *maxfrank, *maxfm, *rank1 ;
    double *flops ;         // [0] reference flop count, [1] trailing-update flops, [2] assembly bytes
    I32 *lvlstat ;          // [4] max actual # rows over the fronts of the level (k_level_maxfm)
    unsigned long long *dbg ;   // [64] cycle counters (only written when built with -DSTMQR_PANEL_TIMING)
    I32 *W ;                // [m] row permutation workspace of qr_hpinv
    const unsigned char *owned ;    // [nf] or null: fronts factorized on this GPU (tree partitioned over GPUs)
    I64 *base1, *base2 ;    // [nf] scans used by qr_hpinv
    // exchange area of k_panel_grid (fronts too tall for a cluster of shared-memory slabs)
    double *gridrec ;       // [slots][2][148][64]
    int4 *gridll ;          // [slots][2][148][128] the records as {lo, tag, hi, tag} lines
    double *gridred ;       // [slots][2][148]
    unsigned *gridctr ;     // [slots][GRID_CTR_STRIDE] arrival counters, zero between launches
    I32 *griderr ;
    // two-level blocked path of the large fronts (kernels_wide.cuh); sized for the widest such level
    double *wVb ;           // [2][slots][ldv*128]   clean Householder vectors of the current outer block
    double *wTbt ;          // [2][slots][128*128]   T of the outer block, transposed
    I32 *wblk ;             // [2][slots][4]         g0, mr, nvtot of the outer block
    double *wWp, *wW2 ;     // [slots][nsplit][ncmax*128] partial V'C,  [slots][ncmax*128] -T' sum
    double *wWpi, *wW2i ;   // the same for the update inside the block (32 reflectors, <= 96 columns)
    double *wGp ;           // [slots][nsplit][128*128] partial Gram matrices Vb'Vb
} ;

} // namespace stmqr
