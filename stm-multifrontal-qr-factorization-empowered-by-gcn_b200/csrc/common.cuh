// common.cuh -- small device helpers shared by the kernels of the B200 engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stmqr {

typedef int32_t I32 ;
typedef int64_t I64 ;

#define STMQR_FULL_MASK 0xffffffffu

__device__ __forceinline__ double warp_sum (double v)
{
#pragma unroll
    for (int o = 16 ; o > 0 ; o >>= 1) v += __shfl_xor_sync (STMQR_FULL_MASK, v, o) ;
    return v ;
}
__device__ __forceinline__ double warp_max (double v)
{
#pragma unroll
    for (int o = 16 ; o > 0 ; o >>= 1) v = fmax (v, __shfl_xor_sync (STMQR_FULL_MASK, v, o)) ;
    return v ;
}
__device__ __forceinline__ int warp_sum_i (int v)
{
#pragma unroll
    for (int o = 16 ; o > 0 ; o >>= 1) v += __shfl_xor_sync (STMQR_FULL_MASK, v, o) ;
    return v ;
}

// Block-wide sum and max of one double each; every thread gets the result.
// sh must hold 2*32 doubles.  Contains two __syncthreads().
__device__ __forceinline__ void block_sum_max (double &s, double &mx, double *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5 ;
    s = warp_sum (s) ;
    mx = warp_max (mx) ;
    if (lane == 0) { sh [w] = s ; sh [32 + w] = mx ; }
    __syncthreads () ;
    double a = (lane < nw) ? sh [lane] : 0.0 ;
    double b = (lane < nw) ? sh [32 + lane] : 0.0 ;
    a = warp_sum (a) ;
    b = warp_max (b) ;
    __syncthreads () ;
    s = a ; mx = b ;
}

// Block-wide exclusive scan of x[0..n) in place (x in global or shared memory), returns the
// total to every thread.  sh must hold 33 ints.  T is int32 or int64.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan (T *x, int n, T *sh)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5 ;
    T carry = 0 ;
    for (int base = 0 ; base < n ; base += blockDim.x)
    {
        int i = base + threadIdx.x ;
        T v = (i < n) ? x [i] : (T) 0 ;
        T inc = v ;
#pragma unroll
        for (int o = 1 ; o < 32 ; o <<= 1)
        {
            T u = __shfl_up_sync (STMQR_FULL_MASK, inc, o) ;
            if (lane >= o) inc += u ;
        }
        if (lane == 31) sh [w] = inc ;
        __syncthreads () ;
        if (w == 0)
        {
            T t = (lane < nw) ? sh [lane] : (T) 0 ;
            T ti = t ;
#pragma unroll
            for (int o = 1 ; o < 32 ; o <<= 1)
            {
                T u = __shfl_up_sync (STMQR_FULL_MASK, ti, o) ;
                if (lane >= o) ti += u ;
            }
            sh [lane] = ti - t ;            // exclusive prefix of warp totals
            if (lane == 31) sh [32] = ti ;  // tile total
        }
        __syncthreads () ;
        if (i < n) x [i] = carry + sh [w] + inc - v ;
        carry += sh [32] ;
        __syncthreads () ;
    }
    return carry ;
}

// offset of column cj inside a packed upper-trapezoidal contribution block with cm rows
// (layout of qr_cpack, SparseQR_factorize.c:1639-1685: column cj holds min(cj+1,cm) entries)
__device__ __forceinline__ I64 cblock_col_offset (I64 cj, I64 cm)
{
    return (cj < cm) ? (cj * (cj + 1)) / 2 : (cm * (cm + 1)) / 2 + (cj - cm) * cm ;
}

} // namespace stmqr
