// kernels_peak.cuh -- FP64 peak microbenchmarks (roofline denominators for the front QR).
// Register-resident loops, no memory traffic: DMMA (mma.sync f64, the only FP64 tensor-core
// path on sm_100a: tcgen05.mma has no f64 kind) in its m8n8k4 and m16n8k16 shapes, and DFMA.
#pragma once
#include "common.cuh"

namespace stmqr {

__device__ __forceinline__ void dmma884 (double &d0, double &d1, double a, double b)
{
    asm volatile ("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d" (d0), "+d" (d1) : "d" (a), "d" (b)) ;
}

__device__ __forceinline__ void dmma16816 (double (&d) [4], const double (&a) [8], const double (&b) [4])
{
    asm volatile ("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 "
        "{%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d" (d [0]), "+d" (d [1]), "+d" (d [2]), "+d" (d [3])
        : "d" (a [0]), "d" (a [1]), "d" (a [2]), "d" (a [3]), "d" (a [4]), "d" (a [5]), "d" (a [6]), "d" (a [7]),
          "d" (b [0]), "d" (b [1]), "d" (b [2]), "d" (b [3])) ;
}

__global__ void k_peak_dmma884 (double *out, int iters)
{
    double acc [16] ;
    for (int i = 0 ; i < 16 ; i++) acc [i] = 0 ;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9 ;
    for (int it = 0 ; it < iters ; it++)
    {
#pragma unroll
        for (int i = 0 ; i < 8 ; i++) dmma884 (acc [2*i], acc [2*i+1], a, b) ;
    }
    double s = 0 ;
    for (int i = 0 ; i < 16 ; i++) s += acc [i] ;
    out [(blockIdx.x * blockDim.x + threadIdx.x) % (148 * 8 * 1024)] = s ;
}

__global__ void k_peak_dmma16816 (double *out, int iters)
{
    double acc [4][4] ;
    for (int i = 0 ; i < 4 ; i++) for (int j = 0 ; j < 4 ; j++) acc [i][j] = 0 ;
    double a [8], b [4] ;
    for (int i = 0 ; i < 8 ; i++) a [i] = 1.0 + (threadIdx.x + i) * 1e-9 ;
    for (int i = 0 ; i < 4 ; i++) b [i] = 1.0 - (threadIdx.x + i) * 1e-9 ;
    for (int it = 0 ; it < iters ; it++)
    {
#pragma unroll
        for (int i = 0 ; i < 4 ; i++) dmma16816 (acc [i], a, b) ;
    }
    double s = 0 ;
    for (int i = 0 ; i < 4 ; i++) for (int j = 0 ; j < 4 ; j++) s += acc [i][j] ;
    out [(blockIdx.x * blockDim.x + threadIdx.x) % (148 * 8 * 1024)] = s ;
}

__global__ void k_peak_dfma (double *out, int iters)
{
    double acc [16] ;
    for (int i = 0 ; i < 16 ; i++) acc [i] = i ;
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9 ;
    for (int it = 0 ; it < iters ; it++)
    {
#pragma unroll
        for (int i = 0 ; i < 16 ; i++) acc [i] = fma (acc [i], a, b) ;
    }
    double s = 0 ;
    for (int i = 0 ; i < 16 ; i++) s += acc [i] ;
    out [(blockIdx.x * blockDim.x + threadIdx.x) % (148 * 8 * 1024)] = s ;
}

} // namespace stmqr
