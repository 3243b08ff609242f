// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tests/micro/dmma_bench.cu -o tests/micro/dmma_bench
// measured on B200 (this round): register loop 36.98 TFLOP/s at 32 warps/SM x 16 accumulators, 30.4 at 8 warps x 1;
// 4x4 tile from shared memory, 8 warps, 1 CTA/SM: 36.5 TFLOP/s (ld 36) vs 34.8 (ld 33, bank conflicts)
// micro-benchmark (not a test): FP64 DMMA (mma.sync.m8n8k4.f64) issue behaviour on B200.
// How many warps per SM and independent accumulators per warp are needed to reach the pipe's peak,
// and what does a 4x4 register tile fed from shared memory (the inner loop of kernels_wide.cuh) reach?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma (double &d0, double &d1, double a, double b)
{
    asm volatile ("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d" (d0), "+d" (d1) : "d" (a), "d" (b)) ;
}
template <int NACC> __global__ void k_reg (double *out, int iters)
{
    double acc [NACC][2] ;
    for (int i = 0 ; i < NACC ; i++) { acc [i][0] = 0 ; acc [i][1] = 0 ; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6 ;
    for (int it = 0 ; it < iters ; it++)
    {
#pragma unroll
        for (int i = 0 ; i < NACC ; i++) dmma (acc [i][0], acc [i][1], a, b) ;
    }
    double s = 0 ;
    for (int i = 0 ; i < NACC ; i++) s += acc [i][0] + acc [i][1] ;
    out [blockIdx.x * blockDim.x + threadIdx.x] = s ;
}
// 4x4 tile: per k-step 4 A fragments + 4 B fragments from shared memory, 16 DMMAs
template <int LD> __global__ void k_tile (double *out, int iters)
{
    extern __shared__ double sm [] ;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5 ;
    const int grp = lane >> 2, tig = lane & 3 ;
    for (int i = tid ; i < 192 * LD ; i += blockDim.x) sm [i] = 1e-3 * (i % 7) ;
    __syncthreads () ;
    const double *Vs = sm, *Cs = sm + 128 * LD ;
    const int wq = w & 3, wc = (w >> 2) & 1 ;
    double acc [4][4][2] ;
    for (int i = 0 ; i < 4 ; i++) for (int j = 0 ; j < 4 ; j++) { acc [i][j][0] = 0 ; acc [i][j][1] = 0 ; }
    for (int it = 0 ; it < iters ; it++)
    {
#pragma unroll
        for (int ks = 0 ; ks < 8 ; ks++)
        {
            const int rr = ks * 4 + tig ;
            double af [4], bf [4] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++) af [mi] = Vs [(wq * 32 + mi * 8 + grp) * LD + rr] ;
#pragma unroll
            for (int ni = 0 ; ni < 4 ; ni++) bf [ni] = Cs [(wc * 32 + ni * 8 + grp) * LD + rr] ;
#pragma unroll
            for (int mi = 0 ; mi < 4 ; mi++)
#pragma unroll
                for (int ni = 0 ; ni < 4 ; ni++) dmma (acc [mi][ni][0], acc [mi][ni][1], af [mi], bf [ni]) ;
        }
    }
    double s = 0 ;
    for (int i = 0 ; i < 4 ; i++) for (int j = 0 ; j < 4 ; j++) s += acc [i][j][0] + acc [i][j][1] ;
    out [blockIdx.x * blockDim.x + threadIdx.x] = s ;
}
template <typename F> double timeit (F f)
{
    cudaEvent_t a, b ; cudaEventCreate (&a) ; cudaEventCreate (&b) ;
    f () ; cudaDeviceSynchronize () ;
    float best = 1e30f ;
    for (int r = 0 ; r < 3 ; r++)
    {
        cudaEventRecord (a) ; f () ; cudaEventRecord (b) ; cudaEventSynchronize (b) ;
        float ms ; cudaEventElapsedTime (&ms, a, b) ; if (ms < best) best = ms ;
    }
    return best * 1e-3 ;
}
int main ()
{
    double *out ; cudaMalloc (&out, 148 * 8 * 1024 * sizeof (double)) ;
    const int iters = 2048 ;
    printf ("register loop: warps/SM x independent accumulators -> TFLOP/s\n") ;
    for (int wps : {4, 8, 16, 32})
    {
        const int block = (wps >= 8) ? 256 : 128, grid = 148 * (wps * 32 / block) ;
        double t1 = timeit ([&] { k_reg<1><<<grid, block>>> (out, iters) ; }) ;
        double t2 = timeit ([&] { k_reg<2><<<grid, block>>> (out, iters) ; }) ;
        double t4 = timeit ([&] { k_reg<4><<<grid, block>>> (out, iters) ; }) ;
        double t8 = timeit ([&] { k_reg<8><<<grid, block>>> (out, iters) ; }) ;
        double t16 = timeit ([&] { k_reg<16><<<grid, block>>> (out, iters) ; }) ;
        const double fl = 148.0 * wps * iters * 512.0 ;
        printf ("  %2d warps/SM: nacc1 %.2f  nacc2 %.2f  nacc4 %.2f  nacc8 %.2f  nacc16 %.2f\n", wps,
            fl * 1 / t1 * 1e-12, fl * 2 / t2 * 1e-12, fl * 4 / t4 * 1e-12, fl * 8 / t8 * 1e-12, fl * 16 / t16 * 1e-12) ;
    }
    printf ("4x4 tile from shared memory (8 LDS.64 + 16 DMMA per k-step), 8 warps per CTA\n") ;
    cudaFuncSetAttribute (k_tile<36>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 36 * 8) ;
    cudaFuncSetAttribute (k_tile<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 36 * 8) ;
    for (int cps : {1, 2, 3})
    {
        const int grid = 148 * cps ;
        double t = timeit ([&] { k_tile<36><<<grid, 256, 192 * 36 * 8>>> (out, iters / 8) ; }) ;
        double u = timeit ([&] { k_tile<33><<<grid, 256, 192 * 36 * 8>>> (out, iters / 8) ; }) ;
        const double fl = (double) grid * 8 * (iters / 8) * 8 * 16 * 512.0 ;
        printf ("  %d CTA/SM: ld 36 %.2f TFLOP/s   ld 33 (bank conflicts) %.2f TFLOP/s\n", cps, fl / t * 1e-12, fl / u * 1e-12) ;
    }
    return 0 ;
}
