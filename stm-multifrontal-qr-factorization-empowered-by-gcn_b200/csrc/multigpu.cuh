// multigpu.cuh -- the data plane of the multi-GPU numeric phase, in C++ behind the C ABI.
//
// Replaces the reference's task tree + TPSM pool (SparseQR_analyze.c:705-1161, SparseQR_multithreads.c:14-115).
// Every front has an owner GPU.  Each GPU walks the etree levels over ITS fronts; after level l, the
// contribution blocks of the level-l fronts whose parent lives on another GPU move there, together with the
// block's row ids (the child's Hii segment) and the three integers the parent's set-up reads (Cm, Hr, Hm).
// All sizes on the wire are the symbolic bounds, so no handshake and no host synchronisation is needed: the
// transfers are ordered on the engine's stream between the pack of level l and the set-up of level l+1.
// At the end the small integer side outputs are merged with element-wise max all-reduces (every entry is
// written by exactly one GPU, the others hold the neutral element) and qr_hpinv finishes on every GPU.
// A level that consists of ONE large front can be factorized cooperatively (stmqr_b200.cu: coop_plan,
// coop_peer_level, the coop_on branch of run_level): the transports that offer plain byte transfers on the
// engine stream (bcast / send / recv below; NCCL today) carry the block reflectors and the column blocks.
//
// Two transports with the same interface:
//   NcclTransport   one process per GPU (torchrun): ncclSend/ncclRecv grouped per level, ncclAllReduce.
//                   libnccl is dlopen'ed, so the single-GPU product has no NCCL dependency.
//   PeerTransport   N handles in ONE process (one host thread each; the drop-in's STMQR_B200_DEVICES, and the
//                   one-GPU tests): cudaMemcpyPeerAsync ordered by events, reductions gathered on handle 0.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace {

using namespace stmqr ;

struct XEdge { I32 c, src, dst ; } ;        // the block of front c moves from GPU src to GPU dst after level(c)

__global__ void k_xpack (const I32 *__restrict__ fronts, I32 count, DNum N, I32 *__restrict__ out)
{
    const I32 i = blockIdx.x * blockDim.x + threadIdx.x ;
    if (i >= count) return ;
    const I32 c = fronts [i] ;
    out [3*i] = N.Cm [c] ; out [3*i+1] = N.Hr [c] ; out [3*i+2] = N.Hm [c] ;
}
__global__ void k_xunpack (const I32 *__restrict__ fronts, I32 count, const I32 *__restrict__ in, DNum N)
{
    const I32 i = blockIdx.x * blockDim.x + threadIdx.x ;
    if (i >= count) return ;
    const I32 c = fronts [i] ;
    N.Cm [c] = in [3*i] ; N.Hr [c] = in [3*i+1] ; N.Hm [c] = in [3*i+2] ;
}
template <typename T, bool MAXOP>
__global__ void k_merge (T *__restrict__ a, const T *__restrict__ b, I64 n)
{
    for (I64 i = (I64) blockIdx.x * blockDim.x + threadIdx.x ; i < n ; i += (I64) gridDim.x * blockDim.x)
        a [i] = MAXOP ? ((a [i] > b [i]) ? a [i] : b [i]) : (a [i] + b [i]) ;
}

enum XType { X_I8 = 0, X_I32 = 1, X_I64 = 2, X_F64 = 3 } ;
inline size_t xsize (XType t) { return t == X_I8 ? 1 : (t == X_I32 ? 4 : 8) ; }

struct Transport
{
    virtual ~Transport () { }
    virtual int nranks () const = 0 ;
    virtual int rank () const = 0 ;
    // point-to-point traffic of one etree level: edges sorted by front id, the same list on both ends
    virtual int exchange (stmqr_handle h, const std::vector<XEdge> &edges, I32 glevel) = 0 ;
    virtual int allreduce (stmqr_handle h, void *p, I64 count, XType t, bool maxop) = 0 ;
    // cooperative fronts (the trailing update of ONE large front spread over the GPUs, see run_level): plain
    // byte transfers ordered on the engine's stream.  A transport without them never takes the cooperative path.
    virtual bool cooperative () const { return false ; }
    virtual int group_begin () { return STMQR_OK ; }
    virtual int group_end () { return STMQR_OK ; }
    virtual int bcast (stmqr_handle, void *, size_t, int) { return STMQR_ERR_INVALID ; }
    virtual int send (stmqr_handle, const void *, size_t, int) { return STMQR_ERR_INVALID ; }
    virtual int recv (stmqr_handle, void *, size_t, int) { return STMQR_ERR_INVALID ; }
    virtual std::string error () const { return err ; }
    std::string err ;
} ;

// ---------------------------------------------------------------------------------------------
// NCCL, loaded at run time
// ---------------------------------------------------------------------------------------------
struct NcclApi
{
    void *lib = nullptr ;
    ncclResult_t (*GetUniqueId) (ncclUniqueId *) = nullptr ;
    ncclResult_t (*CommInitRank) (ncclComm_t *, int, ncclUniqueId, int) = nullptr ;
    ncclResult_t (*CommDestroy) (ncclComm_t) = nullptr ;
    ncclResult_t (*Send) (const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr ;
    ncclResult_t (*Recv) (void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr ;
    ncclResult_t (*AllReduce) (const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr ;
    ncclResult_t (*Broadcast) (const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr ;
    ncclResult_t (*GroupStart) () = nullptr ;
    ncclResult_t (*GroupEnd) () = nullptr ;
    const char *(*GetErrorString) (ncclResult_t) = nullptr ;
    bool load ()
    {
        if (lib) return true ;
        const char *names [] = {"libnccl.so.2", "libnccl.so"} ;
        for (const char *nm : names) { lib = dlopen (nm, RTLD_NOW | RTLD_GLOBAL) ; if (lib) break ; }
        if (!lib) return false ;
#define NCCL_SYM(f) *(void **) (&f) = dlsym (lib, "nccl" #f) ; if (!f) { lib = nullptr ; return false ; }
        NCCL_SYM (GetUniqueId) NCCL_SYM (CommInitRank) NCCL_SYM (CommDestroy) NCCL_SYM (Send) NCCL_SYM (Recv)
        NCCL_SYM (AllReduce) NCCL_SYM (Broadcast) NCCL_SYM (GroupStart) NCCL_SYM (GroupEnd) NCCL_SYM (GetErrorString)
#undef NCCL_SYM
        return true ;
    }
} ;
NcclApi g_nccl ;

struct NcclTransport : Transport
{
    ncclComm_t comm = nullptr ;
    int nr = 1, me = 0 ;
    I32 *d_send = nullptr, *d_recv = nullptr, *d_list = nullptr ;      // scalars of the edges of one level, front lists
    I64 cap = 0 ;
    int nranks () const override { return nr ; }
    int rank () const override { return me ; }
    ~NcclTransport () override
    {
        if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy (comm) ;
        if (d_send) cudaFree (d_send) ;
    }
    bool ok (ncclResult_t r, const char *what)
    {
        if (r == ncclSuccess) return true ;
        err = std::string (what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString (r) : "NCCL error") ;
        return false ;
    }
    int ensure (I64 nedges)
    {
        if (nedges <= cap) return STMQR_OK ;
        if (d_send) cudaFree (d_send) ;
        cap = std::max<I64> (64, 2 * nedges) ;
        // [send scalars 3*cap | recv scalars 3*cap | send list cap | recv list cap]
        if (cudaMalloc ((void **) &d_send, (size_t) cap * 8 * sizeof (I32)) != cudaSuccess) { err = "cudaMalloc" ; cap = 0 ; d_send = nullptr ; return STMQR_ERR_OUT_OF_MEMORY ; }
        d_recv = d_send + 3 * cap ; d_list = d_recv + 3 * cap ;
        return STMQR_OK ;
    }
    int exchange (stmqr_handle h, const std::vector<XEdge> &edges, I32 glevel) override ;
    int allreduce (stmqr_handle h, void *p, I64 count, XType t, bool maxop) override ;
    bool cooperative () const override { return nr > 1 ; }
    int group_begin () override { return ok (g_nccl.GroupStart (), "ncclGroupStart") ? STMQR_OK : STMQR_ERR_CUDA ; }
    int group_end () override { return ok (g_nccl.GroupEnd (), "ncclGroupEnd") ? STMQR_OK : STMQR_ERR_CUDA ; }
    int bcast (stmqr_handle h, void *p, size_t bytes, int root) override ;
    int send (stmqr_handle h, const void *p, size_t bytes, int dst) override ;
    int recv (stmqr_handle h, void *p, size_t bytes, int src) override ;
} ;

// ---------------------------------------------------------------------------------------------
// peer copies between handles of one process
// ---------------------------------------------------------------------------------------------
struct PeerGroup
{
    std::vector<stmqr_handle> hs ;
    std::mutex mu ;
    std::condition_variable cv ;
    int waiting = 0 ;
    unsigned long gen = 0 ;
    int failed = 0 ;
    void barrier ()
    {
        std::unique_lock<std::mutex> lk (mu) ;
        if (failed) return ;                        // a handle gave up: nobody waits for it any more
        const unsigned long g = gen ;
        if (++waiting == (int) hs.size ()) { waiting = 0 ; gen++ ; cv.notify_all () ; }
        else cv.wait (lk, [&] { return gen != g || failed ; }) ;
    }
    void abort ()
    {
        { std::lock_guard<std::mutex> lk (mu) ; failed = 1 ; }
        cv.notify_all () ;
    }
} ;

struct PeerTransport : Transport
{
    PeerGroup *grp = nullptr ;
    int me = 0 ;
    cudaEvent_t evReady = nullptr, evDone = nullptr ;       // my stream reached the exchange / finished reading peers
    void *scratch = nullptr ; size_t scratch_bytes = 0 ;
    int nranks () const override { return (int) grp->hs.size () ; }
    int rank () const override { return me ; }
    ~PeerTransport () override
    {
        if (evReady) cudaEventDestroy (evReady) ;
        if (evDone) cudaEventDestroy (evDone) ;
        if (scratch) cudaFree (scratch) ;
    }
    int exchange (stmqr_handle h, const std::vector<XEdge> &edges, I32 glevel) override ;
    int allreduce (stmqr_handle h, void *p, I64 count, XType t, bool maxop) override ;
} ;

} // namespace
