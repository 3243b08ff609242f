// stmqr_b200.cu -- host side of the B200 multifrontal-QR engine and its C ABI
// (include/stmqr_b200.h).  Replaces the numeric phase of the reference,
// qr_factorize / qr_kernel / qr_multithreads (STMMQR/src/qr/SparseQR_factorize.c:222-985,
// SparseQR_multithreads.c:14-115): instead of one CPU task per etree subtree on per-task
// stacks, the fronts are processed level by level (all fronts of an etree level are
// independent), every level as a handful of batched launches on one CUDA stream, with
// device-side arenas sized from the symbolic bounds.  No host synchronisation between levels:
// all data-dependent shapes (fm, Cm, rank, R+H sizes) stay on the device.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/stmqr_b200.h"
#include "kernels_assembly.cuh"
#include "kernels_panel.cuh"
#include "kernels_update.cuh"
#include "kernels_wide.cuh"
#include "kernels_small.cuh"
#include "kernels_solve.cuh"
#include "kernels_peak.cuh"

using namespace stmqr ;

namespace {

constexpr int PANEL_SLAB_MAX_DOUBLES = 24576 ;   // 192 KB of dynamic shared memory per panel CTA
constexpr int PANEL_CLUSTER_MAX = 8 ;            // portable cluster size

struct Level
{
    I32 first ;         // offset into the level-ordered front list
    I32 count ;
    I32 glevel ;        // etree level (distance from the leaves) in the whole tree
    I32 maxfn ;         // max # columns in the level
    I64 maxFelems ;     // max bound Fm*fn in the level
    I32 maxFm ;         // max bound Fm in the level
    // small-front batching path (kernels_small.cuh): the list is [big fronts | small | tiny], the
    // small ones sorted by shared-memory footprint descending
    I32 nbig ;          // fronts that take the tiled kernels (first in the list)
    I32 nsmall [3] ;    // fronts of the shared-memory classes of k_front_small (> 2048, > 1024, <= 1024 doubles)
    I32 scap [3] ;      // doubles of shared memory per front of the class ((Fm+2) * fn, max)
    I32 srows [3] ;     // max bound on the # rows in the class
    // two-level blocked path (kernels_wide.cuh): few, large fronts
    bool wide ;
    I32 ldv ;           // leading dimension of the clean V buffers
    I32 nsplit ;        // row splits of the K = 128 contraction (WIDE_RS rows each)
    I32 nsplit_in ;     // row splits of the K = 32 update inside an outer block (WIDE_RS_IN rows each)
} ;

constexpr I32 WIDE_RS = 1024, WIDE_RS_IN = 256 ;
constexpr I32 WIDE_MIN_COLS = 384, WIDE_MAX_FRONTS = 64 ;
// # rows of the tallest front from which a level takes the two-level path (tunable:
// STMQR_B200_WIDE_ROWS): buffers are planned when the BOUND reaches it, the path is taken when the
// ACTUAL # rows (known after k_front_setup) does
void plan_wide (Level &L, I32 wide_rows)
{
    L.wide = (L.maxFm >= wide_rows && L.maxfn >= WIDE_MIN_COLS && L.nbig <= WIDE_MAX_FRONTS) ;
    L.ldv = ((L.maxFm + W_RT - 1) / W_RT) * W_RT + W_RT ;
    L.nsplit = (L.ldv + WIDE_RS - 1) / WIDE_RS ;
    L.nsplit_in = (L.ldv + WIDE_RS_IN - 1) / WIDE_RS_IN ;
}

// an ordered list of etree levels over a subset of the fronts (all of them on one GPU; the owned
// subtrees or the top of the tree when the tree is partitioned over several GPUs)
struct LevelSet
{
    std::vector<Level> levels ;
    std::vector<I32> fronts ;           // level-ordered, inside a level sorted by # columns descending
    I32 *d_fronts = nullptr ;
} ;

} // namespace

namespace {

// Host threads that move the downloaded factorization from the pinned staging ring into the
// caller's (pageable, usually freshly malloc'ed) arrays: first-touch page faults and the memcpy
// are spread over several cores while the DMA of the next chunk is in flight.
class CopyPool
{
public:
    explicit CopyPool (int nthreads) : stop_ (false), pending_ (0), gen_ (0)
    {
        for (int t = 0 ; t < nthreads ; t++) workers_.emplace_back ([this, t] { run (t) ; }) ;
    }
    ~CopyPool ()
    {
        { std::lock_guard<std::mutex> lk (mu_) ; stop_ = true ; }
        cv_.notify_all () ;
        for (auto &w : workers_) w.join () ;
    }
    int size () const { return (int) workers_.size () ; }
    // dst[0..bytes) = src[0..bytes), split over the workers; returns when all slices are done
    void copy (char *dst, const char *src, size_t bytes)
    {
        const int nt = size () ;
        if (nt == 0 || bytes < (size_t) (1 << 20)) { memcpy (dst, src, bytes) ; return ; }
        {
            std::lock_guard<std::mutex> lk (mu_) ;
            dst_ = dst ; src_ = src ; bytes_ = bytes ; pending_ = nt ; gen_++ ;
        }
        cv_.notify_all () ;
        std::unique_lock<std::mutex> lk (mu_) ;
        done_.wait (lk, [this] { return pending_ == 0 ; }) ;
    }
private:
    void run (int t)
    {
        unsigned long seen = 0 ;
        for ( ; ; )
        {
            char *dst ; const char *src ; size_t bytes ;
            {
                std::unique_lock<std::mutex> lk (mu_) ;
                cv_.wait (lk, [&] { return stop_ || gen_ != seen ; }) ;
                if (stop_) return ;
                seen = gen_ ; dst = dst_ ; src = src_ ; bytes = bytes_ ;
            }
            const size_t nt = workers_.size () ;
            const size_t slice = ((bytes / nt) + 4095) & ~(size_t) 4095 ;
            const size_t lo = std::min (bytes, slice * t), hi = std::min (bytes, slice * (t + 1)) ;
            if (hi > lo) memcpy (dst + lo, src + lo, hi - lo) ;
            {
                std::lock_guard<std::mutex> lk (mu_) ;
                if (--pending_ == 0) done_.notify_all () ;
            }
        }
    }
    std::vector<std::thread> workers_ ;
    std::mutex mu_ ;
    std::condition_variable cv_, done_ ;
    bool stop_ ;
    int pending_ ;
    unsigned long gen_ ;
    char *dst_ = nullptr ; const char *src_ = nullptr ; size_t bytes_ = 0 ;
} ;

constexpr size_t D2H_CHUNK = (size_t) 32 << 20 ;     // bytes per pinned staging buffer (allocation size)
constexpr int STREAM_MAX_LEVELS = 8192 ;

} // namespace

struct stmqr_handle_s
{
    int device = 0 ;
    bool host_only = false ;                // planner handle: no device, analyze / set_partition only compute the plan
    cudaStream_t stream = nullptr ;         // main stream: set-up, assembly, panels, packing
    cudaStream_t stream2 = nullptr ;        // trailing updates (look-ahead: overlaps the next panel)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr ;
    cudaEvent_t evP [2] = {nullptr, nullptr}, evN [2] = {nullptr, nullptr}, evW = nullptr ;
    std::string err ;
    stmqr_options opt {32, 0, 0, 0} ;
    bool analyzed = false, have_matrix = false, factorized = false ;
    int debug_capture = 0 ;

    // host copy of what the host needs
    I64 m = 0, n = 0, anz = 0, nf = 0, rjsize = 0, hisize = 0, maxfn = 0 ;
    int do_rank_detection = 1 ;
    std::vector<I32> h_Super, h_Rp, h_Hip, h_FmB ;
    LevelSet ls_all, ls_sub, ls_top ;
    std::vector<I32> h_parent, h_owner, h_istop ;   // etree parent; GPU partition (set_partition)
    std::vector<I32> h_Child, h_Childp ;
    int nparts = 1, mypart = 0 ;
    // streamed download (stmqr_b200_factorize_streamed): the R+H blocks of a level go to the host while
    // the next levels are factorized
    bool streaming = false ;
    std::vector<cudaEvent_t> evLvl ;
    unsigned long long *pin_cursor = nullptr ;      // pinned [STREAM_MAX_LEVELS]: R arena cursor after each level
    int stream_levels = 0 ;
    std::mutex smu ;
    std::condition_variable scv ;
    std::deque<int> squeue ;
    bool sdone = false ;
    std::atomic<int> serr {STMQR_OK} ;
    bool stream_overflow = false ;          // a level had no pre-created event: the whole stack is copied at the end
    bool stack_streamed = false ;
    std::thread *sworker = nullptr ;
    I64 stream_cap = 0 ;
    double *stream_dst = nullptr ;
    I32 *pin_lvl = nullptr ;                // pinned: actual max # rows of the level being processed
    bool check_hit = false ;                // STMQR_B200_CHECK: a non-finite value was already reported
    std::vector<cudaEvent_t> evLevelT ;     // STMQR_B200_LEVEL_TIMES=1: one timing event per etree level (real, overlapped schedule)
    std::vector<std::string> lvlNote ;
    unsigned grid_seq = 0 ;                 // launch sequence number of k_panel_grid (tags of its exchange lines)
    int nsm = 148 ;                         // SMs of the device (k_panel_grid: one CTA per SM)
    int cluster_max = 8 ;                   // largest panel cluster (the portable size; 16 measured no gain)
    I32 cluster_rows = 160 ;                // do not split slabs below this many rows
    I32 small_cap = 4096 ;                  // shared-memory doubles up to which a front takes k_front_small (0: off)
    I32 small_cap_used = 0 ;                // the value the current plan was made with
    I32 panel128_rows = 256 ;               // levels of >= 2 fronts per SM with slabs up to this many rows: 128-thread panels
    int update_rsf_max = 8 ;                // max row split (cluster size) of the K = 32 update kernel
    I64 lookahead_elems = 1000000 ;         // levels whose largest front (bound) has at least this many entries
    I32 wide_rows = 4096 ;                  // levels whose tallest front has at least this many rows: two-level path
    I32 grid_maxg = 48 ;                    // most CTAs per front of k_panel_grid (48 measured best on 29k-row fronts: fewer records to poll)
    I32 grid_rows = 6100 ;                  // levels with taller fronts take k_panel_grid
    unsigned char *d_owned = nullptr ;
    double cur_tol = -1 ; I64 cur_ntol = 0 ;
    std::vector<I64> h_Foff, h_Coff, h_Csize ;      // h_Csize: bound on the packed size of each contribution block
    I64 Fcap = 0, Ccap = 0, Rcap = 0 ;
    I64 Ccap_all = 0 ;                              // sum of all bounds (what an arena without recycling would need)
    std::vector<I64> h_Rbound ;                     // bound on the packed R+H size of every front
    I64 Rcap_owned = 0 ;                            // sum over the fronts this GPU owns (set_ownership), + slack
    I64 *d_Coff = nullptr ;                         // device copy of h_Coff (DSym.Coff)
    I64 C_alloc = 0 ;                               // doubles currently allocated for N.C
    I32 maxLevelWidth = 0 ;

    std::vector<void *> allocs ;
    size_t device_bytes = 0 ;
    // download pipeline: device -> pinned ring (DMA) -> caller's arrays (host threads)
    char *pin [2] = {nullptr, nullptr} ;
    cudaEvent_t evPin [2] = {nullptr, nullptr} ;
    cudaStream_t streamCopy = nullptr ;
    CopyPool *pool = nullptr ;

    DSym S {} ;
    DNum N {} ;
    I32 *d_err = nullptr ;
    // matrix
    I64 *d_Ap = nullptr, *d_Ai = nullptr ;
    double *d_Ax = nullptr ;
    I64 a_ncol = 0, a_nnz_cap = 0 ;
    // values-only refactorization: entry p of A -> slot of S (filled by k_build_S), valid for the resident
    // pattern d_Ap/d_Ai; scratch for the pattern of a speculative call and its verdict
    I32 *d_slot = nullptr ;
    bool slot_valid = false ;               // d_slot matches d_Ap/d_Ai
    bool values_only = false ;              // the next factorize_begin scatters d_Ax through d_slot
    I64 *d_vAp = nullptr, *d_vAi = nullptr ;
    I32 *d_vdiff = nullptr ;
    cudaStream_t streamV = nullptr ;        // uploads + compares the caller's pattern beside the numeric phase
    bool verify_pending = false, pattern_mismatch = false ;
    int speculative = 1 ;                   // STMQR_B200_SPECULATIVE=0: always upload the pattern first
    size_t d2h_chunk = D2H_CHUNK ;          // bytes per DMA chunk of the download pipeline
    I64 n_values_only = 0, n_pattern_mismatch = 0 ;
    // int64 staging for the download
    I64 *d_HPinv64 = nullptr, *d_Hii64 = nullptr, *d_wide = nullptr ;
    // debug capture
    double *d_capA = nullptr, *d_capF = nullptr ;
    std::vector<I64> h_capOff ;

    // device Q-apply / R-solve on the resident factorization (kernels_solve.cuh)
    I32 *d_hcol = nullptr, *d_nh = nullptr ;        // Householder table, rebuilt after every factorization
    I32 *d_rlen = nullptr, *d_rcnt = nullptr, *d_roff = nullptr ;   // [rjsize] R part length / non-zeros / offset in its column
    I32 *d_posfront = nullptr ;                     // [rjsize] front of every position of Rj
    I32 *d_RjTp = nullptr, *d_RjTi = nullptr ;      // transpose of Rj: positions of every column, in front order
    I64 *d_Rcolp = nullptr ;                        // [n+1] column pointers of the extracted R
    I64 rcount_econ = -1, rcount_nnz = -1 ;         // what the resident counts were made for
    bool htable_valid = false ;
    double *d_solveZ = nullptr, *d_solveX = nullptr, *d_solveW = nullptr, *d_solveIO = nullptr ;
    I64 solveZ_cap = 0, solveX_cap = 0, solveIO_cap = 0 ;
    double ms_solve = 0 ;

    // multi-GPU (multigpu.cuh): every front has an owner, the fronts of this GPU by level, the blocks that move
    void *transport = nullptr ;                     // Transport *
    LevelSet ls_mine ;
    std::vector<I32> h_level ;                      // etree level of every front
    std::vector<std::vector<int>> xedges ;          // per etree level: indices into xall of the edges leaving that level
    std::vector<I32> xall_c, xall_src, xall_dst ;   // all transfer edges (child front, owner of child, owner of parent)
    void *xptr = nullptr ;                          // PeerTransport: the array this handle contributes to the running all-reduce
    // cooperative front (multi-GPU, one front per level): set by factorize_dist around run_level / coop_peer_level
    struct Coop
    {
        bool armed = false ;                        // the level about to run may go cooperative (every GPU agrees: static test)
        bool on = false ;                           // ... and it does (decided by the home GPU from the actual # rows)
        int home = 0 ;                              // owner of the front
        I32 f = -1, fn = 0 ;
        std::vector<int> chunk_owner ;              // owner GPU of every chunk of COOP_CHUNK columns
        double *stage = nullptr ; I64 stage_cap = 0 ;   // home: the block reflector in the compact layout that travels
        int enabled = 1 ;                           // STMQR_B200_COOP=0 switches the path off, =2 forces it on 2 GPUs too
        int min_parts = 3 ;                         // (measured: on 2 GPUs the home GPU's own share leaves nothing to gain)
        I64 levels_run = 0 ;                        // statistics: cooperative levels of the last factorization
    } coop ;

    stmqr_numeric_info info {} ;
    stmqr_stats stats {} ;
    I64 launches = 0 ;
    // optional per-launch profiling (options.profile_phases)
    std::vector<cudaEvent_t> evpool ;
    std::vector<int> evclass ;
    std::vector<long long> evtag ;        // (level << 32) | first column of the panel step
    long long curtag = 0 ;
    size_t evused = 0 ;
} ;

#include "multigpu.cuh"

namespace {

int fail (stmqr_handle h, int code, const std::string &msg)
{
    if (h) h->err = msg ;
    return code ;
}

#define CK(call) do { cudaError_t e_ = (call) ; if (e_ != cudaSuccess) { \
    return fail (h, (e_ == cudaErrorMemoryAllocation) ? STMQR_ERR_OUT_OF_MEMORY : STMQR_ERR_CUDA, \
        std::string (#call) + ": " + cudaGetErrorString (e_)) ; } } while (0)

template <typename T> int dev_alloc (stmqr_handle h, T **p, size_t count)
{
    *p = nullptr ;
    size_t bytes = std::max<size_t> (count, 1) * sizeof (T) ;
    if (h->host_only) { h->device_bytes += bytes ; return STMQR_OK ; }     // planner: only account for it
    cudaError_t e = cudaMalloc ((void **) p, bytes) ;
    if (e != cudaSuccess)
    {
        return fail (h, STMQR_ERR_OUT_OF_MEMORY, std::string ("cudaMalloc of ") +
            std::to_string (bytes) + " bytes: " + cudaGetErrorString (e)) ;
    }
    h->allocs.push_back ((void *) *p) ;
    h->device_bytes += bytes ;
    return STMQR_OK ;
}
#define ALLOC(ptr, count) do { int s_ = dev_alloc (h, &(ptr), (size_t) (count)) ; if (s_ != STMQR_OK) return s_ ; } while (0)

template <typename T> int upload (stmqr_handle h, T **dst, const std::vector<T> &src)
{
    int s = dev_alloc (h, dst, src.size ()) ;
    if (s != STMQR_OK) return s ;
    if (!src.empty () && !h->host_only)
    {
        cudaError_t e = cudaMemcpyAsync (*dst, src.data (), src.size () * sizeof (T),
            cudaMemcpyHostToDevice, h->stream) ;
        if (e != cudaSuccess) return fail (h, STMQR_ERR_CUDA, cudaGetErrorString (e)) ;
    }
    return STMQR_OK ;
}
#define UPLOAD(ptr, vec) do { int s_ = upload (h, &(ptr), vec) ; if (s_ != STMQR_OK) return s_ ; } while (0)

void free_all (stmqr_handle h)
{
    if (!h->host_only) for (void *p : h->allocs) cudaFree (p) ;
    h->allocs.clear () ;
    h->device_bytes = 0 ;
    h->analyzed = h->have_matrix = h->factorized = false ;
    h->d_Ap = h->d_Ai = nullptr ; h->d_Ax = nullptr ; h->a_ncol = h->a_nnz_cap = 0 ;
    h->d_slot = nullptr ; h->d_vAp = h->d_vAi = nullptr ; h->d_vdiff = nullptr ;
    h->slot_valid = h->values_only = false ;
    h->d_hcol = h->d_nh = nullptr ; h->htable_valid = false ;
    h->d_rlen = h->d_rcnt = h->d_roff = h->d_posfront = h->d_RjTp = h->d_RjTi = nullptr ; h->d_Rcolp = nullptr ;
    h->rcount_econ = h->rcount_nnz = -1 ;
    h->d_solveZ = h->d_solveX = h->d_solveW = h->d_solveIO = nullptr ;
    h->solveZ_cap = h->solveX_cap = h->solveIO_cap = 0 ;
}

// host threads of the planner (the plan is pure host work on the symbolic object: STMQR_B200_PLAN_THREADS, default
// min (16, cores))
inline int plan_threads ()
{
    static const int nt = [] {
        int t = (int) std::min<unsigned> (16u, std::max<unsigned> (1u, std::thread::hardware_concurrency ())) ;
        if (const char *e = getenv ("STMQR_B200_PLAN_THREADS")) t = std::max (1, std::min (64, atoi (e))) ;
        return t ;
    } () ;
    return nt ;
}

// run task (i), i = 0 .. ntasks-1, on the planner's threads (tasks are handed out by an atomic counter)
// task (i, worker) with worker < plan_threads (): per-worker scratch can be allocated once by the caller
template <typename F> void parallel_tasks (I64 ntasks, F &&task)
{
    const int nt = (int) std::min<I64> (plan_threads (), ntasks) ;
    if (nt <= 1) { for (I64 i = 0 ; i < ntasks ; i++) task (i, 0) ; return ; }
    std::atomic<I64> next {0} ;
    auto worker = [&] (int me) { for (I64 i ; (i = next.fetch_add (1)) < ntasks ; ) task (i, me) ; } ;
    std::vector<std::thread> th ;
    for (int t = 1 ; t < nt ; t++) th.emplace_back (worker, t) ;
    worker (0) ;
    for (auto &t : th) t.join () ;
}

// int64 -> int32 with an overflow check (branch-free: the loop vectorises); large arrays in parallel slices
bool narrow (const int64_t *src, I64 count, std::vector<I32> &dst)
{
    dst.resize ((size_t) std::max<I64> (count, 0)) ;
    if (count <= 0) return true ;
    const I64 slice = 1 << 18 ;
    const I64 ns = (count + slice - 1) / slice ;
    std::atomic<int> bad {0} ;
    I32 *d = dst.data () ;
    parallel_tasks (ns, [&] (I64 t, int) {
        const I64 a = t * slice, b = std::min (count, a + slice) ;
        uint64_t acc = 0 ;
        for (I64 i = a ; i < b ; i++)
        {
            const int64_t v = src [i] ;
            acc |= (uint64_t) (v + 0x80000000LL) >> 32 ;
            d [i] = (I32) v ;
        }
        if (acc) bad.store (1) ;
    }) ;
    return bad.load () == 0 ;
}

inline void prof_begin (stmqr_handle h, int cls)
{
    if (!h->opt.profile_phases) return ;
    if (h->evused + 2 > h->evpool.size ())
    {
        cudaEvent_t a, b ;
        cudaEventCreate (&a) ; cudaEventCreate (&b) ;
        h->evpool.push_back (a) ; h->evpool.push_back (b) ;
    }
    cudaEventRecord (h->evpool [h->evused], h->stream) ;
    h->evclass.push_back (cls) ;
    h->evtag.push_back (h->curtag) ;
}
inline void prof_end (stmqr_handle h)
{
    h->launches++ ;
    if (!h->opt.profile_phases) return ;
    cudaEventRecord (h->evpool [h->evused + 1], h->stream) ;
    h->evused += 2 ;
}
#define LAUNCH(cls, ...) do { prof_begin (h, cls) ; __VA_ARGS__ ; prof_end (h) ; } while (0)

inline int grid_for (I64 n, int block, int cap = 148 * 16)
{
    I64 g = (n + block - 1) / block ;
    return (int) std::max<I64> (1, std::min<I64> (g, cap)) ;
}


// dst (pageable host) <- src (device), bytes: DMA into two pinned staging buffers, host threads
// copy each landed chunk into dst while the next chunk's DMA runs.  src must be complete on
// h->stream when this is called (the caller synchronised or this is issued after an event).
int d2h_pipelined (stmqr_handle h, void *dst, const void *src, size_t bytes)
{
    if (!dst || bytes == 0) return STMQR_OK ;
    if (bytes < (size_t) (4 << 20))
    {
        CK (cudaMemcpyAsync (dst, src, bytes, cudaMemcpyDeviceToHost, h->streamCopy)) ;
        CK (cudaStreamSynchronize (h->streamCopy)) ;
        return STMQR_OK ;
    }
    // chunk size of the pipeline (<= the staging buffers): every call pays one chunk of DMA before the first host
    // copy can start and one chunk of host copy after the last DMA, so smaller chunks shorten the fill and drain
    // of the per-level slices of the streamed download (STMQR_B200_D2H_CHUNK_MB, default 32)
    const size_t CH = h->d2h_chunk ;
    const size_t n = (bytes + CH - 1) / CH ;
    auto issue = [&] (size_t c) -> cudaError_t {
        const size_t off = c * CH, len = std::min (CH, bytes - off) ;
        cudaError_t e = cudaMemcpyAsync (h->pin [c & 1], (const char *) src + off, len, cudaMemcpyDeviceToHost,
            h->streamCopy) ;
        if (e != cudaSuccess) return e ;
        return cudaEventRecord (h->evPin [c & 1], h->streamCopy) ;
    } ;
    CK (issue (0)) ;
    for (size_t c = 0 ; c < n ; c++)
    {
        if (c + 1 < n) CK (issue (c + 1)) ;
        CK (cudaEventSynchronize (h->evPin [c & 1])) ;
        const size_t off = c * CH, len = std::min (CH, bytes - off) ;
        h->pool->copy ((char *) dst + off, h->pin [c & 1], len) ;
    }
    return STMQR_OK ;
}


int ensure_copy_pipeline (stmqr_handle h)
{
    if (h->pool) return STMQR_OK ;
    int nt = (int) std::min<unsigned> (8, std::max<unsigned> (1, std::thread::hardware_concurrency () / 2)) ;
    if (const char *e = getenv ("STMQR_B200_COPY_THREADS")) nt = std::max (0, atoi (e)) ;
    if (const char *e = getenv ("STMQR_B200_D2H_CHUNK_MB")) h->d2h_chunk = std::min (D2H_CHUNK, (size_t) std::max (1, atoi (e)) << 20) ;
    h->pool = new CopyPool (nt) ;
    CK (cudaStreamCreateWithFlags (&h->streamCopy, cudaStreamNonBlocking)) ;
    for (int i = 0 ; i < 2 ; i++)
    {
        CK (cudaHostAlloc ((void **) &h->pin [i], D2H_CHUNK, cudaHostAllocDefault)) ;
        CK (cudaEventCreateWithFlags (&h->evPin [i], cudaEventDisableTiming)) ;
    }
    return STMQR_OK ;
}

// the downloader of stmqr_b200_factorize_streamed: per finished level, its slice of the R+H arena
void stream_worker (stmqr_handle h, double *dst, I64 capacity)
{
    cudaSetDevice (h->device) ;
    I64 begin = 0 ;
    for ( ; ; )
    {
        int li ;
        {
            std::unique_lock<std::mutex> lk (h->smu) ;
            h->scv.wait (lk, [&] { return !h->squeue.empty () || h->sdone ; }) ;
            if (h->squeue.empty ()) break ;
            li = h->squeue.front () ; h->squeue.pop_front () ;
        }
        if (h->serr.load () != STMQR_OK) continue ;
        // evLvl is sized in stream_begin, before this thread starts, and never grows while it runs
        if (cudaEventSynchronize (h->evLvl [li]) != cudaSuccess) { h->serr.store (STMQR_ERR_CUDA) ; continue ; }
        const I64 end = (I64) h->pin_cursor [li] ;
        if (end > capacity) { h->serr.store (STMQR_ERR_INVALID) ; continue ; }
        if (end > begin)
        {
            const int s = d2h_pipelined (h, dst + begin, h->N.R + begin, (size_t) (end - begin) * sizeof (double)) ;
            if (s != STMQR_OK) h->serr.store (s) ;
        }
        begin = end ;
    }
}

// Partition of the etree over `nparts` GPUs (host only, deterministic: every rank computes the same
// answer from the same qr_symbolic).  The heaviest subtrees are opened from the roots until no
// subtree outweighs 1/(4 nparts) of the total (the reference cuts its task tree the same way,
// big_flops = total / SPQR_grain, SparseQR_analyze.c:705-859); the opened fronts form the top of the
// tree (part 0), the remaining subtrees are dealt to the parts largest-first onto the lightest part.
int partition_fronts (I64 nf, const int64_t *Childp, const int64_t *Child, const int64_t *Super,
    const int64_t *Rp, const int64_t *Fm, int nparts, int32_t *owner, int32_t *is_top)
{
    if (nf < 0 || nparts < 1 || !owner || !is_top) return STMQR_ERR_INVALID ;
    std::vector<I64> parent ((size_t) nf, -1) ;
    for (I64 f = 0 ; f < nf ; f++)
        for (I64 q = Childp [f] ; q < Childp [f+1] ; q++)
        {
            const I64 c = Child [q] ;
            if (c < 0 || c >= nf) return STMQR_ERR_INVALID ;
            parent [c] = f ;
        }
    for (I64 f = 0 ; f < nf ; f++) { owner [f] = 0 ; is_top [f] = 0 ; }
    if (nparts == 1 || nf == 0) return STMQR_OK ;
    // topological order: children before parents
    std::vector<I64> order ; order.reserve ((size_t) nf) ;
    {
        std::vector<I64> stack ;
        for (I64 r = nf - 1 ; r >= 0 ; r--) if (parent [r] < 0) stack.push_back (r) ;
        while (!stack.empty ())
        {
            const I64 f = stack.back () ; stack.pop_back () ;
            order.push_back (f) ;                       // parents first ...
            for (I64 q = Childp [f] ; q < Childp [f+1] ; q++) stack.push_back (Child [q]) ;
        }
        std::reverse (order.begin (), order.end ()) ;   // ... reversed: children first
    }
    std::vector<double> sub ((size_t) nf, 0.0) ;
    double total = 0 ;
    for (I64 f : order)
    {
        const double fn = (double) (Rp [f+1] - Rp [f]), fm = (double) std::max<int64_t> (Fm [f], 1) ;
        (void) Super ;
        const double w = fm * fn * std::min (fm, fn) + 1.0 ;
        sub [f] += w ; total += w ;
        if (parent [f] >= 0) sub [parent [f]] += sub [f] ;
    }
    const double target = total / (4.0 * nparts) ;
    auto heavier = [&] (I64 a, I64 b) { return sub [a] != sub [b] ? sub [a] > sub [b] : a < b ; } ;
    std::vector<I64> cand ;
    for (I64 f = 0 ; f < nf ; f++) if (parent [f] < 0) cand.push_back (f) ;
    for ( ; ; )
    {
        std::sort (cand.begin (), cand.end (), heavier) ;
        size_t pick = cand.size () ;
        for (size_t i = 0 ; i < cand.size () ; i++)
            if (sub [cand [i]] > target && Childp [cand [i] + 1] > Childp [cand [i]]) { pick = i ; break ; }
        if (pick == cand.size () || (I64) cand.size () > 64 * (I64) nparts) break ;
        const I64 f = cand [pick] ;
        cand.erase (cand.begin () + (std::ptrdiff_t) pick) ;
        is_top [f] = 1 ;
        for (I64 q = Childp [f] ; q < Childp [f+1] ; q++) cand.push_back (Child [q]) ;
    }
    std::sort (cand.begin (), cand.end (), heavier) ;
    std::vector<double> load ((size_t) nparts, 0.0) ;
    std::vector<int32_t> root_owner ((size_t) nf, -1) ;
    for (I64 c : cand)
    {
        int best = 0 ;
        for (int p = 1 ; p < nparts ; p++) if (load [p] < load [best]) best = p ;
        root_owner [c] = best ; load [best] += sub [c] ;
    }
    for (auto it = order.rbegin () ; it != order.rend () ; ++it)       // parents before children
    {
        const I64 f = *it ;
        if (is_top [f]) owner [f] = 0 ;
        else if (root_owner [f] >= 0) owner [f] = root_owner [f] ;
        else owner [f] = owner [parent [f]] ;
    }
    return STMQR_OK ;
}

// -------------------------------------------------------------------------------------------------
// Contribution-block arena with recycling.  The reference keeps the C blocks on its stacks and pops
// a child's block as soon as the parent is assembled (qr_kernel, SparseQR_factorize.c:907-972).
// Here the lifetime of every block is known from the level schedule alone: block c is written by the
// pack of level(c) and last read by the assembly of level(parent(c)), which precedes the pack of that
// level on the engine's stream.  So the offsets are planned once on the host by replaying the level
// sequence through a best-fit free list (sizes are the symbolic bounds: deterministic, no device-side
// allocator, no synchronisation), and the arena is as large as the high-water mark of that replay
// instead of the sum over all fronts (lap3d 96^3: 50 GB of bounds).
// -------------------------------------------------------------------------------------------------
class ArenaReplay
{
public:
    I64 high = 0, peak = 0 ;            // current top of the arena, and its high-water mark (= the capacity needed)
    // holes of the arena, sorted by offset, never adjacent (merged on release)
    std::vector<std::pair<I64, I64>> holes ;        // (offset, size)
    // blocks released since the last merge (the assembly of one level releases all its children at once)
    std::vector<std::pair<I64, I64>> freed ;
    void release (I64 off, I64 size)
    {
        size = (size + 1) & ~(I64) 1 ;
        if (size > 0) freed.emplace_back (off, size) ;
    }
    // fold the released blocks into the hole list: one sort of the level's releases, one linear merge
    void merge_freed ()
    {
        if (freed.empty ()) return ;
        std::sort (freed.begin (), freed.end ()) ;
        std::vector<std::pair<I64, I64>> out ;
        out.reserve (holes.size () + freed.size ()) ;
        size_t a = 0, b = 0 ;
        auto push = [&] (const std::pair<I64, I64> &x) {
            if (!out.empty () && out.back ().first + out.back ().second == x.first) out.back ().second += x.second ;
            else out.push_back (x) ;
        } ;
        while (a < holes.size () || b < freed.size ())
        {
            if (b >= freed.size () || (a < holes.size () && holes [a].first < freed [b].first)) push (holes [a++]) ;
            else push (freed [b++]) ;
        }
        // a hole that touches the top gives the space back
        if (!out.empty () && out.back ().first + out.back ().second == high) { high = out.back ().first ; out.pop_back () ; }
        holes.swap (out) ;
        freed.clear () ;
    }
    // the blocks of one level: first fit decreasing.  Largest block first, every block goes into the LOWEST hole
    // that can hold it (a max-tree over the offset-sorted holes answers that in O(log holes), and the hole shrinks
    // in place), what fits nowhere grows the arena at the top.  Deterministic; O((holes + blocks) log holes) per
    // level on flat arrays (a best-fit search tree of std::map / std::set nodes cost 130 ms on 110 000 fronts).
    void alloc_level (std::vector<std::pair<I64, I32>> &want, std::vector<I64> &Coff)
    {
        merge_freed () ;
        if (want.empty ()) return ;
        std::sort (want.begin (), want.end (), [] (const std::pair<I64, I32> &x, const std::pair<I64, I32> &y) {
            return (x.first != y.first) ? x.first > y.first : x.second < y.second ; }) ;
        const size_t H = holes.size () ;
        size_t base = 1 ;
        while (base < std::max<size_t> (H, 1)) base <<= 1 ;
        tree.assign (2 * base, 0) ;
        for (size_t k = 0 ; k < H ; k++) tree [base + k] = holes [k].second ;
        for (size_t k = base - 1 ; k >= 1 ; k--) tree [k] = std::max (tree [2*k], tree [2*k+1]) ;
        for (const auto &wb : want)
        {
            const I64 size = (wb.first + 1) & ~(I64) 1 ;
            const I32 f = wb.second ;
            if (H > 0 && tree [1] >= size)
            {
                size_t k = 1 ;
                while (k < base) k = (tree [2*k] >= size) ? 2*k : 2*k + 1 ;
                auto &hole = holes [k - base] ;
                Coff [f] = hole.first ;
                hole.first += size ; hole.second -= size ;
                tree [k] = hole.second ;
                for (k >>= 1 ; k >= 1 ; k >>= 1) tree [k] = std::max (tree [2*k], tree [2*k+1]) ;
            }
            else { Coff [f] = high ; high += size ; }
        }
        peak = std::max (peak, high) ;
        // drop the holes that were used up
        size_t o = 0 ;
        for (size_t k = 0 ; k < H ; k++) if (holes [k].second > 0) holes [o++] = holes [k] ;
        holes.resize (o) ;
    }
private:
    std::vector<I64> tree ;
} ;

// Replays `phases` (level sets in processing order) and assigns Coff.  `received`: fronts whose block
// arrives from another GPU before phase `recv_before` starts (cut children of the top of the tree).
void plan_contribution_arena (const std::vector<const LevelSet *> &phases, size_t recv_before,
    const std::vector<I32> &received, const std::vector<I32> &Childp, const std::vector<I32> &Child,
    const std::vector<I64> &Csize, std::vector<I64> &Coff, I64 &cap)
{
    ArenaReplay A ;
    std::vector<unsigned char> live (Coff.size (), 0) ;
    std::fill (Coff.begin (), Coff.end (), (I64) 0) ;
    std::vector<std::pair<I64, I32>> want ;
    for (size_t ph = 0 ; ph < phases.size () ; ph++)
    {
        if (ph == recv_before)
        {
            want.clear () ;
            for (I32 c : received) if (Csize [c] > 0 && !live [c]) { want.emplace_back (Csize [c], c) ; live [c] = 1 ; }
            A.alloc_level (want, Coff) ;
        }
        const LevelSet &LS = *phases [ph] ;
        for (const Level &Lv : LS.levels)
        {
            // assembly of the level: the children's blocks are consumed ...
            for (I32 i = 0 ; i < Lv.count ; i++)
            {
                const I32 f = LS.fronts [Lv.first + i] ;
                for (I32 q = Childp [f] ; q < Childp [f+1] ; q++)
                {
                    const I32 c = Child [q] ;
                    if (live [c]) { A.release (Coff [c], Csize [c]) ; live [c] = 0 ; }
                }
            }
            // ... then the level's own blocks are packed
            want.clear () ;
            for (I32 i = 0 ; i < Lv.count ; i++)
            {
                const I32 f = LS.fronts [Lv.first + i] ;
                if (Csize [f] > 0 && !live [f]) { want.emplace_back (Csize [f], f) ; live [f] = 1 ; }
            }
            A.alloc_level (want, Coff) ;
        }
    }
    cap = A.peak + 2 ;
}

constexpr I32 SMALL_CLASS_ELEMS [2] = {2048, 1024} ;     // boundaries between the shared-memory classes

// shared-memory doubles of a front in k_front_small, or 0 if the front is not eligible
inline I64 small_footprint (I64 FmB, I64 fn, I32 small_cap)
{
    const I64 e = (FmB + 2) * fn ;
    return (small_cap > 0 && fn <= SMALL_MAX_FN && e <= small_cap) ? e : 0 ;
}

// order the fronts of one level ([big by # columns descending | small by footprint descending]) and
// fill the level's statistics
void finish_level (Level &L, std::vector<I32> &v, const std::vector<I32> &Rp, const std::vector<I32> &FmB,
    I32 small_cap, I32 wide_rows)
{
    std::stable_sort (v.begin (), v.end (), [&] (I32 a, I32 b) {
        const I64 fa = Rp [a+1] - Rp [a], fb = Rp [b+1] - Rp [b] ;
        const I64 sa = small_footprint (FmB [a], fa, small_cap), sb = small_footprint (FmB [b], fb, small_cap) ;
        if ((sa == 0) != (sb == 0)) return sa == 0 ;        // big fronts first
        if (sa == 0) return fa > fb ;                       // big: # columns descending
        return sa > sb ; }) ;                               // small: footprint descending
    L.count = (I32) v.size () ;
    L.nbig = 0 ; L.maxfn = 0 ; L.maxFelems = 0 ; L.maxFm = 0 ;
    for (int c = 0 ; c < 3 ; c++) { L.nsmall [c] = 0 ; L.scap [c] = 0 ; L.srows [c] = 0 ; }
    for (I32 f : v)
    {
        const I64 fn = Rp [f+1] - Rp [f] ;
        const I64 sf = small_footprint (FmB [f], fn, small_cap) ;
        if (sf == 0)
        {
            L.nbig++ ;
            L.maxfn = std::max<I32> (L.maxfn, (I32) fn) ;
            L.maxFelems = std::max (L.maxFelems, (I64) FmB [f] * fn) ;
            L.maxFm = std::max (L.maxFm, FmB [f]) ;
        }
        else
        {
            const int c = (sf <= SMALL_CLASS_ELEMS [1]) ? 2 : ((sf <= SMALL_CLASS_ELEMS [0]) ? 1 : 0) ;
            L.nsmall [c]++ ;
            L.scap [c] = std::max<I32> (L.scap [c], (I32) sf) ;
            L.srows [c] = std::max<I32> (L.srows [c], FmB [f] + 2) ;
        }
    }
    plan_wide (L, wide_rows) ;
}

void filter_levels (const LevelSet &all, const std::vector<unsigned char> &keep, const std::vector<I32> &Rp,
    const std::vector<I32> &FmB, I32 small_cap, I32 wide_rows, LevelSet &out)
{
    out.levels.clear () ; out.fronts.clear () ;
    for (const Level &Lv : all.levels)
    {
        std::vector<I32> v ;
        for (I32 i = 0 ; i < Lv.count ; i++)
        {
            const I32 f = all.fronts [Lv.first + i] ;
            if (keep [f]) v.push_back (f) ;
        }
        if (v.empty ()) continue ;
        Level L ; L.first = (I32) out.fronts.size () ; L.glevel = Lv.glevel ;
        finish_level (L, v, Rp, FmB, small_cap, wide_rows) ;
        for (I32 f : v) out.fronts.push_back (f) ;
        out.levels.push_back (L) ;
    }
}

// ---- speculative values-only refactorization ---------------------------------------------------------
// A caller of the reference API (qr_factorize, SparseQR.c:349,371) hands over the whole matrix every time,
// although in a refactorization loop only the values change.  When the handle still holds a pattern of the
// same shape (and the A -> S slot map the first factorization left behind), only the values are uploaded
// before the numeric phase starts; the caller's Ap/Ai travel on a second stream WHILE the fronts are being
// factorized and are compared with the resident pattern on the device.  A mismatch (new pattern with the
// same counts) costs one repeated factorization through the full path; a match saves 16 of the 24 bytes
// per entry of host-to-device traffic on the critical path.
bool can_speculate (stmqr_handle h, const stmqr_csc_view *A)
{
    return h->speculative && h->analyzed && h->slot_valid && h->d_Ax && A && A->p && A->ncol == h->a_ncol &&
        A->nrow == h->m && A->p [A->ncol] == h->anz && h->anz > 0 ;
}

int issue_pattern_check (stmqr_handle h, const stmqr_csc_view *A)
{
    if (!h->streamV) CK (cudaStreamCreateWithFlags (&h->streamV, cudaStreamNonBlocking)) ;
    CK (cudaMemsetAsync (h->d_vdiff, 0, sizeof (I32), h->streamV)) ;
    CK (cudaMemcpyAsync (h->d_vAp, A->p, (size_t) (A->ncol + 1) * sizeof (I64), cudaMemcpyHostToDevice, h->streamV)) ;
    CK (cudaMemcpyAsync (h->d_vAi, A->i, (size_t) h->anz * sizeof (I64), cudaMemcpyHostToDevice, h->streamV)) ;
    k_compare_pattern<<<grid_for (h->anz, 256), 256, 0, h->streamV>>> (A->ncol + 1, h->anz, h->d_Ap, h->d_Ai,
        h->d_vAp, h->d_vAi, h->d_vdiff) ;
    h->verify_pending = true ;
    return STMQR_OK ;
}

// waits for the comparison; -> pattern_mismatch
int finish_pattern_check (stmqr_handle h)
{
    h->pattern_mismatch = false ;
    if (!h->verify_pending) return STMQR_OK ;
    h->verify_pending = false ;
    I32 diff = 1 ;
    CK (cudaMemcpyAsync (&diff, h->d_vdiff, sizeof (I32), cudaMemcpyDeviceToHost, h->streamV)) ;
    CK (cudaStreamSynchronize (h->streamV)) ;
    h->pattern_mismatch = (diff != 0) ;
    if (h->pattern_mismatch) { h->n_pattern_mismatch++ ; h->slot_valid = false ; }
    return STMQR_OK ;
}

// ---- transports (multigpu.cuh) ---------------------------------------------------------------------
int NcclTransport::exchange (stmqr_handle h, const std::vector<XEdge> &edges, I32 glevel)
{
    (void) glevel ;
    std::vector<I32> sl, rl ;
    for (const XEdge &e : edges) { if (e.src == me) sl.push_back (e.c) ; if (e.dst == me) rl.push_back (e.c) ; }
    if (sl.empty () && rl.empty ()) return STMQR_OK ;
    cudaStream_t st = h->stream ;
    int s = ensure ((I64) std::max (sl.size (), rl.size ())) ;
    if (s != STMQR_OK) return s ;
    I32 *d_sl = d_list, *d_rl = d_list + cap ;
    // (pageable source: the copy is staged before the call returns, the vectors may die afterwards)
    if (!sl.empty ()) CK (cudaMemcpyAsync (d_sl, sl.data (), sl.size () * sizeof (I32), cudaMemcpyHostToDevice, st)) ;
    if (!rl.empty ()) CK (cudaMemcpyAsync (d_rl, rl.data (), rl.size () * sizeof (I32), cudaMemcpyHostToDevice, st)) ;
    if (!sl.empty ()) k_xpack<<<(unsigned) ((sl.size () + 127) / 128), 128, 0, st>>> (d_sl, (I32) sl.size (), h->N, d_send) ;
    if (!ok (g_nccl.GroupStart (), "ncclGroupStart")) return STMQR_ERR_CUDA ;
    size_t is = 0, ir = 0 ;
    for (const XEdge &e : edges)
    {
        const I32 c = e.c ;
        const size_t cd = (size_t) h->h_Csize [c], hi = (size_t) (h->h_Hip [c+1] - h->h_Hip [c]) ;
        if (e.src == me)
        {
            if (cd > 0 && !ok (g_nccl.Send (h->N.C + h->h_Coff [c], cd, ncclDouble, e.dst, comm, st), "ncclSend")) return STMQR_ERR_CUDA ;
            if (hi > 0 && !ok (g_nccl.Send (h->N.Hii + h->h_Hip [c], hi, ncclInt32, e.dst, comm, st), "ncclSend")) return STMQR_ERR_CUDA ;
            if (!ok (g_nccl.Send (d_send + 3 * is, 3, ncclInt32, e.dst, comm, st), "ncclSend")) return STMQR_ERR_CUDA ;
            is++ ;
        }
        else if (e.dst == me)
        {
            if (cd > 0 && !ok (g_nccl.Recv (h->N.C + h->h_Coff [c], cd, ncclDouble, e.src, comm, st), "ncclRecv")) return STMQR_ERR_CUDA ;
            if (hi > 0 && !ok (g_nccl.Recv (h->N.Hii + h->h_Hip [c], hi, ncclInt32, e.src, comm, st), "ncclRecv")) return STMQR_ERR_CUDA ;
            if (!ok (g_nccl.Recv (d_recv + 3 * ir, 3, ncclInt32, e.src, comm, st), "ncclRecv")) return STMQR_ERR_CUDA ;
            ir++ ;
        }
    }
    if (!ok (g_nccl.GroupEnd (), "ncclGroupEnd")) return STMQR_ERR_CUDA ;
    if (!rl.empty ()) k_xunpack<<<(unsigned) ((rl.size () + 127) / 128), 128, 0, st>>> (d_rl, (I32) rl.size (), d_recv, h->N) ;
    CK (cudaGetLastError ()) ;
    h->launches += 2 ;
    return STMQR_OK ;
}

int NcclTransport::allreduce (stmqr_handle h, void *p, I64 count, XType t, bool maxop)
{
    if (count <= 0) return STMQR_OK ;
    const ncclDataType_t dt = (t == X_I8) ? ncclInt8 : ((t == X_I32) ? ncclInt32 : ((t == X_I64) ? ncclInt64 : ncclDouble)) ;
    if (!ok (g_nccl.AllReduce (p, p, (size_t) count, dt, maxop ? ncclMax : ncclSum, comm, h->stream), "ncclAllReduce")) return STMQR_ERR_CUDA ;
    return STMQR_OK ;
}

int NcclTransport::bcast (stmqr_handle h, void *p, size_t bytes, int root)
{
    if (bytes == 0) return STMQR_OK ;
    return ok (g_nccl.Broadcast (p, p, bytes, ncclChar, root, comm, h->stream), "ncclBroadcast") ? STMQR_OK : STMQR_ERR_CUDA ;
}
int NcclTransport::send (stmqr_handle h, const void *p, size_t bytes, int dst)
{
    if (bytes == 0) return STMQR_OK ;
    return ok (g_nccl.Send (p, bytes, ncclChar, dst, comm, h->stream), "ncclSend") ? STMQR_OK : STMQR_ERR_CUDA ;
}
int NcclTransport::recv (stmqr_handle h, void *p, size_t bytes, int src)
{
    if (bytes == 0) return STMQR_OK ;
    return ok (g_nccl.Recv (p, bytes, ncclChar, src, comm, h->stream), "ncclRecv") ? STMQR_OK : STMQR_ERR_CUDA ;
}

int PeerTransport::exchange (stmqr_handle h, const std::vector<XEdge> &edges, I32 glevel)
{
    (void) glevel ;
    cudaStream_t st = h->stream ;
    CK (cudaEventRecord (evReady, st)) ;
    grp->barrier () ;                       // every handle has recorded where its level ends
    for (const XEdge &e : edges)
    {
        if (e.dst != me) continue ;
        stmqr_handle src = grp->hs [e.src] ;
        PeerTransport *ts = (PeerTransport *) src->transport ;
        const I32 c = e.c ;
        const size_t cd = (size_t) h->h_Csize [c] * sizeof (double), hi = (size_t) (h->h_Hip [c+1] - h->h_Hip [c]) * sizeof (I32) ;
        CK (cudaStreamWaitEvent (st, ts->evReady, 0)) ;
        if (cd > 0) CK (cudaMemcpyPeerAsync (h->N.C + h->h_Coff [c], h->device, src->N.C + src->h_Coff [c], src->device, cd, st)) ;
        if (hi > 0) CK (cudaMemcpyPeerAsync (h->N.Hii + h->h_Hip [c], h->device, src->N.Hii + src->h_Hip [c], src->device, hi, st)) ;
        CK (cudaMemcpyPeerAsync (h->N.Cm + c, h->device, src->N.Cm + c, src->device, sizeof (I32), st)) ;
        CK (cudaMemcpyPeerAsync (h->N.Hr + c, h->device, src->N.Hr + c, src->device, sizeof (I32), st)) ;
        CK (cudaMemcpyPeerAsync (h->N.Hm + c, h->device, src->N.Hm + c, src->device, sizeof (I32), st)) ;
    }
    return STMQR_OK ;
}

int PeerTransport::allreduce (stmqr_handle h, void *p, I64 count, XType t, bool maxop)
{
    if (count <= 0) return STMQR_OK ;
    cudaStream_t st = h->stream ;
    const size_t bytes = (size_t) count * xsize (t) ;
    const int n = nranks () ;
    h->xptr = p ;
    CK (cudaEventRecord (evReady, st)) ;
    grp->barrier () ;
    if (me == 0)
    {
        if (bytes > scratch_bytes)
        {
            if (scratch) cudaFree (scratch) ;
            scratch = nullptr ; scratch_bytes = 0 ;
            CK (cudaMalloc (&scratch, bytes)) ;
            scratch_bytes = bytes ;
        }
        for (int r = 1 ; r < n ; r++)
        {
            stmqr_handle src = grp->hs [r] ;
            CK (cudaStreamWaitEvent (st, ((PeerTransport *) src->transport)->evReady, 0)) ;
            CK (cudaMemcpyPeerAsync (scratch, h->device, src->xptr, src->device, bytes, st)) ;
            const int g = grid_for (count, 256) ;
            if (t == X_I8) { if (maxop) k_merge<signed char, true><<<g, 256, 0, st>>> ((signed char *) p, (const signed char *) scratch, count) ; else k_merge<signed char, false><<<g, 256, 0, st>>> ((signed char *) p, (const signed char *) scratch, count) ; }
            else if (t == X_I32) { if (maxop) k_merge<I32, true><<<g, 256, 0, st>>> ((I32 *) p, (const I32 *) scratch, count) ; else k_merge<I32, false><<<g, 256, 0, st>>> ((I32 *) p, (const I32 *) scratch, count) ; }
            else if (t == X_I64) { if (maxop) k_merge<I64, true><<<g, 256, 0, st>>> ((I64 *) p, (const I64 *) scratch, count) ; else k_merge<I64, false><<<g, 256, 0, st>>> ((I64 *) p, (const I64 *) scratch, count) ; }
            else { if (maxop) k_merge<double, true><<<g, 256, 0, st>>> ((double *) p, (const double *) scratch, count) ; else k_merge<double, false><<<g, 256, 0, st>>> ((double *) p, (const double *) scratch, count) ; }
        }
        CK (cudaEventRecord (evDone, st)) ;
    }
    grp->barrier () ;                       // handle 0 has enqueued the merge
    if (me != 0)
    {
        stmqr_handle root = grp->hs [0] ;
        CK (cudaStreamWaitEvent (st, ((PeerTransport *) root->transport)->evDone, 0)) ;
        CK (cudaMemcpyPeerAsync (p, h->device, root->xptr, root->device, bytes, st)) ;
    }
    grp->barrier () ;                       // everybody has read handle 0's pointer and event: it may start the next one
    return STMQR_OK ;
}

// Owner of every front (host only, deterministic).  The subtrees below the cut are dealt to the GPUs as
// partition_fronts does (largest first onto the lightest GPU); a front above the cut goes to the GPU of
// its heaviest child, so the largest contribution block never moves and the fronts of one upper level
// spread over the GPUs instead of queueing on GPU 0.
int map_fronts (I64 nf, const int64_t *Childp, const int64_t *Child, const int64_t *Super, const int64_t *Rp,
    const int64_t *Fm, int nparts, int32_t *owner)
{
    std::vector<int32_t> top ((size_t) std::max<I64> (nf, 1), 0) ;
    int s = partition_fronts (nf, Childp, Child, Super, Rp, Fm, nparts, owner, top.data ()) ;
    if (s != STMQR_OK || nparts == 1) return s ;
    std::vector<double> sub ((size_t) nf, 0.0) ;
    std::vector<I64> parent ((size_t) nf, -1) ;
    for (I64 f = 0 ; f < nf ; f++) for (I64 q = Childp [f] ; q < Childp [f+1] ; q++) parent [Child [q]] = f ;
    // children have smaller indices than parents in the reference's postordered front tree; fall back to
    // repeated passes otherwise
    bool monotone = true ;
    for (I64 f = 0 ; f < nf ; f++) if (parent [f] >= 0 && parent [f] <= f) monotone = false ;
    auto weight = [&] (I64 f) { const double fn = (double) (Rp [f+1] - Rp [f]), fm = (double) std::max<int64_t> (Fm [f], 1) ; return fm * fn * std::min (fm, fn) + 1.0 ; } ;
    if (!monotone) return STMQR_OK ;        // keep the top of the tree on GPU 0 (still a valid ownership)
    for (I64 f = 0 ; f < nf ; f++) { sub [f] += weight (f) ; if (parent [f] >= 0) sub [parent [f]] += sub [f] ; }
    for (I64 f = 0 ; f < nf ; f++)
    {
        if (!top [f]) continue ;
        I64 best = -1 ;
        for (I64 q = Childp [f] ; q < Childp [f+1] ; q++) { const I64 c = Child [q] ; if (best < 0 || sub [c] > sub [best]) best = c ; }
        owner [f] = (best >= 0) ? owner [best] : 0 ;     // (children come first: their owners are final)
    }
    return STMQR_OK ;
}

} // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int stmqr_b200_device_count (void)
{
    int n = 0 ;
    if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError () ; return 0 ; }
    return n ;
}

int stmqr_b200_create (int device, stmqr_handle *out)
{
    if (!out) return STMQR_ERR_INVALID ;
    *out = nullptr ;
    int n = stmqr_b200_device_count () ;
    if (n <= 0 || device < 0 || device >= n) return STMQR_ERR_NO_DEVICE ;
    cudaDeviceProp prop ;
    if (cudaGetDeviceProperties (&prop, device) != cudaSuccess) return STMQR_ERR_NO_DEVICE ;
    if (prop.major < 10) return STMQR_ERR_NO_DEVICE ;       // sm_100a code only, no fallback
    stmqr_handle h = new stmqr_handle_s ;
    h->device = device ;
    h->nsm = std::min (148, prop.multiProcessorCount) ;
    if (const char *e = getenv ("STMQR_B200_GRID_ROWS")) h->grid_rows = std::max (256, atoi (e)) ;
    if (const char *e = getenv ("STMQR_B200_GRID_MAXG")) h->grid_maxg = std::max (1, std::min (148, atoi (e))) ;
    if (const char *e = getenv ("STMQR_B200_WIDE_ROWS")) h->wide_rows = std::max (256, atoi (e)) ;
    if (const char *e = getenv ("STMQR_B200_SMALL_ELEMS")) h->small_cap = std::max (0, std::min (5600, atoi (e))) ;
    if (const char *e = getenv ("STMQR_B200_PANEL128_ROWS")) h->panel128_rows = std::max (0, atoi (e)) ;
    if (const char *e = getenv ("STMQR_B200_UPDATE_RSF")) h->update_rsf_max = std::max (1, std::min (8, atoi (e))) ;
    if (const char *e = getenv ("STMQR_B200_LOOKAHEAD_ELEMS")) h->lookahead_elems = std::max (1LL, atoll (e)) ;
    if (const char *e = getenv ("STMQR_B200_CLUSTER_MAX")) h->cluster_max = std::max (1, std::min (PANEL_XR_CTAS, atoi (e))) ;
    if (const char *e = getenv ("STMQR_B200_CLUSTER_ROWS")) h->cluster_rows = std::max (32, atoi (e)) ;
    if (const char *e = getenv ("STMQR_B200_FLAGS")) h->opt.reserved = (int32_t) strtol (e, nullptr, 0) ;
    if (const char *e = getenv ("STMQR_B200_SPECULATIVE")) h->speculative = atoi (e) ;
    if (const char *e = getenv ("STMQR_B200_COOP")) h->coop.enabled = atoi (e) ;
    int prio_lo = 0, prio_hi = 0 ;
    bool ok = cudaSetDevice (device) == cudaSuccess &&
        cudaDeviceGetStreamPriorityRange (&prio_lo, &prio_hi) == cudaSuccess &&
        // panels are the critical path: their stream outranks the trailing updates
        cudaStreamCreateWithPriority (&h->stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
        cudaStreamCreateWithPriority (&h->stream2, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
        cudaEventCreate (&h->ev0) == cudaSuccess && cudaEventCreate (&h->ev1) == cudaSuccess &&
        cudaEventCreate (&h->ev2) == cudaSuccess && cudaEventCreate (&h->ev3) == cudaSuccess &&
        cudaEventCreateWithFlags (&h->evP [0], cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags (&h->evP [1], cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags (&h->evN [0], cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags (&h->evN [1], cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags (&h->evW, cudaEventDisableTiming) == cudaSuccess &&
        cudaHostAlloc ((void **) &h->pin_lvl, 4 * sizeof (I32), cudaHostAllocDefault) == cudaSuccess &&
        cudaFuncSetAttribute (k_panel_cluster<128, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (PANEL_SLAB_MAX_DOUBLES + panel_scratch_doubles (4) + PANEL_XR_DOUBLES) * (int) sizeof (double)) == cudaSuccess &&
        cudaFuncSetAttribute (k_panel_cluster<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (PANEL_SLAB_MAX_DOUBLES + panel_scratch_doubles (8) + PANEL_XR_DOUBLES) * (int) sizeof (double)) == cudaSuccess &&
        cudaFuncSetAttribute (k_panel_cluster<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (PANEL_SLAB_MAX_DOUBLES + panel_scratch_doubles (16)) * (int) sizeof (double)) == cudaSuccess &&
        cudaFuncSetAttribute (k_panel_cluster<128, 6>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute (k_panel_cluster<256, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute (k_panel_cluster<512, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute (k_panel_grid, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (PANEL_SLAB_MAX_DOUBLES + panel_scratch_doubles (16)) * (int) sizeof (double)) == cudaSuccess &&
        cudaFuncSetAttribute (k_update_dmma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) update_smem_bytes<2> ()) == cudaSuccess &&
        cudaFuncSetAttribute (k_update_dmma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) update_smem_bytes<4> ()) == cudaSuccess &&
        cudaFuncSetAttribute (k_wide_vtc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) wide_vtc_smem_bytes<1> ()) == cudaSuccess &&
        cudaFuncSetAttribute (k_wide_vtc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) wide_vtc_smem_bytes<4> ()) == cudaSuccess &&
        cudaFuncSetAttribute (k_wide_apply<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) wide_apply_smem_bytes<1> ()) == cudaSuccess &&
        cudaFuncSetAttribute (k_wide_apply<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) wide_apply_smem_bytes<4> ()) == cudaSuccess &&
        cudaFuncSetAttribute (k_wide_apply_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) wide_apply_rows_smem_bytes ()) == cudaSuccess &&
        cudaFuncSetAttribute (k_wide_tmerge, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int) wide_tmerge_smem_bytes ()) == cudaSuccess ;
    if (!ok)
    {
        cudaGetLastError () ;
        stmqr_b200_destroy (h) ;
        return STMQR_ERR_CUDA ;
    }
    *out = h ;
    return STMQR_OK ;
}

void stmqr_b200_destroy (stmqr_handle h)
{
    if (!h) return ;
    if (h->host_only) { delete h ; return ; }
    cudaSetDevice (h->device) ;
    if (h->transport) { delete (Transport *) h->transport ; h->transport = nullptr ; }
    free_all (h) ;
    if (h->coop.stage) { cudaFree (h->coop.stage) ; h->coop.stage = nullptr ; h->coop.stage_cap = 0 ; }
    if (h->ev0) cudaEventDestroy (h->ev0) ;
    if (h->ev1) cudaEventDestroy (h->ev1) ;
    if (h->ev2) cudaEventDestroy (h->ev2) ;
    if (h->ev3) cudaEventDestroy (h->ev3) ;
    for (cudaEvent_t e : h->evpool) cudaEventDestroy (e) ;
    for (int i = 0 ; i < 2 ; i++)
    {
        if (h->evP [i]) cudaEventDestroy (h->evP [i]) ;
        if (h->evN [i]) cudaEventDestroy (h->evN [i]) ;
    }
    if (h->evW) cudaEventDestroy (h->evW) ;
    if (h->pin_lvl) cudaFreeHost (h->pin_lvl) ;
    if (h->pin_cursor) cudaFreeHost (h->pin_cursor) ;
    for (cudaEvent_t e : h->evLvl) cudaEventDestroy (e) ;
    delete h->pool ;
    for (int i = 0 ; i < 2 ; i++)
    {
        if (h->pin [i]) cudaFreeHost (h->pin [i]) ;
        if (h->evPin [i]) cudaEventDestroy (h->evPin [i]) ;
    }
    if (h->streamCopy) cudaStreamDestroy (h->streamCopy) ;
    if (h->streamV) cudaStreamDestroy (h->streamV) ;
    if (h->stream2) cudaStreamDestroy (h->stream2) ;
    if (h->stream) cudaStreamDestroy (h->stream) ;
    delete h ;
}

int stmqr_b200_set_options (stmqr_handle h, const stmqr_options *opt)
{
    if (!h || !opt) return STMQR_ERR_INVALID ;
    h->opt = *opt ;
    if (const char *e = getenv ("STMQR_B200_FLAGS")) h->opt.reserved |= (int32_t) strtol (e, nullptr, 0) ;
    if (h->opt.panel <= 0 || h->opt.panel > PANEL_MAX) h->opt.panel = PANEL_MAX ;
    return STMQR_OK ;
}

const char *stmqr_b200_last_error (stmqr_handle h) { return h ? h->err.c_str () : "null handle" ; }

int stmqr_b200_set_debug_capture (stmqr_handle h, int on)
{
    if (!h) return STMQR_ERR_INVALID ;
    h->debug_capture = on ;
    return STMQR_OK ;
}

// -------------------------------------------------------------------------------------------------
// analyze: the plan.  Everything here depends only on qr_symbolic.
// -------------------------------------------------------------------------------------------------
int stmqr_b200_analyze (stmqr_handle h, const stmqr_symbolic_view *sym)
{
    if (!h || !sym) return STMQR_ERR_INVALID ;
    auto t0 = std::chrono::steady_clock::now () ;
    static const bool plan_times = getenv ("STMQR_B200_PLAN_TIMES") != nullptr ;
    auto tmark = t0 ;
    auto PLAN_MARK = [&] (const char *what) {
        if (!plan_times) return ;
        auto t = std::chrono::steady_clock::now () ;
        fprintf (stderr, "STMQR_B200_PLAN_TIMES %8.2f ms  %s\n", std::chrono::duration<double, std::milli> (t - tmark).count (), what) ;
        tmark = t ;
    } ;
    if (!h->host_only) cudaSetDevice (h->device) ;
    free_all (h) ;
    PLAN_MARK ("free_all") ;
    h->err.clear () ;
    const I64 m = sym->m, n = sym->n, nf = sym->nf, anz = sym->anz, rjsize = sym->rjsize,
        hisize = sym->hisize ;
    if (m < 0 || n < 0 || nf < 0 || !sym->Sp || !sym->Sleft || !sym->Super || !sym->Rp ||
        !sym->Childp || !sym->Child || !sym->Hip || !sym->Fm || !sym->Cm || !sym->PLinv ||
        (anz > 0 && !sym->Sj) || (rjsize > 0 && !sym->Rj))
        return fail (h, STMQR_ERR_INVALID, "analyze: missing symbolic arrays") ;
    if (m > INT32_MAX - 2 || n > INT32_MAX - 2 || anz > INT32_MAX - 2 || rjsize > INT32_MAX - 2 ||
        hisize > INT32_MAX - 2)
        return fail (h, STMQR_ERR_TOO_LARGE, "analyze: index range exceeds int32 device indices") ;
    h->m = m ; h->n = n ; h->nf = nf ; h->anz = anz ; h->rjsize = rjsize ; h->hisize = hisize ;
    h->maxfn = sym->maxfn ;
    h->do_rank_detection = (int) sym->do_rank_detection ;
    // the engine always packs R+H (qr_rhpack's keepH branch, :1726-1780): an R-only numeric object
    // would be laid out differently, so refuse it instead of returning the wrong layout.  The
    // reference always analyses with keepH = TRUE (SparseQR_analyze.c:205).
    if (!sym->keepH) return fail (h, STMQR_ERR_INVALID, "analyze: keepH = 0 (R-only packing) is not supported") ;

    std::vector<I32> Super, Rp, Rj, Sleft, Sp, Sj, Child, Childp, Hip, PLinv, FmB, CmB ;
    bool ok = narrow (sym->Super, nf+1, Super) && narrow (sym->Rp, nf+1, Rp) &&
        narrow (sym->Rj, rjsize, Rj) && narrow (sym->Sleft, n+2, Sleft) &&
        narrow (sym->Sp, m+1, Sp) && narrow (sym->Sj, anz, Sj) &&
        narrow (sym->Child, nf+1, Child) && narrow (sym->Childp, nf+2, Childp) &&
        narrow (sym->Hip, nf+1, Hip) && narrow (sym->PLinv, m, PLinv) &&
        narrow (sym->Fm, nf, FmB) && narrow (sym->Cm, nf, CmB) ;
    if (!ok) return fail (h, STMQR_ERR_TOO_LARGE, "analyze: symbolic value exceeds int32") ;
    std::vector<I32> Qinv ((size_t) n), Qfill32 ((size_t) n) ;
    for (I64 k = 0 ; k < n ; k++)
    {
        I64 j = sym->Qfill ? sym->Qfill [k] : k ;
        if (j < 0 || j >= n) return fail (h, STMQR_ERR_INVALID, "analyze: Qfill is not a permutation") ;
        Qinv [(size_t) j] = (I32) k ;
        Qfill32 [(size_t) k] = (I32) j ;
    }

    PLAN_MARK ("narrow + Qinv") ;
    // ---- parent of each front, etree levels (all fronts of a level are independent) --------------
    std::vector<I32> parent ((size_t) nf, -1), level ((size_t) nf, 0) ;
    for (I64 f = 0 ; f < nf ; f++)
        for (I32 q = Childp [f] ; q < Childp [f+1] ; q++)
        {
            I32 c = Child [q] ;
            if (c < 0 || c >= nf) return fail (h, STMQR_ERR_INVALID, "analyze: bad Child") ;
            parent [c] = (I32) f ;
        }
    {
        // children have smaller postorder index than parents; Post may be absent -> iterate to fixpoint
        // using the fact that in SPQR's supernodal tree parent index > child index.
        bool monotone = true ;
        for (I64 f = 0 ; f < nf ; f++) if (parent [f] >= 0 && parent [f] <= f) monotone = false ;
        if (monotone)
        {
            for (I64 f = 0 ; f < nf ; f++)
                if (parent [f] >= 0) level [parent [f]] = std::max (level [parent [f]], level [f] + 1) ;
        }
        else
        {
            if (!sym->Post) return fail (h, STMQR_ERR_INVALID, "analyze: Post required") ;
            for (I64 k = 0 ; k < nf ; k++)
            {
                I64 f = sym->Post [k] ;
                if (parent [f] >= 0) level [parent [f]] = std::max (level [parent [f]], level [f] + 1) ;
            }
        }
    }
    PLAN_MARK ("parent + levels") ;
    h->h_Foff.assign ((size_t) nf, 0) ; h->h_Coff.assign ((size_t) nf, 0) ; h->h_Csize.assign ((size_t) nf, 0) ;
    h->h_Rbound.assign ((size_t) nf, 0) ; h->Rcap_owned = 0 ;
    // ---- one pass over the fronts (in parallel slices, every slice with its own column map):
    //   * contribution-block bound: qr_analyze's Cm[f] rows by cn columns (SparseQR_analyze.c:536-550)
    //   * R+H bound per front: sum_j min (max (j+1, Stair_j), fm) with the bound staircase (:559-573)
    //   * the symbolic maps Cj (child column -> column of the parent front) and Sjf (S entry -> front column)
    std::vector<I32> Cj ((size_t) std::max<I64> (rjsize, 1), 0), Sjf ((size_t) std::max<I64> (anz, 1), 0) ;
    {
        const I64 nslice = std::max<I64> (1, std::min<I64> (nf, 8 * (I64) plan_threads ())) ;
        std::vector<I64> cut ((size_t) nslice + 1, nf) ;
        cut [0] = 0 ;
        {
            // slices of about equal work: by position in Rj (columns of the fronts) plus the front count
            const double total = (double) rjsize + (double) nf ;
            I64 f = 0 ;
            for (I64 t = 1 ; t < nslice ; t++)
            {
                const double goal = total * (double) t / (double) nslice ;
                while (f < nf && (double) Rp [f] + (double) f < goal) f++ ;
                cut [(size_t) t] = f ;
            }
        }
        const I64 maxfn1 = std::max<I64> (sym->maxfn, 1) ;
        const bool rbound_full = getenv ("STMQR_B200_RBOUND_FULL") != nullptr ;    // (A/B: the staircase bound without the C block taken off)
        // per-worker scratch, allocated without initialisation (every entry is written before it is read)
        const int nwk = plan_threads () ;
        std::vector<std::unique_ptr<I32 []>> FmapW ((size_t) nwk), stairW ((size_t) nwk) ;
        parallel_tasks (nslice, [&] (I64 t, int me) {
            if (!FmapW [(size_t) me])
            {
                FmapW [(size_t) me].reset (new I32 [(size_t) std::max<I64> (n, 1)]) ;
                stairW [(size_t) me].reset (new I32 [(size_t) maxfn1]) ;
            }
            I32 *const Fmap = FmapW [(size_t) me].get (), *const stairB = stairW [(size_t) me].get () ;
            for (I64 f = cut [(size_t) t] ; f < cut [(size_t) t + 1] ; f++)
            {
                const I64 p1 = Rp [f], col1 = Super [f] ;
                const I64 fp = Super [f+1] - col1, fn = Rp [f+1] - p1 ;
                const I64 cn = fn - fp, cm = std::min<I64> (CmB [f], cn) ;
                h->h_Csize [f] = (cm * (cm + 1)) / 2 + cm * (cn - cm) ;
                for (I64 j = 0 ; j < fn ; j++) Fmap [Rj [p1 + j]] = (I32) j ;
                for (I64 j = 0 ; j < fn ; j++)
                    stairB [j] = (j < fp) ? (Sleft [col1+j+1] - Sleft [col1+j]) : 0 ;
                for (I32 r = Sleft [col1] ; r < Sleft [col1+fp] ; r++)
                    for (I32 pz = Sp [r] ; pz < Sp [r+1] ; pz++) Sjf [pz] = Fmap [Sj [pz]] ;
                for (I32 q = Childp [f] ; q < Childp [f+1] ; q++)
                {
                    const I32 c = Child [q] ;
                    const I64 fpc = Super [c+1] - Super [c] ;
                    const I64 cnc = (Rp [c+1] - Rp [c]) - fpc ;
                    const I64 cmc = std::min<I64> (CmB [c], cnc) ;
                    const I64 pc = Rp [c] + fpc ;
                    for (I64 ci = 0 ; ci < cnc ; ci++)
                    {
                        const I32 j = Fmap [Rj [pc + ci]] ;
                        Cj [pc + ci] = j ;
                        if (ci < cmc) stairB [j]++ ;
                    }
                }
                I64 fm = 0, rh = 0, run = 0 ;
                for (I64 j = 0 ; j < fn ; j++) fm += stairB [j] ;
                for (I64 j = 0 ; j < fn ; j++)
                {
                    run += stairB [j] ;
                    rh += std::min<I64> (std::max<I64> (j + 1, run), fm) ;
                }
                if (fm > FmB [f]) FmB [f] = (I32) fm ;     // never trust a smaller bound
                if (!rbound_full)
                {
                    // the C block is not part of R+H: the reference's own bound takes off the smallest block the
                    // front can leave (rhsize -= csize_min, :559-573) -- its stacks rely on exactly this figure
                    const I64 rm = std::min<I64> (fm, fp) ;
                    const I64 cmn = std::min<I64> (std::max<I64> (fm - rm, 0), cn) ;
                    rh -= cmn * (cmn + 1) / 2 + cmn * (cn - cmn) ;
                }
                h->h_Rbound [(size_t) f] = rh ;
            }
        }) ;
        I64 coff = 0, rcap = 0 ;
        for (I64 f = 0 ; f < nf ; f++) { coff += (h->h_Csize [f] + 1) & ~(I64) 1 ; rcap += h->h_Rbound [(size_t) f] ; }
        h->Ccap_all = coff ;
        h->Rcap = rcap + 16 ;
    }

    PLAN_MARK ("bounds + Cj + Sjf maps") ;
    I32 nlev = 0 ;
    for (I64 f = 0 ; f < nf ; f++) nlev = std::max (nlev, level [f] + 1) ;
    std::vector<std::vector<I32>> byLevel ((size_t) nlev) ;
    for (I64 f = 0 ; f < nf ; f++) byLevel [level [f]].push_back ((I32) f) ;
    h->ls_all = LevelSet () ; h->ls_sub = LevelSet () ; h->ls_top = LevelSet () ;
    h->h_parent = parent ; h->nparts = 1 ; h->mypart = 0 ; h->h_owner.clear () ; h->h_istop.clear () ;
    h->h_level = level ; h->ls_mine = LevelSet () ; h->xedges.clear () ; h->xall_c.clear () ; h->xall_src.clear () ; h->xall_dst.clear () ;
    h->Fcap = 0 ; h->maxLevelWidth = 0 ;
    // the fused shared-memory path of the small fronts (off under debug capture: the assembled F of
    // every front is tapped from the front arena)
    h->small_cap_used = h->debug_capture ? 0 : h->small_cap ;
    if (h->opt.small_elems > 0) h->small_cap_used = h->debug_capture ? 0 : std::min<I32> (h->opt.small_elems, 5600) ;
    if (h->opt.reserved & 8) h->small_cap_used = 0 ;
    I64 nsmall_tot = 0 ;
    for (I32 l = 0 ; l < nlev ; l++)
    {
        auto &v = byLevel [l] ;
        Level L ; L.first = (I32) h->ls_all.fronts.size () ; L.glevel = l ;
        finish_level (L, v, Rp, FmB, h->small_cap_used, h->wide_rows) ;
        I64 off = 0 ;
        for (I32 f : v)
        {
            const I64 fn = Rp [f+1] - Rp [f] ;
            const I64 fe = (I64) FmB [f] * fn ;
            h->h_Foff [f] = off ;
            off += (fe + 1) & ~(I64) 1 ;            // keep every front 16-byte aligned
            h->ls_all.fronts.push_back (f) ;
        }
        h->Fcap = std::max (h->Fcap, off) ;
        h->maxLevelWidth = std::max (h->maxLevelWidth, std::max<I32> (L.nbig, 1)) ;
        nsmall_tot += L.nsmall [0] + L.nsmall [1] + L.nsmall [2] ;
        h->ls_all.levels.push_back (L) ;
    }

    PLAN_MARK ("level sets") ;
    // ---- device memory -----------------------------------------------------------------------------
    h->h_Super = Super ; h->h_Rp = Rp ; h->h_Hip = Hip ; h->h_FmB = FmB ;
    DSym &S = h->S ;
    S.m = (I32) m ; S.n = (I32) n ; S.nf = (I32) nf ;
    I32 *p ;
    UPLOAD (p, Super) ; S.Super = p ;   UPLOAD (p, Rp) ; S.Rp = p ;       UPLOAD (p, Rj) ; S.Rj = p ;
    UPLOAD (p, Sleft) ; S.Sleft = p ;   UPLOAD (p, Sp) ; S.Sp = p ;       UPLOAD (p, Sj) ; S.Sj = p ;
    UPLOAD (p, Child) ; S.Child = p ;   UPLOAD (p, Childp) ; S.Childp = p ; UPLOAD (p, Hip) ; S.Hip = p ;
    UPLOAD (p, PLinv) ; S.PLinv = p ;   UPLOAD (p, Qinv) ; S.Qinv = p ;     UPLOAD (p, Qfill32) ; S.Qfill = p ;
    UPLOAD (p, Cj) ; S.Cj = p ;         UPLOAD (p, Sjf) ; S.Sjf = p ;
    I64 *p64 ;
    UPLOAD (p64, h->h_Foff) ; S.Foff = p64 ;
    h->h_Child = Child ; h->h_Childp = Childp ;
    {
        // contribution-block offsets: replay of the level schedule (one GPU: every level of the tree)
        std::vector<const LevelSet *> phases {&h->ls_all} ;
        plan_contribution_arena (phases, phases.size (), std::vector<I32> (), Childp, Child, h->h_Csize, h->h_Coff, h->Ccap) ;
        if (h->opt.reserved & 64)
        {
            // A/B switch: no recycling, every block has its own place (round-1 layout)
            I64 coff = 0 ;
            for (I64 f = 0 ; f < nf ; f++) { h->h_Coff [f] = coff ; coff += (h->h_Csize [f] + 1) & ~(I64) 1 ; }
            h->Ccap = coff + 2 ;
        }
    }
    PLAN_MARK ("uploads + contribution arena plan") ;
    UPLOAD (p64, h->h_Coff) ; S.Coff = p64 ; h->d_Coff = p64 ;
    UPLOAD (h->ls_all.d_fronts, h->ls_all.fronts) ;
    h->d_owned = nullptr ; h->N.owned = nullptr ;

    DNum &N = h->N ;
    ALLOC (N.Sx, anz) ;
    ALLOC (N.F, h->Fcap) ;
    ALLOC (N.C, h->Ccap) ; h->C_alloc = h->Ccap ;
    ALLOC (N.R, h->Rcap) ;
    ALLOC (N.HTau, rjsize) ;
    // per-slot panel outputs: 2 parities (look-ahead), 4 on the levels that take the two-level path
    I64 pslots = 2 * (I64) h->maxLevelWidth ;
    {
        I64 vb = 0, tb = 0, wp = 0, w2 = 0, wpi = 0, w2i = 0, gp = 0, nb = 0 ;
        for (const Level &Lv : h->ls_all.levels)
        {
            if (!Lv.wide) continue ;
            const I64 c = Lv.nbig ;
            pslots = std::max (pslots, (I64) WB_PANELS * c) ;
            vb = std::max (vb, 2 * c * (I64) Lv.ldv * WB) ;
            tb = std::max (tb, 2 * c * (I64) (WB * WB)) ;
            nb = std::max (nb, 2 * c * 4) ;
            wp = std::max (wp, c * Lv.nsplit * (I64) Lv.maxfn * WB) ;
            w2 = std::max (w2, c * (I64) Lv.maxfn * WB) ;
            wpi = std::max (wpi, c * Lv.nsplit_in * (I64) (WB * PANEL_MAX)) ;
            w2i = std::max (w2i, c * (I64) (WB * PANEL_MAX)) ;
            gp = std::max (gp, c * Lv.nsplit_in * (I64) ((WB_PANELS - 1) * 96 * PANEL_MAX)) ;
        }
        ALLOC (N.wVb, vb) ; ALLOC (N.wTbt, tb) ; ALLOC (N.wblk, nb) ;
        ALLOC (N.wWp, wp) ; ALLOC (N.wW2, w2) ; ALLOC (N.wWpi, wpi) ; ALLOC (N.wW2i, w2i) ; ALLOC (N.wGp, gp) ;
    }
    ALLOC (N.Tws, pslots * PANEL_MAX * PANEL_MAX) ;
    {
        I64 gslots = 1 ;
        for (const Level &Lv : h->ls_all.levels) if (Lv.maxFm >= h->grid_rows) gslots = std::max<I64> (gslots, Lv.nbig) ;
        ALLOC (N.gridrec, gslots * 2 * 148 * 64) ;
        ALLOC (N.gridred, gslots * 2 * 148) ;
        ALLOC (N.gridll, gslots * 2 * 148 * 128) ;
        ALLOC (N.gridctr, gslots * GRID_CTR_STRIDE) ;
        ALLOC (N.griderr, 2) ;              // [0] a grid panel did not fit its slabs, [1] the R+H arena overflowed
        if (!h->host_only)
        {
            CK (cudaMemsetAsync (N.gridll, 0, gslots * 2 * 148 * 128 * sizeof (int4), h->stream)) ;
            CK (cudaMemsetAsync (N.gridctr, 0, gslots * GRID_CTR_STRIDE * sizeof (unsigned), h->stream)) ;
            CK (cudaMemsetAsync (N.griderr, 0, 2 * sizeof (I32), h->stream)) ;
        }
    }
    PLAN_MARK ("arena allocations") ;
    ALLOC (N.stair, rjsize) ;
    ALLOC (N.Cmap, rjsize) ;
    ALLOC (N.rowpos, m) ;
    ALLOC (N.Hii, hisize) ;
    ALLOC (N.Hm, 3 * std::max<I64> (nf, 1)) ; N.Hr = N.Hm + std::max<I64> (nf, 1) ; N.Cm = N.Hr + std::max<I64> (nf, 1) ;   // one block: merged over the GPUs by ONE all-reduce
    ALLOC (N.rank, nf) ;
    ALLOC (N.colp, rjsize) ;
    ALLOC (N.rsize, nf) ; ALLOC (N.Roff, nf) ;
    ALLOC (N.Rdead, n) ;
    ALLOC (N.g, h->maxLevelWidth) ; ALLOC (N.done, h->maxLevelWidth) ;
    // panel outputs are double buffered (parity of the panel step): the trailing update of step j
    // reads buffer j&1 while the panel of step j+1 fills the other one
    ALLOC (N.pnl_g1, pslots) ; ALLOC (N.pnl_nv, pslots) ;
    ALLOC (N.pnl_tend, pslots) ;
    ALLOC (N.pnl_cols, pslots * PANEL_MAX) ;
    ALLOC (N.rcursor, 1) ;
    ALLOC (N.sumrank, 4) ; N.maxfrank = N.sumrank + 1 ; N.maxfm = N.sumrank + 2 ; N.rank1 = N.sumrank + 3 ;
    ALLOC (N.flops, 4) ;
    ALLOC (N.W, m) ;
    ALLOC (N.dbg, 64) ;
    ALLOC (N.lvlstat, 4) ;
    ALLOC (N.base1, nf) ; ALLOC (N.base2, nf) ;
    ALLOC (h->d_err, 1) ;
    ALLOC (h->d_hcol, rjsize) ; ALLOC (h->d_nh, nf) ;
    ALLOC (h->d_rlen, rjsize) ;         // R part length of every front column (k_htable; the R extraction tables are lazy)
    ALLOC (h->d_HPinv64, m) ;
    ALLOC (h->d_Hii64, hisize) ;
    ALLOC (h->d_wide, rjsize + 2 * nf + 2) ;
    if (h->debug_capture)
    {
        h->h_capOff.assign ((size_t) nf + 1, 0) ;
        for (I64 f = 0 ; f < nf ; f++)
            h->h_capOff [f+1] = h->h_capOff [f] + (I64) FmB [f] * (Rp [f+1] - Rp [f]) ;
        ALLOC (h->d_capA, h->h_capOff [nf]) ;
        ALLOC (h->d_capF, h->h_capOff [nf]) ;
    }
    PLAN_MARK ("remaining allocations") ;
    if (!h->host_only) CK (cudaStreamSynchronize (h->stream)) ;
    PLAN_MARK ("stream sync") ;
    h->analyzed = true ;
    memset (&h->stats, 0, sizeof (h->stats)) ;
    h->stats.ms_plan = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - t0).count () ;
    h->stats.nlevels = (I64) h->ls_all.levels.size () ;
    h->stats.nf_small = nsmall_tot ; h->stats.nf_big = nf - nsmall_tot ;
    h->stats.device_bytes = (I64) h->device_bytes ;
    return STMQR_OK ;
}

// -------------------------------------------------------------------------------------------------
int stmqr_b200_upload_matrix (stmqr_handle h, const stmqr_csc_view *A)
{
    if (!h || !A || !h->analyzed) return fail (h, STMQR_ERR_INVALID, "upload_matrix: analyze first") ;
    if (h->host_only) return fail (h, STMQR_ERR_NO_DEVICE, "upload_matrix: planner handle (no device)") ;
    cudaSetDevice (h->device) ;
    if (A->nrow != h->m || A->ncol != h->n || !A->p || (A->p [A->ncol] > 0 && (!A->i || !A->x)))
        return fail (h, STMQR_ERR_INVALID, "upload_matrix: matrix does not match the analysis") ;
    const I64 nnz = A->p [A->ncol] ;
    if (nnz != h->anz) return fail (h, STMQR_ERR_INVALID, "upload_matrix: nnz(A) != anz of the analysis") ;
    if (!h->d_Ap || h->a_ncol != A->ncol || h->a_nnz_cap < nnz)
    {
        ALLOC (h->d_Ap, A->ncol + 1) ;
        ALLOC (h->d_Ai, nnz) ;
        ALLOC (h->d_Ax, nnz) ;
        ALLOC (h->d_slot, nnz) ;
        ALLOC (h->d_vAp, A->ncol + 1) ;
        ALLOC (h->d_vAi, nnz) ;
        ALLOC (h->d_vdiff, 1) ;
        h->a_ncol = A->ncol ; h->a_nnz_cap = nnz ;
    }
    h->slot_valid = false ; h->values_only = false ;        // a new pattern: k_build_S refills the slot map
    CK (cudaEventRecord (h->ev0, h->stream)) ;
    CK (cudaMemcpyAsync (h->d_Ap, A->p, (A->ncol + 1) * sizeof (I64), cudaMemcpyHostToDevice, h->stream)) ;
    if (nnz > 0)
    {
        CK (cudaMemcpyAsync (h->d_Ai, A->i, nnz * sizeof (I64), cudaMemcpyHostToDevice, h->stream)) ;
        CK (cudaMemcpyAsync (h->d_Ax, A->x, nnz * sizeof (double), cudaMemcpyHostToDevice, h->stream)) ;
    }
    CK (cudaEventRecord (h->ev1, h->stream)) ;
    CK (cudaStreamSynchronize (h->stream)) ;
    float ms = 0 ;
    cudaEventElapsedTime (&ms, h->ev0, h->ev1) ;
    h->stats.ms_h2d = ms ;
    h->have_matrix = true ;
    return STMQR_OK ;
}

// New values on the RESIDENT pattern (the matrix last given to upload_matrix, whose A -> S slot map the
// first factorization left on the device): 8 bytes per entry go to the device instead of 24, and S is built
// by a scatter instead of a search.  The caller vouches that Ap/Ai are unchanged.
int stmqr_b200_upload_values (stmqr_handle h, const double *Ax, int64_t nnz)
{
    if (!h || !h->analyzed || h->host_only) return fail (h, STMQR_ERR_INVALID, "upload_values: analyze first") ;
    if (!h->d_Ax || !h->slot_valid || nnz != h->anz || (nnz > 0 && !Ax))
        return fail (h, STMQR_ERR_INVALID, "upload_values: no resident pattern with this many entries (upload_matrix + one factorization first)") ;
    cudaSetDevice (h->device) ;
    CK (cudaEventRecord (h->ev0, h->stream)) ;
    if (nnz > 0) CK (cudaMemcpyAsync (h->d_Ax, Ax, nnz * sizeof (double), cudaMemcpyHostToDevice, h->stream)) ;
    CK (cudaEventRecord (h->ev1, h->stream)) ;
    CK (cudaStreamSynchronize (h->stream)) ;
    float ms = 0 ;
    cudaEventElapsedTime (&ms, h->ev0, h->ev1) ;
    h->stats.ms_h2d = ms ;
    h->have_matrix = true ;
    h->values_only = true ;
    h->n_values_only++ ;
    return STMQR_OK ;
}

// -------------------------------------------------------------------------------------------------
// ---- the numeric phase in pieces (one GPU: begin, all levels, hpinv_a, hpinv_b; several GPUs: the
// host interleaves the exchange of the cut contribution blocks and the merges, see py/stmqr_b200/dist.py)
int stmqr_b200_factorize_begin (stmqr_handle h, double tol, int64_t ntol)
{
    if (!h || !h->analyzed || !h->have_matrix || h->host_only)
        return fail (h, STMQR_ERR_INVALID, "factorize: analyze and upload_matrix first") ;
    cudaSetDevice (h->device) ;
    cudaStream_t st = h->stream ;
    DSym &S = h->S ; DNum &N = h->N ;
    if (!h->do_rank_detection) tol = -1 ;           // SparseQR_factorize.c:285-289
    h->cur_tol = tol ; h->cur_ntol = ntol ;
    h->launches = 0 ;
    h->factorized = false ;
    h->htable_valid = false ;
    h->rcount_econ = h->rcount_nnz = -1 ;
    for (cudaEvent_t e : h->evLevelT) cudaEventDestroy (e) ;
    h->evLevelT.clear () ; h->lvlNote.clear () ;
    CK (cudaEventRecord (h->ev0, st)) ;
    CK (cudaMemsetAsync (N.Rdead, 0, std::max<I64> (h->n, 1), st)) ;
    CK (cudaMemsetAsync (N.rcursor, 0, sizeof (unsigned long long), st)) ;
    CK (cudaMemsetAsync (N.sumrank, 0, 4 * sizeof (I32), st)) ;
    CK (cudaMemsetAsync (N.flops, 0, 4 * sizeof (double), st)) ;
    h->evused = 0 ; h->evclass.clear () ; h->evtag.clear () ; h->curtag = 0 ;
    CK (cudaMemsetAsync (h->d_err, 0, sizeof (I32), st)) ;
    CK (cudaMemsetAsync (N.dbg, 0, 64 * sizeof (unsigned long long), st)) ;
    CK (cudaMemsetAsync (N.HTau, 0, std::max<I64> (h->rjsize, 1) * sizeof (double), st)) ;
    // every slot of S is written by k_build_S when A matches the analysed pattern; a slot that no
    // entry of A maps to must read as an explicit zero, not as stale memory
    CK (cudaMemsetAsync (N.Sx, 0, std::max<I64> (h->anz, 1) * sizeof (double), st)) ;
    if (h->nparts > 1)
    {
        // arrays that are merged over the GPUs with an element-wise max: neutral element first
        CK (cudaMemsetAsync (N.Hm, 0, std::max<I64> (h->nf, 1) * sizeof (I32), st)) ;
        CK (cudaMemsetAsync (N.Hr, 0, std::max<I64> (h->nf, 1) * sizeof (I32), st)) ;
        CK (cudaMemsetAsync (N.Cm, 0, std::max<I64> (h->nf, 1) * sizeof (I32), st)) ;
        CK (cudaMemsetAsync (N.W, 0xff, std::max<I64> (h->m, 1) * sizeof (I32), st)) ;
        // per-front outputs that stmqr_b200_gather_outputs merges: zero where this GPU does not write
        CK (cudaMemsetAsync (N.stair, 0, std::max<I64> (h->rjsize, 1) * sizeof (I32), st)) ;
        CK (cudaMemsetAsync (N.Roff, 0, std::max<I64> (h->nf, 1) * sizeof (I64), st)) ;
        CK (cudaMemsetAsync (h->d_Hii64, 0, std::max<I64> (h->hisize, 1) * sizeof (I64), st)) ;
    }
    if (h->anz > 0)
    {
        if (h->values_only && h->slot_valid)
        {
            LAUNCH (0, k_scatter_values<<<grid_for (h->anz, 256), 256, 0, st>>> (h->anz, h->d_Ax, h->d_slot, N.Sx)) ;
        }
        else
        {
            LAUNCH (0, k_build_S<<<grid_for (h->n * 32, 256), 256, 0, st>>> ((I32) h->n, h->d_Ap, h->d_Ai, h->d_Ax, S, N.Sx, h->d_slot, h->d_err)) ;
            h->slot_valid = true ;          // (an entry outside the pattern raises d_err: the factorization fails, and
        }                                   //  upload_matrix resets the flag before the next pattern)
    }
    return STMQR_OK ;
}

// The K = 128 trailing update of columns [cb,ce) of the first nfronts fronts of a wide level with the block
// reflector of the outer block WA.buf (kernels_wide.cuh): W = Vb'C over row splits, W2 = -T' sum W, C += Vb W2.
// Every column tile is independent of the others, so any partition of the columns into ranges (look-ahead, or
// column blocks of one front spread over several GPUs) performs bit for bit the same arithmetic per column.
static void wide_outer_update (stmqr_handle h, const WideArgs &WA, I32 actFm, I32 nsplit_max, cudaStream_t su,
    I32 nfronts, I32 cb, I32 ce)
{
    if (nfronts <= 0 || cb >= ce) return ;
    DSym &S = h->S ; DNum &N = h->N ;
    const I32 nrt = std::min<I32> (WA.ldv / W_RT, (actFm + W_RT - 1) / W_RT) ;
    const I32 nsplA = std::min<I32> (nsplit_max, (actFm + WIDE_RS - 1) / WIDE_RS) ;
    const I32 nct = (ce - cb + W_NC - 1) / W_NC ;
    LAUNCH (14, k_wide_vtc<4><<<dim3 (nct * nsplA, nfronts), 256, wide_vtc_smem_bytes<4> (), su>>> (WA, S, N, WIDE_OUTER, 0, cb, ce, nct)) ;
    LAUNCH (15, k_wide_wt<4><<<dim3 ((ce - cb + 15) / 16, nfronts), 256, 0, su>>> (WA, S, N, WIDE_OUTER, 0, cb, ce)) ;
    if (h->opt.reserved & 16)
    {
        LAUNCH (16, k_wide_apply<4><<<dim3 (nct * nrt, nfronts), 256, wide_apply_smem_bytes<4> (), su>>> (WA, S, N, WIDE_OUTER, 0, cb, ce, nct)) ;
    }
    else
    {
        // row groups: about 3 CTAs per SM in total, every CTA walks down its share of the row tiles
        // (the count that fills whole waves of one CTA per SM best, every CTA at least 4 row tiles)
        I32 nrg = 1 ;
        {
            const I64 base = (I64) nct * nfronts ;
            double best = -1 ;
            const I32 hi = std::max<I32> (1, std::min<I32> (nrt / 4, (I32) ((8 * (I64) h->nsm + base - 1) / base))) ;
            for (I32 c = 1 ; c <= hi ; c++)
            {
                const I64 tot = base * c, waves = (tot + h->nsm - 1) / h->nsm ;
                const double eff = (double) tot / (double) (waves * h->nsm) ;
                if (eff >= best - 1e-9) { best = eff ; nrg = c ; }
            }
        }
        LAUNCH (16, k_wide_apply_rows<<<dim3 (nct * nrg, nfronts), 256, wide_apply_rows_smem_bytes (), su>>> (WA, S, N, cb, ce, nct, nrg)) ;
    }
}

// ---------------------------------------------------------------------------------------------
// Cooperative front (several GPUs): the top of the etree of a 3-D problem is a chain of single
// fronts that holds most of the flops; subtree ownership leaves it on one GPU.  Here the front's HOME GPU keeps
// the whole critical path (assembly, every panel, the block reflectors, the pack), and the K = 128 trailing
// update -- the bulk of the flops -- is spread over all GPUs by column chunks:
//   * home assembles F and sends every other GPU the column chunks it owns (chunks of 512 columns, dealt so that
//     home, which also runs the panels, gets a smaller share);
//   * after the four panels of outer block J, home broadcasts the block reflector (V in the compact layout,
//     T, the row window) and everybody applies it to the columns it still holds;
//   * the owner of block J+2 applies it to that block FIRST and sends the block home, where it arrives while
//     the panels of block J+1 run: home applies the last reflector (V_{J+1}) itself and factorizes it.
// Every column sees the same reflectors in the same order through the same kernels as on one GPU, so the
// factorization is bit for bit the single-GPU one.  All transfers are ordered on the engine streams (NCCL
// send/recv/broadcast); the only host synchronisation is one 16-byte header per cooperative level.
// ---------------------------------------------------------------------------------------------
constexpr I32 COOP_CHUNK = 4 * WB ;

static void coop_plan (stmqr_handle_s::Coop &C, int np, int home, I32 f, I32 fn)
{
    C.home = home ; C.f = f ; C.fn = fn ;
    const int nch = (int) ((fn + COOP_CHUNK - 1) / COOP_CHUNK) ;
    // share of the home GPU: the panels of a block cost about rho of its trailing update on one GPU
    const double rho = 0.14 ;
    const double hs = std::max (0.0, (1.0 - rho * (np - 1)) / np) ;
    std::vector<double> target ((size_t) np, (np > 1) ? (1.0 - hs) / (np - 1) : 1.0), have ((size_t) np, 0.0) ;
    target [(size_t) home] = hs ;
    C.chunk_owner.assign ((size_t) std::max (nch, 1), home) ;      // chunk 0 (blocks 0 .. 3) never leaves home
    for (int c = 1 ; c < nch ; c++)
    {
        int best = -1 ; double bd = -1e300 ;
        for (int q = 0 ; q < np ; q++)
        {
            const int p = (home + 1 + q) % np ;                    // ties: the GPUs after home first
            const double d = target [(size_t) p] * c - have [(size_t) p] ;
            if (target [(size_t) p] > 0 && d > bd + 1e-12) { bd = d ; best = p ; }
        }
        if (best < 0) best = home ;
        C.chunk_owner [(size_t) c] = best ; have [(size_t) best] += 1.0 ;
    }
}
static inline int coop_owner (const stmqr_handle_s::Coop &C, I32 col) { return C.chunk_owner [(size_t) (col / COOP_CHUNK)] ; }
// the column ranges [a,b) that GPU `me` holds at or after column `from`
static void coop_ranges (const stmqr_handle_s::Coop &C, int me, I32 from, std::vector<std::pair<I32, I32>> &out)
{
    out.clear () ;
    for (I32 c = from / COOP_CHUNK ; (I64) c * COOP_CHUNK < C.fn ; c++)
    {
        if (C.chunk_owner [(size_t) c] != me) continue ;
        const I32 a = std::max<I32> (from, c * COOP_CHUNK), b = std::min<I32> (C.fn, (c + 1) * COOP_CHUNK) ;
        if (a >= b) continue ;
        if (!out.empty () && out.back ().second == a) out.back ().second = b ; else out.emplace_back (a, b) ;
    }
}
// rows of the block reflector that travel: every row of the front, padded like k_wide_vextract pads them
static inline I32 coop_ldv (I32 fm, I32 ldv) { return std::min<I32> (ldv, ((fm + W_RT - 1) / W_RT) * W_RT + W_RT) ; }

// a GPU that does not own the front of a cooperative level: holds column chunks of F and applies the broadcast
// block reflectors to them
static int coop_peer_level (stmqr_handle h, I64 gl)
{
    Transport *T = (Transport *) h->transport ;
    stmqr_handle_s::Coop &C = h->coop ;
    cudaStream_t st = h->stream ;
    DNum &N = h->N ;
    const Level &Lv = h->ls_all.levels [(size_t) gl] ;
    const int me = T->rank () ;
    int s ;
    // header from home: does the level go cooperative (actual # rows), and the # rows
    if ((s = T->bcast (h, N.lvlstat, 4 * sizeof (I32), C.home)) != STMQR_OK) return s ;
    I32 hdr [4] = {0, 0, 0, 0} ;
    CK (cudaMemcpyAsync (hdr, N.lvlstat, sizeof (hdr), cudaMemcpyDeviceToHost, st)) ;
    CK (cudaStreamSynchronize (st)) ;
    if (!hdr [1]) return STMQR_OK ;
    C.on = true ; C.levels_run++ ;
    const I32 fm = hdr [2], fn = C.fn, f = C.f ;
    CK (cudaMemcpyAsync (N.Hm + f, hdr + 2, sizeof (I32), cudaMemcpyHostToDevice, st)) ;   // (ld of F in the kernels; max-merged later: same value as home's)
    double *F = N.F + h->h_Foff [f] ;
    std::vector<std::pair<I32, I32>> mine ;
    coop_ranges (C, me, 0, mine) ;
    if ((s = T->group_begin ()) != STMQR_OK) return s ;
    for (auto &r : mine)
        if ((s = T->recv (h, F + (I64) r.first * fm, (size_t) fm * (size_t) (r.second - r.first) * sizeof (double), C.home)) != STMQR_OK) return s ;
    if ((s = T->group_end ()) != STMQR_OK) return s ;
    const I32 ldvC = coop_ldv (fm, Lv.ldv) ;
    WideArgs WA ; WA.fronts = h->ls_all.d_fronts + Lv.first ; WA.count = 1 ; WA.buf = 0 ; WA.ldv = ldvC ;
    WA.rs = WIDE_RS ; WA.nsplit = Lv.nsplit ; WA.ncmax = Lv.maxfn ;
    const I32 nblk = (fn + WB - 1) / WB ;
    for (I32 J = 0 ; J < nblk ; J++)
    {
        const I32 cb = (J + 1) * WB ;
        if (cb >= fn) break ;
        const int buf = J & 1 ;
        WA.buf = buf ;
        h->curtag = ((long long) gl << 32) | (long long) (J * WB) ;
        if ((s = T->group_begin ()) != STMQR_OK) return s ;
        if ((s = T->bcast (h, N.wVb + (I64) buf * ((I64) ldvC * WB), (size_t) ldvC * WB * sizeof (double), C.home)) != STMQR_OK) return s ;
        if ((s = T->bcast (h, N.wTbt + (I64) buf * (WB * WB), (size_t) WB * WB * sizeof (double), C.home)) != STMQR_OK) return s ;
        if ((s = T->bcast (h, N.wblk + (I64) buf * 4, 4 * sizeof (I32), C.home)) != STMQR_OK) return s ;
        if ((s = T->group_end ()) != STMQR_OK) return s ;
        // block J+1 is at home already; block J+2 goes there after this update
        const I32 c2 = cb + WB ;
        if (c2 >= fn) continue ;
        const I32 e2 = std::min<I32> (c2 + WB, fn) ;
        const bool send2 = (coop_owner (C, c2) == me) ;
        if (send2) wide_outer_update (h, WA, fm, Lv.nsplit, st, 1, c2, e2) ;
        coop_ranges (C, me, send2 ? e2 : c2, mine) ;
        for (auto &r : mine) wide_outer_update (h, WA, fm, Lv.nsplit, st, 1, r.first, r.second) ;
        if (send2 && (s = T->send (h, F + (I64) c2 * fm, (size_t) fm * (size_t) (e2 - c2) * sizeof (double), C.home)) != STMQR_OK) return s ;
    }
    CK (cudaGetLastError ()) ;
    return STMQR_OK ;
}

// part: 0 = every front (one GPU), 1 = the etree subtrees this GPU owns, 2 = the top of the tree
// one etree level of the level set LS on this handle's streams (everything asynchronous except the one small
// read-back on levels with very large fronts)
static int run_level (stmqr_handle h, const LevelSet &LS, const Level &Lv, long long levelno)
{
    cudaStream_t st = h->stream, st2 = h->stream2 ;
    DSym &S = h->S ; DNum &N = h->N ;
    const double tol = h->cur_tol ; const I64 ntol = h->cur_ntol ;
    const int nc_update = (h->opt.reserved >> 8) & 0xff ;   // 0 auto, 2 or 4: ring stages of the update kernel
    const int PB = (h->opt.panel > 0 && h->opt.panel <= PANEL_MAX) ? h->opt.panel : PANEL_MAX ;
    static const bool level_times = getenv ("STMQR_B200_LEVEL_TIMES") != nullptr ;
    auto mark_level = [&] (const std::string &note) {
        if (!level_times) return ;
        cudaEvent_t e ;
        cudaEventCreate (&e) ;
        cudaEventRecord (e, st) ;
        h->evLevelT.push_back (e) ; h->lvlNote.push_back (note) ;
    } ;
    if (level_times && h->evLevelT.empty ()) mark_level ("start") ;
    {
        h->curtag = levelno << 32 ;
        const I32 *fr = LS.d_fronts + Lv.first ;
        const I32 nbig = Lv.nbig ;
        const int nsl = (int) std::min<I64> (148, std::max<I64> (1, Lv.maxFelems / 8192)) ;
        auto big_path = [&] () -> int {
        // one CTA per front: as many threads as the widest front of the level can use (scans and per-column loops)
        const int fthreads = (Lv.maxfn >= 2048) ? 1024 : ((Lv.maxfn >= 512) ? 512 : ((Lv.maxfn >= 192) ? 256 : 128)) ;
        LAUNCH (1, k_front_setup<<<nbig, fthreads, 0, st>>> (fr, S, N)) ;
        // levels with large fronts: the ACTUAL # rows of the tallest front decides which kernels run
        // (the symbolic bound is ~2x too big under rank detection).  One tiny read-back per such level.
        I32 actFm = Lv.maxFm ;
        if (Lv.wide || Lv.maxFm >= h->grid_rows)
        {
            k_level_maxfm<<<1, 256, 0, st>>> (fr, nbig, N) ; h->launches++ ;
            CK (cudaMemcpyAsync (h->pin_lvl, N.lvlstat, sizeof (I32), cudaMemcpyDeviceToHost, st)) ;
            CK (cudaStreamSynchronize (st)) ;
            actFm = std::min (Lv.maxFm, std::max (1, h->pin_lvl [0])) ;
        }
        {
            // zero-fill at ~8 CTAs of 16-byte stores per SM over the whole level, then the scatter
            const int nz = (int) std::min<I64> (148, std::max<I64> (1, std::min<I64> (Lv.maxFelems / 4096, (8 * (I64) h->nsm + nbig - 1) / nbig))) ;
            LAUNCH (2, k_zero_fronts<<<dim3 (nbig, nz), 256, 0, st>>> (fr, S, N)) ;
            LAUNCH (2, k_assemble<<<dim3 (nbig, nsl), 256, 0, st>>> (fr, S, N)) ;
        }
        if (h->debug_capture)
        {
            CK (cudaStreamSynchronize (st)) ;
            std::vector<I32> hm ((size_t) h->nf) ;
            CK (cudaMemcpy (hm.data (), N.Hm, h->nf * sizeof (I32), cudaMemcpyDeviceToHost)) ;
            for (I32 i = 0 ; i < nbig ; i++)
            {
                I32 f = LS.fronts [Lv.first + i] ;
                I64 cnt = (I64) hm [f] * (h->h_Rp [f+1] - h->h_Rp [f]) ;
                if (cnt > 0) CK (cudaMemcpy (h->d_capA + h->h_capOff [f], N.F + h->h_Foff [f],
                    cnt * sizeof (double), cudaMemcpyDeviceToDevice)) ;
            }
        }
        static const bool checking = getenv ("STMQR_B200_CHECK") != nullptr ;
        auto check = [&] (const char *stage, long long a, long long b) {
            if (!checking || h->check_hit) return ;
            cudaStreamSynchronize (st) ; cudaStreamSynchronize (st2) ;
            I32 init [4] = {-1, 0, 0, 0} ;
            cudaMemcpy (N.lvlstat, init, sizeof (init), cudaMemcpyHostToDevice) ;
            k_check_finite<<<nbig, 256, 0, st>>> (fr, nbig, S, N, N.lvlstat) ;
            I32 res [4] ;
            cudaMemcpy (res, N.lvlstat, sizeof (res), cudaMemcpyDeviceToHost) ;
            if (res [0] >= 0)
            {
                h->check_hit = true ;
                fprintf (stderr, "STMQR_B200_CHECK: first |F| >= 1e100 or non-finite after %s (level %lld, %lld, %lld): front %d row %d col %d\n",
                    stage, levelno, a, b, res [0], res [1], res [2]) ;
            }
        } ;
        check ("assemble", 0, 0) ;
        LevelArgs L ; L.fronts = fr ; L.count = nbig ; L.tol = tol ; L.ntol = ntol ;
        L.flags = (h->opt.reserved & 32) ? 1 : 0 ;          // bit 5: one column per panel exchange
        // ---- front QR of the level: panel steps of PB columns over all active fronts ---------------
        // cluster size: the row slab of one CTA (rows / CS x PB doubles) should fit in shared memory
        int CS = 1 ;
        while (CS < PANEL_CLUSTER_MAX && ((I64) (actFm + CS - 1) / CS + 4) * PB > PANEL_SLAB_MAX_DOUBLES) CS *= 2 ;
        // few fronts in the level: idle SMs are better spent on shorter slabs (a column step sweeps the
        // slab three times through shared memory); up to the non-portable cluster size 16
        while (CS < h->cluster_max && (I64) nbig * CS * 2 <= h->nsm && (actFm + CS - 1) / CS > h->cluster_rows) CS *= 2 ;
        // fronts too tall for a cluster of 8 shared-memory slabs: G CTAs per front with a global-memory
        // exchange (k_panel_grid), as many fronts per launch as fit one CTA per SM
        const bool gridpanel = (actFm >= h->grid_rows) && PB == PANEL_MAX && !(h->opt.reserved & 4) ;
        const I64 rowsPerCta = ((I64) (actFm + CS - 1) / CS + 7) & ~(I64) 3 ;
        // (at least 2 x 32 x 33 doubles: the leader builds T in the slab after writing it back)
        const I32 slabCap = (I32) std::max<I64> (2 * PANEL_MAX * (PANEL_MAX + 1),
            std::min<I64> (PANEL_SLAB_MAX_DOUBLES, rowsPerCta * PB)) ;
        // 8 warps per CTA beat 16 even on 768-row slabs (measured: a column step is latency bound, and
        // two CTAs of different fronts per SM hide each other's exchanges); 16 only on request
        int pthreads = (rowsPerCta >= 64) ? 256 : 128 ;
        // many short fronts: 4-warp CTAs, up to 5-6 of them per SM (tunable: STMQR_B200_PANEL128_ROWS)
        if (rowsPerCta <= h->panel128_rows && (I64) nbig >= 2 * (I64) h->nsm) pthreads = 128 ;
        if (((h->opt.reserved >> 16) & 0xff) >= 16 && rowsPerCta >= 256 && CS == 1) pthreads = 512 ;
        if ((h->opt.reserved >> 16) & 0xff) pthreads = std::min (pthreads, 32 * ((h->opt.reserved >> 16) & 0xff)) ;   // tuning
        // number of fronts of the level with more than k columns (sorted by # columns descending)
        auto active_at = [&] (I32 k, I32 hi) -> I32 {
            I32 lo = 0 ;
            while (lo < hi)
            {
                I32 mid = (lo + hi) / 2 ;
                I32 f = LS.fronts [Lv.first + mid] ;
                if (h->h_Rp [f+1] - h->h_Rp [f] > k) lo = mid + 1 ; else hi = mid ;
            }
            return lo ;
        } ;
        const I32 gneed = (I32) (((I64) actFm + 8 + 759) / 760) ;
        auto launch_panel = [&] (I32 active, I32 k1, I32 parity) -> cudaError_t {
            if (gridpanel && gneed <= h->nsm)
            {
                const I32 per = std::max<I32> (1, h->nsm / gneed) ;
                const size_t smem = (size_t) (PANEL_SLAB_MAX_DOUBLES + panel_scratch_doubles (16)) * sizeof (double) ;
                for (I32 s0 = 0 ; s0 < active ; s0 += per)
                {
                    const I32 nb = std::min<I32> (per, active - s0) ;
                    // no more CTAs than the step-cost model of k_panel_grid can use (they all have to
                    // become resident before the panel starts)
                    const I32 G = std::min<I32> (std::min<I32> (h->nsm / nb, std::max<I32> (gneed, h->grid_maxg)),
                        std::max<I32> (gneed, (I32) std::sqrt (0.84 * (double) actFm) + 1)) ;
                    k_panel_grid<<<(unsigned) (G * nb), 512, smem, st>>> (L, S, N, k1, (I32) PB, parity,
                        (I32) PANEL_SLAB_MAX_DOUBLES, G, s0, ++h->grid_seq) ;
                    if (s0 + per < active) h->launches++ ;
                }
                return cudaGetLastError () ;
            }
            cudaLaunchConfig_t cfg = {} ;
            cfg.gridDim = dim3 ((unsigned) active * CS, 1, 1) ;
            cfg.blockDim = dim3 (pthreads, 1, 1) ;
            cfg.dynamicSmemBytes = (size_t) (slabCap + panel_scratch_doubles (pthreads / 32) + (CS > 1 ? PANEL_XR_DOUBLES : 0)) * sizeof (double) ;
            cfg.stream = st ;
            cudaLaunchAttribute at [1] ;
            at [0].id = cudaLaunchAttributeClusterDimension ;
            at [0].val.clusterDim.x = CS ; at [0].val.clusterDim.y = 1 ; at [0].val.clusterDim.z = 1 ;
            cfg.attrs = at ; cfg.numAttrs = 1 ;
            if (pthreads == 128) return cudaLaunchKernelEx (&cfg, k_panel_cluster<128, 6>, L, S, N, k1, (I32) PB, parity, slabCap) ;
            if (pthreads == 256) return cudaLaunchKernelEx (&cfg, k_panel_cluster<256, 2>, L, S, N, k1, (I32) PB, parity, slabCap) ;
            return cudaLaunchKernelEx (&cfg, k_panel_cluster<512, 1>, L, S, N, k1, (I32) PB, parity, slabCap) ;
        } ;
        auto launch_update = [&] (cudaStream_t su, I32 nfronts, I32 cbeg, I32 cend, I32 parity) {
            if (nfronts <= 0 || cbeg >= cend) return ;
            // few CTAs (less than one per SM): deeper ring per CTA; many: two CTAs per SM
            const I32 ntile = (cend - cbeg + UPD_NC - 1) / UPD_NC ;
            const I64 ctas = (I64) nfronts * ntile ;
            // too few column tiles to fill the GPU and tall fronts: split the rows over a cluster
            I32 rsf = 1 ;
            while (rsf < h->update_rsf_max && ctas * rsf * 2 <= 4 * (I64) h->nsm && actFm / (rsf * 2) >= 256) rsf *= 2 ;
            cudaLaunchConfig_t cfg = {} ;
            cfg.gridDim = dim3 (nfronts, ntile, rsf) ;
            cfg.blockDim = dim3 (256, 1, 1) ;
            cfg.stream = su ;
            cudaLaunchAttribute at [1] ;
            at [0].id = cudaLaunchAttributeClusterDimension ;
            at [0].val.clusterDim.x = 1 ; at [0].val.clusterDim.y = 1 ; at [0].val.clusterDim.z = rsf ;
            cfg.attrs = at ; cfg.numAttrs = (rsf > 1) ? 1 : 0 ;
            if (nc_update == 4 || (nc_update == 0 && ctas * rsf <= h->nsm))
            {
                cfg.dynamicSmemBytes = update_smem_bytes<4> () ;
                cudaLaunchKernelEx (&cfg, k_update_dmma<4>, L, S, N, cbeg, cend, parity, rsf) ;
            }
            else
            {
                cfg.dynamicSmemBytes = update_smem_bytes<2> () ;
                cudaLaunchKernelEx (&cfg, k_update_dmma<2>, L, S, N, cbeg, cend, parity, rsf) ;
            }
        } ;
        // look-ahead pays only when the trailing update is much bigger than its first 32 columns
        const bool lookahead = !h->opt.profile_phases && !(h->opt.reserved & 1) && Lv.maxFelems >= h->lookahead_elems ;
        const bool wide = Lv.wide && actFm >= h->wide_rows && !(h->opt.reserved & 2) && PB == PANEL_MAX ;
        // cooperative front: this GPU is the home of the level's only front (factorize_dist armed the level on
        // every GPU); tell the others whether it goes ahead, and hand out their column chunks of the assembled F
        Transport *const coopT = h->coop.armed ? (Transport *) h->transport : nullptr ;
        const int coop_me = coopT ? coopT->rank () : 0 ;
        bool coop_on = false ;
        if (coopT)
        {
            coop_on = wide && nbig == 1 && Lv.count == 1 ;
            int cs ;
            I32 hdr [4] = {0, coop_on ? 1 : 0, actFm, 0} ;
            CK (cudaMemcpyAsync (N.lvlstat, hdr, sizeof (hdr), cudaMemcpyHostToDevice, st)) ;     // (pageable source: staged before the call returns)
            if ((cs = coopT->bcast (h, N.lvlstat, sizeof (hdr), coop_me)) != STMQR_OK) return cs ;
            if (coop_on)
            {
                h->coop.on = true ; h->coop.levels_run++ ;
                const I32 f0 = LS.fronts [Lv.first] ;
                double *F0 = N.F + h->h_Foff [f0] ;
                if ((cs = coopT->group_begin ()) != STMQR_OK) return cs ;
                for (int p = 0 ; p < coopT->nranks () ; p++)
                {
                    if (p == coop_me) continue ;
                    std::vector<std::pair<I32, I32>> theirs ;
                    coop_ranges (h->coop, p, 0, theirs) ;
                    for (auto &r : theirs)
                        if ((cs = coopT->send (h, F0 + (I64) r.first * actFm, (size_t) actFm * (size_t) (r.second - r.first) * sizeof (double), p)) != STMQR_OK) return cs ;
                }
                if ((cs = coopT->group_end ()) != STMQR_OK) return cs ;
            }
        }
        if (wide)
        {
            // ---- two-level blocking (kernels_wide.cuh): outer blocks of 128 columns = 4 panels -------
            WideArgs WA ; WA.fronts = fr ; WA.count = nbig ; WA.buf = 0 ; WA.ldv = Lv.ldv ;
            WA.rs = WIDE_RS ; WA.nsplit = Lv.nsplit ; WA.ncmax = Lv.maxfn ;
            WideArgs WI = WA ; WI.rs = WIDE_RS_IN ; WI.nsplit = Lv.nsplit_in ;
            const I32 nrt = std::min<I32> (Lv.ldv / W_RT, (actFm + W_RT - 1) / W_RT) ;
            const I32 nsplA = std::min<I32> (Lv.nsplit, (actFm + WIDE_RS - 1) / WIDE_RS) ;
            const I32 nsplI = std::min<I32> (Lv.nsplit_in, (actFm + WIDE_RS_IN - 1) / WIDE_RS_IN) ;
            // the K = 32 update of columns [cb,ce) (inside the block) by panel p of the block
            // after panel p: Gram blocks of V_p with the earlier panels of the block (p >= 1) and the
            // K = 32 update of the columns [cb,ce) that are left in the block (p < 3), one launch
            auto inner_update = [&] (I32 nfronts, I32 p, I32 cb, I32 ce) {
                const I32 nct = (ce > cb) ? (ce - cb + W_NC - 1) / W_NC : 0 ;
                const I32 ncg = (p * PANEL_MAX + W_NC - 1) / W_NC ;
                if (nct + ncg == 0) return ;
                LAUNCH (9, k_wide_vtc<1><<<dim3 ((nct + ncg) * nsplI, nfronts), 256, wide_vtc_smem_bytes<1> (), st>>> (WI, S, N, WIDE_INNER, p, cb, ce, nct)) ;
                if (nct == 0) return ;
                LAUNCH (10, k_wide_wt_inner<<<dim3 ((ce - cb + 7) / 8, nfronts), 256, 0, st>>> (WI, S, N, p, cb, ce)) ;
                LAUNCH (11, k_wide_apply<1><<<dim3 (nct * nrt, nfronts), 256, wide_apply_smem_bytes<1> (), st>>> (WI, S, N, WIDE_INNER, p, cb, ce, nct)) ;
            } ;
            // (events of LAUNCH are recorded on the main stream: only meaningful when su == st)
            auto outer_update = [&] (cudaStream_t su, I32 nfronts, I32 cb, I32 ce) {
                wide_outer_update (h, WA, actFm, Lv.nsplit, su, nfronts, cb, ce) ;
            } ;
            const I32 nblk = (Lv.maxfn + WB - 1) / WB ;
            bool pending [2] = {false, false} ;     // a "rest of the trailing matrix" update in flight on stream2
            for (I32 J = 0 ; J < nblk ; J++)
            {
                const I32 j0 = J * WB ;
                if (active_at (j0, nbig) == 0) break ;
                WA.buf = WI.buf = J & 1 ;
                for (I32 p = 0 ; p < WB_PANELS ; p++)
                {
                    const I32 k1 = j0 + p * PANEL_MAX ;
                    if (k1 >= Lv.maxfn) break ;
                    const I32 act = active_at (k1, nbig) ;
                    if (act == 0) break ;
                    h->curtag = (levelno << 32) | (long long) k1 ;
                    LAUNCH (3, CK (launch_panel (act, k1, p))) ;
                    LAUNCH (8, k_wide_vextract<<<dim3 ((std::min<I32> (Lv.ldv, actFm + 2 * W_RT) + 255) / 256, act), 256, 0, st>>> (WI, S, N, p)) ;
                    const I32 cb = k1 + PANEL_MAX, ce = std::min<I32> (j0 + WB, Lv.maxfn) ;
                    const I32 act2 = (cb < Lv.maxfn) ? active_at (cb, act) : 0 ;      // fronts with columns right of the panel
                    if (act2 > 0) inner_update (act2, p, cb, std::max (cb, ce)) ;
                }
                const I32 cb = j0 + WB ;
                if (cb >= Lv.maxfn) break ;
                const I32 act2 = active_at (cb, nbig) ;
                if (act2 == 0) break ;
                LAUNCH (13, k_wide_tmerge<<<act2, 1024, wide_tmerge_smem_bytes (), st>>> (WI, S, N)) ;
                if (coop_on)
                {
                    int cs ;
                    const I32 f0 = LS.fronts [Lv.first] ;
                    double *F0 = N.F + h->h_Foff [f0] ;
                    const I32 cm = std::min<I32> (cb + WB, Lv.maxfn) ;
                    const int buf = J & 1 ;
                    // block J+1 comes home with the reflectors of blocks 0 .. J-1 applied (its owner posted the send
                    // before it enters the broadcast below: same order on both sides)
                    const int own1 = coop_owner (h->coop, cb) ;
                    if (own1 != coop_me &&
                        (cs = coopT->recv (h, F0 + (I64) cb * actFm, (size_t) actFm * (size_t) (cm - cb) * sizeof (double), own1)) != STMQR_OK) return cs ;
                    // the block reflector of block J to everybody: V in the compact layout (actual rows), T, the row window
                    const I32 ldvC = coop_ldv (actFm, Lv.ldv) ;
                    const I64 need = (I64) ldvC * WB ;
                    if (need > h->coop.stage_cap)
                    {
                        if (h->coop.stage) { CK (cudaStreamSynchronize (st)) ; cudaFree (h->coop.stage) ; h->coop.stage = nullptr ; h->coop.stage_cap = 0 ; }
                        CK (cudaMalloc ((void **) &h->coop.stage, (size_t) need * sizeof (double))) ;
                        h->coop.stage_cap = need ;
                    }
                    CK (cudaMemcpy2DAsync (h->coop.stage, (size_t) ldvC * sizeof (double),
                        N.wVb + (I64) buf * nbig * ((I64) Lv.ldv * WB), (size_t) Lv.ldv * sizeof (double),
                        (size_t) ldvC * sizeof (double), WB, cudaMemcpyDeviceToDevice, st)) ;
                    if ((cs = coopT->group_begin ()) != STMQR_OK) return cs ;
                    if ((cs = coopT->bcast (h, h->coop.stage, (size_t) need * sizeof (double), coop_me)) != STMQR_OK) return cs ;
                    if ((cs = coopT->bcast (h, N.wTbt + (I64) buf * nbig * (WB * WB), (size_t) WB * WB * sizeof (double), coop_me)) != STMQR_OK) return cs ;
                    if ((cs = coopT->bcast (h, N.wblk + (I64) buf * nbig * 4, 4 * sizeof (I32), coop_me)) != STMQR_OK) return cs ;
                    if ((cs = coopT->group_end ()) != STMQR_OK) return cs ;
                    // V_J -> block J+1 here, then my own chunks beyond it on the second stream beside the next panels
                    if (pending [(J + 1) & 1]) { CK (cudaStreamWaitEvent (st, h->evN [(J + 1) & 1], 0)) ; pending [(J + 1) & 1] = false ; }
                    outer_update (st, act2, cb, cm) ;
                    std::vector<std::pair<I32, I32>> mine ;
                    coop_ranges (h->coop, coop_me, cm, mine) ;
                    if (!mine.empty ())
                    {
                        CK (cudaEventRecord (h->evP [J & 1], st)) ;
                        CK (cudaStreamWaitEvent (st2, h->evP [J & 1], 0)) ;
                        for (auto &r : mine) outer_update (st2, act2, r.first, r.second) ;
                        CK (cudaEventRecord (h->evN [J & 1], st2)) ;
                        pending [J & 1] = true ;
                    }
                }
                else if (lookahead)
                {
                    // the next block's columns first (main stream), the rest of the trailing matrix on the
                    // second stream while the next block's panels run
                    if (pending [(J + 1) & 1]) { CK (cudaStreamWaitEvent (st, h->evN [(J + 1) & 1], 0)) ; pending [(J + 1) & 1] = false ; }
                    const I32 cm = std::min<I32> (cb + WB, Lv.maxfn) ;
                    outer_update (st, act2, cb, cm) ;
                    if (cm < Lv.maxfn)
                    {
                        const I32 act3 = active_at (cm, act2) ;
                        if (act3 > 0)
                        {
                            CK (cudaEventRecord (h->evP [J & 1], st)) ;
                            CK (cudaStreamWaitEvent (st2, h->evP [J & 1], 0)) ;
                            outer_update (st2, act3, cm, Lv.maxfn) ;
                            CK (cudaEventRecord (h->evN [J & 1], st2)) ;
                            pending [J & 1] = true ;
                        }
                    }
                }
                else
                {
                    outer_update (st, act2, cb, Lv.maxfn) ;
                }
            }
            for (int b = 0 ; b < 2 ; b++)
                if (pending [b]) CK (cudaStreamWaitEvent (st, h->evN [b], 0)) ;
        }
        else
        {
            I32 active = active_at (0, nbig) ;
            bool rest_pending [2] = {false, false} ;        // a rest-of-the-trailing-matrix update in flight on stream2
            if (active > 0) { LAUNCH (3, CK (launch_panel (active, 0, 0))) ; }
            for (I32 j = 0 ; active > 0 ; j++)
            {
                const I32 k2 = (j + 1) * PB ;
                if (k2 >= Lv.maxfn) break ;
                h->curtag = (levelno << 32) | (long long) k2 ;
                const I32 active2 = active_at (k2, active) ;
                if (active2 == 0) break ;
                const I32 par = j & 1 ;
                if (lookahead)
                {
                    // The critical chain panel j -> update of the next panel's 32 columns -> panel j+1 stays on
                    // the main stream (no event hop between streams on it); the rest of the trailing update
                    // runs on the second stream beside panel j+1.  The rest update of step j also writes the
                    // columns of the narrow update of step j+1 and reads the panel buffers of parity j & 1 that
                    // panel j+2 overwrites: the main stream waits for it before the next narrow update.
                    if (k2 + PB < Lv.maxfn)
                    {
                        CK (cudaEventRecord (h->evP [par], st)) ;
                        CK (cudaStreamWaitEvent (st2, h->evP [par], 0)) ;
                        launch_update (st2, active2, k2 + PB, Lv.maxfn, par) ; h->launches++ ;
                        CK (cudaEventRecord (h->evN [par], st2)) ;
                        rest_pending [par] = true ;
                    }
                    if (rest_pending [par ^ 1]) { CK (cudaStreamWaitEvent (st, h->evN [par ^ 1], 0)) ; rest_pending [par ^ 1] = false ; }
                    launch_update (st, active2, k2, std::min<I32> (k2 + PB, Lv.maxfn), par) ; h->launches++ ;
                }
                else
                {
                    LAUNCH (4, launch_update (st, active2, k2, Lv.maxfn, par)) ;
                }
                LAUNCH (3, CK (launch_panel (active2, k2, par ^ 1))) ;
                active = active2 ;
            }
            if (lookahead)
            {
                CK (cudaEventRecord (h->evW, st2)) ;
                CK (cudaStreamWaitEvent (st, h->evW, 0)) ;
                (void) rest_pending ;
            }
        }
        if (h->debug_capture)
        {
            CK (cudaStreamSynchronize (st)) ;
            std::vector<I32> hm ((size_t) h->nf) ;
            CK (cudaMemcpy (hm.data (), N.Hm, h->nf * sizeof (I32), cudaMemcpyDeviceToHost)) ;
            for (I32 i = 0 ; i < nbig ; i++)
            {
                I32 f = LS.fronts [Lv.first + i] ;
                I64 cnt = (I64) hm [f] * (h->h_Rp [f+1] - h->h_Rp [f]) ;
                if (cnt > 0) CK (cudaMemcpy (h->d_capF + h->h_capOff [f], N.F + h->h_Foff [f],
                    cnt * sizeof (double), cudaMemcpyDeviceToDevice)) ;
            }
        }
        check ("front QR", 0, 0) ;
        LAUNCH (5, k_front_finish<<<nbig, fthreads, 0, st>>> (fr, S, N)) ;
        return STMQR_OK ;
        } ;
        if (nbig > 0) { const int sb = big_path () ; if (sb != STMQR_OK) return sb ; }
        // ---- the small fronts of the level: one warp per front, everything in shared memory -----------
        {
            I32 first = nbig ;
            for (int c = 0 ; c < 3 ; c++)
            {
                if (Lv.nsmall [c] == 0) continue ;
                const size_t smem = small_smem_bytes (Lv.scap [c], Lv.srows [c]) ;
                LAUNCH (3, k_front_small<<<Lv.nsmall [c], 32, smem, st>>> (fr + first, S, N, tol, ntol, Lv.scap [c], Lv.srows [c])) ;
                first += Lv.nsmall [c] ;
            }
        }
        LAUNCH (5, k_level_alloc<<<1, 1024, 0, st>>> (fr, Lv.count, N, h->Rcap)) ;
        LAUNCH (6, k_pack<<<dim3 (Lv.count, nsl), 256, 0, st>>> (fr, S, N)) ;
        mark_level ("level " + std::to_string (levelno) + ": " + std::to_string (Lv.count) + " fronts (" + std::to_string (nbig) +
            " tiled), max " + std::to_string (Lv.maxFm) + " x " + std::to_string (Lv.maxfn) + (Lv.wide ? " wide" : "")) ;
        if (h->streaming && h->stream_levels >= (int) h->evLvl.size ()) h->stream_overflow = true ;
        if (h->streaming && !h->stream_overflow)
        {
            // the level's R+H blocks are final: hand their slice of the arena to the downloader
            // (the events were all created in stream_begin: the downloader thread reads evLvl)
            const int li = h->stream_levels++ ;
            CK (cudaMemcpyAsync (h->pin_cursor + li, N.rcursor, sizeof (unsigned long long), cudaMemcpyDeviceToHost, st)) ;
            CK (cudaEventRecord (h->evLvl [li], st)) ;
            { std::lock_guard<std::mutex> lk (h->smu) ; h->squeue.push_back (li) ; }
            h->scv.notify_one () ;
        }
    }
    return STMQR_OK ;
}

int stmqr_b200_factorize_levels (stmqr_handle h, int part)
{
    if (!h || !h->analyzed || !h->have_matrix) return fail (h, STMQR_ERR_INVALID, "factorize_levels: not ready") ;
    if (part != 0 && h->nparts <= 1) return fail (h, STMQR_ERR_INVALID, "factorize_levels: no partition set") ;
    cudaSetDevice (h->device) ;
    const LevelSet &LS = (part == 0) ? h->ls_all : ((part == 1) ? h->ls_sub : h->ls_top) ;
    long long levelno = -1 ;
    for (const Level &Lv : LS.levels)
    {
        levelno++ ;
        const int s = run_level (h, LS, Lv, levelno) ;
        if (s != STMQR_OK) return s ;
    }
    return STMQR_OK ;
}

// qr_hpinv part 1: needs Hm, Hr, Cm of EVERY front on this device (merged over the GPUs)
int stmqr_b200_factorize_hpinv_a (stmqr_handle h)
{
    if (!h || !h->analyzed) return STMQR_ERR_INVALID ;
    cudaSetDevice (h->device) ;
    cudaStream_t st = h->stream ;
    DSym &S = h->S ; DNum &N = h->N ;
    if (h->nf > 0)
    {
        LAUNCH (7, k_hpinv_counts<<<grid_for (h->nf, 256, 1 << 20), 256, 0, st>>> (S, N)) ;
        LAUNCH (7, k_scan_i64<<<1, 1024, 0, st>>> (N.base1, N.base2, (I32) h->nf)) ;
        LAUNCH (7, k_hpinv_rows<<<grid_for (h->nf * 32, 256, 1 << 22), 256, 0, st>>> (S, N)) ;
    }
    LAUNCH (7, k_hpinv_empty<<<grid_for (std::max<I64> (h->m, 1), 256, 1 << 22), 256, 0, st>>> (S, N)) ;
    return STMQR_OK ;
}

// qr_hpinv part 2 (needs the row permutation W merged over the GPUs, and Rdead) + the scalars
int stmqr_b200_factorize_hpinv_b (stmqr_handle h, stmqr_numeric_info *info)
{
    if (!h || !h->analyzed) return STMQR_ERR_INVALID ;
    cudaSetDevice (h->device) ;
    cudaStream_t st = h->stream ;
    DSym &S = h->S ; DNum &N = h->N ;
    const I64 ntol = h->cur_ntol ;
    {
        LAUNCH (7, k_hpinv_apply<<<grid_for (std::max<I64> (std::max (h->m, h->nf * 32), 1), 256), 256, 0, st>>> (S, N, h->d_HPinv64, h->d_Hii64)) ;
        const I64 nt = std::min<I64> (std::max<I64> (ntol, 0), h->n) ;
        if (nt > 0) LAUNCH (7, k_rank1<<<(unsigned) ((nt + 255) / 256), 256, 0, st>>> (N.Rdead, nt, N.rank1)) ;
    }
    CK (cudaEventRecord (h->ev1, st)) ;

    // scalars back
    unsigned long long rcur = 0 ; I32 sc [4] = {0, 0, 0, 0} ; double fl3 [4] = {0, 0, 0, 0} ; I32 err = 0 ;
    CK (cudaMemcpyAsync (&rcur, N.rcursor, sizeof (rcur), cudaMemcpyDeviceToHost, st)) ;
    CK (cudaMemcpyAsync (sc, N.sumrank, sizeof (sc), cudaMemcpyDeviceToHost, st)) ;
    CK (cudaMemcpyAsync (fl3, N.flops, sizeof (fl3), cudaMemcpyDeviceToHost, st)) ;
    CK (cudaMemcpyAsync (&err, h->d_err, sizeof (err), cudaMemcpyDeviceToHost, st)) ;
    I32 gerr2 [2] = {0, 0} ;
    CK (cudaMemcpyAsync (gerr2, N.griderr, sizeof (gerr2), cudaMemcpyDeviceToHost, st)) ;
    CK (cudaStreamSynchronize (st)) ;
    CK (cudaGetLastError ()) ;
    if (err) return fail (h, STMQR_ERR_INVALID, "factorize: an entry of A is not in the pattern of S") ;
    if (gerr2 [0]) return fail (h, STMQR_ERR_INVALID, "factorize: a panel did not fit the shared-memory slabs of k_panel_grid") ;
    if (gerr2 [1]) return fail (h, STMQR_ERR_INVALID, "factorize: R+H arena bound exceeded (the packed blocks of the overflowing levels were not written)") ;
#ifdef STMQR_PANEL_TIMING
    {
        unsigned long long dbg [64] ;
        cudaMemcpy (dbg, N.dbg, sizeof (dbg), cudaMemcpyDeviceToHost) ;
        fprintf (stderr, "panel rare (rescale) path taken %llu times: ss==0 %llu, 0<ss<=1e-280 %llu, ss>=1e280 %llu, nan %llu, |alpha|<=1e-120 %llu, alpha==0 %llu, |alpha|>=1e140 %llu\n",
            dbg [63], dbg [62], dbg [61], dbg [60], dbg [59], dbg [58], dbg [57], dbg [56]) ;
        fprintf (stderr, "panel look-ahead: %llu steps took two columns, %llu candidates fell back to one\n", dbg [54], dbg [55]) ;
        const char *nm [8] = {"dots", "bar", "reduce+xchg", "scalar", "update", "endsync", "epilogue", "looptop"} ;
        for (int b = 0 ; b < 48 ; b += 8)
        {
            fprintf (stderr, "panel cycles [%s, %s]:", b % 24 == 0 ? "16 warps" : (b % 24 == 8 ? "8 warps" : "4 warps"),
                b >= 24 ? "cluster" : "single") ;
            for (int j = 0 ; j < 8 ; j++) fprintf (stderr, " %s=%.3fM", nm [j], dbg [b+j] * 1e-6) ;
            fprintf (stderr, "\n") ;
        }
    }
#endif
    if ((I64) rcur > h->Rcap) return fail (h, STMQR_ERR_INVALID, "factorize: R+H arena bound exceeded") ;
    if (h->evLevelT.size () > 1)
    {
        for (size_t i = 1 ; i < h->evLevelT.size () ; i++)
        {
            float t = 0 ;
            if (h->lvlNote [i] == "start") continue ;
            cudaEventElapsedTime (&t, h->evLevelT [i-1], h->evLevelT [i]) ;
            fprintf (stderr, "STMQR_B200_LEVEL_TIMES %8.3f ms  %s\n", t, h->lvlNote [i].c_str ()) ;
        }
    }
    float ms = 0 ;
    cudaEventElapsedTime (&ms, h->ev0, h->ev1) ;
    h->stats.ms_numeric = ms ;
    const double fl = fl3 [0] ;
    h->stats.launches = h->launches ;
    h->stats.flops = fl ;
    h->stats.update_flops = fl3 [1] ;
    h->stats.bytes_assemble = fl3 [2] ;
    for (int c = 0 ; c < 8 ; c++) { h->stats.ms_class [c] = 0 ; h->stats.launches_class [c] = 0 ; }
    if (h->opt.profile_phases)
    {
        // STMQR_B200_TRACE=<file>: one line per launch (class, etree level, first column, microseconds)
        const char *tracefile = getenv ("STMQR_B200_TRACE") ;
        FILE *trace = tracefile ? fopen (tracefile, "w") : nullptr ;
        if (trace) fprintf (trace, "class,level,k1,us\n") ;
        for (size_t e = 0 ; e < h->evclass.size () ; e++)
        {
            float t = 0 ;
            cudaEventElapsedTime (&t, h->evpool [2*e], h->evpool [2*e+1]) ;
            // classes >= 8 are the kernels of the two-level path (trace detail): 8 V extract (panel
            // class), 9-11 inner vtc/wt/apply, 12 Gram, 13 T merge, 14-16 outer vtc/wt/apply (update class)
            const int cl = (h->evclass [e] < 8) ? h->evclass [e] : ((h->evclass [e] == 8) ? 3 : 4) ;
            h->stats.ms_class [cl] += t ;
            h->stats.launches_class [cl] += 1 ;
            if (trace) fprintf (trace, "%d,%lld,%lld,%.4f\n", h->evclass [e], h->evtag [e] >> 32,
                h->evtag [e] & 0xffffffffLL, t * 1e3) ;
        }
        if (trace) fclose (trace) ;
        h->stats.ms_assemble = h->stats.ms_class [1] + h->stats.ms_class [2] + h->stats.ms_class [5] + h->stats.ms_class [6] ;
        h->stats.ms_front = h->stats.ms_class [3] + h->stats.ms_class [4] ;
    }
    h->info.rank = sc [0] ;
    h->info.maxfrank = std::max<I32> (1, sc [1]) ;       // maxfrank starts at 1 (:555)
    h->info.maxfm = sc [2] ;
    h->info.rank1 = (ntol >= h->n) ? sc [0] : sc [3] ;
    h->info.rh_size = (I64) rcur ;
    h->info.flops = fl ;
    if (info) *info = h->info ;
    h->factorized = true ;
    return STMQR_OK ;
}

int stmqr_b200_factorize_resident (stmqr_handle h, double tol, int64_t ntol, stmqr_numeric_info *info)
{
    int s ;
    if ((s = stmqr_b200_factorize_begin (h, tol, ntol)) != STMQR_OK) return s ;
    if ((s = stmqr_b200_factorize_levels (h, 0)) != STMQR_OK) return s ;
    if ((s = stmqr_b200_factorize_hpinv_a (h)) != STMQR_OK) return s ;
    return stmqr_b200_factorize_hpinv_b (h, info) ;
}

int stmqr_b200_sync (stmqr_handle h)
{
    if (!h) return STMQR_ERR_INVALID ;
    cudaSetDevice (h->device) ;
    CK (cudaStreamSynchronize (h->stream)) ;
    CK (cudaStreamSynchronize (h->stream2)) ;
    return STMQR_OK ;
}

int stmqr_b200_factorize (stmqr_handle h, const stmqr_csc_view *A, double tol, int64_t ntol,
    stmqr_numeric_info *info)
{
    if (!h) return STMQR_ERR_INVALID ;
    int s ;
    if (can_speculate (h, A))
    {
        // values first, the pattern beside the numeric phase (see issue_pattern_check)
        if ((s = stmqr_b200_upload_values (h, A->x, h->anz)) != STMQR_OK) return s ;
        if ((s = stmqr_b200_factorize_begin (h, tol, ntol)) != STMQR_OK) return s ;
        if ((s = stmqr_b200_factorize_levels (h, 0)) != STMQR_OK) return s ;
        if ((s = stmqr_b200_factorize_hpinv_a (h)) != STMQR_OK) return s ;
        if ((s = issue_pattern_check (h, A)) != STMQR_OK) return s ;
        s = stmqr_b200_factorize_hpinv_b (h, info) ;
        const int sv = finish_pattern_check (h) ;
        if (sv != STMQR_OK) return sv ;
        if (!h->pattern_mismatch) return s ;
    }
    s = stmqr_b200_upload_matrix (h, A) ;
    if (s != STMQR_OK) return s ;
    return stmqr_b200_factorize_resident (h, tol, ntol, info) ;
}

int stmqr_b200_refactorize_values (stmqr_handle h, const double *Ax, int64_t nnz, double tol, int64_t ntol,
    stmqr_numeric_info *info)
{
    int s = stmqr_b200_upload_values (h, Ax, nnz) ;
    if (s != STMQR_OK) return s ;
    return stmqr_b200_factorize_resident (h, tol, ntol, info) ;
}

int stmqr_b200_rh_bound (stmqr_handle h, int64_t *doubles)
{
    if (!h || !h->analyzed || !doubles) return STMQR_ERR_INVALID ;
    // (with an ownership set: only the blocks of this GPU's fronts land in its stack)
    *doubles = (h->nparts > 1 && h->Rcap_owned > 0) ? h->Rcap_owned : h->Rcap ;
    return STMQR_OK ;
}

int stmqr_b200_stream_begin (stmqr_handle h, double *stack, int64_t capacity)
{
    if (!h || !h->analyzed || !stack || capacity < 1) return fail (h, STMQR_ERR_INVALID, "stream_begin: no destination") ;
    if (h->streaming) return fail (h, STMQR_ERR_INVALID, "stream_begin: already streaming") ;
    cudaSetDevice (h->device) ;
    int s ;
    if ((s = ensure_copy_pipeline (h)) != STMQR_OK) return s ;
    if (!h->pin_cursor) CK (cudaHostAlloc ((void **) &h->pin_cursor, STREAM_MAX_LEVELS * sizeof (unsigned long long), cudaHostAllocDefault)) ;
    h->squeue.clear () ; h->sdone = false ; h->serr.store (STMQR_OK) ; h->stream_levels = 0 ;
    h->stream_overflow = false ;
    {
        // one event per level that can be streamed, created BEFORE the downloader thread exists
        const size_t need = std::min<size_t> ((size_t) STREAM_MAX_LEVELS,
            h->ls_all.levels.size () + h->ls_sub.levels.size () + h->ls_top.levels.size () + 1) ;
        while (h->evLvl.size () < need)
        {
            cudaEvent_t e ;
            CK (cudaEventCreateWithFlags (&e, cudaEventDisableTiming)) ;
            h->evLvl.push_back (e) ;
        }
    }
    h->stream_cap = capacity ; h->stream_dst = stack ;
    h->streaming = true ;
    h->sworker = new std::thread (stream_worker, h, stack, (I64) capacity) ;
    return STMQR_OK ;
}

int stmqr_b200_stream_end (stmqr_handle h)
{
    if (!h || !h->streaming) return fail (h, STMQR_ERR_INVALID, "stream_end: not streaming") ;
    cudaSetDevice (h->device) ;
    { std::lock_guard<std::mutex> lk (h->smu) ; h->sdone = true ; }
    h->scv.notify_all () ;
    h->sworker->join () ;
    delete h->sworker ; h->sworker = nullptr ;
    h->streaming = false ;
    if (h->serr.load () != STMQR_OK) return fail (h, h->serr.load (), "stream_end: download of the R+H stack failed") ;
    if (!h->factorized) return STMQR_OK ;              // the factorization itself failed: its status counts
    if (h->info.rh_size > h->stream_cap) return fail (h, STMQR_ERR_INVALID, "stream_end: stack capacity too small") ;
    if (h->stream_overflow)
    {
        // (more etree levels than events: the tail was not streamed) copy everything again
        const int s = d2h_pipelined (h, h->stream_dst, h->N.R, h->info.rh_size * sizeof (double)) ;
        if (s != STMQR_OK) return s ;
    }
    return STMQR_OK ;
}

int stmqr_b200_factorize_streamed (stmqr_handle h, const stmqr_csc_view *A, double tol, int64_t ntol,
    double *stack, int64_t capacity, stmqr_numeric_info *info)
{
    if (!h || !stack || capacity < 1) return fail (h, STMQR_ERR_INVALID, "factorize_streamed: no destination") ;
    int s ;
    if (can_speculate (h, A))
    {
        if ((s = stmqr_b200_upload_values (h, A->x, h->anz)) != STMQR_OK) return s ;
        if ((s = stmqr_b200_stream_begin (h, stack, capacity)) != STMQR_OK) return s ;
        if ((s = stmqr_b200_factorize_begin (h, tol, ntol)) == STMQR_OK &&
            (s = stmqr_b200_factorize_levels (h, 0)) == STMQR_OK &&
            (s = stmqr_b200_factorize_hpinv_a (h)) == STMQR_OK &&
            (s = issue_pattern_check (h, A)) == STMQR_OK)
            s = stmqr_b200_factorize_hpinv_b (h, info) ;
        const int sv = finish_pattern_check (h) ;
        const int s2 = stmqr_b200_stream_end (h) ;
        if (sv != STMQR_OK) return sv ;
        if (!h->pattern_mismatch) return (s != STMQR_OK) ? s : s2 ;
        // the pattern changed: everything streamed so far belongs to the wrong matrix; start over
    }
    s = stmqr_b200_upload_matrix (h, A) ;
    if (s != STMQR_OK) return s ;
    if ((s = stmqr_b200_stream_begin (h, stack, capacity)) != STMQR_OK) return s ;
    if ((s = stmqr_b200_factorize_begin (h, tol, ntol)) == STMQR_OK &&
        (s = stmqr_b200_factorize_levels (h, 0)) == STMQR_OK &&
        (s = stmqr_b200_factorize_hpinv_a (h)) == STMQR_OK)
        s = stmqr_b200_factorize_hpinv_b (h, info) ;
    const int s2 = stmqr_b200_stream_end (h) ;
    return (s != STMQR_OK) ? s : s2 ;
}

// -------------------------------------------------------------------------------------------------
int stmqr_b200_download (stmqr_handle h, const stmqr_numeric_view *out)
{
    if (!h || !out || !h->factorized) return fail (h, STMQR_ERR_INVALID, "download: factorize first") ;
    cudaSetDevice (h->device) ;
    cudaStream_t st = h->stream ;
    DNum &N = h->N ;
    { int s0 = ensure_copy_pipeline (h) ; if (s0 != STMQR_OK) return s0 ; }
    auto t0 = std::chrono::steady_clock::now () ;
    // widen the int32 device arrays once (HStair | Hm | Hr back to back in d_wide)
    const I64 nw1 = std::max<I64> (h->rjsize, 0), nw2 = std::max<I64> (h->nf, 0) ;
    if (nw1 > 0) k_widen<<<grid_for (nw1, 256), 256, 0, st>>> (N.stair, h->d_wide, nw1) ;
    if (nw2 > 0)
    {
        k_widen<<<grid_for (nw2, 256), 256, 0, st>>> (N.Hm, h->d_wide + nw1, nw2) ;
        k_widen<<<grid_for (nw2, 256), 256, 0, st>>> (N.Hr, h->d_wide + nw1 + nw2, nw2) ;
    }
    CK (cudaStreamSynchronize (st)) ;
    int s ;
    if (h->info.rh_size > 0 && out->stack && (s = d2h_pipelined (h, out->stack, N.R, h->info.rh_size * sizeof (double))) != STMQR_OK) return s ;
    if ((s = d2h_pipelined (h, out->Roff, N.Roff, h->nf * sizeof (I64))) != STMQR_OK) return s ;
    if ((s = d2h_pipelined (h, out->Rdead, N.Rdead, h->n)) != STMQR_OK) return s ;
    if ((s = d2h_pipelined (h, out->HTau, N.HTau, h->rjsize * sizeof (double))) != STMQR_OK) return s ;
    if ((s = d2h_pipelined (h, out->HPinv, h->d_HPinv64, h->m * sizeof (I64))) != STMQR_OK) return s ;
    if ((s = d2h_pipelined (h, out->Hii, h->d_Hii64, h->hisize * sizeof (I64))) != STMQR_OK) return s ;
    if ((s = d2h_pipelined (h, out->HStair, h->d_wide, nw1 * sizeof (I64))) != STMQR_OK) return s ;
    if ((s = d2h_pipelined (h, out->Hm, h->d_wide + nw1, nw2 * sizeof (I64))) != STMQR_OK) return s ;
    if ((s = d2h_pipelined (h, out->Hr, h->d_wide + nw1 + nw2, nw2 * sizeof (I64))) != STMQR_OK) return s ;
    h->stats.ms_d2h = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - t0).count () ;
    return STMQR_OK ;
}

// -------------------------------------------------------------------------------------------------
// multi-GPU: one handle per GPU, the etree partitioned over the handles
// -------------------------------------------------------------------------------------------------
int stmqr_b200_partition_fronts (const stmqr_symbolic_view *sym, int nparts, int32_t *owner, int32_t *is_top)
{
    if (!sym || !sym->Childp || !sym->Child || !sym->Rp || !sym->Fm) return STMQR_ERR_INVALID ;
    return partition_fronts (sym->nf, sym->Childp, sym->Child, sym->Super, sym->Rp, sym->Fm, nparts, owner, is_top) ;
}

int stmqr_b200_set_partition (stmqr_handle h, int nparts, int mypart, const int32_t *owner, const int32_t *is_top)
{
    if (!h || !h->analyzed || nparts < 1 || mypart < 0 || mypart >= nparts || !owner || !is_top)
        return fail (h, STMQR_ERR_INVALID, "set_partition: analyze first; 0 <= mypart < nparts") ;
    if (!h->host_only) cudaSetDevice (h->device) ;
    const I64 nf = h->nf ;
    h->nparts = nparts ; h->mypart = mypart ;
    h->h_owner.assign (owner, owner + nf) ; h->h_istop.assign (is_top, is_top + nf) ;
    std::vector<unsigned char> owned ((size_t) std::max<I64> (nf, 1), 0), sub ((size_t) std::max<I64> (nf, 1), 0),
        top ((size_t) std::max<I64> (nf, 1), 0) ;
    for (I64 f = 0 ; f < nf ; f++)
    {
        if (is_top [f] && owner [f] != 0) return fail (h, STMQR_ERR_INVALID, "set_partition: top fronts belong to part 0") ;
        if (is_top [f] && h->h_parent [f] >= 0 && !is_top [h->h_parent [f]])
            return fail (h, STMQR_ERR_INVALID, "set_partition: the top of the tree must be closed upwards") ;
        owned [f] = (owner [f] == mypart) ;
        sub [f] = owned [f] && !is_top [f] ;
        top [f] = owned [f] && is_top [f] ;
    }
    filter_levels (h->ls_all, sub, h->h_Rp, h->h_FmB, h->small_cap_used, h->wide_rows, h->ls_sub) ;
    filter_levels (h->ls_all, top, h->h_Rp, h->h_FmB, h->small_cap_used, h->wide_rows, h->ls_top) ;
    UPLOAD (h->ls_sub.d_fronts, h->ls_sub.fronts) ;
    UPLOAD (h->ls_top.d_fronts, h->ls_top.fronts) ;
    if (nparts > 1) { UPLOAD (h->d_owned, owned) ; h->N.owned = h->d_owned ; }
    else { h->d_owned = nullptr ; h->N.owned = nullptr ; }
    if (!(h->opt.reserved & 64))
    {
        // contribution-block offsets for THIS GPU's schedule: its subtrees level by level, then (part 0) the
        // blocks of the cut children owned by other GPUs arrive, then the top of the tree level by level
        std::vector<I32> received ;
        if (mypart == 0)
            for (I64 f = 0 ; f < nf ; f++)
                if (is_top [f])
                    for (I32 q = h->h_Childp [f] ; q < h->h_Childp [f+1] ; q++)
                    {
                        const I32 c = h->h_Child [q] ;
                        if (!is_top [c] && owner [c] != 0) received.push_back (c) ;
                    }
        std::vector<const LevelSet *> phases {&h->ls_sub, &h->ls_top} ;
        I64 cap = 0 ;
        if (nparts == 1) { phases.assign (1, &h->ls_all) ; }
        plan_contribution_arena (phases, (nparts == 1) ? phases.size () : 1, received, h->h_Childp, h->h_Child,
            h->h_Csize, h->h_Coff, cap) ;
        h->Ccap = cap ;
        if (h->host_only)
        {
            h->device_bytes += (size_t) std::max<I64> (0, cap - h->C_alloc) * sizeof (double) ;
            h->C_alloc = std::max (h->C_alloc, cap) ;
        }
        else if (cap > h->C_alloc)
        {
            // (another schedule can have a higher high-water mark than the one-GPU schedule)
            for (auto &pp : h->allocs) if (pp == (void *) h->N.C) { cudaFree (pp) ; pp = nullptr ; }
            h->device_bytes -= (size_t) h->C_alloc * sizeof (double) ;
            h->N.C = nullptr ;
            double *nc = nullptr ;
            if (cudaMalloc ((void **) &nc, (size_t) cap * sizeof (double)) != cudaSuccess)
                return fail (h, STMQR_ERR_OUT_OF_MEMORY, "set_partition: contribution-block arena") ;
            for (auto &pp : h->allocs) if (pp == nullptr) { pp = (void *) nc ; break ; }
            h->N.C = nc ; h->C_alloc = cap ;
            h->device_bytes += (size_t) cap * sizeof (double) ;
        }
        if (!h->host_only) CK (cudaMemcpyAsync (h->d_Coff, h->h_Coff.data (), (size_t) nf * sizeof (I64), cudaMemcpyHostToDevice, h->stream)) ;
    }
    if (!h->host_only) CK (cudaStreamSynchronize (h->stream)) ;
    return STMQR_OK ;
}

// ---- general ownership: every front on its owner GPU, blocks move after the child's level ---------------
int stmqr_b200_map_fronts (const stmqr_symbolic_view *sym, int nparts, int32_t *owner)
{
    if (!sym || !sym->Childp || !sym->Child || !sym->Rp || !sym->Fm || !owner || nparts < 1) return STMQR_ERR_INVALID ;
    return map_fronts (sym->nf, sym->Childp, sym->Child, sym->Super, sym->Rp, sym->Fm, nparts, owner) ;
}

int stmqr_b200_set_ownership (stmqr_handle h, int nparts, int mypart, const int32_t *owner)
{
    if (!h || !h->analyzed || nparts < 1 || mypart < 0 || mypart >= nparts || !owner)
        return fail (h, STMQR_ERR_INVALID, "set_ownership: analyze first; 0 <= mypart < nparts") ;
    if (!h->host_only) cudaSetDevice (h->device) ;
    const I64 nf = h->nf ;
    h->nparts = nparts ; h->mypart = mypart ;
    h->h_owner.assign (owner, owner + nf) ;
    h->h_istop.assign ((size_t) nf, 0) ;
    std::vector<unsigned char> owned ((size_t) std::max<I64> (nf, 1), 0) ;
    for (I64 f = 0 ; f < nf ; f++)
    {
        if (owner [f] < 0 || owner [f] >= nparts) return fail (h, STMQR_ERR_INVALID, "set_ownership: owner out of range") ;
        owned [f] = (owner [f] == mypart) ;
    }
    filter_levels (h->ls_all, owned, h->h_Rp, h->h_FmB, h->small_cap_used, h->wide_rows, h->ls_mine) ;
    h->ls_sub = LevelSet () ; h->ls_top = LevelSet () ;
    h->Rcap_owned = 16 ;
    for (I64 f = 0 ; f < nf ; f++) if (owned [f]) h->Rcap_owned += h->h_Rbound [(size_t) f] ;
    // the blocks that move: child and parent on different GPUs, after the child's level, sorted by front
    const I64 nlev = (I64) h->ls_all.levels.size () ;
    h->xedges.assign ((size_t) nlev, std::vector<int> ()) ;
    h->xall_c.clear () ; h->xall_src.clear () ; h->xall_dst.clear () ;
    for (I64 c = 0 ; c < nf ; c++)
    {
        const I32 p = h->h_parent [c] ;
        if (p < 0 || owner [p] == owner [c]) continue ;
        h->xedges [h->h_level [c]].push_back ((int) h->xall_c.size ()) ;
        h->xall_c.push_back ((I32) c) ; h->xall_src.push_back (owner [c]) ; h->xall_dst.push_back (owner [p]) ;
    }
    // contribution-block arena for this GPU's schedule: per etree level, the children of my fronts are
    // consumed, my fronts' blocks are packed, then the blocks arriving from other GPUs are placed.  (A block
    // whose parent lives elsewhere is kept: the copy that takes it away is asynchronous.)
    if (!(h->opt.reserved & 64))
    {
        ArenaReplay A ;
        std::vector<unsigned char> live ((size_t) std::max<I64> (nf, 1), 0) ;
        std::vector<std::pair<I64, I32>> want ;
        std::fill (h->h_Coff.begin (), h->h_Coff.end (), (I64) 0) ;
        size_t li = 0 ;
        for (I64 gl = 0 ; gl < nlev ; gl++)
        {
            if (li < h->ls_mine.levels.size () && h->ls_mine.levels [li].glevel == gl)
            {
                const Level &Lv = h->ls_mine.levels [li++] ;
                for (I32 i = 0 ; i < Lv.count ; i++)
                {
                    const I32 f = h->ls_mine.fronts [Lv.first + i] ;
                    for (I32 q = h->h_Childp [f] ; q < h->h_Childp [f+1] ; q++)
                    {
                        const I32 c = h->h_Child [q] ;
                        if (live [c]) { A.release (h->h_Coff [c], h->h_Csize [c]) ; live [c] = 0 ; }
                    }
                }
                want.clear () ;
                for (I32 i = 0 ; i < Lv.count ; i++)
                {
                    const I32 f = h->ls_mine.fronts [Lv.first + i] ;
                    if (h->h_Csize [f] > 0) { want.emplace_back (h->h_Csize [f], f) ; live [f] = 1 ; }
                }
                A.alloc_level (want, h->h_Coff) ;
            }
            want.clear () ;
            for (int e : h->xedges [(size_t) gl])
                if (h->xall_dst [e] == mypart)
                {
                    const I32 c = h->xall_c [e] ;
                    if (h->h_Csize [c] > 0) { want.emplace_back (h->h_Csize [c], c) ; live [c] = 1 ; }
                }
            A.alloc_level (want, h->h_Coff) ;
            // (blocks sent away stay allocated: live[] is only cleared by a local parent)
            for (int e : h->xedges [(size_t) gl]) if (h->xall_src [e] == mypart) live [h->xall_c [e]] = 0 ;
        }
        const I64 cap = A.peak + 2 ;
        h->Ccap = cap ;
        if (h->host_only)
        {
            h->device_bytes += (size_t) std::max<I64> (0, cap - h->C_alloc) * sizeof (double) ;
            h->C_alloc = std::max (h->C_alloc, cap) ;
        }
        else if (cap > h->C_alloc)
        {
            for (auto &pp : h->allocs) if (pp == (void *) h->N.C) { cudaFree (pp) ; pp = nullptr ; }
            h->device_bytes -= (size_t) h->C_alloc * sizeof (double) ;
            h->N.C = nullptr ;
            double *nc = nullptr ;
            if (cudaMalloc ((void **) &nc, (size_t) cap * sizeof (double)) != cudaSuccess)
                return fail (h, STMQR_ERR_OUT_OF_MEMORY, "set_ownership: contribution-block arena") ;
            for (auto &pp : h->allocs) if (pp == nullptr) { pp = (void *) nc ; break ; }
            h->N.C = nc ; h->C_alloc = cap ;
            h->device_bytes += (size_t) cap * sizeof (double) ;
        }
        if (!h->host_only) CK (cudaMemcpyAsync (h->d_Coff, h->h_Coff.data (), (size_t) nf * sizeof (I64), cudaMemcpyHostToDevice, h->stream)) ;
    }
    UPLOAD (h->ls_mine.d_fronts, h->ls_mine.fronts) ;
    if (nparts > 1) { UPLOAD (h->d_owned, owned) ; h->N.owned = h->d_owned ; }
    else { h->d_owned = nullptr ; h->N.owned = nullptr ; }
    if (!h->host_only) CK (cudaStreamSynchronize (h->stream)) ;
    return STMQR_OK ;
}

int stmqr_b200_nccl_unique_id (void *id128)
{
    if (!id128) return STMQR_ERR_INVALID ;
    if (!g_nccl.load ()) return STMQR_ERR_INVALID ;
    ncclUniqueId id ;
    if (g_nccl.GetUniqueId (&id) != ncclSuccess) return STMQR_ERR_CUDA ;
    static_assert (sizeof (ncclUniqueId) == 128, "ncclUniqueId") ;
    memcpy (id128, &id, 128) ;
    return STMQR_OK ;
}

int stmqr_b200_comm_init (stmqr_handle h, int nranks, int rank, const void *id128)
{
    if (!h || h->host_only || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return STMQR_ERR_INVALID ;
    if (!g_nccl.load ()) return fail (h, STMQR_ERR_INVALID, "comm_init: libnccl.so.2 not found") ;
    cudaSetDevice (h->device) ;
    if (h->transport) { delete (Transport *) h->transport ; h->transport = nullptr ; }
    NcclTransport *t = new NcclTransport ;
    ncclUniqueId id ;
    memcpy (&id, id128, 128) ;
    t->nr = nranks ; t->me = rank ;
    const ncclResult_t r = g_nccl.CommInitRank (&t->comm, nranks, id, rank) ;
    if (r != ncclSuccess)
    {
        const std::string msg = std::string ("ncclCommInitRank: ") + g_nccl.GetErrorString (r) ;
        t->comm = nullptr ; delete t ;
        return fail (h, STMQR_ERR_CUDA, msg) ;
    }
    h->transport = t ;
    return STMQR_OK ;
}

// N handles of this process as one group (peer copies; the handles may sit on different devices or, for
// tests, on the same one).  The group owns nothing but the synchronisation state.
int stmqr_b200_peer_group_create (stmqr_handle *hs, int n, void **group)
{
    if (!hs || n < 1 || !group) return STMQR_ERR_INVALID ;
    PeerGroup *g = new PeerGroup ;
    for (int i = 0 ; i < n ; i++)
    {
        stmqr_handle h = hs [i] ;
        if (!h || h->host_only) { delete g ; return STMQR_ERR_INVALID ; }
        g->hs.push_back (h) ;
    }
    for (int i = 0 ; i < n ; i++)
    {
        stmqr_handle h = hs [i] ;
        cudaSetDevice (h->device) ;
        for (int j = 0 ; j < n ; j++)
            if (hs [j]->device != h->device)
            {
                int can = 0 ;
                cudaDeviceCanAccessPeer (&can, h->device, hs [j]->device) ;
                if (can) { cudaError_t e = cudaDeviceEnablePeerAccess (hs [j]->device, 0) ; if (e != cudaSuccess) cudaGetLastError () ; }
            }
        if (h->transport) delete (Transport *) h->transport ;
        PeerTransport *t = new PeerTransport ;
        t->grp = g ; t->me = i ;
        if (cudaEventCreateWithFlags (&t->evReady, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags (&t->evDone, cudaEventDisableTiming) != cudaSuccess) { delete t ; delete g ; return STMQR_ERR_CUDA ; }
        h->transport = t ;
    }
    *group = g ;
    return STMQR_OK ;
}

void stmqr_b200_peer_group_destroy (void *group)
{
    PeerGroup *g = (PeerGroup *) group ;
    if (!g) return ;
    for (stmqr_handle h : g->hs) if (h->transport) { cudaSetDevice (h->device) ; delete (Transport *) h->transport ; h->transport = nullptr ; }
    delete g ;
}

// The numeric phase of ONE GPU of a group (set_ownership + a transport first); every GPU of the group calls
// it at the same time (one process per GPU with NCCL, or one host thread per handle with peer copies).
int stmqr_b200_factorize_dist (stmqr_handle h, double tol, int64_t ntol, stmqr_numeric_info *info)
{
    if (!h || !h->analyzed || !h->have_matrix) return fail (h, STMQR_ERR_INVALID, "factorize_dist: analyze and upload_matrix first") ;
    Transport *T = (Transport *) h->transport ;
    if (!T || T->nranks () != h->nparts || T->rank () != h->mypart || h->xedges.size () != h->ls_all.levels.size ())
        return fail (h, STMQR_ERR_INVALID, "factorize_dist: set_ownership and a transport of the same size first") ;
    int s ;
    auto tfail = [&] (int code) { return T->error ().empty () ? code : fail (h, code, "factorize_dist: " + T->error ()) ; } ;
    if ((s = stmqr_b200_factorize_begin (h, tol, ntol)) != STMQR_OK) return s ;
    cudaSetDevice (h->device) ;
    DNum &N = h->N ;
    const I64 nlev = (I64) h->ls_all.levels.size () ;
    size_t li = 0 ;
    std::vector<XEdge> edges ;
    // STMQR_B200_DEBUG_SYNC=1: synchronise and check after every stage (names the stage of an asynchronous fault)
    static const bool dbg_sync = getenv ("STMQR_B200_DEBUG_SYNC") != nullptr ;
    auto stage = [&] (const char *what, I64 gl) -> int {
        if (!dbg_sync) return STMQR_OK ;
        cudaError_t e1 = cudaStreamSynchronize (h->stream), e2 = cudaStreamSynchronize (h->stream2), e3 = cudaGetLastError () ;
        const cudaError_t e = (e1 != cudaSuccess) ? e1 : ((e2 != cudaSuccess) ? e2 : e3) ;
        if (e == cudaSuccess) return STMQR_OK ;
        return fail (h, STMQR_ERR_CUDA, std::string ("factorize_dist: part ") + std::to_string (h->mypart) + " after " + what +
            " of level " + std::to_string (gl) + ": " + cudaGetErrorString (e)) ;
    } ;
    if ((s = stage ("begin", -1)) != STMQR_OK) return s ;
    h->coop.levels_run = 0 ;
    for (I64 gl = 0 ; gl < nlev ; gl++)
    {
        // a level that consists of ONE large front may be factorized cooperatively (static test: the same on
        // every GPU; the home GPU confirms with the actual # rows, see run_level / coop_peer_level)
        const Level &Lg = h->ls_all.levels [(size_t) gl] ;
        const bool coop_lvl = h->coop.enabled && T->cooperative () && h->nparts >= ((h->coop.enabled >= 2) ? 2 : h->coop.min_parts) &&
            Lg.count == 1 && Lg.nbig == 1 && Lg.wide ;
        h->coop.armed = false ; h->coop.on = false ;
        if (coop_lvl)
        {
            const I32 cf = h->ls_all.fronts [Lg.first] ;
            coop_plan (h->coop, h->nparts, h->h_owner [cf], cf, h->h_Rp [cf+1] - h->h_Rp [cf]) ;
            h->coop.armed = true ;
        }
        if (li < h->ls_mine.levels.size () && h->ls_mine.levels [li].glevel == gl)
        {
            if ((s = run_level (h, h->ls_mine, h->ls_mine.levels [li], gl)) != STMQR_OK) return tfail (s) ;
            li++ ;
            if ((s = stage ("the kernels", gl)) != STMQR_OK) return s ;
        }
        else if (coop_lvl)
        {
            if ((s = coop_peer_level (h, gl)) != STMQR_OK) return tfail (s) ;
            if ((s = stage ("the cooperative updates", gl)) != STMQR_OK) return s ;
        }
        h->coop.armed = false ; h->coop.on = false ;
        if (h->xedges [(size_t) gl].empty ()) continue ;
        edges.clear () ;
        for (int e : h->xedges [(size_t) gl]) edges.push_back (XEdge {h->xall_c [e], h->xall_src [e], h->xall_dst [e]}) ;
        if ((s = T->exchange (h, edges, (I32) gl)) != STMQR_OK) return tfail (s) ;
        if ((s = stage ("the exchange", gl)) != STMQR_OK) return s ;
    }
    // merge the integer side outputs (every entry has one writer, the others hold the neutral element)
    if ((s = T->allreduce (h, N.Hm, 3 * std::max<I64> (h->nf, 1), X_I32, true)) != STMQR_OK) return tfail (s) ;
    if ((s = T->allreduce (h, N.Rdead, h->n, X_I8, true)) != STMQR_OK) return tfail (s) ;
    if ((s = stmqr_b200_factorize_hpinv_a (h)) != STMQR_OK) return s ;
    if ((s = T->allreduce (h, N.W, h->m, X_I32, true)) != STMQR_OK) return tfail (s) ;
    if ((s = T->allreduce (h, N.sumrank, 1, X_I32, false)) != STMQR_OK) return tfail (s) ;
    if ((s = T->allreduce (h, N.sumrank + 1, 2, X_I32, true)) != STMQR_OK) return tfail (s) ;
    if ((s = T->allreduce (h, N.flops, 3, X_F64, false)) != STMQR_OK) return tfail (s) ;
    static const bool coop_verbose = getenv ("STMQR_B200_COOP_VERBOSE") != nullptr ;
    if (coop_verbose) fprintf (stderr, "stmqr_b200 factorize_dist: part %d of %d, %lld cooperative level(s)\n", h->mypart, h->nparts, (long long) h->coop.levels_run) ;
    return stmqr_b200_factorize_hpinv_b (h, info) ;
}

// After factorize_dist: make the per-front outputs (HStair, HTau, the permuted Hii, the offset of every packed
// block in its owner's stack) global on every GPU of the group, so that ONE of them can hand the whole integer
// side of the qr_numeric to the host.  Collective (every GPU of the group calls it).
int stmqr_b200_gather_outputs (stmqr_handle h)
{
    if (!h || !h->factorized) return fail (h, STMQR_ERR_INVALID, "gather_outputs: factorize_dist first") ;
    Transport *T = (Transport *) h->transport ;
    if (!T) return fail (h, STMQR_ERR_INVALID, "gather_outputs: no transport") ;
    cudaSetDevice (h->device) ;
    int s ;
    if ((s = T->allreduce (h, h->N.stair, h->rjsize, X_I32, true)) != STMQR_OK) return s ;
    if ((s = T->allreduce (h, h->N.HTau, h->rjsize, X_F64, false)) != STMQR_OK) return s ;
    if ((s = T->allreduce (h, h->N.Roff, h->nf, X_I64, true)) != STMQR_OK) return s ;
    if ((s = T->allreduce (h, h->d_Hii64, h->hisize, X_I64, true)) != STMQR_OK) return s ;
    CK (cudaStreamSynchronize (h->stream)) ;
    return STMQR_OK ;
}

int stmqr_b200_coop_chunks (int nparts, int home, int64_t fn, int32_t *chunk_owner, int64_t *nchunks)
{
    if (nparts < 1 || home < 0 || home >= nparts || fn < 0 || fn > INT32_MAX - 2 * COOP_CHUNK) return STMQR_ERR_INVALID ;
    stmqr_handle_s::Coop C ;
    coop_plan (C, nparts, home, 0, (I32) fn) ;
    const int64_t nch = (fn + COOP_CHUNK - 1) / COOP_CHUNK ;
    if (nchunks) *nchunks = nch ;
    if (chunk_owner) for (int64_t c = 0 ; c < nch ; c++) chunk_owner [c] = C.chunk_owner [(size_t) c] ;
    return STMQR_OK ;
}

// All handles of a peer group, one host thread each.
int stmqr_b200_factorize_multi (void *group, double tol, int64_t ntol, stmqr_numeric_info *infos)
{
    return stmqr_b200_factorize_multi_ex (group, tol, ntol, infos, 0) ;
}

int stmqr_b200_factorize_multi_ex (void *group, double tol, int64_t ntol, stmqr_numeric_info *infos, int gather)
{
    PeerGroup *g = (PeerGroup *) group ;
    if (!g || g->hs.empty ()) return STMQR_ERR_INVALID ;
    const int n = (int) g->hs.size () ;
    { std::lock_guard<std::mutex> lk (g->mu) ; g->failed = 0 ; g->waiting = 0 ; }
    std::vector<int> st ((size_t) n, STMQR_OK) ;
    std::vector<std::thread> th ;
    for (int i = 0 ; i < n ; i++)
        th.emplace_back ([&, i] {
            st [i] = stmqr_b200_factorize_dist (g->hs [i], tol, ntol, infos ? infos + i : nullptr) ;
            if (st [i] == STMQR_OK && gather) st [i] = stmqr_b200_gather_outputs (g->hs [i]) ;
            if (st [i] != STMQR_OK) g->abort () ;       // the others must not wait for this handle at a barrier
        }) ;
    for (auto &t : th) t.join () ;
    for (int i = 0 ; i < n ; i++) if (st [i] != STMQR_OK) return st [i] ;
    return STMQR_OK ;
}

// ---- host-only planner: the plan (level schedule, arenas, partition) without a device ------------------
int stmqr_b200_create_planner (stmqr_handle *out)
{
    if (!out) return STMQR_ERR_INVALID ;
    stmqr_handle h = new stmqr_handle_s ;
    h->host_only = true ;
    if (const char *e = getenv ("STMQR_B200_GRID_ROWS")) h->grid_rows = std::max (256, atoi (e)) ;
    if (const char *e = getenv ("STMQR_B200_WIDE_ROWS")) h->wide_rows = std::max (256, atoi (e)) ;
    if (const char *e = getenv ("STMQR_B200_SMALL_ELEMS")) h->small_cap = std::max (0, std::min (5600, atoi (e))) ;
    *out = h ;
    return STMQR_OK ;
}

int stmqr_b200_plan_info (stmqr_handle h, stmqr_plan_info *out, int64_t *Coff, int64_t *Csize, int32_t *level)
{
    if (!h || !h->analyzed || !out) return STMQR_ERR_INVALID ;
    out->nlevels = (int64_t) h->ls_all.levels.size () ;
    out->F_doubles = h->Fcap ; out->C_doubles = h->Ccap ; out->C_doubles_unrecycled = h->Ccap_all ;
    out->R_doubles = h->Rcap ; out->device_bytes = (int64_t) h->device_bytes ;
    out->nparts = h->nparts ; out->mypart = h->mypart ;
    for (I64 f = 0 ; f < h->nf ; f++)
    {
        if (Coff) Coff [f] = h->h_Coff [f] ;
        if (Csize) Csize [f] = h->h_Csize [f] ;
    }
    if (level)
    {
        I32 l = 0 ;
        for (const Level &Lv : h->ls_all.levels)
        {
            for (I32 i = 0 ; i < Lv.count ; i++) level [h->ls_all.fronts [Lv.first + i]] = l ;
            l++ ;
        }
    }
    return STMQR_OK ;
}

int stmqr_b200_device_array (stmqr_handle h, int which, void **ptr, int64_t *count, int32_t *elem_bytes)
{
    if (!h || !h->analyzed || !ptr || !count || !elem_bytes) return STMQR_ERR_INVALID ;
    DNum &N = h->N ;
    switch (which)
    {
        case STMQR_ARRAY_HM:    *ptr = N.Hm ;    *count = h->nf ; *elem_bytes = 4 ; break ;
        case STMQR_ARRAY_HR:    *ptr = N.Hr ;    *count = h->nf ; *elem_bytes = 4 ; break ;
        case STMQR_ARRAY_CM:    *ptr = N.Cm ;    *count = h->nf ; *elem_bytes = 4 ; break ;
        case STMQR_ARRAY_RDEAD: *ptr = N.Rdead ; *count = h->n ;  *elem_bytes = 1 ; break ;
        case STMQR_ARRAY_W:     *ptr = N.W ;     *count = h->m ;  *elem_bytes = 4 ; break ;
        default: return fail (h, STMQR_ERR_INVALID, "device_array: unknown array") ;
    }
    return STMQR_OK ;
}

// what a parent on another GPU needs of front f: its packed contribution block and the row ids
// of the block's rows.  cm, hr < 0: read Cm[f], Hr[f] from this device (the owner, after a sync);
// otherwise they are the owner's values and are stored on this device (the receiver).
int stmqr_b200_front_regions (stmqr_handle h, int64_t f, int64_t cm, int64_t hr, int64_t hm,
    stmqr_front_regions *out)
{
    if (!h || !h->analyzed || !out || f < 0 || f >= h->nf) return STMQR_ERR_INVALID ;
    cudaSetDevice (h->device) ;
    DNum &N = h->N ;
    I32 v [3] ;
    if (cm < 0 || hr < 0)
    {
        CK (cudaMemcpy (&v [0], N.Cm + f, sizeof (I32), cudaMemcpyDeviceToHost)) ;
        CK (cudaMemcpy (&v [1], N.Hr + f, sizeof (I32), cudaMemcpyDeviceToHost)) ;
        CK (cudaMemcpy (&v [2], N.Hm + f, sizeof (I32), cudaMemcpyDeviceToHost)) ;
    }
    else
    {
        v [0] = (I32) cm ; v [1] = (I32) hr ; v [2] = (I32) hm ;
        CK (cudaMemcpy (N.Cm + f, &v [0], sizeof (I32), cudaMemcpyHostToDevice)) ;
        CK (cudaMemcpy (N.Hr + f, &v [1], sizeof (I32), cudaMemcpyHostToDevice)) ;
        CK (cudaMemcpy (N.Hm + f, &v [2], sizeof (I32), cudaMemcpyHostToDevice)) ;
    }
    const I64 fp = h->h_Super [f+1] - h->h_Super [f], fn = h->h_Rp [f+1] - h->h_Rp [f], cn = fn - fp ;
    const I64 c = v [0] ;
    out->cm = c ; out->hr = v [1] ; out->hm = v [2] ;
    out->C = N.C + h->h_Coff [f] ;
    out->C_doubles = (c * (c + 1)) / 2 + c * (cn - c) ;
    out->Hii = N.Hii + h->h_Hip [f] + v [1] ;
    out->Hii_ints = c ;
    return STMQR_OK ;
}

// -------------------------------------------------------------------------------------------------
// Q-apply and R-solve on the resident factorization (kernels_solve.cuh)
// -------------------------------------------------------------------------------------------------
namespace {

int solve_ready (stmqr_handle h, const char *what)
{
    if (!h || !h->factorized || h->host_only) return fail (h, STMQR_ERR_INVALID, std::string (what) + ": factorize first") ;
    if (h->nparts > 1) return fail (h, STMQR_ERR_INVALID, std::string (what) + ": needs the whole factorization on one GPU") ;
    cudaSetDevice (h->device) ;
    if (!h->htable_valid)
    {
        if (h->nf > 0) k_htable<<<grid_for (h->nf * 32, 256, 1 << 22), 256, 0, h->stream>>> (h->S, h->N, h->d_hcol, h->d_nh, h->d_rlen) ;
        CK (cudaGetLastError ()) ;
        h->htable_valid = true ;
    }
    return STMQR_OK ;
}

// grows a device scratch array (outside the plan's allocation list: it survives until destroy / re-analyze)
int grow (stmqr_handle h, double **p, I64 *cap, I64 need)
{
    if (need <= *cap) return STMQR_OK ;
    if (*p) { for (auto &q : h->allocs) if (q == (void *) *p) q = nullptr ; cudaFree (*p) ; h->device_bytes -= (size_t) *cap * sizeof (double) ; *p = nullptr ; *cap = 0 ; }
    ALLOC (*p, need) ;
    *cap = need ;
    return STMQR_OK ;
}

// Z <- Q'Z or QZ, Z m-by-nx on the device in the factorization's row order
int device_qapply (stmqr_handle h, int method, I64 nx, double *Z)
{
    const LevelSet &LS = h->ls_all ;
    const I64 nl = (I64) LS.levels.size () ;
    for (I64 li = 0 ; li < nl ; li++)
    {
        const Level &Lv = LS.levels [(size_t) ((method == 0) ? li : (nl - 1 - li))] ;
        const I32 *fr = LS.d_fronts + Lv.first ;
        const I32 nsm = Lv.count - Lv.nbig ;
        for (I64 c0 = 0 ; c0 < nx ; )
        {
            const bool four = (nx - c0 >= 4) ;
            if (Lv.nbig > 0)
            {
                if (four) k_qapply<256, 4><<<Lv.nbig, 256, 0, h->stream>>> (fr, Lv.nbig, h->S, h->N, h->d_Hii64, h->d_hcol, h->d_nh, method, (I32) c0, Z) ;
                else k_qapply<256, 1><<<Lv.nbig, 256, 0, h->stream>>> (fr, Lv.nbig, h->S, h->N, h->d_Hii64, h->d_hcol, h->d_nh, method, (I32) c0, Z) ;
            }
            if (nsm > 0)
            {
                const int g = (nsm + 7) / 8 ;
                if (four) k_qapply<32, 4><<<g, 256, 0, h->stream>>> (fr + Lv.nbig, nsm, h->S, h->N, h->d_Hii64, h->d_hcol, h->d_nh, method, (I32) c0, Z) ;
                else k_qapply<32, 1><<<g, 256, 0, h->stream>>> (fr + Lv.nbig, nsm, h->S, h->N, h->d_Hii64, h->d_hcol, h->d_nh, method, (I32) c0, Z) ;
            }
            c0 += four ? 4 : 1 ;
        }
    }
    CK (cudaGetLastError ()) ;
    return STMQR_OK ;
}

// X <- R \ B or E (R \ B); B m-by-nrhs (device), X n-by-nrhs (device, overwritten), W n-by-nrhs scratch
int device_rsolve (stmqr_handle h, int use_Qfill, I64 nrhs, const double *B, double *X, double *W)
{
    CK (cudaMemsetAsync (X, 0, (size_t) std::max<I64> (h->n * nrhs, 1) * sizeof (double), h->stream)) ;
    const LevelSet &LS = h->ls_all ;
    const I32 *Qf = use_Qfill ? h->S.Qfill : nullptr ;
    for (I64 li = (I64) LS.levels.size () - 1 ; li >= 0 ; li--)
    {
        const Level &Lv = LS.levels [(size_t) li] ;
        const I32 *fr = LS.d_fronts + Lv.first ;
        const I32 nsm = Lv.count - Lv.nbig ;
        for (I64 c0 = 0 ; c0 < nrhs ; )
        {
            const bool four = (nrhs - c0 >= 4) ;
            if (Lv.nbig > 0)
            {
                if (four) k_rsolve<256, 4><<<Lv.nbig, 256, 0, h->stream>>> (fr, Lv.nbig, h->S, h->N, h->d_hcol, Qf, h->info.rank, (I32) c0, B, X, W) ;
                else k_rsolve<256, 1><<<Lv.nbig, 256, 0, h->stream>>> (fr, Lv.nbig, h->S, h->N, h->d_hcol, Qf, h->info.rank, (I32) c0, B, X, W) ;
            }
            if (nsm > 0)
            {
                const int g = (nsm + 7) / 8 ;
                if (four) k_rsolve<32, 4><<<g, 256, 0, h->stream>>> (fr + Lv.nbig, nsm, h->S, h->N, h->d_hcol, Qf, h->info.rank, (I32) c0, B, X, W) ;
                else k_rsolve<32, 1><<<g, 256, 0, h->stream>>> (fr + Lv.nbig, nsm, h->S, h->N, h->d_hcol, Qf, h->info.rank, (I32) c0, B, X, W) ;
            }
            c0 += four ? 4 : 1 ;
        }
    }
    CK (cudaGetLastError ()) ;
    return STMQR_OK ;
}

} // namespace

int stmqr_b200_qmult (stmqr_handle h, int method, int64_t nx, const double *X, double *Y)
{
    int s = solve_ready (h, "qmult") ;
    if (s != STMQR_OK) return s ;
    if ((method != 0 && method != 1) || nx < 0 || (nx > 0 && (!X || !Y))) return fail (h, STMQR_ERR_INVALID, "qmult: bad arguments") ;
    if (nx == 0 || h->m == 0) return STMQR_OK ;
    const I64 cnt = h->m * nx ;
    if ((s = grow (h, &h->d_solveZ, &h->solveZ_cap, cnt)) != STMQR_OK) return s ;
    if ((s = grow (h, &h->d_solveIO, &h->solveIO_cap, std::max (cnt, h->n * nx))) != STMQR_OK) return s ;
    cudaStream_t st = h->stream ;
    CK (cudaEventRecord (h->ev2, st)) ;
    CK (cudaMemcpyAsync (h->d_solveIO, X, (size_t) cnt * sizeof (double), cudaMemcpyHostToDevice, st)) ;
    // Q'X works on Z (HPinv [i], :) = X (i, :), QX on X itself and permutes at the end (:2004-2048)
    if (method == 0) k_permute_rows<<<grid_for (cnt, 256), 256, 0, st>>> (h->m, nx, h->d_HPinv64, h->d_solveIO, h->d_solveZ, 1) ;
    else CK (cudaMemcpyAsync (h->d_solveZ, h->d_solveIO, (size_t) cnt * sizeof (double), cudaMemcpyDeviceToDevice, st)) ;
    if ((s = device_qapply (h, method, nx, h->d_solveZ)) != STMQR_OK) return s ;
    const double *res = h->d_solveZ ;
    if (method == 1) { k_permute_rows<<<grid_for (cnt, 256), 256, 0, st>>> (h->m, nx, h->d_HPinv64, h->d_solveZ, h->d_solveIO, 0) ; res = h->d_solveIO ; }
    CK (cudaMemcpyAsync (Y, res, (size_t) cnt * sizeof (double), cudaMemcpyDeviceToHost, st)) ;
    CK (cudaEventRecord (h->ev3, st)) ;
    CK (cudaStreamSynchronize (st)) ;
    float ms = 0 ; cudaEventElapsedTime (&ms, h->ev2, h->ev3) ; h->ms_solve = ms ;
    return STMQR_OK ;
}

int stmqr_b200_rsolve (stmqr_handle h, int use_Qfill, int64_t nrhs, const double *B, double *X)
{
    int s = solve_ready (h, "rsolve") ;
    if (s != STMQR_OK) return s ;
    if (nrhs < 0 || (nrhs > 0 && (!B || !X))) return fail (h, STMQR_ERR_INVALID, "rsolve: bad arguments") ;
    if (nrhs == 0 || h->n == 0) return STMQR_OK ;
    if ((s = grow (h, &h->d_solveIO, &h->solveIO_cap, std::max (h->m, h->n) * nrhs)) != STMQR_OK) return s ;
    if ((s = grow (h, &h->d_solveX, &h->solveX_cap, 2 * h->n * nrhs)) != STMQR_OK) return s ;
    cudaStream_t st = h->stream ;
    CK (cudaEventRecord (h->ev2, st)) ;
    CK (cudaMemcpyAsync (h->d_solveIO, B, (size_t) (h->m * nrhs) * sizeof (double), cudaMemcpyHostToDevice, st)) ;
    if ((s = device_rsolve (h, use_Qfill, nrhs, h->d_solveIO, h->d_solveX, h->d_solveX + h->n * nrhs)) != STMQR_OK) return s ;
    CK (cudaMemcpyAsync (X, h->d_solveX, (size_t) (h->n * nrhs) * sizeof (double), cudaMemcpyDeviceToHost, st)) ;
    CK (cudaEventRecord (h->ev3, st)) ;
    CK (cudaStreamSynchronize (st)) ;
    float ms = 0 ; cudaEventElapsedTime (&ms, h->ev2, h->ev3) ; h->ms_solve = ms ;
    return STMQR_OK ;
}

// x = E * (R \ (Q'b)): the least-squares solve of qrtest.c:11-53 (QR_qmult (QR_QTX) + QR_solve
// (QR_RETX_EQUALS_B)) without the factor ever leaving the device
int stmqr_b200_solve_ls (stmqr_handle h, int64_t nrhs, const double *B, double *X, double *device_ms)
{
    int s = solve_ready (h, "solve_ls") ;
    if (s != STMQR_OK) return s ;
    if (nrhs < 0 || (nrhs > 0 && (!B || !X))) return fail (h, STMQR_ERR_INVALID, "solve_ls: bad arguments") ;
    if (nrhs == 0 || h->n == 0 || h->m == 0) return STMQR_OK ;
    const I64 cnt = h->m * nrhs ;
    if ((s = grow (h, &h->d_solveZ, &h->solveZ_cap, cnt)) != STMQR_OK) return s ;
    if ((s = grow (h, &h->d_solveIO, &h->solveIO_cap, std::max (h->m, h->n) * nrhs)) != STMQR_OK) return s ;
    if ((s = grow (h, &h->d_solveX, &h->solveX_cap, 2 * h->n * nrhs)) != STMQR_OK) return s ;
    cudaStream_t st = h->stream ;
    CK (cudaEventRecord (h->ev2, st)) ;
    CK (cudaMemcpyAsync (h->d_solveIO, B, (size_t) cnt * sizeof (double), cudaMemcpyHostToDevice, st)) ;
    k_permute_rows<<<grid_for (cnt, 256), 256, 0, st>>> (h->m, nrhs, h->d_HPinv64, h->d_solveIO, h->d_solveZ, 1) ;
    if ((s = device_qapply (h, 0, nrhs, h->d_solveZ)) != STMQR_OK) return s ;
    if ((s = device_rsolve (h, 1, nrhs, h->d_solveZ, h->d_solveX, h->d_solveX + h->n * nrhs)) != STMQR_OK) return s ;
    CK (cudaMemcpyAsync (X, h->d_solveX, (size_t) (h->n * nrhs) * sizeof (double), cudaMemcpyDeviceToHost, st)) ;
    CK (cudaEventRecord (h->ev3, st)) ;
    CK (cudaStreamSynchronize (st)) ;
    float ms = 0 ; cudaEventElapsedTime (&ms, h->ev2, h->ev3) ; h->ms_solve = ms ;
    if (device_ms) *device_ms = ms ;
    return STMQR_OK ;
}

// Symbolic inputs of the device R extraction, built on first use (the drop-in's numeric phase never needs them):
// the transpose of Rj (column of R -> its positions in Rj, ascending = in front order) and the front of every
// position (k_rcol_scan, k_rcount, k_rfill).
static int ensure_rconvert_tables (stmqr_handle h)
{
    if (h->d_RjTp) return STMQR_OK ;
    const I64 n = h->n, nf = h->nf, rjsize = h->rjsize ;
    std::vector<I32> Rj ((size_t) std::max<I64> (rjsize, 1)) ;
    if (rjsize > 0) CK (cudaMemcpy (Rj.data (), h->S.Rj, (size_t) rjsize * sizeof (I32), cudaMemcpyDeviceToHost)) ;
    const std::vector<I32> &Rp = h->h_Rp ;
    std::vector<I32> RjTp ((size_t) n + 1, 0), RjTi ((size_t) std::max<I64> (rjsize, 1)), posfront ((size_t) std::max<I64> (rjsize, 1)) ;
    for (I64 pz = 0 ; pz < rjsize ; pz++) RjTp [(size_t) Rj [pz] + 1]++ ;
    for (I64 j = 0 ; j < n ; j++) RjTp [(size_t) j + 1] += RjTp [(size_t) j] ;
    std::vector<I32> cur (RjTp.begin (), RjTp.end () - 1) ;
    for (I64 f = 0 ; f < nf ; f++)
        for (I32 pz = Rp [f] ; pz < Rp [f+1] ; pz++)
        {
            posfront [pz] = (I32) f ;
            RjTi [cur [Rj [pz]]++] = pz ;
        }
    UPLOAD (h->d_RjTp, RjTp) ; UPLOAD (h->d_RjTi, RjTi) ; UPLOAD (h->d_posfront, posfront) ;
    ALLOC (h->d_rcnt, rjsize) ; ALLOC (h->d_roff, rjsize) ;
    ALLOC (h->d_Rcolp, n + 2) ;
    CK (cudaStreamSynchronize (h->stream)) ;        // (the host vectors go out of scope)
    return STMQR_OK ;
}

// -------------------------------------------------------------------------------------------------
// R as a compressed-column matrix (qr_rcount / qr_rconvert on the device, kernels_solve.cuh)
// -------------------------------------------------------------------------------------------------
int stmqr_b200_rcount (stmqr_handle h, int64_t econ, int64_t *Rp_out, int64_t *nnzR)
{
    int s = solve_ready (h, "rcount") ;
    if (s != STMQR_OK) return s ;
    if (econ < 0) return fail (h, STMQR_ERR_INVALID, "rcount: econ < 0") ;
    if ((s = ensure_rconvert_tables (h)) != STMQR_OK) return s ;
    cudaStream_t st = h->stream ;
    const I64 rj = h->rjsize, n = h->n ;
    if (h->rcount_econ != econ)
    {
        if (rj > 0) k_rcount<<<(unsigned) ((rj * 32 + 255) / 256), 256, 0, st>>> (h->S, h->N, h->d_posfront, h->d_rlen, rj, econ, h->d_rcnt) ;
        if (n > 0) k_rcol_scan<<<(unsigned) ((n * 32 + 255) / 256), 256, 0, st>>> (n, h->d_RjTp, h->d_RjTi, h->d_rcnt, h->d_roff, h->d_Rcolp) ;
        // column totals -> column pointers: exclusive scan over n+1 entries (the last one is the total)
        CK (cudaMemsetAsync (h->d_Rcolp + n, 0, sizeof (I64), st)) ;
        k_scan1_i64<<<1, 1024, 0, st>>> (h->d_Rcolp, (I32) (n + 1)) ;
        I64 tot = 0 ;
        CK (cudaMemcpyAsync (&tot, h->d_Rcolp + n, sizeof (I64), cudaMemcpyDeviceToHost, st)) ;
        CK (cudaStreamSynchronize (st)) ;
        CK (cudaGetLastError ()) ;
        h->rcount_econ = econ ; h->rcount_nnz = tot ;
    }
    if (Rp_out) { CK (cudaMemcpy (Rp_out, h->d_Rcolp, (size_t) (n + 1) * sizeof (I64), cudaMemcpyDeviceToHost)) ; }
    if (nnzR) *nnzR = h->rcount_nnz ;
    return STMQR_OK ;
}

int stmqr_b200_rconvert (stmqr_handle h, int64_t econ, int64_t *Rp_out, int64_t *Ri, double *Rx)
{
    int64_t nnz = 0 ;
    int s = stmqr_b200_rcount (h, econ, Rp_out, &nnz) ;
    if (s != STMQR_OK) return s ;
    if (nnz > 0 && (!Ri || !Rx)) return fail (h, STMQR_ERR_INVALID, "rconvert: no destination") ;
    if (nnz == 0) return STMQR_OK ;
    cudaStream_t st = h->stream ;
    // the extracted factor is staged on the device: Ri in the I64 scratch of the download, Rx in a solve buffer
    I64 *dRi = nullptr ;
    if ((s = grow (h, &h->d_solveIO, &h->solveIO_cap, 2 * nnz)) != STMQR_OK) return s ;
    double *dRx = h->d_solveIO ;
    dRi = (I64 *) (h->d_solveIO + nnz) ;
    k_rfill<<<(unsigned) ((h->rjsize * 32 + 255) / 256), 256, 0, st>>> (h->S, h->N, h->d_posfront, h->d_rlen, h->d_roff, h->d_Rcolp,
        h->rjsize, econ, dRi, dRx) ;
    CK (cudaGetLastError ()) ;
    CK (cudaStreamSynchronize (st)) ;
    { int s1 = ensure_copy_pipeline (h) ; if (s1 != STMQR_OK) return s1 ; }
    if ((s = d2h_pipelined (h, Ri, dRi, (size_t) nnz * sizeof (I64))) != STMQR_OK) return s ;
    if ((s = d2h_pipelined (h, Rx, dRx, (size_t) nnz * sizeof (double))) != STMQR_OK) return s ;
    return STMQR_OK ;
}

int stmqr_b200_get_stats (stmqr_handle h, stmqr_stats *out)
{
    if (!h || !out) return STMQR_ERR_INVALID ;
    h->stats.device_bytes = (I64) h->device_bytes ;
    *out = h->stats ;
    return STMQR_OK ;
}

int stmqr_b200_measure_fp64_peak (stmqr_handle h, double *dmma_tflops, double *dfma_tflops)
{
    if (!h) return STMQR_ERR_INVALID ;
    cudaSetDevice (h->device) ;
    double *d_out = nullptr ;
    CK (cudaMalloc ((void **) &d_out, 148 * 8 * 1024 * sizeof (double))) ;
    const int iters = 4096 ;
    const int grid = 148 * 4, block = 256 ;
    double best [3] = {0, 0, 0} ;
    for (int variant = 0 ; variant < 3 ; variant++)
    {
        for (int rep = 0 ; rep < 4 ; rep++)
        {
            CK (cudaEventRecord (h->ev0, h->stream)) ;
            if (variant == 0) k_peak_dmma884<<<grid, block, 0, h->stream>>> (d_out, iters) ;
            else if (variant == 1) k_peak_dmma16816<<<grid, block, 0, h->stream>>> (d_out, iters) ;
            else k_peak_dfma<<<grid, block, 0, h->stream>>> (d_out, iters) ;
            CK (cudaEventRecord (h->ev1, h->stream)) ;
            CK (cudaStreamSynchronize (h->stream)) ;
            float ms = 0 ;
            cudaEventElapsedTime (&ms, h->ev0, h->ev1) ;
            const double warps = (double) grid * block / 32 ;
            double fl ;
            if (variant == 0) fl = warps * iters * 8.0 * 512.0 ;          // 8 independent m8n8k4
            else if (variant == 1) fl = warps * iters * 4.0 * 4096.0 ;    // 4 independent m16n8k16
            else fl = (double) grid * block * iters * 16.0 * 2.0 ;        // 16 independent DFMA
            if (rep > 0) best [variant] = std::max (best [variant], fl / (ms * 1e-3) * 1e-12) ;
        }
    }
    cudaFree (d_out) ;
    if (dmma_tflops) *dmma_tflops = std::max (best [0], best [1]) ;
    if (dfma_tflops) *dfma_tflops = best [2] ;
    return STMQR_OK ;
}

int stmqr_b200_get_front (stmqr_handle h, int64_t f, int which, double *F, int64_t capacity,
    int64_t *fm, int64_t *fn)
{
    if (!h || !h->factorized || !h->debug_capture || f < 0 || f >= h->nf)
        return fail (h, STMQR_ERR_INVALID, "get_front: needs debug capture and a factorization") ;
    cudaSetDevice (h->device) ;
    I32 hm = 0 ;
    CK (cudaMemcpy (&hm, h->N.Hm + f, sizeof (I32), cudaMemcpyDeviceToHost)) ;
    const I64 n = h->h_Rp [f+1] - h->h_Rp [f] ;
    if (fm) *fm = hm ;
    if (fn) *fn = n ;
    const I64 cnt = (I64) hm * n ;
    if (cnt > capacity) return fail (h, STMQR_ERR_INVALID, "get_front: buffer too small") ;
    if (cnt > 0)
        CK (cudaMemcpy (F, (which ? h->d_capF : h->d_capA) + h->h_capOff [f], cnt * sizeof (double),
            cudaMemcpyDeviceToHost)) ;
    return STMQR_OK ;
}

} // extern "C"
