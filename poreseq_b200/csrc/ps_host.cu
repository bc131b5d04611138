// ps_host.cu -- host side of the B200-native PoreSeq scoring path: context / region objects, the
// batch scheduler that lays regions out in HBM and launches the kernels of ps_device.cuh, the host
// drivers that stay sequential by nature (MakeMutations accept loop, FindPointMutations), and the
// C-ABI of include/poreseq_b200.h.
//
// There is deliberately no CPU implementation of the DP here: if CUDA is unavailable every compute
// entry point fails with PS_E_CUDA.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <pthread.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/poreseq_b200.h"
#include "ps_device.cuh"
#include "ps_fast.cuh"
#include "ps_score32.cuh"
#include "ps_fill2.cuh"
#include "ps_internal.h"

using namespace psdev;

#ifndef PS_FILL_MINB
#define PS_FILL_MINB 3            // CTAs of the 160-thread fill per SM (128 registers); 4 = 96 registers, measured slower
#endif

// ------------------------------------------------------------------------------------------
// errors
static std::string g_create_error;

// cudaStreamSynchronize spins on a host core; with more driving threads than cores (the lanes and groups of
// ps_consensus_batch: 16-48 threads on the 4 cores a rank gets on an 8-GPU box) the spinning takes the cores from the
// threads that have work.  A blocking-sync event puts the waiting thread to sleep instead.
int ps_stream_wait(ps_ctx* ctx)
{
    if (!ctx->blocking_wait) return (int)cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaEventRecord(ctx->wait_ev, ctx->stream);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaEventSynchronize(ctx->wait_ev);
}

void ps_set_error(ps_ctx* ctx, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    // several library threads may refuse their regions at once (ps_regions_create, ps_pack_regions_create)
    static std::mutex error_lock;
    std::lock_guard<std::mutex> hold(error_lock);
    if (ctx) ctx->error = buf; else g_create_error = buf;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess)                                                                 \
        {                                                                                         \
            ps_set_error(ctx, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(err__), __FILE__, \
                         __LINE__, #call);                                                        \
            return PS_E_CUDA;                                                                     \
        }                                                                                         \
    } while (0)

// ------------------------------------------------------------------------------------------
// recycling allocator of the level arrays (ps_internal.h)
namespace
{
constexpr size_t PS_POOL_GRAIN = 1024, PS_POOL_CLASSES = 256, PS_POOL_CAP = (size_t)768 << 20;
constexpr int PS_POOL_SHARDS = 16;              // a thread works on the shard its id hashes to, and steals from the others
struct alignas(64) PoolShard
{
    std::mutex mu;
    std::vector<void*> free_list[PS_POOL_CLASSES + 1];
};
struct LevelPool
{
    PoolShard shard[PS_POOL_SHARDS];
    std::atomic<size_t> held{0};
};
std::atomic<LevelPool*> g_level_pool{nullptr};
// after fork() the child starts with fresh free lists: a shard mutex may have been held by a thread that does not
// exist in the child (the parent's blocks stay valid memory, they are just not recycled there)
void level_pool_forget() { g_level_pool.store(new LevelPool(), std::memory_order_release); }
LevelPool& level_pool()
{
    LevelPool* p = g_level_pool.load(std::memory_order_acquire);
    if (!p)
    {
        static std::once_flag once;
        std::call_once(once, [] {
            g_level_pool.store(new LevelPool(), std::memory_order_release);   // never destroyed: regions may be released during process exit
            pthread_atfork(nullptr, nullptr, level_pool_forget);
        });
        p = g_level_pool.load(std::memory_order_acquire);
    }
    return *p;
}
inline size_t pool_class(size_t bytes) { return (bytes + PS_POOL_GRAIN - 1) / PS_POOL_GRAIN; }
inline int pool_home()
{
    static std::atomic<int> next{0};
    thread_local int home = next.fetch_add(1, std::memory_order_relaxed) % PS_POOL_SHARDS;
    return home;
}
}

void* ps_pool_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    const size_t c = pool_class(bytes);
    if (c <= PS_POOL_CLASSES)
    {
        LevelPool& P = level_pool();
        const int h = pool_home();
        for (int k = 0; k < PS_POOL_SHARDS; k++)
        {
            PoolShard& S = P.shard[(h + k) % PS_POOL_SHARDS];
            std::unique_lock<std::mutex> g(S.mu, std::defer_lock);
            if (k == 0) g.lock();
            else if (!g.try_lock()) continue;
            std::vector<void*>& f = S.free_list[c];
            if (!f.empty())
            {
                void* p = f.back();
                f.pop_back();
                P.held.fetch_sub(c * PS_POOL_GRAIN, std::memory_order_relaxed);
                return p;
            }
        }
        bytes = c * PS_POOL_GRAIN;
    }
    void* p = malloc(bytes);
    if (!p) throw std::bad_alloc();
    return p;
}

void ps_pool_free(void* p, size_t bytes) noexcept
{
    if (!p) return;
    if (bytes == 0) bytes = 1;
    const size_t c = pool_class(bytes);
    if (c <= PS_POOL_CLASSES)
    {
        LevelPool& P = level_pool();
        if (P.held.load(std::memory_order_relaxed) + c * PS_POOL_GRAIN <= PS_POOL_CAP)
        {
            PoolShard& S = P.shard[pool_home()];
            std::lock_guard<std::mutex> g(S.mu);
            try
            {
                S.free_list[c].push_back(p);
                P.held.fetch_add(c * PS_POOL_GRAIN, std::memory_order_relaxed);
                return;
            }
            catch (...) {}
        }
    }
    free(p);
}

size_t ps_pool_held() { return level_pool().held.load(); }

// ------------------------------------------------------------------------------------------
// host worker threads: the per-event staging of a batch (log(stdv), band planning, copies into the
// pinned buffers, result scatter) is independent per event
namespace {
struct Pool
{
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable wake, done;
    const std::function<void(int)>* fn = nullptr;
    std::atomic<int> next{0};
    int n = 0, busy = 0;
    unsigned long long generation = 0;
    bool stop = false;

    explicit Pool(int threads)
    {
        for (int t = 0; t < threads; t++) workers.emplace_back([this] { loop(); });
    }
    void drain()
    {
        for (;;)
        {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            (*fn)(i);
        }
    }
    void loop()
    {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(m);
        for (;;)
        {
            wake.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            lk.unlock();
            drain();
            lk.lock();
            if (--busy == 0) done.notify_all();
        }
    }
    void run(int count, const std::function<void(int)>& f)
    {
        std::unique_lock<std::mutex> lk(m);
        fn = &f; n = count; next.store(0);
        busy = (int)workers.size();
        generation++;
        wake.notify_all();
        lk.unlock();
        drain();
        lk.lock();
        done.wait(lk, [&] { return busy == 0; });
        fn = nullptr;
    }
};
Pool* g_pool = nullptr;
std::once_flag g_pool_fork;
std::mutex g_pool_make;
void pool_forget() { g_pool = nullptr; }          // worker threads do not survive fork(): start over in the child
}

void ps_parallel_for(int n, const std::function<void(int)>& fn)
{
    if (n <= 0) return;
    if (n < 4) { for (int i = 0; i < n; i++) fn(i); return; }
    {
        std::lock_guard<std::mutex> g(g_pool_make);
        if (!g_pool)
        {
            std::call_once(g_pool_fork, [] { pthread_atfork(nullptr, nullptr, pool_forget); });
            int threads = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
            if (const char* e = getenv("PORESEQ_B200_THREADS")) threads = std::max(1, atoi(e));
            g_pool = new Pool(threads - 1);
        }
    }
    // one parallel loop at a time: a caller that finds the workers busy (several host threads driving
    // their own contexts, or a nested loop) runs its loop inline
    static std::mutex busy;
    std::unique_lock<std::mutex> turn(busy, std::try_to_lock);
    if (g_pool->workers.empty() || !turn.owns_lock()) { for (int i = 0; i < n; i++) fn(i); return; }
    g_pool->run(n, fn);
}

// ------------------------------------------------------------------------------------------
// context
int ps_ctx::init()
{
    ps_ctx* ctx = this;
    if (ready) return PS_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
    {
        ps_set_error(ctx, "no CUDA device available (%s); poreseq_b200 has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return PS_E_CUDA;
    }
    if (device < 0 || device >= n) { ps_set_error(ctx, "device %d out of range (0..%d)", device, n - 1); return PS_E_ARG; }
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
    {
        ps_set_error(ctx, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return PS_E_CUDA;
    }
    sm_count = prop.multiProcessorCount;
    total_mem = prop.totalGlobalMem;
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));   // never waits for (or holds up) the legacy default stream of the host program
    CU(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&join_ev, cudaEventDisableTiming));
    for (int i = 0; i <= PS_T_COUNT; i++) CU(cudaEventCreate(&tev[i]));
    CU(cudaEventCreateWithFlags(&wait_ev, cudaEventBlockingSync | cudaEventDisableTiming));
    CU(cudaFuncSetAttribute(k_backtrace, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024));
    CU(cudaFuncSetAttribute(k_fill<160, PS_FILL_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    CU(cudaFuncSetAttribute(k_fill<192, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(k_fill<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024));
    CU(cudaFuncSetAttribute(k_mutscore_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(k_mutscore_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(k_mutscore<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(k_mutscore<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CU(cudaFuncSetAttribute(k_fill2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_fill2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_score_f32<true, false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_score_f32<true, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_score_f32<false, false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_score_f32<false, false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_score_f32<false, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_score_f32<false, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // the per-thread rings of the exact mutation kernel are what limits its occupancy: ask for the largest shared-memory carve-out
    CU(cudaFuncSetAttribute(k_mutscore<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_mutscore<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ready = true;
    return PS_OK;
}

int ps_ctx::ensure(DevBuf& b, size_t bytes)
{
    ps_ctx* ctx = this;
    if (bytes <= b.cap) return PS_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess)
    {
        ps_set_error(ctx, "out of device memory allocating %zu bytes (%s)", bytes, cudaGetErrorString(e));
        b.p = nullptr;
        return PS_E_CUDA;
    }
    b.cap = want;
    return PS_OK;
}

bool ps_pin_reserve(PinBuf* b, size_t bytes, size_t keep)
{
    if (bytes <= b->cap) return true;
    size_t want = bytes + bytes / 2 + 4096;
    void* q = nullptr;
    if (cudaHostAlloc(&q, want, cudaHostAllocDefault) != cudaSuccess)
    {
        cudaGetLastError();
        q = malloc(want);                       // pageable staging still works, only slower
        if (!q) return false;
        if (b->p && keep) memcpy(q, b->p, keep);
        // pageable blocks are leaked into the context map as cap with the low bit set
        if (b->p) { if (b->cap & 1) free(b->p); else cudaFreeHost(b->p); }
        b->p = q; b->cap = want | 1;
        return true;
    }
    if (b->p && keep) memcpy(q, b->p, keep);
    if (b->p) { if (b->cap & 1) free(b->p); else cudaFreeHost(b->p); }
    b->p = q; b->cap = want & ~(size_t)1;
    return true;
}

ps_ctx::~ps_ctx()
{
    for (ps_ctx* h : helpers) delete h;
    helpers.clear();
    for (ps_ctx* h : group_ctx) delete h;
    group_ctx.clear();
    if (!ready) return;
    cudaSetDevice(device);
    ps_comm_destroy(this);
    for (auto& kv : bufs) if (kv.second.p) cudaFree(kv.second.p);
    for (auto& kv : pins) if (kv.second.p) { if (kv.second.cap & 1) free(kv.second.p); else cudaFreeHost(kv.second.p); }
    for (int i = 0; i <= PS_T_COUNT; i++) cudaEventDestroy(tev[i]);
    cudaEventDestroy(fork_ev); cudaEventDestroy(join_ev); cudaEventDestroy(wait_ev);
    cudaStreamDestroy(side);
    cudaStreamDestroy(stream);
}

// ------------------------------------------------------------------------------------------
// host-side sequence / event helpers
std::vector<int> ps_states_of(const std::string& bases)     // cpp/Sequence.h:69-100
{
    std::vector<int> st;
    const int n = (int)bases.size();
    if (n < 5) return st;
    st.reserve(n - 4);
    auto code = [](char ch) -> int { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : (int)ch; };
    int window = 0;
    for (int i = 0; i < 4; i++) window = (window << 2) + code(bases[i]);
    for (int i = 4; i < n; i++)
    {
        if (code(bases[i - 4]) < 4)
        {
            window = (N_STATES - 1) & ((window << 2) + code(bases[i]));
            st.push_back(window);
        }
        else { window = 0; st.push_back(-1); }
    }
    return st;
}

std::string ps_apply_mutation(const std::string& bases, int start, const std::string& orig, const std::string& mut)
{   // cpp/Sequence.h:38-59
    if ((size_t)start >= bases.size()) return bases;
    std::string out(bases, 0, start);
    out += mut;
    size_t rest = (size_t)start + orig.size();
    if (rest < bases.size()) out.append(bases, rest, std::string::npos);
    return out;
}

void HostEvent::update_refs() { update_refs_from(ref_align.data()); }

void HostEvent::update_refs_from(const double* ref_align)     // cpp/EventData.h:110-169
{
    const int n = n0;
    ri_stale = false;
    refstart = refend = -1;
    int lo = 0, hi = n - 1;
    while (lo < n && !(ref_align[lo] > 0)) lo++;
    while (hi >= 0 && !(ref_align[hi] > 0)) hi--;
    if (lo == n || hi < 0) { ri_empty = true; return; }
    ri_empty = false;
    refstart = (int)ref_align[lo];
    refend = (int)ref_align[hi];
    ref_index.assign(ref_align, ref_align + n);
    const double slope = (ref_align[hi] - ref_align[lo]) / (double)(hi - lo);
    const double icpt = ref_align[lo] - slope * lo;
    int anchor = -1;
    for (int i = 0; i < n; i++)
    {
        if (i < lo || i > hi) { ref_index[i] = slope * i + icpt; continue; }
        if (!(ref_align[i] > 0)) continue;
        if (anchor > 0)
        {
            const double m = (ref_align[i] - ref_align[anchor]) / (i - anchor);
            for (int j = anchor + 1; j < i; j++) ref_index[j] = m * (j - anchor) + ref_align[anchor];
        }
        anchor = i;
    }
}

void HostEvent::ensure_levrec()
{
    if (levrec.size() == (size_t)n0 * 3) return;
    levrec.resize((size_t)n0 * 3);
    for (int i = 0; i < n0; i++)
    {
        // staged level record (psdev::LevIn): mean, stdv, 3*log(stdv)  (cpp/EventData.h:218-220)
        levrec[3 * i] = mean[i]; levrec[3 * i + 1] = stdv[i]; levrec[3 * i + 2] = 3 * std::log(stdv[i]);
    }
}

void ps_build_model(const HostModel& hm, ModelDev& md)     // cpp/EventData.h:48-73
{
    for (int s = 0; s < N_STATES; s++)
    {
        StateParams& p = md.st[s];
        p.lev_mean = hm.raw[0][s];
        p.lev_stdv = hm.raw[1][s];
        p.log_lev = std::log(hm.raw[1][s]);
        p.sd_mean = hm.raw[2][s];
        p.sd_lambda = std::pow(hm.raw[2][s], 3) / std::pow(hm.raw[3][s], 2);
        p.log_lambda = std::log(p.sd_lambda);
        p.r_lev_stdv = 1.0 / p.lev_stdv;
        p.r_sd_mean = 1.0 / p.sd_mean;
    }
    md.lskip = std::log(hm.trans[0]);
    md.lstay = std::log(hm.trans[1]);
    md.lext = std::log(hm.trans[2]);
    md.lins = std::log(hm.trans[3]);
}

// ------------------------------------------------------------------------------------------
// Job: one launch sequence over a batch of regions
struct MutSpec                      // the mutations of one region: an explicit list or every point edit
{
    const std::vector<HostMut>* list = nullptr;
    bool points = false;
};

typedef RegTabDev RegTab;

// absolute error bound of one FP32 (mutation, event) delta of k_mutscore_rows_f32, see Job::upload
constexpr double PS_FAST_PAIR_ERR = 5e-5;     // 3x the largest error seen (1.55e-5 per event, profiles/r2_fast_error.txt)

struct Job
{
    ps_ctx* ctx;
    std::vector<ps_region*> regs;
    std::vector<MutSpec> muts;                   // per region (empty for alignment-only jobs)
    bool want_muts;

    // host staging (pinned, owned by the context)
    PinVec<EvDesc> ev;
    PinVec<int> states;
    PinVec<char> bases;
    PinVec<LevIn> lev;
    bool fast = false;                           // FP32 pass + exact re-score (PS_PRECISION_FAST)
    bool scores_only = false;                    // PSAlign.ScorePoints / ScoreMutations semantics (pyx:278-342): scores out, the realignment is dropped --
                                                 // no D2H of the alignments, nothing scattered into the regions
    std::vector<ps_region*> owned;               // regions built for this job alone (ps_score_points_direct), deleted with it
    ~Job() { for (ps_region* r : owned) delete r; }
    bool sharded = false;                        // this job scores ONE rank's block of a region's events (ps_comm.cu): sums combined over NCCL
    int total_events = 0;                        // events of the region over all ranks (sharded FAST: the re-score threshold counts them all)
    double* d_scores2 = nullptr;                 // sharded FAST: totals of the exact pass (dense, only the flagged entries mean anything)
    bool score32 = false;                        // ScoreEvents in FAST mode: k_score_f32, no matrices, no backtrace, events untouched
    PinVec<int> s32_list;                        // events by launch class of k_score_f32: (staged, plain), (staged, inv), (global, plain), (global, inv)
    int s32_count[4] = {0, 0, 0, 0}, s32_max_n0 = 0;
    int max_ev = 1;
    PinVec<double> ref_align, ref_like, ref_index;
    PinVec<int> ri_empty, mono, cen_old;
    PinVec<MutDev> mdev;
    PinVec<char> mut_str;
    PinVec<RegTab> regtab;
    std::vector<const HostModel*> model_src;
    std::vector<int> wave_need, wave_need_lazy;
    // wide-fill launch classes: events whose wavefront fits 160 threads (3 CTAs per SM), wider ones, serial ones
    PinVec<int> fill_list;
    int fill_count[4] = {0, 0, 0, 0}, fill_threads[4] = {160, 192, 224, 32};
    double bias = -1e-6;                         // start value of every mutation's sum over events
    double wide_cells_fwd = 0, narrow_cells = 0;
    long long n_levels, n_cols, n_cen, n_tasks, n_muts, n_band, n_strips = 0;
    int cen_pad;
    Batch b;
    RegTab* d_regtab = nullptr;
    int* d_s32_list = nullptr;
    double* d_evbest = nullptr;

    Job(ps_ctx* c) : ctx(c), want_muts(false), n_levels(0), n_cols(0), n_cen(0), n_tasks(0), n_muts(0), n_band(0), cen_pad(8)
    {
        ev = c->pinned<EvDesc>("ev"); states = c->pinned<int>("states"); bases = c->pinned<char>("bases");
        lev = c->pinned<LevIn>("lev"); ref_align = c->pinned<double>("ref_align");
        ref_like = c->pinned<double>("ref_like"); ref_index = c->pinned<double>("ref_index");
        ri_empty = c->pinned<int>("ri_empty"); mono = c->pinned<int>("mono"); cen_old = c->pinned<int>("cen_old");
        mdev = c->pinned<MutDev>("mdev"); mut_str = c->pinned<char>("mut_str"); regtab = c->pinned<RegTab>("regtab");
        fill_list = c->pinned<int>("fill_list"); s32_list = c->pinned<int>("s32_list");
    }

    void plan_event(const HostEvent& he, const EvDesc& d, int rw, int* cen, int* ok_out, int* need_out, int* need_lazy_out, double* cells_out);
    int build();
    int upload();
    int run(bool backward_and_muts);
    int download_enqueue();                      // D2H copies + completion event on the stream
    int finish(std::vector<double>* align_scores, std::vector<double>* mut_scores);   // sync + scatter
    PinVec<int> rs, re;
    PinVec<double> best, msc;
    bool have_scores = false;
    bool host_muts = false, dev_points = false;   // some mutation tables come from the host / are written by k_points
    double* raw_scores = nullptr;
};

// getrefstate(c) for c = 0 .. n_cols-1 (cpp/EventData.h:172-183: std::lower_bound over ref_index); returns whether
// the centres are nondecreasing.
int ps_host_centres(const double* ri, int n0, int n_cols, int* cen)
{
    // (an event with ONE aligned level has a 0/0 slope: every other entry of ref_index is NaN, cpp/EventData.h:143-144;
    // std::lower_bound's probes then decide, so a NaN anywhere must take the binary search below)
    int ok = 1;
    bool sorted = true;
    for (int i = 1; i < n0 && sorted; i++) sorted = ri[i] >= ri[i - 1];
    if (n0 > 0 && ri[0] != ri[0]) sorted = false;
    if (sorted)
    {
        // on sorted data lower_bound is "first element >= c": one linear merge for all columns
        int idx = 0;
        for (int c = 0; c < n_cols; c++)
        {
            while (idx < n0 && ri[idx] < (double)c) idx++;
            cen[c] = idx;
        }
    }
    else
        for (int c = 0; c < n_cols; c++)
        {
            const int v = (int)(std::lower_bound(ri, ri + n0, (double)c) - ri);
            cen[c] = v;
            if (c > 0 && v < cen[c - 1]) ok = 0;
        }
    return ok;
}

// What the planner sees of one event (host only): ref_index as updaterefs leaves it and the band centres of the first
// n_cols columns.  For the CPU tests of the band planning; ref_index_out may be null.
extern "C" int ps_band_centres(ps_region* R, int event, int n_cols, int* centres, double* ref_index_out, int* ri_empty, int* monotone)
{
    if (!R || event < 0 || event >= (int)R->events.size() || n_cols < 0 || (n_cols > 0 && !centres))
        return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_band_centres");
    HostEvent& he = R->events[event];
    he.ensure_refs();
    int ok = 1;
    if (he.ri_empty) for (int c = 0; c < n_cols; c++) centres[c] = 1;            // cpp/Alignment.cpp:129-132
    else
    {
        ok = ps_host_centres(he.ref_index.data(), he.n0, n_cols, centres);
        if (ref_index_out) memcpy(ref_index_out, he.ref_index.data(), sizeof(double) * (size_t)he.n0);
    }
    if (monotone) *monotone = ok;
    if (ri_empty) *ri_empty = he.ri_empty ? 1 : 0;
    return PS_OK;
}

// Band centres of the wide fill (cpp/EventData.h:172-183: lower_bound over ref_index; 1 when the
// event has no alignment, cpp/Alignment.cpp:129-132), whether they are nondecreasing, the forward
// band cell count, and the number of threads the wavefront needs so that a thread's next strip
// starts after its current one has ended.  Writes cen_old[d.cen_off ..]; one call per event, any thread.
void Job::plan_event(const HostEvent& he, const EvDesc& d, int rw, int* cen, int* ok_out, int* need_out, int* need_lazy_out, double* cells_out)
{
    const int N = d.N, n0 = he.n0;
    for (int c = 0; c <= N + cen_pad; c++) cen[c] = 1;
    int ok = 1, need = 32, need_lazy[3] = {32, 32, 32};
    double cells = 0;
    if (!he.ri_empty)
    {
        ok = ps_host_centres(he.ref_index.data(), n0, N + cen_pad + 1, cen);
    }
    if (d.usable && ok)
    {
        // strips of CW columns in processing order: strip j runs row pairs [lo_j, hi_j] at steps j + pair
        const int J = (N + CW - 1) / CW;
        std::vector<int> lo(J + 1), hi(J + 1);
        for (int dir = 0; dir < 2; dir++)
        {
            for (int j = 0; j < J; j++) { lo[j] = 1 << 30; hi[j] = 0; }
            for (int k = 1; k <= N; k++)
            {
                const int c = dir ? N - k + 1 : k;
                int mid = dir ? n0 - cen[c] + 1 : cen[c];
                mid = std::min(std::max(mid, 1), n0);
                const int i0 = std::max(1, mid - rw), i1 = std::min(n0, mid + rw);
                const int j = (k - 1) / CW;
                lo[j] = std::min(lo[j], j + ((i0 - 1) >> 1)); hi[j] = std::max(hi[j], j + ((i1 - 1) >> 1));
                if (!dir) cells += i1 - i0 + 1;
            }
            // a thread's next strip (j+T) must start after its current one (j) has ended; with P + 1 idle steps in
            // between, a warp may take up its finished lanes' next strips together every P-th step (fill_wave):
            // need_lazy[q] = threads needed for P = 2, 4, 8
            int jj = 1, jl[3] = {1, 1, 1};
            for (int j = 0; j < J; j++)
            {
                if (jj <= j) jj = j + 1;
                while (jj < J && lo[jj] <= hi[j]) jj++;
                need = std::max(need, jj - j);
                for (int q = 0; q < 3; q++)
                {
                    if (jl[q] <= j) jl[q] = j + 1;
                    while (jl[q] < J && lo[jl[q]] <= hi[j] + (2 << q) + 1) jl[q]++;
                    need_lazy[q] = std::max(need_lazy[q], jl[q] - j);
                }
            }
        }
    }
    *ok_out = ok; *need_out = need; for (int q = 0; q < 3; q++) need_lazy_out[q] = need_lazy[q]; *cells_out = cells;
}

int Job::build()
{
    const ps_params& P = regs[0]->params;
    for (size_t r = 0; r < regs.size(); r++)
    {
        const ps_params& Q = regs[r]->params;
        if (Q.lik_offset != P.lik_offset || Q.realign_width != P.realign_width || Q.scoring_width != P.scoring_width)
        {
            ps_set_error(ctx, "all regions of a batch must share lik_offset / realign_width / scoring_width");
            return PS_E_ARG;
        }
    }
    if (P.realign_width < 0 || P.realign_width > 511 || P.scoring_width < 0)
    {
        ps_set_error(ctx, "realign_width must be in 0..511 and scoring_width >= 0 (got %d, %d)", P.realign_width, P.scoring_width);
        return PS_E_ARG;
    }
    // longest net insertion decides how far past N the post-backtrace centre table must reach
    cen_pad = 8;
    size_t tot_levels = 0, tot_states = 0, tot_bases = 0, tot_events = 0, tot_muts = 0, tot_cen = 0;
    for (size_t r = 0; r < regs.size(); r++)
        if (want_muts && muts[r].list)
            for (const HostMut& m : *muts[r].list)
                cen_pad = std::max(cen_pad, (int)m.mut.size() - (int)m.orig.size() + 8);
    for (size_t r = 0; r < regs.size(); r++)
    {
        for (const HostEvent& he : regs[r]->events) tot_levels += he.n0;
        tot_states += regs[r]->states.size(); tot_bases += regs[r]->bases.size(); tot_events += regs[r]->events.size();
        tot_cen += regs[r]->events.size() * (regs[r]->states.size() + cen_pad + 1);
        if (want_muts) tot_muts += muts[r].points ? regs[r]->states.size() * 9 : muts[r].list->size();
    }
    if (!lev.resize(tot_levels) || !ref_align.resize(tot_levels) ||
        !ref_like.resize(tot_levels) || !states.reserve(tot_states) ||
        !bases.reserve(tot_bases) || !ev.reserve(tot_events) || !mdev.reserve(tot_muts) || !ri_empty.resize(tot_events) ||
        !mono.resize(tot_events) || !cen_old.resize(tot_cen) || !regtab.reserve(regs.size()))
    {
        ps_set_error(ctx, "out of host memory staging the batch");
        return PS_E_INTERNAL;
    }
    mut_str.append("ACGT", 4);                  // single-base replacement strings live at offsets 0..3

    const bool trace_b = ctx->trace;
    auto nowb = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tb0 = nowb();
    // pass 1 (serial, cheap): region tables, list-mutation tables, event descriptors and offsets.  The
    // point-mutation tables (8-9 entries per state) are only sized here and written in pass 2.
    std::vector<const HostEvent*> hev;
    std::vector<double> ev_cols;                  // per event: narrow columns of its region's mutations
    std::vector<long long> reg_mut_off(regs.size(), 0);
    hev.reserve(tot_events); ev_cols.reserve(tot_events);
    for (size_t r = 0; r < regs.size(); r++)
    {
        ps_region* R = regs[r];
        const long long state_off = (long long)states.size();
        const long long base_off = (long long)bases.size();
        states.append(R->states.data(), R->states.size());
        bases.append(R->bases.data(), R->bases.size());
        const long long mut_off = (long long)mdev.size();
        reg_mut_off[r] = mut_off;
        const int ev0 = (int)ev.size();
        double cols = 0;
        int region_inv = 0;               // any base that is not ACGT (an invalid state, or a polluted one near the end)
        long long odd = 0;                // non-ACGT bases that start a 5-mer window (4 substitutions instead of 3)
        {
            const size_t ns = R->states.size();
            for (size_t i = 0; i < R->bases.size(); i++)
            {
                const char ch = R->bases[i];
                if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') { region_inv = 1; if (i < ns) odd++; }
            }
        }
        if (want_muts && muts[r].points)
        {
            // FindPointMutations (cpp/FindMutations.cpp:191-234): del, 3 subs (4 for a non-ACGT base), 4 ins per state
            const long long ns = (long long)R->states.size();
            mdev.resize((size_t)(mut_off + 8 * ns + odd));
            cols = (double)(ns - odd) * (5 + 6.0 * 3 + 6.0 * 4) + (double)odd * (5 + 6.0 * 4 + 6.0 * 4);
        }
        else if (want_muts)
        {
            for (const HostMut& hm : *muts[r].list)
            {
                MutDev d;
                d.start = hm.start; d.n_orig = (int)hm.orig.size(); d.n_mut = (int)hm.mut.size();
                d.str_off = (int)mut_str.size();
                mut_str.append(hm.mut.data(), hm.mut.size());
                mdev.push_back(d);
                if (!((size_t)hm.start > R->bases.size())) cols += (double)hm.mut.size() + 5;
            }
        }
        const int nm = (int)(mdev.size() - mut_off);
        RegTab rt; rt.mut_off = mut_off; rt.ev0 = ev0; rt.nev = (int)R->events.size();
        rt.plain_points = (want_muts && muts[r].points && odd == 0) ? 1 : 0; rt.pad = 0;
        if (want_muts && !rt.plain_points) host_muts = true;
        if (rt.plain_points) dev_points = true;
        max_ev = std::max(max_ev, rt.nev);
        regtab.push_back(rt);
        // model de-duplication across the whole batch (events usually share two models): once per
        // (region, model), not per event
        std::vector<int> model_of(R->models.size(), -1);
        for (size_t k = 0; k < R->events.size(); k++)
        {
            const HostEvent& he = R->events[k];
            EvDesc d;
            memset(&d, 0, sizeof d);
            d.region = (int)r;
            d.n0 = he.n0;
            d.N = (int)R->states.size();
            d.L = (int)R->bases.size();
            d.usable = (!he.ri_empty && R->params.realign_width != 0 && he.n0 > 0) ? 1 : 0;
            d.inv = region_inv;
            int& mi = model_of[he.model];
            if (mi < 0)
            {
                const HostModel* hm = &R->models[he.model];
                for (size_t q = 0; q < model_src.size(); q++)
                    if (model_src[q] == hm || memcmp(model_src[q], hm, sizeof(HostModel)) == 0) { mi = (int)q; break; }
                if (mi < 0) { mi = (int)model_src.size(); model_src.push_back(hm); }
            }
            d.model = mi;
            d.n_muts = nm;
            d.lev_off = n_levels;
            d.col_off = n_cols;
            d.state_off = state_off;
            d.base_off = base_off;
            d.cen_off = n_cen;
            d.mut_off = mut_off;
            d.task_off = n_tasks;
            n_levels += he.n0;
            n_cols += d.N;                       // columns 1..N live at col_off+1 .. col_off+N
            n_cen += d.N + cen_pad + 1;
            n_tasks += nm;
            ev.push_back(d);
            hev.push_back(&he);
            ev_cols.push_back(cols);
        }
    }
    n_cols += 1;                                  // index 0 of the first event is never used
    n_muts = (long long)mdev.size();

    const double tb1 = nowb();
    // pass 2a (parallel over regions): the implicit point-mutation tables in FindPointMutations order
    if (want_muts)
        ps_parallel_for((int)regs.size(), [&](int r) {
            if (!muts[r].points || regtab[r].plain_points) return;      // plain ACGT regions: k_points on the device
            const ps_region* R = regs[r];
            MutDev* out = mdev.data() + reg_mut_off[r];
            for (int i = 0; i < (int)R->states.size(); i++)
            {
                const char here = R->bases[i];
                MutDev d;
                d.start = i; d.n_orig = 1; d.n_mut = 0; d.str_off = 0;
                *out++ = d;
                d.n_mut = 1;
                for (int j = 0; j < 4; j++)
                    if ("ACGT"[j] != here) { d.str_off = j; *out++ = d; }
                d.n_orig = 0;
                for (int j = 0; j < 4; j++) { d.str_off = j; *out++ = d; }
            }
        });
    const double tb2 = nowb();
    // pass 2b (parallel over events): level records, alignment arrays, band centres, wavefront plan
    const int ne = (int)ev.size();
    wave_need.assign(ne, 32);
    wave_need_lazy.assign((size_t)ne * 3, 1 << 30);
    std::vector<double> ev_cells(ne, 0.0);
    const int rw = P.realign_width;
    ps_parallel_for(ne, [&](int e) {
        HostEvent& he = *const_cast<HostEvent*>(hev[e]);
        const EvDesc& d = ev[e];
        he.ensure_refs();
        const size_t at = (size_t)d.lev_off, n = (size_t)he.n0;
        if (he.ext_levrec)
            memcpy(lev.data() + at, he.ext_levrec, n * sizeof(LevIn));     // a shadow region: the records of the event it shadows
        else if ((he.staged++ == 0 && he.levrec.empty()) || he.ext_mean)
        {
            // first batch of this event: (mean, stdv, 3 log stdv) straight into the staging buffer (the log is the only
            // transcendental of the path, cpp/EventData.h:218-220); an event that comes back (Refine's recursion, the
            // consensus loop) caches them next time.  1/stdv and the FP32 records are derived on the device (k_rows).
            LevIn* out = lev.data() + at;
            const double* mean = he.ext_mean ? he.ext_mean : he.mean.data();
            const double* stdv = he.ext_stdv ? he.ext_stdv : he.stdv.data();
            for (size_t k = 0; k < n; k++) { out[k].mean = mean[k]; out[k].stdv = stdv[k]; out[k].lsd3 = 3 * std::log(stdv[k]); }
        }
        else
        {
            he.ensure_levrec();
            memcpy(lev.data() + at, he.levrec.data(), n * sizeof(LevIn));
        }
        // ref_align / ref_like / ref_index are outputs of the device (k_backtrace rewrites them for every usable
        // event, the band centres of the fill are planned here on the host): nothing to stage or upload
        ri_empty[e] = (he.ri_empty || !d.usable) ? 1 : 0;
        int ok = 1;
        plan_event(he, d, rw, cen_old.data() + d.cen_off, &ok, &wave_need[e], &wave_need_lazy[(size_t)e * 3], &ev_cells[e]);
        // the wavefront fill assumes log(prob_skip) <= 0 and log(prob_insert) <= 0 (see fill_wave / cell_pre)
        if (!(model_src[d.model]->trans[0] <= 1.0) || !(model_src[d.model]->trans[3] <= 1.0)) ok = 0;
        mono[e] = ok;
    });
    const double tb3 = nowb();
    // wavefront-major band storage: slots per step = wavefront width of the event's launch class (at
    // most that many strips are live on one step), one slot per strip for the serially filled events
    std::vector<int> cls(ne, -1);
    int wide_t = 224;
    for (int e = 0; e < ne; e++)
    {
        const EvDesc& d = ev[e];
        if (d.usable && want_muts) narrow_cells += ev_cols[e] * std::min(d.n0, 2 * P.scoring_width + 1);
        if (!d.usable) continue;
        if (mono[e] && wave_need[e] > 512) mono[e] = 0;
        if (mono[e]) wide_cells_fwd += ev_cells[e];
        cls[e] = !mono[e] ? 3 : wave_need[e] <= 160 ? 0 : wave_need[e] <= 192 ? 1 : 2;
        if (cls[e] == 2) wide_t = std::max(wide_t, ((wave_need[e] + 31) / 32) * 32);
        fill_count[cls[e]]++;
    }
    fill_threads[2] = wide_t;
    if (!fill_list.resize((size_t)std::max(ne, 1))) { ps_set_error(ctx, "out of host memory staging the batch"); return PS_E_INTERNAL; }
    {
        int at[4] = {0, fill_count[0], fill_count[0] + fill_count[1], fill_count[0] + fill_count[1] + fill_count[2]};
        for (int e = 0; e < ne; e++) if (cls[e] >= 0) fill_list[at[cls[e]]++] = e;
    }
    for (int e = 0; e < ne; e++)
    {
        EvDesc& d = ev[e];
        if (!d.usable) { d.ts = 1; d.rs = 4; d.band_off = n_band; continue; }
        const int J = (d.N + CW - 1) / CW;
        d.ts = cls[e] == 3 ? J + 1 : fill_threads[cls[e]];
        d.lazy = 0;                               // strip-switch period - 1: the longest of 8, 4, 2 the event's idle steps allow
        if (cls[e] != 3)
            for (int q = 0; q < 3; q++)
                if (wave_need_lazy[(size_t)e * 3 + q] <= fill_threads[cls[e]]) d.lazy = (2 << q) - 1;
        d.rs = d.ts * 4;                          // one 2x2 tile per slot
        d.band_off = n_band;                      // multiple of 4: keeps the 32-byte tiles aligned
        n_band += (long long)(J + (d.n0 + 1) / 2 + 2) * d.rs;
        d.strip_off = n_strips;                   // forward strips 0..J, then reverse strips 0..J
        n_strips += 2 * (J + 1);
    }
    if (trace_b) fprintf(stderr, "[ps] build: tables %.2f ms, point tables %.2f, per-event staging %.2f, layout %.2f\n", tb1 - tb0, tb2 - tb1, tb3 - tb2, nowb() - tb3);
    return PS_OK;
}

template <class T>
static int up(ps_ctx* ctx, const char* name, const T* src, size_t count, T** out)
{
    DevBuf& buf = ctx->bufs[name];
    int rc = ctx->ensure(buf, std::max<size_t>(count, 1) * sizeof(T));
    if (rc) return rc;
    if (count) CU(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (long long)(count * sizeof(T));
    *out = (T*)buf.p;
    return PS_OK;
}

template <class T>
static int room(ps_ctx* ctx, const char* name, size_t count, T** out)
{
    DevBuf& buf = ctx->bufs[name];
    int rc = ctx->ensure(buf, std::max<size_t>(count, 1) * sizeof(T));
    if (rc) return rc;
    *out = (T*)buf.p;
    return PS_OK;
}

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

int Job::upload()
{
    memset(&b, 0, sizeof b);
    const ps_params& P = regs[0]->params;
    b.n_events = (int)ev.size();
    b.lik_offset = P.lik_offset;
    b.log2pi = std::log(2 * M_PI);                 // cpp/AlignUtil.h:24
    b.realign_width = P.realign_width;
    b.scoring_width = P.scoring_width;
    b.cen_pad = cen_pad;
    // exact pass with one warp per (mutation, event) pair when the pairs are few and no replacement string needs
    // more than 32 narrow columns (k_mutscore_warp)
    b.warp_limit = ((size_t)(2 * P.scoring_width + 2) * 8 * sizeof(double) <= 96 * 1024 && !ctx->no_warp) ? 32768 : 0;
    b.RS = ((2 * P.realign_width + 1) + 3) & ~3;
    b.n_tasks = n_tasks;

    PinVec<ModelDev> models = ctx->pinned<ModelDev>("models");
    if (!models.resize(model_src.size())) { ps_set_error(ctx, "out of host memory staging the models"); return PS_E_INTERNAL; }
    ps_parallel_for((int)model_src.size(), [&](int q) { ps_build_model(*model_src[q], models[q]); });

    EvDesc* d_ev; ModelDev* d_models; int* d_states; char* d_bases;
    LevIn* d_lev;
    TRY(up(ctx, "ev", ev.data(), ev.size(), &d_ev));
    TRY(up(ctx, "models", models.data(), models.size(), &d_models));
    TRY(up(ctx, "states", states.data(), states.size(), &d_states));
    TRY(up(ctx, "bases", bases.data(), bases.size(), &d_bases));
    TRY(up(ctx, "lev", lev.data(), lev.size(), &d_lev));
    TRY(room(ctx, "ref_align", ref_align.size(), &b.ref_align));
    TRY(room(ctx, "ref_like", ref_like.size(), &b.ref_like));
    TRY(room(ctx, "ref_index", ref_align.size(), &b.ref_index));
    TRY(up(ctx, "ri_empty", ri_empty.data(), ri_empty.size(), &b.ri_empty));
    TRY(up(ctx, "mono", mono.data(), mono.size(), &b.mono));
    { int* fl; TRY(up(ctx, "fill_list", fill_list.data(), fill_list.size(), &fl)); b.fill_list = fl; }
    b.ev = d_ev; b.models = d_models; b.states = d_states; b.bases = d_bases;
    b.lev_in = d_lev;
    TRY(room(ctx, "lev_rec", (size_t)std::max<long long>(n_levels, 1), &b.lev));
    TRY(room(ctx, "bt_src", (size_t)n_levels, &b.bt_src));
    TRY(room(ctx, "refstart", ev.size(), &b.refstart));
    TRY(room(ctx, "refend", ev.size(), &b.refend));
    TRY(up(ctx, "cen_old", cen_old.data(), cen_old.size(), &b.cen_old));
    TRY(room(ctx, "cen_new", (size_t)n_cen, &b.cen_new));

    const size_t cells = score32 ? 1 : (size_t)std::max<long long>(n_band, 1);     // k_score_f32 stores no band matrices
    TRY(room(ctx, "Fm", cells, &b.Fm));
    TRY(room(ctx, "Fs", cells, &b.Fs));
    TRY(room(ctx, "Fstep", cells, &b.Fstep));
    TRY(room(ctx, "strips", (size_t)std::max<long long>(n_strips, 1), &b.strips));
    TRY(room(ctx, "rowF", (size_t)std::max<long long>(n_levels, 1) + 1, &b.rowF));     // + 1: load_rows reads rows ia, ia + 1
    TRY(room(ctx, "rowB", (size_t)std::max<long long>(n_levels, 1) + 1, &b.rowB));
    TRY(room(ctx, "Fi0", (size_t)n_cols, &b.Fi0));
    TRY(room(ctx, "Flen", (size_t)n_cols, &b.Flen));
    TRY(room(ctx, "Fcb", (size_t)n_cols, &b.Fcb));
    TRY(room(ctx, "Fcbi", (size_t)n_cols, &b.Fcbi));
    TRY(room(ctx, "Fbest", (size_t)n_cols, &b.Fbest));
    TRY(room(ctx, "Fbi", (size_t)n_cols, &b.Fbi));
    TRY(room(ctx, "Fbj", (size_t)n_cols, &b.Fbj));
    if (want_muts)
    {
        TRY(room(ctx, "Bm", cells, &b.Bm));
        TRY(room(ctx, "Bi0", (size_t)n_cols, &b.Bi0));
        TRY(room(ctx, "Blen", (size_t)n_cols, &b.Blen));
        TRY(room(ctx, "Bcb", (size_t)n_cols, &b.Bcb));
        TRY(room(ctx, "Bcbi", (size_t)n_cols, &b.Bcbi));
        TRY(room(ctx, "Bbest", (size_t)n_cols, &b.Bbest));
        TRY(room(ctx, "old", (size_t)n_cols, &b.old));
        MutDev* d_m; char* d_ms; RegTab* d_rt;
        if (host_muts) TRY(up(ctx, "muts", mdev.data(), mdev.size(), &d_m));
        else TRY(room(ctx, "muts", mdev.size(), &d_m));
        TRY(up(ctx, "mut_str", mut_str.data(), mut_str.size(), &d_ms));
        TRY(up(ctx, "regtab", regtab.data(), regtab.size(), &d_rt));
        b.muts = d_m; b.mut_str = d_ms;
        d_regtab = d_rt;
        TRY(room(ctx, "delta", (size_t)n_tasks, &b.delta));
        TRY(room(ctx, "scores", (size_t)n_muts, &b.scores));
        b.regs = d_rt; b.n_regs = (int)regs.size(); b.max_ev = max_ev;
        TRY(room(ctx, "flag_list", (size_t)n_muts + 1, &b.flag_list));
        TRY(room(ctx, "flag_count", 1, &b.flag_count));
    }
    if (fast || score32)
    {
        // FP32 twins: fused emission coefficients per state, log transition costs per model
        PinVec<StateParamsF> stf = ctx->pinned<StateParamsF>("stf");
        PinVec<float4> trf = ctx->pinned<float4>("trf");
        if (!stf.resize(model_src.size() * N_STATES) || !trf.resize(model_src.size()))
        {
            ps_set_error(ctx, "out of host memory staging the models");
            return PS_E_INTERNAL;
        }
        const double l2p = std::log(2 * M_PI);
        for (size_t q = 0; q < model_src.size(); q++)
        {
            const ModelDev& md = models[q];
            for (int st = 0; st < N_STATES; st++)
            {
                const StateParams& sp = md.st[st];
                StateParamsF& f = stf[q * N_STATES + st];
                f.mu = (float)sp.lev_mean;
                f.a_s = (float)(-0.5 / (sp.lev_stdv * sp.lev_stdv));
                f.c_s = (float)(-0.5 * l2p - sp.log_lev + 0.5 * (sp.log_lambda - l2p) + P.lik_offset);
                f.mu2 = (float)sp.sd_mean;
                f.f_s = (float)(-0.5 * sp.sd_lambda / (sp.sd_mean * sp.sd_mean));
                f.pad0 = f.pad1 = f.pad2 = 0.f;
            }
            trf[q] = make_float4((float)md.lskip, (float)md.lstay, (float)md.lext, (float)md.lins);
        }
        StateParamsF* d_stf; float4* d_trf;
        TRY(up(ctx, "stf", stf.data(), stf.size(), &d_stf));
        TRY(up(ctx, "trf", trf.data(), trf.size(), &d_trf));
        TRY(room(ctx, "levf", (size_t)std::max<long long>(n_levels, 1) + 1, &b.levf));   // + 1: k_score_f32 requests one row ahead
        b.stf = d_stf; b.trf = d_trf;
        // Re-score threshold of the FAST mode, per EVENT of the mutation's region (k_flag multiplies by the region's
        // event count).  PS_FAST_PAIR_ERR bounds the absolute error of one rebased FP32 (mutation, event) delta
        // (measured: scripts/fast_error.py -> profiles/r2_fast_error.txt); a total over E events is then off by at
        // most E * PS_FAST_PAIR_ERR, and every mutation whose FP32 total is above -tau = -E * PS_FAST_PAIR_ERR / 1e-4
        // is re-scored exactly -- so what keeps its FP32 value is within 1e-4 RELATIVE of the reference
        // (BASELINE.json north_star), with no absolute slack.
        b.tau = ctx->tau_override >= 0 ? ctx->tau_override : PS_FAST_PAIR_ERR * 1e4;
        b.tau_events = sharded ? total_events : 0;
    }
    if (score32)
    {
        // launch classes of k_score_f32: level records staged whole by one TMA bulk copy (short events) or read through
        // L1, plain ACGT regions or regions with invalid states
        const int ne = (int)ev.size();
        if (!s32_list.resize((size_t)std::max(ne, 1))) { ps_set_error(ctx, "out of host memory staging the batch"); return PS_E_INTERNAL; }
        auto cls = [&](const EvDesc& d) { return ((d.n0 <= PS_SCORE32_STAGE_LEVELS && !ctx->no_stage) ? 0 : 2) + (d.inv ? 1 : 0); };
        for (int e = 0; e < ne; e++) s32_count[cls(ev[e])]++;
        int at[4] = {0, s32_count[0], s32_count[0] + s32_count[1], s32_count[0] + s32_count[1] + s32_count[2]};
        for (int e = 0; e < ne; e++)
        {
            const int c = cls(ev[e]);
            s32_list[at[c]++] = e;
            if (c < 2) s32_max_n0 = std::max(s32_max_n0, ev[e].n0);
        }
        int* d_list;
        TRY(up(ctx, "s32_list", s32_list.data(), s32_list.size(), &d_list));
        d_s32_list = d_list;
    }
    return PS_OK;
}

#define MARK(i) CU(cudaEventRecord(ctx->tev[i], ctx->stream))
#define LAUNCHED() do { ctx->launches++; CU(cudaGetLastError()); } while (0)

int Job::run(bool full)
{
    const int nev = (int)ev.size();
    int maxN = 0;
    for (const EvDesc& d : ev) maxN = std::max(maxN, d.N);
    if (nev == 0) return PS_OK;
    MARK(PS_T_CENTRES);    // (pre-call band centres are planned on the host, see plan_event)
    MARK(PS_T_FORWARD);
    if (score32)
    {
        int maxn0 = 0;
        for (const EvDesc& d : ev) maxn0 = std::max(maxn0, d.n0);
        k_rows<<<dim3(std::max(1, (maxn0 + 127) / 128), nev), 128, 0, ctx->stream>>>(b);      // (a batch of events without levels)
        LAUNCHED();
        Score32Args a;
        float* d_out32;
        TRY(room(ctx, "evbest32", ev.size(), &d_out32));
        CU(cudaMemsetAsync(d_out32, 0, ev.size() * sizeof(float), ctx->stream));
        a.out32 = d_out32;
        a.strip = (2 * b.realign_width + 1 + 8 + 3) & ~3;            // a band's rows + the row requested one step ahead, 16-byte multiple
        // Staged form (level records of one event brought into shared memory by one TMA bulk copy; one event per CTA): few
        // events, many warps per event -- the sweep of one event is a pipeline of blocks ~90 steps apart.  Persistent form
        // (level records through L1; a CTA walks through its events without a barrier): many events, 4 warps per CTA.
        // Measured equal per cell (profiles/r2_score32_stage_ab.txt); the persistent form saves the pipeline's fill and
        // drain per event, the staged form is kept for the small batches where an event has a CTA to itself anyway.
        const bool persist = nev >= 4 * ctx->sm_count || ctx->no_stage;
        int warps = persist ? 4 : nev >= 2 * ctx->sm_count ? 8 : 16;
        int warps_long = nev >= 4 * ctx->sm_count ? 4 : nev >= 2 * ctx->sm_count ? 8 : 16;
        if (ctx->s32_warps) { warps = ctx->s32_warps; warps_long = ctx->s32_warps; }
        const size_t strips_s = (size_t)(warps + 1) * a.strip * sizeof(float);
        const size_t staged = strips_s + (size_t)(s32_max_n0 + 1) * sizeof(LevelRecF);
        const size_t strips = (size_t)(warps_long + 1) * a.strip * sizeof(float);
        const int threads = 32 * warps, threads_long = 32 * warps_long;
        auto grid_for = [&](int count, int thr, size_t smem) {
            // persistent: as many CTAs as fit at once (registers: 64 per thread; shared memory), never more than events
            const int by_regs = 65536 / (80 * thr), by_smem = (int)((220 * 1024) / (smem + 1024));
            return std::max(1, std::min(count, ctx->sm_count * std::max(1, std::min(by_regs, by_smem))));
        };
        int off = 0;
        for (int c = 0; c < 4; c++)
        {
            a.list = d_s32_list + off; a.count = s32_count[c];
            off += s32_count[c];
            if (!a.count) continue;
            const bool inv = c & 1, short_ev = c < 2;
            if (short_ev && !persist)
            {
                if (inv) k_score_f32<true, true, 512><<<a.count, threads, staged, ctx->stream>>>(b, a);
                else k_score_f32<true, false, 512><<<a.count, threads, staged, ctx->stream>>>(b, a);
            }
            else
            {
                const int thr = short_ev ? threads : threads_long;
                const size_t sm = short_ev ? strips_s : strips;
                const int g = grid_for(a.count, thr, sm);
                if (thr <= 256)
                {
                    if (inv) k_score_f32<false, true, 256><<<g, thr, sm, ctx->stream>>>(b, a);
                    else k_score_f32<false, false, 256><<<g, thr, sm, ctx->stream>>>(b, a);
                }
                else
                {
                    if (inv) k_score_f32<false, true, 512><<<g, thr, sm, ctx->stream>>>(b, a);
                    else k_score_f32<false, false, 512><<<g, thr, sm, ctx->stream>>>(b, a);
                }
            }
            LAUNCHED();
        }
        TRY(room(ctx, "evbest", ev.size(), &d_evbest));
        k_score_f32_out<<<(nev + 127) / 128, 128, 0, ctx->stream>>>(d_out32, d_evbest, nev);
        LAUNCHED();
        MARK(PS_T_BACKWARD); MARK(PS_T_BACKTRACE); MARK(PS_T_JOIN); MARK(PS_T_MUTSCORE); MARK(PS_T_REDUCE); MARK(PS_T_D2H);
        return PS_OK;
    }
    // wavefront fill: one CTA per (event, direction), forward and reverse in one launch (grid.y = 2),
    // one launch per width class
    {
        if (ctx->trace)
            fprintf(stderr, "[ps] fill: %d events: %d at 160 threads, %d at 192, %d at %d, %d serial; band cells %lld\n", nev,
                    fill_count[0], fill_count[1], fill_count[2], fill_threads[2], fill_count[3], n_band);
        const int dirs = full ? 2 : 1;
        {
            int maxn0 = 0;
            for (const EvDesc& d : ev) maxn0 = std::max(maxn0, d.n0);
            k_strips<<<dim3(((maxN + CW - 1) / CW + 1 + 127) / 128, nev, dirs), 128, 0, ctx->stream>>>(b);
            LAUNCHED();
            k_rows<<<dim3(std::max(1, (maxn0 + 127) / 128), nev), 128, 0, ctx->stream>>>(b);      // (a batch of events without levels)
            LAUNCHED();
        }
        // the majority class on the main stream, the others beside it on the side stream
        const int off1 = fill_count[0], off2 = off1 + fill_count[1], off3 = off2 + fill_count[2];
        const bool forked = fill_count[0] > 0 && fill_count[1] + fill_count[2] + fill_count[3] > 0;
        cudaStream_t other = forked ? ctx->side : ctx->stream;
        if (forked)
        {
            CU(cudaEventRecord(ctx->fork_ev, ctx->stream));
            CU(cudaStreamWaitEvent(ctx->side, ctx->fork_ev, 0));
        }
        // 8-deep rings (2 x 8 x 16 B per thread) + next strip record (11 x 16 B) for the point-to-point classes
        const size_t smem160 = std::max<size_t>(54 * 160, 2 * b.RS) * sizeof(double);
        const size_t smem192 = std::max<size_t>(54 * 192, 2 * b.RS) * sizeof(double);
        if (ctx->fill2 && off3 > 0)
        {
            // every wavefront-capable event (the three width classes are one list: k_fill2 has no width classes)
            Fill2Args fa;
            fa.list = b.fill_list; fa.count = off3; fa.dirs = dirs;
            fa.slots = (2 * b.realign_width + 1 + 8 + 3) & ~3;
            const int warps = 4;
            const size_t smem = (size_t)(warps + 1) * fa.slots * sizeof(double2);
            const int per_sm = std::max(1, std::min(4, (int)((220 * 1024) / (smem + 1024))));
            const int grid = std::max(1, std::min(off3 * dirs, ctx->sm_count * per_sm));
            bool any_inv = false;
            for (const EvDesc& d : ev) any_inv = any_inv || d.inv;
            if (any_inv) k_fill2<true><<<grid, 32 * warps, smem, ctx->stream>>>(b, fa);
            else k_fill2<false><<<grid, 32 * warps, smem, ctx->stream>>>(b, fa);
            LAUNCHED();
            k_fill_best<<<dim3(off3, dirs), 256, 0, ctx->stream>>>(b, fa);
            LAUNCHED();
            if (fill_count[3]) { k_fill<160, PS_FILL_MINB><<<dim3(fill_count[3], dirs), 32, smem160, ctx->stream>>>(b, off3); LAUNCHED(); }
        }
        else
        {
        if (fill_count[2])
        {
            const size_t smem = std::max<size_t>(40 * 512, 2 * b.RS) * sizeof(double);
            k_fill<512, 1><<<dim3(fill_count[2], dirs), fill_threads[2], smem, other>>>(b, off2);
            LAUNCHED();
        }
        if (fill_count[1]) { k_fill<192, 2><<<dim3(fill_count[1], dirs), 192, smem192, other>>>(b, off1); LAUNCHED(); }
        if (fill_count[3]) { k_fill<160, PS_FILL_MINB><<<dim3(fill_count[3], dirs), 32, smem160, other>>>(b, off3); LAUNCHED(); }
        if (fill_count[0]) { k_fill<160, PS_FILL_MINB><<<dim3(fill_count[0], dirs), 160, smem160, ctx->stream>>>(b, 0); LAUNCHED(); }
        }
        if (forked)
        {
            CU(cudaEventRecord(ctx->join_ev, ctx->side));
            CU(cudaStreamWaitEvent(ctx->stream, ctx->join_ev, 0));
        }
    }
    MARK(PS_T_BACKWARD);   // (reverse fill shares the launch above; kept as a phase marker)
    MARK(PS_T_BACKTRACE);
    {
        int max_n0 = 0;
        for (const EvDesc& d : ev) max_n0 = std::max(max_n0, d.n0);
        // one warp per event, up to BT_WARPS events per CTA: 8 B value + 4 B source per level and event in shared memory
        const int budget = 160 * 1024 / 12;
        int wpc = BT_WARPS;
        while (wpc > 1 && (long long)wpc * ((max_n0 + 1) & ~1) > budget) wpc >>= 1;
        int smem_levels = std::min((max_n0 + 1) & ~1, (budget / wpc) & ~1);
        k_backtrace<<<(nev + wpc - 1) / wpc, 32 * BT_WARPS, (size_t)wpc * smem_levels * 12, ctx->stream>>>(b, smem_levels, wpc);
    }
    LAUNCHED();
    MARK(PS_T_JOIN);
    if (full && n_tasks > 0)
    {
        {
            dim3 grid((maxN + cen_pad + 1 + 127) / 128, nev);
            k_centres<<<grid, 128, 0, ctx->stream>>>(b, b.cen_new, 0);
            LAUNCHED();
        }
        if (dev_points)
        {
            k_points<<<dim3((maxN + 127) / 128, (unsigned)regs.size()), 128, 0, ctx->stream>>>(b);
            LAUNCHED();
        }
        {
            dim3 grid((maxN + 31) / 32, nev);
            k_join<<<grid, 256, 0, ctx->stream>>>(b);
            LAUNCHED();
        }
        // the sum over events: score[m] = bias + sum_e delta(m, e) in event order.  Event-sharded jobs combine the ranks'
        // sums on the stream: ordered (rank r continues the running sums of the ranks before it: bit-identical to one
        // GPU) or by one all-reduce of partial sums.
        const unsigned rblk = (unsigned)((n_muts + 127) / 128);
        auto reduce_stage = [&](int from_list, double* out) -> int {
            const RegTabDev* rt = (const RegTabDev*)d_regtab;
            const int nr = (int)regs.size();
            if (!sharded || ctx->comm_ranks <= 1)
            {
                k_reduce<<<rblk, 128, 0, ctx->stream>>>(b, rt, nr, n_muts, bias, from_list, nullptr, out);
                LAUNCHED();
                return PS_OK;
            }
            if (ctx->comm_ordered)
            {
                const bool first = ctx->comm_rank == 0, last = ctx->comm_rank == ctx->comm_ranks - 1;
                if (!first) TRY(psi_comm_recv_prev(ctx, out, (size_t)n_muts));
                k_reduce<<<rblk, 128, 0, ctx->stream>>>(b, rt, nr, n_muts, bias, from_list, first ? nullptr : out, out);
                LAUNCHED();
                if (!last) TRY(psi_comm_send_next(ctx, out, (size_t)n_muts));
                TRY(psi_comm_bcast_last(ctx, out, (size_t)n_muts));
            }
            else
            {
                if (from_list) CU(cudaMemsetAsync(out, 0, (size_t)n_muts * sizeof(double), ctx->stream));
                k_reduce<<<rblk, 128, 0, ctx->stream>>>(b, rt, nr, n_muts, 0.0, from_list, nullptr, out);
                LAUNCHED();
                TRY(psi_comm_allreduce_sum(ctx, out, (size_t)n_muts));
                k_add_scalar<<<rblk, 128, 0, ctx->stream>>>(out, n_muts, bias);
                LAUNCHED();
            }
            return PS_OK;
        };
        MARK(PS_T_MUTSCORE);
        {
            // previous-column ring: shared memory when 2W+2 doubles per thread fit, else global scratch
            const size_t ring = (size_t)(2 * b.scoring_width + 2) * sizeof(double);
            const int threads = 128;
            const bool in_smem = ring * 128 <= 96 * 1024;            // the smem ring is laid out for 128 threads
            long long blocks = std::min<long long>((n_tasks + threads - 1) / threads, (long long)ctx->sm_count * 16);
            if (fast)
            {
                // pass 1: every pair in rebased FP32; pass 2: exact FP64 for the mutations that matter
                k_mutscore_rows_f32<<<(unsigned)blocks, threads, 0, ctx->stream>>>(b);
                LAUNCHED();
                TRY(reduce_stage(0, b.scores));
                CU(cudaMemsetAsync(b.flag_count, 0, sizeof(int), ctx->stream));
                k_flag<<<(unsigned)((n_muts + 127) / 128), 128, 0, ctx->stream>>>(b, n_muts, 1);
                LAUNCHED();
                const unsigned rblocks = (unsigned)std::min<long long>(blocks, (long long)ctx->sm_count * 4);
                // few flagged pairs (the usual case): one warp per pair; many: one thread per pair.  The count only
                // exists on the device, so both are launched and one of them returns at once.
                if (b.warp_limit > 0)
                {
                    k_mutscore_warp<true><<<(unsigned)ctx->sm_count * 8, threads, ring * 8, ctx->stream>>>(b);
                    LAUNCHED();
                }
                if (in_smem) k_mutscore<true, true><<<rblocks, threads, ring * 128, ctx->stream>>>(b);
                else
                {
                    b.scratch_slots = (long long)rblocks * threads;
                    TRY(room(ctx, "scratch", (size_t)b.scratch_slots * (2 * b.scoring_width + 2), &b.scratch));
                    k_mutscore<false, true><<<rblocks, threads, 0, ctx->stream>>>(b);
                }
                LAUNCHED();
                MARK(PS_T_REDUCE);
                if (sharded && ctx->comm_ranks > 1)
                {
                    // every rank flagged the same mutations (the totals are the same everywhere); their exact sums travel
                    // through a second dense array and replace the FP32 totals
                    TRY(room(ctx, "scores2", (size_t)n_muts, &d_scores2));
                    TRY(reduce_stage(1, d_scores2));
                    k_merge_flagged<<<rblk, 128, 0, ctx->stream>>>(b, d_scores2, 0.0);
                    LAUNCHED();
                }
                else TRY(reduce_stage(1, b.scores));
            }
            else
            {
                if (b.warp_limit > 0 && n_tasks <= b.warp_limit)
                    k_mutscore_warp<false><<<(unsigned)std::min<long long>((n_tasks + 3) / 4, (long long)ctx->sm_count * 8), threads, ring * 8, ctx->stream>>>(b);
                else if (in_smem) k_mutscore<true, false><<<(unsigned)blocks, threads, ring * 128, ctx->stream>>>(b);
                else
                {
                    b.scratch_slots = blocks * threads;
                    TRY(room(ctx, "scratch", (size_t)b.scratch_slots * (2 * b.scoring_width + 2), &b.scratch));
                    k_mutscore<false, false><<<(unsigned)blocks, threads, 0, ctx->stream>>>(b);
                }
                LAUNCHED();
                MARK(PS_T_REDUCE);
                TRY(reduce_stage(0, b.scores));
            }
        }
    }
    else { MARK(PS_T_MUTSCORE); MARK(PS_T_REDUCE); }
    TRY(room(ctx, "evbest", ev.size(), &d_evbest));
    k_event_scores<<<(nev + 127) / 128, 128, 0, ctx->stream>>>(b, d_evbest);
    LAUNCHED();
    MARK(PS_T_D2H);
    return PS_OK;
}

int Job::download_enqueue()
{
    const size_t nl = (size_t)n_levels, ne = ev.size();
    rs = ctx->pinned<int>("refstart"); re = ctx->pinned<int>("refend");
    best = ctx->pinned<double>("evbest"); msc = ctx->pinned<double>("mscores");
    if (!rs.resize(ne) || !re.resize(ne) || !best.resize(ne) || !msc.resize((size_t)n_muts))
    {
        ps_set_error(ctx, "out of host memory staging the results");
        return PS_E_INTERNAL;
    }
    if (score32)
    {
        // scores only: the events keep their alignments (poreseq/_poreseqcpp.pyx:273-276)
        if (ne) CU(cudaMemcpyAsync(best.data(), d_evbest, ne * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h_bytes += (long long)(ne * sizeof(double));
        have_scores = false;
        MARK(PS_T_TOTAL);
        return PS_OK;
    }
    if (nl && !scores_only)
    {
        CU(cudaMemcpyAsync(ref_align.data(), b.ref_align, nl * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ref_like.data(), b.ref_like, nl * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (ne)
    {
        CU(cudaMemcpyAsync(ri_empty.data(), b.ri_empty, ne * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(rs.data(), b.refstart, ne * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(re.data(), b.refend, ne * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(best.data(), d_evbest, ne * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    ctx->d2h_bytes += (long long)((scores_only ? 0 : 2 * nl * sizeof(double)) + ne * (3 * sizeof(int) + sizeof(double)));
    have_scores = want_muts && n_muts && n_tasks;
    if (have_scores) ctx->d2h_bytes += (long long)((size_t)n_muts * sizeof(double));
    if (have_scores)
        CU(cudaMemcpyAsync(msc.data(), b.scores, (size_t)n_muts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MARK(PS_T_TOTAL);
    return PS_OK;
}

int Job::finish(std::vector<double>* align_scores, std::vector<double>* mut_scores)
{
    const size_t ne = ev.size();
    CU((cudaError_t)ps_stream_wait(ctx));
    // scatter the realigned events back into their regions
    std::vector<HostEvent*> hev;
    hev.reserve(ne);
    for (ps_region* R : regs)
        for (HostEvent& he : R->events) hev.push_back(&he);
    if (!score32 && !scores_only) ps_parallel_for((int)ne, [&](int e) {
        const EvDesc& d = ev[e];
        if (!d.usable) return;
        HostEvent& he = *hev[e];
        std::copy(ref_align.data() + d.lev_off, ref_align.data() + d.lev_off + d.n0, he.ref_align.begin());
        std::copy(ref_like.data() + d.lev_off, ref_like.data() + d.lev_off + d.n0, he.ref_like.begin());
        he.ri_empty = ri_empty[e] != 0;
        he.refstart = rs[e]; he.refend = re[e];
        // ref_index stays on the device; the host rebuilds it from ref_align (same arithmetic, cpp/EventData.h:110-169)
        // if this event is staged again or ViterbiMutate looks at it
        if (he.ri_empty) { he.ref_index.clear(); he.ri_stale = false; }
        else he.ri_stale = true;
    });
    if (align_scores)
    {
        align_scores->resize(ne);
        for (size_t k = 0; k < ne; k++) (*align_scores)[k] = std::max(best[k], 0.0);   // cpp/Alignment.h:127-130
    }
    if (mut_scores)
    {
        if (have_scores) mut_scores->assign(msc.data(), msc.data() + n_muts);
        else mut_scores->assign((size_t)n_muts, bias);
    }
    if (raw_scores)                               // straight into the caller's array
    {
        if (have_scores) memcpy(raw_scores, msc.data(), (size_t)n_muts * sizeof(double));
        else std::fill(raw_scores, raw_scores + n_muts, bias);
    }
    // timings + algorithmic cell counts (wide = band cells of the usable events, both
    // directions when the reverse fill ran; narrow = (|mut|+5) x band rows per (mutation, usable event))
    for (int i = 0; i < PS_T_TOTAL; i++)
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->tev[i], ctx->tev[i + 1]);
        ctx->timing[i] = ms;
    }
    float tot = 0;
    cudaEventElapsedTime(&tot, ctx->tev[PS_T_H2D], ctx->tev[PS_T_TOTAL]);
    ctx->timing[PS_T_TOTAL] = tot;
    ctx->wide_cells = want_muts ? 2 * wide_cells_fwd : wide_cells_fwd;
    ctx->narrow_cells = narrow_cells;
    return PS_OK;
}

// ------------------------------------------------------------------------------------------
// drivers shared by the C entry points
// One job = build (host staging) + upload + kernels + result copies, all enqueued on the context's
// stream by job_begin; job_end waits for the stream and scatters the results into the regions.
// Between the two the host is free (e.g. to stage the next batch on another context).
static int job_begin(ps_ctx* ctx, const std::vector<ps_region*>& regs, const std::vector<MutSpec>* muts, double bias, bool score32 = false,
                     int shard_total_events = 0, bool scores_only = false, bool owns_regions = false)
{
    TRY(ctx->init());
    CU(cudaSetDevice(ctx->device));
    if (ctx->pending) { ps_set_error(ctx, "a batch is already in flight on this context"); return PS_E_ARG; }
    for (ps_region* R : regs)
        if (R->bases.size() < 5) { ps_set_error(ctx, "sequences shorter than 5 bases are not supported"); return PS_E_ARG; }
    {
        // the events of a batch are staged (and realigned) in parallel: a region can be in a batch only once
        std::vector<ps_region*> seen(regs);
        std::sort(seen.begin(), seen.end());
        if (std::adjacent_find(seen.begin(), seen.end()) != seen.end())
        {
            ps_set_error(ctx, "the same region handle appears twice in one batch");
            return PS_E_ARG;
        }
    }
    Job* job = new Job(ctx);
    job->regs = regs;
    job->want_muts = muts != nullptr;
    if (muts) job->muts = *muts;
    job->bias = bias;
    job->score32 = score32 && muts == nullptr;
    job->scores_only = scores_only;
    job->sharded = shard_total_events > 0 && muts != nullptr;
    job->total_events = shard_total_events;
    // FAST flags mutations by their TOTAL over all events; an event shard (bias 0, ps_score_mutations_partial) only has
    // its part of the sum, so partial sums are always exact
    job->fast = ctx->precision == PS_PRECISION_FAST && muts != nullptr && bias != 0.0;
    const bool trace = ctx->trace;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    ctx->h2d_bytes = 0; ctx->d2h_bytes = 0;
    int rc = job->build();
    const double t1 = now();
    if (!rc) rc = (cudaEventRecord(ctx->tev[PS_T_H2D], ctx->stream) == cudaSuccess) ? PS_OK : PS_E_CUDA;
    if (!rc) rc = job->upload();
    const double t2 = now();
    if (!rc) rc = job->run(muts != nullptr);
    if (!rc) rc = job->download_enqueue();
    const double t3 = now();
    if (trace) fprintf(stderr, "[ps] host: build %.2f ms, upload(enqueue) %.2f, kernels+d2h(enqueue) %.2f\n", t1 - t0, t2 - t1, t3 - t2);
    if (rc) { delete job; return rc; }
    if (owns_regions) job->owned = regs;               // (only a job that was enqueued takes the regions over)
    ctx->pending = job;
    return PS_OK;
}

static int job_end(ps_ctx* ctx, std::vector<double>* align_scores, std::vector<double>* mut_scores, double* raw_scores = nullptr)
{
    Job* job = (Job*)ctx->pending;
    if (!job) { ps_set_error(ctx, "no batch in flight on this context"); return PS_E_ARG; }
    job->raw_scores = raw_scores;
    CU(cudaSetDevice(ctx->device));
    ctx->pending = nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    int rc = job->finish(align_scores, mut_scores);
    if (ctx->trace) fprintf(stderr, "[ps] host: wait+scatter %.2f ms (device total %.2f)\n", now() - t0, ctx->timing[PS_T_TOTAL]);
    delete job;
    return rc;
}

// Band storage a region's events need (upper estimate: widest common wavefront class, both directions
// when mutations are scored).  Decides how a large job is cut into consecutive sub-batches.
static double region_band_bytes(const ps_region* R, bool full)
{
    const double J = (R->states.size() + 1) / 2;
    double cells = 0;
    for (const HostEvent& he : R->events) cells += (J + (he.n0 + 1) / 2 + 2) * 192.0 * 4.0;
    return cells * (full ? 33.0 : 17.0);
}

static int run_job(ps_ctx* ctx, const std::vector<ps_region*>& regs, const std::vector<MutSpec>* muts,
                   std::vector<double>* align_scores, std::vector<double>* mut_scores, double bias = -1e-6, int shard_total_events = 0)
{
    TRY(ctx->init());
    // a job whose band matrices would not fit (e.g. FindMutations: seeds x events wide fills of a 10 kb
    // region) runs as consecutive sub-batches of whole regions; results are concatenated in region order
    // At most 32 GB of band storage per sub-batch: growing the band buffers is a synchronous cudaFree + cudaMalloc whose
    // cost rises with the size -- the 30 seed realignments of a 10 kb x 30x region (1800 events) took 1411 ms as three
    // sub-batches of 76 GB and 432 ms as eight of 30 GB (240 events = one wave of the 192-thread fill class each),
    // measured on a B200 (gpurun_out -> profiles/r1_consensus_10kb_trace.txt).
    double budget = std::min(0.40 * (double)ctx->total_mem, 32e9);
    if (ctx->band_budget > 0) budget = ctx->band_budget;      // bytes; tests force the split path with it
    double total = 0;
    for (const ps_region* R : regs) total += region_band_bytes(R, muts != nullptr);
    if (total <= budget || regs.size() <= 1)
    {
        TRY(job_begin(ctx, regs, muts, bias, false, shard_total_events));
        return job_end(ctx, align_scores, mut_scores);
    }
    if (align_scores) align_scores->clear();
    if (mut_scores) mut_scores->clear();
    size_t a = 0;
    while (a < regs.size())
    {
        size_t b = a;
        double acc = 0;
        while (b < regs.size() && (b == a || acc + region_band_bytes(regs[b], muts != nullptr) <= budget))
            acc += region_band_bytes(regs[b++], muts != nullptr);
        std::vector<ps_region*> sub(regs.begin() + a, regs.begin() + b);
        std::vector<MutSpec> subm;
        if (muts) subm.assign(muts->begin() + a, muts->begin() + b);
        std::vector<double> as, ms;
        TRY(job_begin(ctx, sub, muts ? &subm : nullptr, bias));
        TRY(job_end(ctx, align_scores ? &as : nullptr, mut_scores ? &ms : nullptr));
        if (align_scores) align_scores->insert(align_scores->end(), as.begin(), as.end());
        if (mut_scores) mut_scores->insert(mut_scores->end(), ms.begin(), ms.end());
        a = b;
    }
    return PS_OK;
}

int ps_run_alignments(ps_ctx* ctx, const std::vector<ps_region*>& regs,
                      std::vector<std::vector<double>>* scores, std::vector<std::vector<double>>* likes)
{
    std::vector<double> flat;
    TRY(run_job(ctx, regs, nullptr, &flat, nullptr));
    size_t e = 0;
    if (scores) scores->assign(regs.size(), std::vector<double>());
    if (likes) likes->assign(regs.size(), std::vector<double>());
    for (size_t r = 0; r < regs.size(); r++)
    {
        const ps_region* R = regs[r];
        if (scores) (*scores)[r].assign(flat.begin() + e, flat.begin() + e + R->events.size());
        e += R->events.size();
        if (!likes) continue;
        // per-base likelihood profile, cpp/MakeMutations.cpp:168-189 (events in order)
        std::vector<double>& lk = (*likes)[r];
        lk.assign(R->bases.size(), 0.0);
        const int N = (int)R->states.size();
        for (const HostEvent& he : R->events)
        {
            double carried = 0;
            int at = 1;
            for (int j = 0; j < he.n0; j++)
                if (he.ref_align[j] > 0)
                {
                    for (int k = at; k < he.ref_align[j]; k++) lk[k + 1] += carried;
                    carried = he.ref_like[j];
                    at = (int)he.ref_align[j];
                }
            for (int k = at; k < N + 3; k++) lk[k + 1] += carried;
        }
    }
    return PS_OK;
}

// PSAlign.ScoreEvents (poreseq/_poreseqcpp.pyx:263-276): ScoreAlignments(data, NULL) -- one score per event, nothing else
// leaves the call (the realignment is not propagated, A.3-10), so the regions' events are left as they were.
// FAST precision: k_score_f32 (log-space FP32, no matrices, no backtrace); EXACT: the FP64 forward fill.
int ps_run_event_scores(ps_ctx* ctx, const std::vector<ps_region*>& regs, std::vector<double>* flat)
{
    TRY(ctx->init());
    bool f32 = ctx->precision == PS_PRECISION_FAST;
    for (const ps_region* R : regs)
        for (const HostModel& hm : R->models)                       // k_score_f32 folds implicit moves into the floor: needs log p <= 0
            if (!(hm.trans[0] <= 1.0) || !(hm.trans[3] <= 1.0)) f32 = false;
    if (f32)
    {
        TRY(job_begin(ctx, regs, nullptr, -1e-6, true));
        return job_end(ctx, flat, nullptr);
    }
    // exact: the job realigns the events in place; put the alignments back afterwards
    struct Keep { LevelVec ref_align, ref_like; bool ri_empty; int refstart, refend; };
    std::vector<Keep> keep;
    for (ps_region* R : regs)
        for (HostEvent& he : R->events)
        {
            he.ensure_refs();
            keep.push_back(Keep{he.ref_align, he.ref_like, he.ri_empty, he.refstart, he.refend});
        }
    const int rc = run_job(ctx, regs, nullptr, flat, nullptr);
    size_t k = 0;
    for (ps_region* R : regs)
        for (HostEvent& he : R->events)
        {
            Keep& q = keep[k++];
            he.ref_align.swap(q.ref_align); he.ref_like.swap(q.ref_like);
            he.ri_empty = q.ri_empty; he.refstart = q.refstart; he.refend = q.refend;
            he.ri_stale = !he.ri_empty;
            if (he.ri_empty) he.ref_index.clear();
        }
    return rc;
}

int ps_refine_region(ps_region* R, int* nbases)                  // PSAlign.Refine, poreseq/_poreseqcpp.pyx:437-472
{
    std::vector<HostMut> v = ps_point_mutations(R);
    TRY(ps_score_mutation_list(R, v));
    int nb = 0;
    TRY(ps_make_mutation_list(R, v, &nb));
    if (nbases) *nbases = nb;
    return PS_OK;
}

std::vector<HostMut> ps_point_mutations(const ps_region* R)      // cpp/FindMutations.cpp:191-234
{
    static const char acgt[] = "ACGT";
    std::vector<HostMut> v;
    v.reserve(R->states.size() * 8);
    for (int i = 0; i < (int)R->states.size(); i++)
    {
        const char here = R->bases[i];
        HostMut m;
        m.start = i; m.score = -1e-6;
        m.orig.assign(1, here);
        v.push_back(m);                                   // deletion
        for (int j = 0; j < 4; j++)
            if (acgt[j] != here) { m.mut.assign(1, acgt[j]); v.push_back(m); }    // substitutions
        m.orig.clear();
        for (int j = 0; j < 4; j++) { m.mut.assign(1, acgt[j]); v.push_back(m); } // insertions
    }
    return v;
}

int ps_score_mutation_list(ps_region* R, std::vector<HostMut>& muts, double bias, int shard_total_events)
{
    std::vector<MutSpec> per(1);
    per[0].list = &muts;
    std::vector<double> sc;
    ps_ctx* ctx = R->ctx;
    const bool fast = ctx->precision == PS_PRECISION_FAST && bias != 0.0;
    // FAST keeps approximate values for the clearly negative scores.  MakeMutations sorts the whole list with an UNSTABLE
    // std::sort (cpp/MakeMutations.cpp:83) before it drops the negative ones, so where two exactly tied non-negative
    // scores end up can depend on the negative keys around them.  Such ties get the whole list re-scored exactly, from
    // the alignments the first pass started from (the pass realigns the events in place).
    struct Seed { LevelVec ref_align; bool ri_empty; int refstart, refend; };
    std::vector<Seed> seeds;
    if (fast)
    {
        seeds.resize(R->events.size());
        for (size_t e = 0; e < R->events.size(); e++)
        {
            const HostEvent& he = R->events[e];
            seeds[e].ref_align = he.ref_align; seeds[e].ri_empty = he.ri_empty;
            seeds[e].refstart = he.refstart; seeds[e].refend = he.refend;
        }
    }
    TRY(run_job(ctx, std::vector<ps_region*>(1, R), &per, nullptr, &sc, bias, shard_total_events));
    if (fast)
    {
        std::vector<double> keep;
        for (double v : sc) if (v >= 0) keep.push_back(v);
        std::sort(keep.begin(), keep.end());
        if (std::adjacent_find(keep.begin(), keep.end()) != keep.end())
        {
            for (size_t e = 0; e < R->events.size(); e++)
            {
                HostEvent& he = R->events[e];
                he.ref_align = seeds[e].ref_align; he.ri_empty = seeds[e].ri_empty;
                he.refstart = seeds[e].refstart; he.refend = seeds[e].refend;
                he.ri_stale = !he.ri_empty;
                if (he.ri_empty) he.ref_index.clear();
            }
            ctx->precision = PS_PRECISION_EXACT;
            const int rc = run_job(ctx, std::vector<ps_region*>(1, R), &per, nullptr, &sc, bias, shard_total_events);
            ctx->precision = PS_PRECISION_FAST;
            if (rc) return rc;
            ctx->exact_reruns++;
        }
    }
    for (size_t i = 0; i < muts.size(); i++) muts[i].score = sc[i];
    return PS_OK;
}

// The same for several regions in ONE job (the consensus loop in lockstep over regions, ps_lockstep.cu): lists[r] is
// scored against regs[r].  FAST precision: a region whose list ends up with exactly tied non-negative scores is re-scored
// exactly from the alignments it started with (see ps_score_mutation_list), in one second job for all such regions.
int ps_score_mutation_lists(ps_ctx* ctx, const std::vector<ps_region*>& regs, const std::vector<std::vector<HostMut>*>& lists)
{
    if (regs.empty()) return PS_OK;
    std::vector<MutSpec> per(regs.size());
    for (size_t r = 0; r < regs.size(); r++) per[r].list = lists[r];
    const bool fast = ctx->precision == PS_PRECISION_FAST;
    struct Seed { LevelVec ref_align; bool ri_empty; int refstart, refend; };
    std::vector<std::vector<Seed>> seeds;
    if (fast)
    {
        seeds.resize(regs.size());
        ps_parallel_for((int)regs.size(), [&](int r) {
            seeds[r].resize(regs[r]->events.size());
            for (size_t e = 0; e < regs[r]->events.size(); e++)
            {
                const HostEvent& he = regs[r]->events[e];
                seeds[r][e].ref_align = he.ref_align; seeds[r][e].ri_empty = he.ri_empty;
                seeds[r][e].refstart = he.refstart; seeds[r][e].refend = he.refend;
            }
        });
    }
    std::vector<double> sc;
    TRY(run_job(ctx, regs, &per, nullptr, &sc));
    size_t at = 0;
    std::vector<ps_region*> again;
    std::vector<std::vector<HostMut>*> again_lists;
    for (size_t r = 0; r < regs.size(); r++)
    {
        std::vector<HostMut>& v = *lists[r];
        for (size_t i = 0; i < v.size(); i++) v[i].score = sc[at + i];
        at += v.size();
        if (!fast) continue;
        std::vector<double> keep;
        for (const HostMut& m : v) if (m.score >= 0) keep.push_back(m.score);
        std::sort(keep.begin(), keep.end());
        if (std::adjacent_find(keep.begin(), keep.end()) == keep.end()) continue;
        for (size_t e = 0; e < regs[r]->events.size(); e++)
        {
            HostEvent& he = regs[r]->events[e];
            he.ref_align = seeds[r][e].ref_align; he.ri_empty = seeds[r][e].ri_empty;
            he.refstart = seeds[r][e].refstart; he.refend = seeds[r][e].refend;
            he.ri_stale = !he.ri_empty;
            if (he.ri_empty) he.ref_index.clear();
        }
        again.push_back(regs[r]); again_lists.push_back(lists[r]);
    }
    if (!again.empty())
    {
        ctx->precision = PS_PRECISION_EXACT;
        const int rc = ps_score_mutation_lists(ctx, again, again_lists);
        ctx->precision = PS_PRECISION_FAST;
        if (rc) return rc;
        ctx->exact_reruns += (long long)again.size();
    }
    return PS_OK;
}

static bool by_score_desc(const HostMut& a, const HostMut& b) { return a.score > b.score; }

// cpp/MakeMutations.cpp:74-146.  std::sort with the same ordering predicate on the same element
// order reproduces the reference's (unstable) tie placement.
// One pass (:83-139): accepts in score order, invalidates neighbours, shifts later starts; `deferred` receives the
// mutations the reference would score again (:104-108, :142-143) -- the caller does so when there are more than ten.
void ps_make_mutation_pass(ps_region* R, std::vector<HostMut> muts, int* changed_out, std::vector<HostMut>* deferred)
{
    const int spacing = 10;
    int changed = 0;
    deferred->clear();
    std::sort(muts.begin(), muts.end(), by_score_desc);
    while (!muts.empty() && muts.back().score < 0) muts.pop_back();
    *changed_out = 0;
    if (muts.empty()) return;
    // the accepted edits are applied to a working copy of the bases; the region's sequence (and its 5-mer states) is
    // replaced once after the loop -- nothing inside the loop reads the states (0.08 ms per accept at 10 kb otherwise)
    std::string seq = R->bases;
    for (size_t i = 0; i < muts.size(); i++)
    {
        HostMut& a = muts[i];
        if (a.score < 0) { deferred->push_back(a); continue; }
        seq = ps_apply_mutation(seq, a.start, a.orig, a.mut);
        changed += (int)std::max(a.orig.size(), a.mut.size());
        for (size_t j = i + 1; j < muts.size(); j++)
        {
            HostMut& c = muts[j];
            const int lo = std::max(a.start, c.start);
            const int hi = (int)std::min((size_t)a.start + a.mut.size(), (size_t)c.start + c.mut.size());
            if (lo < hi + spacing && c.score > 0) { c.score = -1; continue; }
            if ((size_t)c.start >= (size_t)a.start + a.orig.size())
                c.start += (int)(a.mut.size() - a.orig.size());
        }
    }
    R->set_sequence(seq);
    *changed_out = changed;
}

int ps_make_mutation_list(ps_region* R, std::vector<HostMut> muts, int* nbases)
{
    int changed = 0;
    std::vector<HostMut> deferred;
    ps_make_mutation_pass(R, std::move(muts), &changed, &deferred);
    if (deferred.size() > 10)
    {
        int more = 0;
        TRY(ps_score_mutation_list(R, deferred));
        TRY(ps_make_mutation_list(R, deferred, &more));
        changed += more;
    }
    *nbases = changed;
    return PS_OK;
}

ps_region* ps_shadow_region(const ps_region* R)
{
    ps_region* nd = new ps_region();
    nd->ctx = R->ctx;
    nd->bases = R->bases; nd->states = R->states;
    nd->models = R->models;
    nd->params = R->params;
    nd->events.resize(R->events.size());
    for (size_t e = 0; e < R->events.size(); e++)
    {
        const HostEvent& src = R->events[e];
        HostEvent& he = nd->events[e];
        he.n0 = src.n0; he.model = src.model; he.complement = src.complement;
        he.ri_empty = src.ri_empty; he.refstart = src.refstart; he.refend = src.refend;
        he.ref_align = src.ref_align;                     // remapped by the caller, then update_refs
        he.ref_like.assign((size_t)src.n0, 0.0);          // written by the realignment
        he.ext_mean = src.mean.data(); he.ext_stdv = src.stdv.data();
        he.ext_levrec = src.levrec.size() == (size_t)src.n0 * 3 ? src.levrec.data() : nullptr;
    }
    return nd;
}

void ps_region::rng_seed(unsigned seed)
{
    memset(&rng_data, 0, sizeof rng_data);
    memset(rng_state, 0, sizeof rng_state);
    initstate_r(seed, rng_state, sizeof rng_state, &rng_data);
    own_rng = true;
}

double ps_region::next_uniform()
{
    int32_t v = 0;
    random_r(&rng_data, &v);
    return v / (double(RAND_MAX) + 1);
}

void ps_region::set_sequence(const std::string& s)
{
    bases = s;
    states = ps_states_of(s);
}

// ------------------------------------------------------------------------------------------
// C-ABI
extern "C" {

const char* ps_version(void) { return "poreseq_b200 0.1 (sm_100a, fp64-exact)"; }

ps_ctx* ps_create(int device)
{
    ps_ctx* ctx = new ps_ctx();
    ctx->device = device;
    ctx->trace = getenv("PORESEQ_B200_TRACE") != nullptr;
    ctx->no_warp = getenv("PORESEQ_B200_NO_WARP") != nullptr;
    ctx->sw_host = getenv("PORESEQ_B200_SW_HOST") != nullptr;
    ctx->no_stage = getenv("PORESEQ_B200_NO_STAGE") != nullptr;
    if (const char* e = getenv("PORESEQ_B200_VIT_CLUSTER")) ctx->vit_cluster = atoi(e) != 0;
    if (const char* e = getenv("PORESEQ_B200_FILL2")) ctx->fill2 = atoi(e) != 0;
    if (const char* e = getenv("PORESEQ_B200_BLOCKING_WAIT")) ctx->blocking_wait = atoi(e) != 0;
    if (const char* e = getenv("PORESEQ_B200_CONSENSUS")) ctx->threads_consensus = std::string(e) == "threads";
    if (const char* e = getenv("PORESEQ_B200_GROUPS")) ctx->consensus_groups = std::max(1, atoi(e));
    if (const char* e = getenv("PORESEQ_B200_S32_WARPS")) ctx->s32_warps = std::max(2, std::min(atoi(e), PS_SCORE32_MAX_WARPS));
    if (const char* e = getenv("PORESEQ_B200_BAND_BUDGET")) ctx->band_budget = atof(e);
    if (const char* e = getenv("PORESEQ_B200_TAU")) ctx->tau_override = atof(e);
    return ctx;                       // CUDA is initialised lazily (fork-safe)
}

void ps_destroy(ps_ctx* ctx) { delete ctx; }

const char* ps_last_error(ps_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

long long ps_launch_count(ps_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ps_set_precision(ps_ctx* ctx, int mode)
{
    if (!ctx || (mode != PS_PRECISION_EXACT && mode != PS_PRECISION_FAST)) return PS_BAD_ARGS(ctx, "ps_set_precision");
    ctx->precision = mode;
    return PS_OK;
}

int ps_last_timing(ps_ctx* ctx, double* ms)
{
    if (!ctx || !ms) return PS_BAD_ARGS(ctx, "ps_last_timing");
    for (int i = 0; i < PS_T_COUNT; i++) ms[i] = ctx->timing[i];
    return PS_OK;
}

int ps_last_bytes(ps_ctx* ctx, long long* h2d, long long* d2h)
{
    if (!ctx) return PS_BAD_ARGS(ctx, "ps_last_bytes");
    if (h2d) *h2d = ctx->h2d_bytes;
    if (d2h) *d2h = ctx->d2h_bytes;
    return PS_OK;
}

int ps_last_cells(ps_ctx* ctx, double* wide, double* narrow)
{
    if (!ctx) return PS_BAD_ARGS(ctx, "ps_last_cells");
    if (wide) *wide = ctx->wide_cells;
    if (narrow) *narrow = ctx->narrow_cells;
    return PS_OK;
}

ps_region* ps_region_create(ps_ctx* ctx, const char* bases, int len, const ps_params* params)
{
    if (!ctx || !bases || len < 0) { ps_set_error(ctx, "ps_region_create: bad arguments"); return nullptr; }
    ps_region* R = new ps_region();
    R->ctx = ctx;
    R->set_sequence(std::string(bases, len));
    if (params) R->params = *params;
    else { R->params.lik_offset = 4.5; R->params.scoring_width = 150; R->params.realign_width = 300; R->params.verbose = 0; }
    return R;
}

void ps_region_destroy(ps_region* r) { delete r; }

void ps_regions_destroy(ps_region* const* regions, int n_regions)
{
    if (!regions || n_regions <= 0) return;
    ps_parallel_for(n_regions, [&](int k) { delete regions[k]; });
}

int ps_region_add_event(ps_region* R, int n0, const double* mean, const double* stdv, const double* ref_align,
                        const double* ref_like, const double* level_mean, const double* level_stdv,
                        const double* sd_mean, const double* sd_stdv, int complement, double prob_skip,
                        double prob_stay, double prob_extend, double prob_insert, const char* seq2d)
{
    if (!R) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_region_add_event");
    ps_ctx* ctx = R->ctx;
    if (n0 < 0 || (n0 > 0 && (!mean || !stdv || !ref_align || !ref_like)) || !level_mean || !level_stdv || !sd_mean || !sd_stdv)
    {
        ps_set_error(ctx, "ps_region_add_event: bad arguments");
        return PS_E_ARG;
    }
    HostModel hm;
    memset(&hm, 0, sizeof hm);
    memcpy(hm.raw[0], level_mean, sizeof hm.raw[0]);
    memcpy(hm.raw[1], level_stdv, sizeof hm.raw[1]);
    memcpy(hm.raw[2], sd_mean, sizeof hm.raw[2]);
    memcpy(hm.raw[3], sd_stdv, sizeof hm.raw[3]);
    hm.trans[0] = prob_skip; hm.trans[1] = prob_stay; hm.trans[2] = prob_extend; hm.trans[3] = prob_insert;
    int mi = -1;
    for (size_t q = 0; q < R->models.size(); q++)
        if (memcmp(&R->models[q], &hm, sizeof hm) == 0) { mi = (int)q; break; }
    if (mi < 0) { mi = (int)R->models.size(); R->models.push_back(hm); }
    HostEvent he;
    he.n0 = n0;
    he.model = mi;
    he.complement = complement != 0;
    he.mean.assign(mean, mean + n0);
    he.stdv.assign(stdv, stdv + n0);
    he.ref_align.assign(ref_align, ref_align + n0);
    he.ref_like.assign(ref_like, ref_like + n0);
    if (seq2d) he.seq2d = seq2d;
    he.update_refs();
    R->events.push_back(std::move(he));
    return PS_OK;
}

int ps_region_add_events(ps_region* R, int n_events, const int* n0, const double* mean, const double* stdv,
                         const double* ref_align, const double* ref_like, const int* model_index, int n_models,
                         const double* models, const double* probs, const int* complement, const char* const* seq2d)
{
    if (!R) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_region_add_events");
    ps_ctx* ctx = R->ctx;
    if (n_events < 0 || n_models < 0 || (n_events > 0 && (!n0 || !model_index || !models || !probs || n_models == 0)))
    {
        ps_set_error(ctx, "ps_region_add_events: bad arguments");
        return PS_E_ARG;
    }
    // model table -> region models (de-duplicated against the ones already there)
    std::vector<int> map(n_models, -1);
    for (int q = 0; q < n_models; q++)
    {
        HostModel hm;
        memcpy(hm.raw, models + (size_t)q * 4 * PS_N_STATES, sizeof hm.raw);
        memcpy(hm.trans, probs + (size_t)q * 4, sizeof hm.trans);
        for (size_t k = 0; k < R->models.size() && map[q] < 0; k++)
            if (memcmp(&R->models[k], &hm, sizeof hm) == 0) map[q] = (int)k;
        if (map[q] < 0) { map[q] = (int)R->models.size(); R->models.push_back(hm); }
    }
    const size_t first = R->events.size();
    std::vector<size_t> at(n_events + 1, 0);
    for (int e = 0; e < n_events; e++)
    {
        if (n0[e] < 0 || model_index[e] < 0 || model_index[e] >= n_models || (n0[e] > 0 && (!mean || !stdv || !ref_align || !ref_like)))
        {
            ps_set_error(ctx, "ps_region_add_events: bad event %d", e);
            return PS_E_ARG;
        }
        at[e + 1] = at[e] + (size_t)n0[e];
    }
    R->events.resize(first + n_events);
    for (int e = 0; e < n_events; e++)                   // (too little work per event to farm out)
    {
        const int n = n0[e];
        const size_t a = at[e];
        HostEvent& he = R->events[first + e];
        he.n0 = n;
        he.model = map[model_index[e]];
        he.complement = complement ? complement[e] != 0 : false;
        he.mean.assign(mean + a, mean + a + n);
        he.stdv.assign(stdv + a, stdv + a + n);
        he.ref_align.assign(ref_align + a, ref_align + a + n);
        he.ref_like.assign(ref_like + a, ref_like + a + n);
        if (seq2d && seq2d[e]) he.seq2d = seq2d[e];
        he.update_refs();
    }
    return PS_OK;
}

int ps_regions_create(ps_ctx* ctx, int n_regions, const ps_region_desc* desc, ps_region** out)
{
    if (!ctx || n_regions < 0 || (n_regions > 0 && (!desc || !out))) return PS_BAD_ARGS(ctx, "ps_regions_create");
    std::vector<int> rc(n_regions, PS_OK);
    for (int k = 0; k < n_regions; k++) out[k] = nullptr;
    ps_parallel_for(n_regions, [&](int k) {
        const ps_region_desc& d = desc[k];
        ps_region* R = ps_region_create(ctx, d.bases, d.len, &d.params);
        if (!R) { rc[k] = PS_E_ARG; return; }
        rc[k] = ps_region_add_events(R, d.n_events, d.n0, d.mean, d.stdv, d.ref_align, d.ref_like, d.model_index,
                                     d.n_models, d.models, d.probs, d.complement, d.seq2d);
        if (rc[k]) { delete R; return; }
        out[k] = R;
    });
    for (int k = 0; k < n_regions; k++)
        if (rc[k])
        {
            for (int q = 0; q < n_regions; q++) { delete out[q]; out[q] = nullptr; }
            ps_set_error(ctx, "ps_regions_create: region %d was refused", k);
            return rc[k];
        }
    return PS_OK;
}

int ps_region_set_params(ps_region* R, const ps_params* p)
{
    if (!R || !p) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_region_set_params");
    R->params = *p;
    return PS_OK;
}

int ps_region_num_events(ps_region* R) { return R ? (int)R->events.size() : PS_E_ARG; }
int ps_region_sequence_length(ps_region* R) { return R ? (int)R->bases.size() : PS_E_ARG; }

int ps_region_get_sequence(ps_region* R, char* out, int cap)
{
    if (!R || !out) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_region_get_sequence");
    if ((int)R->bases.size() + 1 > cap) { ps_set_error(R->ctx, "sequence buffer too small"); return PS_E_CAPACITY; }
    memcpy(out, R->bases.c_str(), R->bases.size() + 1);
    return PS_OK;
}

int ps_region_get_event_align(ps_region* R, int e, double* ref_align, double* ref_like)
{
    if (!R || e < 0 || e >= (int)R->events.size()) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_region_get_event_align");
    const HostEvent& he = R->events[e];
    if (ref_align) std::copy(he.ref_align.begin(), he.ref_align.end(), ref_align);
    if (ref_like) std::copy(he.ref_like.begin(), he.ref_like.end(), ref_like);
    return PS_OK;
}

int ps_score_alignments(ps_region* R, double* scores, double* likes)
{
    if (!R || !scores) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_score_alignments");
    std::vector<std::vector<double>> sc, lk;
    TRY(ps_run_alignments(R->ctx, std::vector<ps_region*>(1, R), &sc, likes ? &lk : nullptr));
    std::copy(sc[0].begin(), sc[0].end(), scores);
    if (likes) for (size_t k = 0; k < lk[0].size(); k++) likes[k] += lk[0][k];
    return PS_OK;
}

int ps_score_events(ps_region* R, double* scores)
{
    if (!R || !scores) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_score_events");
    std::vector<double> flat;
    TRY(ps_run_event_scores(R->ctx, std::vector<ps_region*>(1, R), &flat));
    std::copy(flat.begin(), flat.end(), scores);
    return PS_OK;
}

int ps_score_events_batch(ps_region* const* regions, int n_regions, double* scores)
{
    if (n_regions < 0 || (n_regions > 0 && (!regions || !scores))) return PS_BAD_ARGS(nullptr, "ps_score_events_batch");
    if (n_regions == 0) return PS_OK;
    for (int k = 0; k < n_regions; k++)
        if (!regions[k] || regions[k]->ctx != regions[0]->ctx) return PS_BAD_ARGS(regions[0] ? regions[0]->ctx : nullptr, "ps_score_events_batch");
    std::vector<double> flat;
    TRY(ps_run_event_scores(regions[0]->ctx, std::vector<ps_region*>(regions, regions + n_regions), &flat));
    std::copy(flat.begin(), flat.end(), scores);
    return PS_OK;
}

static std::vector<HostMut> gather_muts(int n, const int* start, const char* const* orig, const char* const* mut,
                                        const double* scores)
{
    std::vector<HostMut> v(n);
    for (int i = 0; i < n; i++)
    {
        v[i].start = start[i];
        v[i].orig = orig[i] ? orig[i] : "";
        v[i].mut = mut[i] ? mut[i] : "";
        v[i].score = scores ? scores[i] : -1e-6;
    }
    return v;
}

int ps_score_mutations(ps_region* R, int n, const int* start, const char* const* orig, const char* const* mut, double* scores)
{
    if (!R || n < 0 || (n > 0 && (!start || !orig || !mut || !scores))) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_score_mutations");
    std::vector<HostMut> v = gather_muts(n, start, orig, mut, nullptr);
    TRY(ps_score_mutation_list(R, v));
    for (int i = 0; i < n; i++) scores[i] = v[i].score;
    return PS_OK;
}

int ps_score_mutations_partial(ps_region* R, int n, const int* start, const char* const* orig, const char* const* mut, double* partial)
{
    if (!R || n < 0 || (n > 0 && (!start || !orig || !mut || !partial))) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_score_mutations_partial");
    std::vector<HostMut> v = gather_muts(n, start, orig, mut, nullptr);
    TRY(ps_score_mutation_list(R, v, 0.0));
    for (int i = 0; i < n; i++) partial[i] = v[i].score;
    return PS_OK;
}

// The events of ONE region split across the ranks of ps_comm_init: this handle holds this rank's block of the events (in
// event order: rank 0 the first block, and so on), every rank passes the same mutations and gets the same, complete
// scores back.  The sums over events are combined on the GPUs (ps_comm.cu).  poreseq/Variant.py:71-76.
int ps_score_mutations_sharded(ps_region* R, int n, const int* start, const char* const* orig, const char* const* mut, double* scores)
{
    if (!R || n < 0 || (n > 0 && (!start || !orig || !mut || !scores))) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_score_mutations_sharded");
    ps_ctx* ctx = R->ctx;
    if (!ctx->comm) { ps_set_error(ctx, "ps_score_mutations_sharded: the context has no communicator (ps_comm_init)"); return PS_E_ARG; }
    if (R->events.empty()) { ps_set_error(ctx, "ps_score_mutations_sharded: every rank must hold at least one event"); return PS_E_ARG; }
    TRY(ctx->init());
    CU(cudaSetDevice(ctx->device));
    // events of the region over all ranks (the FAST mode's re-score threshold is per event of the TOTAL)
    double* d_count;
    TRY(room(ctx, "shard_count", 1, &d_count));
    double local = (double)R->events.size(), total = 0;
    CU(cudaMemcpyAsync(d_count, &local, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    TRY(psi_comm_allreduce_sum(ctx, d_count, 1));
    CU(cudaMemcpyAsync(&total, d_count, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU((cudaError_t)ps_stream_wait(ctx));
    std::vector<HostMut> v = gather_muts(n, start, orig, mut, nullptr);
    TRY(ps_score_mutation_list(R, v, -1e-6, (int)total));
    for (int i = 0; i < n; i++) scores[i] = v[i].score;
    return PS_OK;
}

static int emit_points(ps_ctx* ctx, const std::vector<HostMut>& v, int cap, int* n, int* start, char* orig, char* mut, double* scores)
{
    if (n) *n = (int)v.size();
    if ((int)v.size() > cap) { ps_set_error(ctx, "output capacity %d < %zu point mutations", cap, v.size()); return PS_E_CAPACITY; }
    for (size_t i = 0; i < v.size(); i++)
    {
        if (start) start[i] = v[i].start;
        if (orig) orig[i] = v[i].orig.empty() ? 0 : v[i].orig[0];
        if (mut) mut[i] = v[i].mut.empty() ? 0 : v[i].mut[0];
        if (scores) scores[i] = v[i].score;
    }
    return PS_OK;
}

int ps_find_point_mutations(ps_region* R, int cap, int* n, int* start, char* orig, char* mut)
{
    if (!R) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_find_point_mutations");
    return emit_points(R->ctx, ps_point_mutations(R), cap, n, start, orig, mut, nullptr);
}

// Enumerates the point edits of one region in FindPointMutations order straight into the flat
// output arrays (no per-edit objects), returning how many there are.
static long long write_points(const ps_region* R, long long at, long long cap, int* start, char* orig, char* mut)
{
    long long n = 0;
    for (int i = 0; i < (int)R->states.size(); i++)
    {
        const char here = R->bases[i];
        for (int kind = 0; kind < 9; kind++)
        {
            // kind 0: deletion; 1..4: substitution by ACGT[kind-1] (skipping the base itself); 5..8: insertion
            if (kind >= 1 && kind <= 4 && "ACGT"[kind - 1] == here) continue;
            if (at + n < cap)
            {
                if (start) start[at + n] = i;
                if (orig) orig[at + n] = kind <= 4 ? here : 0;
                if (mut) mut[at + n] = kind == 0 ? 0 : "ACGT"[(kind - 1) & 3];
            }
            n++;
        }
    }
    return n;
}

int ps_score_points_batch_begin(ps_region* const* regions, int n_regions, int cap, int* n_out, long long* off_out,
                                int* start, char* orig, char* mut)
{
    if (!regions || n_regions <= 0 || !regions[0]) return PS_E_ARG;
    ps_ctx* ctx = regions[0]->ctx;
    std::vector<ps_region*> regs(regions, regions + n_regions);
    std::vector<MutSpec> per(n_regions);
    long long at = 0;
    std::vector<long long> offs(n_regions), cnt(n_regions);
    for (int r = 0; r < n_regions; r++)
    {
        if (!regs[r] || regs[r]->ctx != ctx) { ps_set_error(ctx, "all regions of a batch must belong to one context"); return PS_E_ARG; }
        per[r].points = true;
        // 8 edits per state, 9 where the base is not ACGT (no substitution is skipped)
        const ps_region* R = regs[r];
        long long n = 8 * (long long)R->states.size();
        for (size_t i = 0; i < R->states.size(); i++)
        {
            const char ch = R->bases[i];
            if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') n++;
        }
        offs[r] = at; cnt[r] = n;
        if (n_out) n_out[r] = (int)n;
        if (off_out) off_out[r] = at;
        at += n;
    }
    if (at > cap) { ps_set_error(ctx, "output capacity %d < %lld point mutations", cap, at); return PS_E_CAPACITY; }
    if (start || orig || mut)
        ps_parallel_for(n_regions, [&](int r) { write_points(regs[r], offs[r], cap, start, orig, mut); });
    return job_begin(ctx, regs, &per, -1e-6);
}

// PSAlign.ScorePoints straight from the caller's arrays (poreseq/_poreseqcpp.pyx:278-308: PythonToAlignData ->
// FindPointMutations -> ScoreMutations -> the scores; the realignment is dropped, nothing is written back).  No region
// handles: the level arrays are read where they lie (mean, stdv once into the pinned staging records, ref_align once for
// the band centres), no copy of them is kept, no alignment comes back from the device.
static ps_region* borrowed_region(ps_ctx* ctx, const ps_region_desc& d)
{
    if (!d.bases || d.len < 5 || d.n_events < 0 || d.n_models <= 0 || !d.n0 || !d.model_index || !d.models || !d.probs) return nullptr;
    ps_region* R = new ps_region();
    R->ctx = ctx;
    R->params = d.params;
    R->set_sequence(std::string(d.bases, d.len));
    std::vector<int> map(d.n_models, -1);
    for (int q = 0; q < d.n_models; q++)
    {
        HostModel hm;
        memcpy(hm.raw, d.models + (size_t)q * 4 * PS_N_STATES, sizeof hm.raw);
        memcpy(hm.trans, d.probs + (size_t)q * 4, sizeof hm.trans);
        for (size_t k = 0; k < R->models.size() && map[q] < 0; k++)
            if (memcmp(&R->models[k], &hm, sizeof hm) == 0) map[q] = (int)k;
        if (map[q] < 0) { map[q] = (int)R->models.size(); R->models.push_back(hm); }
    }
    R->events.resize(d.n_events);
    size_t at = 0;
    for (int e = 0; e < d.n_events; e++)
    {
        const int n = d.n0[e];
        if (n < 0 || d.model_index[e] < 0 || d.model_index[e] >= d.n_models || (n > 0 && (!d.mean || !d.stdv || !d.ref_align))) { delete R; return nullptr; }
        HostEvent& he = R->events[e];
        he.n0 = n;
        he.model = map[d.model_index[e]];
        he.complement = d.complement ? d.complement[e] != 0 : false;
        he.ext_mean = d.mean + at; he.ext_stdv = d.stdv + at;
        he.update_refs_from(d.ref_align + at);
        at += (size_t)n;
    }
    return R;
}

int ps_score_points_direct_begin(ps_ctx* ctx, int n_regions, const ps_region_desc* desc, int cap, int* n_out, long long* off_out,
                                 int* start, char* orig, char* mut)
{
    if (!ctx || n_regions <= 0 || !desc) return PS_BAD_ARGS(ctx, "ps_score_points_direct_begin");
    std::vector<ps_region*> regs(n_regions, nullptr);
    ps_parallel_for(n_regions, [&](int r) { regs[r] = borrowed_region(ctx, desc[r]); });
    auto drop = [&] { for (ps_region* R : regs) delete R; };
    for (int r = 0; r < n_regions; r++)
        if (!regs[r]) { drop(); ps_set_error(ctx, "ps_score_points_direct: region %d was refused", r); return PS_E_ARG; }
    std::vector<MutSpec> per(n_regions);
    long long at = 0;
    std::vector<long long> offs(n_regions);
    for (int r = 0; r < n_regions; r++)
    {
        per[r].points = true;
        const ps_region* R = regs[r];
        long long n = 8 * (long long)R->states.size();
        for (size_t i = 0; i < R->states.size(); i++)
        {
            const char ch = R->bases[i];
            if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') n++;
        }
        offs[r] = at;
        if (n_out) n_out[r] = (int)n;
        if (off_out) off_out[r] = at;
        at += n;
    }
    if (at > cap) { drop(); ps_set_error(ctx, "output capacity %d < %lld point mutations", cap, at); return PS_E_CAPACITY; }
    if (start || orig || mut)
        ps_parallel_for(n_regions, [&](int r) { write_points(regs[r], offs[r], cap, start, orig, mut); });
    const int rc = job_begin(ctx, regs, &per, -1e-6, false, 0, /*scores_only=*/true, /*owns_regions=*/true);
    if (rc) drop();
    return rc;
}

int ps_score_points_direct_end(ps_ctx* ctx, double* scores)
{
    if (!ctx) return PS_BAD_ARGS(ctx, "ps_score_points_direct_end");
    return job_end(ctx, nullptr, nullptr, scores);
}

int ps_score_points_direct(ps_ctx* ctx, int n_regions, const ps_region_desc* desc, int cap, int* n_out, long long* off_out,
                           int* start, char* orig, char* mut, double* scores)
{
    TRY(ps_score_points_direct_begin(ctx, n_regions, desc, cap, n_out, off_out, start, orig, mut));
    return ps_score_points_direct_end(ctx, scores);
}

int ps_score_points_batch_end(ps_ctx* ctx, double* scores)
{
    if (!ctx) return PS_BAD_ARGS(ctx, "ps_score_points_batch_end");
    return job_end(ctx, nullptr, nullptr, scores);
}

int ps_score_points_batch(ps_region* const* regions, int n_regions, int cap, int* n_out, long long* off_out,
                          int* start, char* orig, char* mut, double* scores)
{
    TRY(ps_score_points_batch_begin(regions, n_regions, cap, n_out, off_out, start, orig, mut));
    return ps_score_points_batch_end(regions[0]->ctx, scores);
}

int ps_score_points(ps_region* R, int cap, int* n, int* start, char* orig, char* mut, double* scores)
{
    if (!R) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_score_points");
    int count = 0;
    long long off = 0;
    ps_region* one = R;
    int rc = ps_score_points_batch(&one, 1, cap, &count, &off, start, orig, mut, scores);
    if (n) *n = count;
    return rc;
}

int ps_make_mutations(ps_region* R, int n, const int* start, const char* const* orig, const char* const* mut,
                      const double* scores, int* nbases)
{
    if (!R || n < 0 || (n > 0 && (!start || !orig || !mut || !scores))) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_make_mutations");
    int nb = 0;
    TRY(ps_make_mutation_list(R, gather_muts(n, start, orig, mut, scores), &nb));
    if (nbases) *nbases = nb;
    return PS_OK;
}

int ps_refine(ps_region* R, int* nbases)
{
    if (!R) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_refine");
    return ps_refine_region(R, nbases);
}

int ps_seq_to_states(const char* seq, int len, int* states)
{
    if (!seq || len < 0) return PS_E_ARG;
    std::vector<int> st = ps_states_of(std::string(seq, len));
    if (states) std::copy(st.begin(), st.end(), states);
    return (int)st.size();
}

} // extern "C"
