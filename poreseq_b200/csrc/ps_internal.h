// ps_internal.h -- host-side objects behind the opaque handles of include/poreseq_b200.h.
#pragma once
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/poreseq_b200.h"

struct DevBuf { void* p = nullptr; size_t cap = 0; };

// Growable array in pinned host memory (owned by the context, reused across calls): the staging
// area the H2D / D2H copies run from at full PCIe rate.
struct PinBuf { void* p = nullptr; size_t cap = 0; };
template <class T>
struct PinVec
{
    PinBuf* buf = nullptr;
    size_t n = 0;
    T* data() const { return (T*)buf->p; }
    size_t size() const { return n; }
    T& operator[](size_t i) const { return ((T*)buf->p)[i]; }
    T* begin() const { return (T*)buf->p; }
    T* end() const { return (T*)buf->p + n; }
    bool reserve(size_t count);                 // false on allocation failure
    bool resize(size_t count) { if (!reserve(count)) return false; n = count; return true; }
    void push_back(const T& v) { if (n + 1 > buf->cap / sizeof(T)) reserve((n + 1) * 2); ((T*)buf->p)[n++] = v; }
    void append(const T* src, size_t count) { reserve(n + count); memcpy((T*)buf->p + n, src, count * sizeof(T)); n += count; }
    void fill(size_t count, const T& v) { reserve(n + count); for (size_t i = 0; i < count; i++) ((T*)buf->p)[n + i] = v; n += count; }
};
bool ps_pin_reserve(PinBuf* b, size_t bytes, size_t keep);
template <class T> bool PinVec<T>::reserve(size_t count) { return ps_pin_reserve(buf, count * sizeof(T), n * sizeof(T)); }

struct ps_ctx
{
    int device = 0;
    bool ready = false;
    int sm_count = 0;
    size_t total_mem = 0;
    int precision = 0;                        // PS_PRECISION_EXACT / PS_PRECISION_FAST
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;              // runs the minority launch classes of the wide fill beside the main one
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    cudaEvent_t wait_ev = nullptr;            // blocking-sync event (see ps_stream_wait)
    bool blocking_wait = false;               // wait for the stream by sleeping on an event instead of spinning (PORESEQ_B200_BLOCKING_WAIT,
                                              // and every context ps_consensus_batch drives from its own threads: more threads than cores)
    cudaEvent_t tev[PS_T_COUNT + 1];
    double timing[PS_T_COUNT] = {0};
    double wide_cells = 0, narrow_cells = 0;
    long long h2d_bytes = 0, d2h_bytes = 0;   // bytes the last batch copied to / from the device
    long long launches = 0;
    long long exact_reruns = 0;               // FAST-mode lists re-scored exactly because of tied non-negative scores
    void* pending = nullptr;                  // the batch in flight between *_begin and *_end (a Job)
    // event-shard communicator (ps_comm.cu): the ranks of one box that share the events of a region
    void* comm = nullptr;                     // ncclComm_t
    int comm_rank = 0, comm_ranks = 1;
    bool comm_ordered = true;                 // ordered chain (bit-exact) or one all-reduce
    // environment knobs, read ONCE when the context is created (ps_create): a per-call getenv() races with a
    // setenv() of the host program
    bool trace = false;                       // PORESEQ_B200_TRACE: phase timings on stderr
    bool no_warp = false;                     // PORESEQ_B200_NO_WARP: never use the warp-per-pair exact kernel
    bool sw_host = false;                     // PORESEQ_B200_SW_HOST: FindMutations' Smith-Waterman maps on the host
    double band_budget = 0;                   // PORESEQ_B200_BAND_BUDGET: bytes of band storage per sub-batch (0: default)
    int consensus_groups = 0;                 // PORESEQ_B200_GROUPS: lockstep groups of ps_consensus_batch running side by side (0: chosen by batch size)
    bool threads_consensus = false;           // PORESEQ_B200_CONSENSUS=threads: ps_consensus_batch as regions in flight on threads instead of lockstep (A/B)
    bool no_stage = false;                    // PORESEQ_B200_NO_STAGE: k_score_f32 reads level records through L1 instead of a TMA-staged copy (A/B)
    bool fill2 = false;                       // PORESEQ_B200_FILL2=1: the wide fill on the warp-block schedule (k_fill2) instead of k_fill (A/B)
    bool vit_cluster = true;                  // PORESEQ_B200_VIT_CLUSTER=0: the Viterbi chain on one CTA instead of a cluster of 8 (A/B)
    int s32_warps = 0;                        // PORESEQ_B200_S32_WARPS: warps per CTA of k_score_f32 (0: chosen by batch size)
    double tau_override = -1;                 // PORESEQ_B200_TAU: FAST-mode re-score threshold (diagnostics; < 0: derived)
    std::string error;
    std::map<std::string, DevBuf> bufs;       // grow-only named device buffers, reused across calls
    std::map<std::string, PinBuf> pins;       // grow-only named pinned host buffers
    template <class T> PinVec<T> pinned(const char* name) { PinVec<T> v; v.buf = &pins[name]; v.n = 0; return v; }

    std::vector<ps_ctx*> group_ctx;           // ps_consensus_batch: the contexts of the lockstep groups 1.. (group 0 is this context); each has helpers of its own
    std::vector<ps_ctx*> helpers;             // ps_consensus_batch: further contexts (stream + buffers) on the same device, one per region in flight
    int init();                               // lazy CUDA initialisation
    int ensure(DevBuf& b, size_t bytes);
    ~ps_ctx();
};

struct HostModel                              // raw inputs of ModelData::setData/setParams
{
    double raw[4][PS_N_STATES];               // level_mean, level_stdv, sd_mean, sd_stdv
    double trans[4];                          // prob_skip, prob_stay, prob_extend, prob_insert
};

// Recycling allocator of the per-event level arrays.  A batch of regions is created, scored and destroyed per
// step; with the default allocator the 30-60 MB of level arrays freed by a destroyed batch are trimmed from the heap
// and faulted in again by the next one (2/3 of the time of ps_regions_create).  Blocks of up to 256 KB are kept on
// per-size free lists instead (1 KB size classes, at most PS_POOL_CAP bytes held), larger ones go to malloc.
void* ps_pool_alloc(size_t bytes);
void  ps_pool_free(void* p, size_t bytes) noexcept;
size_t ps_pool_held();                           // bytes on the free lists (tests)

template <class T>
struct PoolAlloc
{
    using value_type = T;
    PoolAlloc() = default;
    template <class U> PoolAlloc(const PoolAlloc<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(ps_pool_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t n) noexcept { ps_pool_free(p, n * sizeof(T)); }
    template <class U> bool operator==(const PoolAlloc<U>&) const { return true; }
    template <class U> bool operator!=(const PoolAlloc<U>&) const { return false; }
};
typedef std::vector<double, PoolAlloc<double>> LevelVec;

struct HostEvent                              // cpp/EventData.h:78-229
{
    int n0 = 0;
    int model = 0;
    bool complement = false;
    bool ri_empty = true;
    int refstart = -1, refend = -1;
    LevelVec mean, stdv, ref_align, ref_like, ref_index;
    LevelVec levrec;               // 3 doubles per level, the staged LevIn layout (mean, stdv, 3 log stdv), cached
    bool ri_stale = false;                    // ref_align was rewritten by a batch; ref_index is rebuilt from it on first use
    int staged = 0;                           // batches this event was staged for (the level records are cached from the second on)
    std::string seq2d;
    // borrowed level arrays (ps_score_points_direct: the caller's buffers for the duration of one call, nothing copied);
    // mean / stdv / ref_align / ref_like above stay empty then
    const double* ext_mean = nullptr;
    const double* ext_stdv = nullptr;
    const double* ext_levrec = nullptr;       // borrowed staged level records (3 doubles per level) of the event this one shadows
    void update_refs();
    void update_refs_from(const double* ra);  // the same from an array that is not this event's own
    void ensure_refs() { if (ri_stale) update_refs(); }
    void ensure_levrec();                     // log(stdv) etc. (cpp/EventData.h:218-220), cached
};

struct HostMut                                // cpp/AlignUtil.h:69-92 MutInfo / MutScore
{
    int start = 0;
    std::string orig, mut;
    double score = -1e-6;
};

struct ps_region                              // cpp/AlignData.h:24-34
{
    ps_ctx* ctx = nullptr;
    std::string bases;
    std::vector<int> states;
    std::vector<HostEvent> events;
    std::vector<HostModel> models;
    ps_params params;
    std::map<std::string, std::vector<double>> seqlikes;   // FindMutations cache (cpp/AlignData.h:34)
    std::vector<HostMut> found;                            // result of the last ps_find_mutations
    std::vector<std::string> viterbi;                      // result of the last ps_viterbi_mutate
    // ps_consensus: the reference runs one process per region, so every region's ViterbiMutate calls draw from a
    // rand() stream of their own that starts at glibc's default seed 1 (cpp/Viterbi.cpp:108, never seeded).  A region
    // with `own_rng` draws from a private copy of that generator (random_r: the same TYPE_3 sequence as rand()) instead
    // of the process-global one, so regions in flight side by side do not interleave their draws.
    bool own_rng = false;
    char rng_state[128];
    struct random_data rng_data;
    void rng_seed(unsigned seed);
    double next_uniform();                                 // rand() / (RAND_MAX + 1.0)
    std::vector<std::pair<std::string, std::string>> stage_log;   // ps_consensus: (stage, sequence after it)
    std::vector<int> stage_nbases;
    void set_sequence(const std::string& s);
};

void ps_set_error(ps_ctx* ctx, const char* fmt, ...);
// wait until everything enqueued on the context's stream is done; cudaError_t
int ps_stream_wait(ps_ctx* ctx);
// PS_E_ARG with a message naming the entry point (ctx may be null: the message is then what ps_last_error(NULL) returns)
#define PS_BAD_ARGS(ctx_, fn_) (ps_set_error((ctx_), "%s: bad arguments", (fn_)), PS_E_ARG)
// fn(i) for i in [0, n) on the library's host worker threads (PORESEQ_B200_THREADS, default
// min(cores, 8)); the caller takes part.  Used for the per-event staging work of a batch.
void ps_parallel_for(int n, const std::function<void(int)>& fn);
std::vector<int> ps_states_of(const std::string& bases);
std::string ps_apply_mutation(const std::string& bases, int start, const std::string& orig, const std::string& mut);
std::vector<HostMut> ps_point_mutations(const ps_region* R);
int ps_score_mutation_list(ps_region* R, std::vector<HostMut>& muts, double bias = -1e-6, int shard_total_events = 0);
int ps_make_mutation_list(ps_region* R, std::vector<HostMut> muts, int* nbases);
void ps_make_mutation_pass(ps_region* R, std::vector<HostMut> muts, int* changed_out, std::vector<HostMut>* deferred);
int ps_score_mutation_lists(ps_ctx* ctx, const std::vector<ps_region*>& regs, const std::vector<std::vector<HostMut>*>& lists);
// A copy of R for FindMutations' seed realignments: its own sequence, alignments and models, the level data (mean, stdv,
// staged level records) borrowed from R's events, which must outlive it and have their level records made (ensure_levrec)
ps_region* ps_shadow_region(const ps_region* R);
int ps_consensus_lockstep(ps_ctx* ctx, ps_region* const* regions, int n_regions, int reps, int point_width, int in_flight);

struct SWResult                               // cpp/swlib.h:25-33
{
    int score = 0;
    double accuracy = 0;
    std::vector<int> inds1, inds2;
};

SWResult psi_swfull(const std::string& s1, const std::string& s2);
// the same on the GPU for one sequence against many (ps_sw.cu); PS_E_ARG when a sequence is too long for it
int psi_swfull_batch(ps_ctx* ctx, const std::string& s1, const std::vector<std::string>& others, std::vector<SWResult>& out);
SWResult psi_map_alignments_with(ps_region* R, const std::string& newseq, SWResult al);
void psi_fillinds(SWResult& al);
void psi_pick_candidates(const std::string& bases, const std::vector<double>& base, const std::vector<std::string>& seeds,
                         const std::vector<const std::vector<double>*>& profs, std::vector<SWResult>& als, std::vector<HostMut>& found);
SWResult psi_map_alignments(ps_region* R, const std::string& newseq);
// forward fill + backtrace of every event of every region (ScoreAlignments); per-region score
// vectors and per-base likelihood profiles on request
int ps_run_alignments(ps_ctx* ctx, const std::vector<ps_region*>& regs,
                      std::vector<std::vector<double>>* scores, std::vector<std::vector<double>>* likes);
int ps_find_mutation_list(ps_region* R, const std::vector<std::string>& seeds, std::vector<HostMut>& found);
int ps_mutate_loop(ps_region* R, const std::vector<std::string>& seeds, int reps, int* totbases);
int ps_viterbi_list(ps_region* R, int nkeep, double skip_prob, double stay_prob, double mut_min, double mut_max,
                    std::vector<std::string>& out);
int ps_refine_region(ps_region* R, int* nbases);
// NCCL steps of the event-sharded sum, enqueued on the context's stream (ps_comm.cu)
int psi_comm_allreduce_sum(ps_ctx* ctx, double* buf, size_t count);
int psi_comm_recv_prev(ps_ctx* ctx, double* buf, size_t count);
int psi_comm_send_next(ps_ctx* ctx, const double* buf, size_t count);
int psi_comm_bcast_last(ps_ctx* ctx, double* buf, size_t count);
