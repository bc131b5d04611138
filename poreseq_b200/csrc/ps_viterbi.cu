// ps_viterbi.cu -- ViterbiMutate (cpp/Viterbi.cpp:239-426): per reference position, pooled
// trimmed-mean emissions over the reads, a 1024-state Viterbi + normalised forward step with
// 1/2/3-base advances and stays, then the best path or `nkeep` forward-weighted random back-samples,
// turned back into sequences.
//
// Split of work:
//   host   which levels of which read sit on each position (getrefstates, cpp/EventData.h:187-204),
//          their mean level / stdv and log(stdv) (libm), the skip/stop rule (:310-325), the glibc
//          rand() stream (:108), StatesToSequence (:171-237)
//   K8     k_vit_obs      1024 x n_reads pdfs per position, per-state sort + trimmed mean (:300-343)
//   K7     k_vit_chain    one 1024-thread CTA walks the positions: max-plus (Viterbi, first maximum
//                         wins) and sum-product (forward) over 84 predecessors + stay from shared
//                         memory, x exp(obs), normalise (:39-102)
//          k_vit_sample   one CTA per sample: T-row x fwd^atten, normalise, inverse CDF (:105-131)
//
// Exactness: emissions, the sort/trimmed mean and the Viterbi recursion are IEEE double in the
// reference's order, so liks / backptrs (hence the nkeep=0 path) are bit-identical.  The forward
// probabilities use the device exp() and a tree sum, the sampler the device pow(): they differ from
// glibc in the last ulps, which can move an inverse-CDF boundary by ~1e-16 -- a sampled state flips
// with probability ~1e-13 per draw.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <vector>

#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include "ps_internal.h"
#include "ps_types.cuh"

using namespace psdev;

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)
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

struct VitSlot { int model; int pad; double lvl, sd, logsd; };   // one read on one position

// ------------------------------------------------------------------------------------------
// K8: obs[t][state] = trimmed mean over the reads on position t of lognormpdf + logigpdf
// (no lik_offset, cpp/Viterbi.cpp:300-306).  One thread per (position, state); its per-read values
// sit in shared memory (vals[s * blockDim + tid]) for the ascending sort.
__global__ void k_vit_obs(const ModelDev* models, const VitSlot* slots, const int* slot_off, int n_pos,
                          double log2pi, double* obs, double* eobs)
{
    extern __shared__ double vals[];
    const int t = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int T = blockDim.x, tid = threadIdx.x;
    const int s0 = slot_off[t], nlik = slot_off[t + 1] - s0;
    for (int s = 0; s < nlik; s++)
    {
        const VitSlot sl = slots[s0 + s];
        const StateParams& p = models[sl.model].st[j];
        double d = (sl.lvl - p.lev_mean) / p.lev_stdv;
        double l = -0.5 * (d * d + log2pi) - p.log_lev;
        double g = (sl.sd - p.sd_mean) / p.sd_mean;
        l += 0.5 * (p.log_lambda - 3 * sl.logsd - log2pi - g * g * p.sd_lambda / sl.sd);
        vals[s * T + tid] = l;
    }
    double o;
    if (nlik > 1)
    {
        for (int a = 1; a < nlik; a++)              // insertion sort, ascending
        {
            double v = vals[a * T + tid];
            int b = a - 1;
            while (b >= 0 && vals[b * T + tid] > v) { vals[(b + 1) * T + tid] = vals[b * T + tid]; b--; }
            vals[(b + 1) * T + tid] = v;
        }
        int nskip = (int)floor(nlik * 0.25);
        if (nskip > nlik - 2) nskip = 0;
        double lik = 0.0;
        for (int k = nskip; k < nlik; k++) lik += vals[k * T + tid];
        o = lik / (nlik - nskip);
    }
    else o = vals[tid];
    obs[(size_t)t * N_STATES + j] = o;
    eobs[(size_t)t * N_STATES + j] = exp(o);
}

// ------------------------------------------------------------------------------------------
// K7: the chain.  liks/backptrs exactly as V_LIK::V_LIK; fwd normalised with a tree sum.
struct VitConst { double lsp[3], sp[3], stay_lik, stay_prob; };

__global__ void __launch_bounds__(1024) k_vit_chain(const double* obs, const double* eobs, int n_pos, VitConst vc,
                                                     double* fwd, int* backptr, double* last_liks)
{
    __shared__ double lk[2][N_STATES];
    __shared__ double fw[2][N_STATES];
    __shared__ double red[32];
    const int d = threadIdx.x, lane = d & 31, wid = d >> 5;
    lk[0][d] = 0.0;
    fw[0][d] = 1.0 / N_STATES;
    __syncthreads();
    for (int t = 0; t < n_pos; t++)
    {
        const double* pl = lk[t & 1];
        const double* pf = fw[t & 1];
        const double o = obs[(size_t)t * N_STATES + d];
        double best = NEG, f = 0.0;
        int ptr = -1;
#pragma unroll
        for (int j = 1; j <= 3; j++)
        {
            const double lsp = vc.lsp[j - 1], sp = vc.sp[j - 1];
            const double base = o + lsp;
            const int lowbits = d >> (2 * j), shift = 10 - 2 * j;
#pragma unroll 4
            for (int k = 0; k < (1 << (2 * j)); k++)
            {
                const int p = lowbits + (k << shift);
                const double l = base + pl[p];
                f += sp * pf[p];
                if (l > best) { best = l; ptr = p; }
            }
        }
        {
            const double l = o + vc.stay_lik + pl[d];
            if (l > best) { best = l; ptr = d; }
            f += vc.stay_prob * pf[d];
        }
        f *= eobs[(size_t)t * N_STATES + d];
        // normalise forward probabilities (tree sum over the block)
        double s = f;
        for (int o2 = 16; o2; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
        if (lane == 0) red[wid] = s;
        __syncthreads();
        if (wid == 0)
        {
            double w = red[lane];
            for (int o2 = 16; o2; o2 >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o2);
            if (lane == 0) red[0] = 1.0 / w;
        }
        __syncthreads();
        f *= red[0];
        lk[(t + 1) & 1][d] = best;
        fw[(t + 1) & 1][d] = f;
        fwd[(size_t)t * N_STATES + d] = f;
        backptr[(size_t)t * N_STATES + d] = ptr;
        __syncthreads();
    }
    last_liks[d] = lk[n_pos & 1][d];
}

// The same chain on a CLUSTER of 8 CTAs (thread-block clusters + distributed shared memory): the 1024 destination states
// are split over 8 SMs (128 threads each), every CTA keeps a full copy of the previous position's Viterbi scores and
// forward probabilities in its own shared memory, writes its 128 new values into all eight copies (st.shared::cluster
// through cluster.map_shared_rank) and the cluster meets once per position.  The single-CTA form is bound by ONE SM's FP64
// pipe (336 FP64 instructions per state and position x 32 warps: 6.5 us per position measured); here a position costs one
// warp per scheduler's worth of that plus the cluster barrier.  Per-state arithmetic and its order are unchanged; the
// normalising sum is the sum of the eight CTAs' tree sums (the forward probabilities were never bit-identical to the
// reference's index-order sum).
constexpr int VIT_CL = 8;

__global__ void __cluster_dims__(VIT_CL, 1, 1) __launch_bounds__(N_STATES / VIT_CL)
k_vit_chain_cluster(const double* obs, const double* eobs, int n_pos, VitConst vc, double* fwd, int* backptr, double* last_liks)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int PER = N_STATES / VIT_CL;                    // states per CTA = threads per CTA
    __shared__ double lk[2][N_STATES];
    __shared__ double fw[2][N_STATES];
    __shared__ double part[2][VIT_CL];
    __shared__ double red[PER / 32];
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int d = rank * PER + tid;                           // this thread's destination state
    for (int i = tid; i < N_STATES; i += PER) { lk[0][i] = 0.0; fw[0][i] = 1.0 / N_STATES; }
    // the eight copies of every array, as this thread sees them
    double* r_lk[VIT_CL]; double* r_fw[VIT_CL]; double* r_part[VIT_CL];
#pragma unroll
    for (int q = 0; q < VIT_CL; q++)
    {
        r_lk[q] = cluster.map_shared_rank(&lk[0][0], q);
        r_fw[q] = cluster.map_shared_rank(&fw[0][0], q);
        r_part[q] = cluster.map_shared_rank(&part[0][0], q);
    }
    cluster.sync();
    for (int t = 0; t < n_pos; t++)
    {
        const int cur = t & 1, nxt = cur ^ 1;
        const double* pl = lk[cur];
        const double* pf = fw[cur];
        const double o = obs[(size_t)t * N_STATES + d];
        double best = NEG, f = 0.0;
        int ptr = -1;
#pragma unroll
        for (int j = 1; j <= 3; j++)
        {
            const double lsp = vc.lsp[j - 1], sp = vc.sp[j - 1];
            const double base = o + lsp;
            const int lowbits = d >> (2 * j), shift = 10 - 2 * j;
#pragma unroll 4
            for (int k = 0; k < (1 << (2 * j)); k++)
            {
                const int p = lowbits + (k << shift);
                const double l = base + pl[p];
                f += sp * pf[p];
                if (l > best) { best = l; ptr = p; }
            }
        }
        {
            const double l = o + vc.stay_lik + pl[d];
            if (l > best) { best = l; ptr = d; }
            f += vc.stay_prob * pf[d];
        }
        f *= eobs[(size_t)t * N_STATES + d];
        // this CTA's part of the normalising sum
        double s = f;
        for (int o2 = 16; o2; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
        if (lane == 0) red[wid] = s;
        __syncthreads();
        // new values (forward probability not yet normalised) and the partial sum into all eight copies
#pragma unroll
        for (int q = 0; q < VIT_CL; q++)
        {
            r_lk[q][nxt * N_STATES + d] = best;
            r_fw[q][nxt * N_STATES + d] = f;
        }
        if (tid < VIT_CL)
        {
            double w = 0.0;
#pragma unroll
            for (int k = 0; k < PER / 32; k++) w += red[k];
            r_part[tid][cur * VIT_CL + rank] = w;
        }
        cluster.sync();
        double w = 0.0;
#pragma unroll
        for (int q = 0; q < VIT_CL; q++) w += part[cur][q];
        const double inv = 1.0 / w;
        // every CTA normalises its own copy (the same products everywhere)
#pragma unroll
        for (int q = 0; q < VIT_CL; q++) fw[nxt][q * PER + tid] *= inv;
        fwd[(size_t)t * N_STATES + d] = f * inv;
        backptr[(size_t)t * N_STATES + d] = ptr;
        __syncthreads();
    }
    last_liks[d] = lk[n_pos & 1][d];
    cluster.sync();                                           // nobody leaves while its shared memory may still be written
}

// ------------------------------------------------------------------------------------------
// Sampler (V_LIK::randbp + buildT): sample k walks back from `startst`; at chain step t the
// predecessor is drawn from  T[cur][i] * fwd[t][i]^atten  normalised, by inverse CDF on r[k][n-1-t].
// T[cur][i] = sum over advances j=1..4 of 0.25 * (0.25*skip)^(j-1) for which i is a j-step
// predecessor of cur, diagonal overwritten by stay (cpp/Viterbi.cpp:134-169).
__global__ void __launch_bounds__(1024) k_vit_sample(const double* fwd, int n_pos, int startst, const double* rnd,
                                                      double skip_prob, double stay_prob, double mut_min,
                                                      double mut_max, int nkeep, int* paths)
{
    __shared__ double red[32];
    __shared__ double scan[32];
    __shared__ int chosen;
    const int i = threadIdx.x, lane = i & 31, wid = i >> 5, k = blockIdx.x;
    const double atten = mut_min + (mut_max - mut_min) * k / (double)nkeep;
    double spj[4];
    spj[0] = 0.25;
    for (int j = 1; j < 4; j++) spj[j] = spj[j - 1] * 0.25 * skip_prob;
    int cur = startst;
    for (int t = n_pos - 1; t >= 0; t--)
    {
        if (i == 0) { paths[(size_t)k * n_pos + (n_pos - 1 - t)] = cur; chosen = N_STATES - 1; }
        double w = 0.0;
#pragma unroll
        for (int j = 1; j <= 4; j++)
            if ((i & ((1 << (10 - 2 * j)) - 1)) == (cur >> (2 * j))) w += spj[j - 1];
        if (i == cur) w = stay_prob;
        double p = 0.0;
        if (w != 0.0) p = w * pow(fwd[(size_t)t * N_STATES + i], atten);
        // total
        double s = p;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[wid] = s;
        __syncthreads();
        if (wid == 0)
        {
            double v = red[lane];
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[0] = 1.0 / v;
        }
        __syncthreads();
        p *= red[0];
        // inclusive prefix sum
        double c = p;
        for (int o = 1; o < 32; o <<= 1)
        {
            double v = __shfl_up_sync(0xffffffffu, c, o);
            if (lane >= o) c += v;
        }
        if (lane == 31) scan[wid] = c;
        __syncthreads();
        if (wid == 0)
        {
            double v = scan[lane];
            for (int o = 1; o < 32; o <<= 1)
            {
                double u = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += u;
            }
            scan[lane] = v;
        }
        __syncthreads();
        if (wid > 0) c += scan[wid - 1];
        const double r = rnd[(size_t)k * n_pos + (n_pos - 1 - t)];
        // first index whose running sum exceeds r: lowest lane per warp, then lowest warp
        const unsigned hit = __ballot_sync(0xffffffffu, r < c);
        if (hit && lane == 0) atomicMin(&chosen, wid * 32 + (__ffs(hit) - 1));
        __syncthreads();
        cur = chosen;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// host side
void ps_build_model(const HostModel& hm, ModelDev& md);      // ps_host.cu

static char base_of(int state, int ind) { return "ACGT"[3 & (state >> (2 * (4 - ind)))]; }

// StatesToSequence (cpp/Viterbi.cpp:171-237)
static std::string states_to_sequence(const std::vector<int>& st)
{
    std::string seq;
    int cur = st[0];
    seq.push_back(base_of(cur, 0));
    for (size_t i = 1; i < st.size(); i++)
    {
        const int nxt = st[i];
        if (nxt == cur) continue;                                   // stay
        bool linked = false;
        for (int n = 1; n <= 4 && !linked; n++)
            for (int ind = 0; ind < (1 << (2 * n)); ind++)
                if ((((cur << (2 * n)) & (N_STATES - 1)) + ind) == nxt)
                {
                    for (int j = 1; j <= n; j++) seq.push_back(base_of(cur, j));
                    cur = nxt;
                    linked = true;
                    break;
                }
        if (!linked) { cur = nxt; seq.push_back(base_of(cur, 0)); }   // mismatch: jump
    }
    for (int j = 1; j <= 4; j++) seq.push_back(base_of(cur, j));
    return seq;
}

template <class T>
static int vroom(ps_ctx* ctx, const char* name, size_t count, T** out)
{
    DevBuf& buf = ctx->bufs[name];
    int rc = ctx->ensure(buf, std::max<size_t>(count, 1) * sizeof(T));
    if (rc) return rc;
    *out = (T*)buf.p;
    return PS_OK;
}

// The host half of ViterbiMutate (cpp/Viterbi.cpp:262-325): which reads sit on which position (getrefstates,
// cpp/EventData.h:187-204), their pooled level / stdv, and the reference's sequential skip / stop rule over the positions.
// Every event must carry an alignment (ensure_refs done).  `positions` (optional) receives the positions kept.
static void vit_positions(ps_region* R, std::vector<VitSlot>& slots, std::vector<int>& slot_off, int& max_lik, std::vector<int>* positions)
{
    const int E = (int)R->events.size();
    int refind = R->events[0].refstart;
    int maxref = 0;
    for (const HostEvent& ev : R->events) { refind = std::min(refind, ev.refstart); maxref = std::max(maxref, ev.refend); }
    // first level whose ref_index equals a given integer exactly (std::find in getrefstates).  ref_index is extrapolated
    // past the last aligned level (cpp/EventData.h:148-152), so a read can still "sit" on positions beyond every read's
    // refend when its tail slope makes the extrapolated values integers: those (rare) positions are kept in `beyond`.
    std::vector<std::vector<int>> first(E);
    std::vector<std::map<int, int>> beyond(E);
    for (int k = 0; k < E; k++)
    {
        const HostEvent& ev = R->events[k];
        first[k].assign(maxref + 2, -1);
        for (int i = 0; i < ev.n0; i++)
        {
            const double v = ev.ref_index[i];
            if (!(v >= 0) || v != std::floor(v)) continue;
            if (v <= maxref + 1)
            {
                int& slot = first[k][(int)v];
                if (slot < 0) slot = i;
            }
            else if (v < 2147483647.0)
                beyond[k].emplace((int)v, i);                 // emplace keeps the first level of a value
        }
    }
    slots.clear();
    slot_off.assign(1, 0);
    max_lik = 1;
    // What each candidate position holds (the reads sitting on it, pooled level / stdv; how many reads span it) does not
    // depend on the positions before it: computed for all candidates on the worker threads, then the reference's
    // sequential skip / stop rule (cpp/Viterbi.cpp:310-325) walks over the results.
    const int ref0 = refind;
    const int n_cand = (ref0 >= 0 && ref0 <= maxref + 1) ? maxref + 2 - ref0 : 0;
    std::vector<std::vector<VitSlot>> cand(n_cand);
    std::vector<int> cand_nal(n_cand, 0);
    auto position = [&](int ri, std::vector<VitSlot>& out, int& nal_out) {
        int nal = 0;
        for (int k = 0; k < E; k++)
        {
            const HostEvent& ev = R->events[k];
            if (ri >= ev.refstart && ri <= ev.refend) nal++;
            int f = -1;
            if (ri >= 0 && ri <= maxref + 1) f = first[k][ri];
            else { auto it = beyond[k].find(ri); if (it != beyond[k].end()) f = it->second; }
            if (f < 0) continue;
            double lvl = ev.mean[f], sd = ev.stdv[f];
            int cnt = 1;
            for (int i = f + 1; i < ev.n0 && ev.ref_align[i] <= ri; i++)
                if (ev.ref_align[i] > 0) { lvl += ev.mean[i]; sd += ev.stdv[i]; cnt++; }
            lvl = lvl / cnt;
            sd = sd / cnt;
            VitSlot s;
            s.model = ev.model; s.pad = 0; s.lvl = lvl; s.sd = sd; s.logsd = std::log(sd);
            out.push_back(s);
        }
        nal_out = nal;
    };
    ps_parallel_for(n_cand, [&](int q) { position(ref0 + q, cand[q], cand_nal[q]); });
    std::vector<VitSlot> tail;
    while (true)
    {
        int nlik = 0, nal = 0;
        const std::vector<VitSlot>* here = nullptr;
        if (refind >= ref0 && refind - ref0 < n_cand) { here = &cand[refind - ref0]; nal = cand_nal[refind - ref0]; }
        else { tail.clear(); position(refind, tail, nal); here = &tail; }
        nlik = (int)here->size();
        if (nlik <= nal * 0.2)
        {
            if (nal == 0) break;
            refind++;
            continue;
        }
        slots.insert(slots.end(), here->begin(), here->end());
        slot_off.push_back((int)slots.size());
        if (positions) positions->push_back(refind);
        max_lik = std::max(max_lik, nlik);
        refind++;
    }
}

int ps_viterbi_list(ps_region* R, int nkeep, double skip_prob, double stay_prob, double mut_min, double mut_max,
                    std::vector<std::string>& out)
{
    out.clear();
    ps_ctx* ctx = R->ctx;
    TRY(ctx->init());
    CU(cudaSetDevice(ctx->device));
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    const int E = (int)R->events.size();
    if (E == 0) { ps_set_error(ctx, "ViterbiMutate needs at least one event"); return PS_E_ARG; }
    for (HostEvent& ev : R->events) ev.ensure_refs();
    for (const HostEvent& ev : R->events)
        if (ev.ri_empty)
        {
            // the reference dereferences an empty path here (cpp/Viterbi.cpp:262-264, 385-400)
            ps_set_error(ctx, "ViterbiMutate needs every event to carry an alignment");
            return PS_E_ARG;
        }
    // ---- host: which reads sit on which position -------------------------------------------
    std::vector<VitSlot> slots;
    std::vector<int> slot_off;
    int max_lik = 1;
    vit_positions(R, slots, slot_off, max_lik, nullptr);
    const int n_pos = (int)slot_off.size() - 1;
    if (n_pos == 0) { ps_set_error(ctx, "ViterbiMutate: no position is covered by the events"); return PS_E_ARG; }

    // ---- device ------------------------------------------------------------------------------
    // every host side of a copy is pinned memory owned by the context: a copy from or to pageable memory
    // waits for the stream inside the driver call, which stalls the CUDA calls of the other host threads
    // (other regions in flight on their own contexts) for as long as the chain kernel runs
    PinVec<ModelDev> models = ctx->pinned<ModelDev>("vit_models_h");
    PinVec<VitSlot> h_slots = ctx->pinned<VitSlot>("vit_slots_h");
    PinVec<int> h_off = ctx->pinned<int>("vit_off_h");
    if (!models.resize(R->models.size()) || !h_slots.resize(slots.size()) || !h_off.resize(slot_off.size()))
    {
        ps_set_error(ctx, "out of host memory staging ViterbiMutate");
        return PS_E_INTERNAL;
    }
    for (size_t q = 0; q < models.size(); q++) ps_build_model(R->models[q], models[q]);
    std::copy(slots.begin(), slots.end(), h_slots.data());
    std::copy(slot_off.begin(), slot_off.end(), h_off.data());
    ModelDev* d_models; VitSlot* d_slots; int* d_off;
    double *d_obs, *d_eobs, *d_fwd, *d_last, *d_rnd;
    int *d_bp, *d_paths;
    TRY(vroom(ctx, "vit_models", models.size(), &d_models));
    TRY(vroom(ctx, "vit_slots", slots.size(), &d_slots));
    TRY(vroom(ctx, "vit_off", slot_off.size(), &d_off));
    TRY(vroom(ctx, "vit_obs", (size_t)n_pos * N_STATES, &d_obs));
    TRY(vroom(ctx, "vit_eobs", (size_t)n_pos * N_STATES, &d_eobs));
    TRY(vroom(ctx, "vit_fwd", (size_t)n_pos * N_STATES, &d_fwd));
    TRY(vroom(ctx, "vit_bp", (size_t)n_pos * N_STATES, &d_bp));
    TRY(vroom(ctx, "vit_last", (size_t)N_STATES, &d_last));
    CU(cudaMemcpyAsync(d_models, models.data(), models.size() * sizeof(ModelDev), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_slots, h_slots.data(), slots.size() * sizeof(VitSlot), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_off, h_off.data(), slot_off.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    {
        int threads = 128;
        while (threads > 32 && (size_t)threads * max_lik * sizeof(double) > 200 * 1024) threads >>= 1;
        const size_t smem = (size_t)threads * max_lik * sizeof(double);
        if (smem > 200 * 1024) { ps_set_error(ctx, "ViterbiMutate: %d reads on one position exceed shared memory", max_lik); return PS_E_ARG; }
        CU(cudaFuncSetAttribute(k_vit_obs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
        dim3 grid(N_STATES / threads, n_pos);
        k_vit_obs<<<grid, threads, smem, ctx->stream>>>(d_models, d_slots, d_off, n_pos, std::log(2 * M_PI), d_obs, d_eobs);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    VitConst vc;
    {
        // cpp/Viterbi.cpp:42-44, 57-58, 76-77: sp_1 = .25, sp_{j+1} = sp_j*.25*skip; lsp likewise in logs
        const double skip_lik = std::log(skip_prob);
        double sp = 0.25, lsp = std::log(0.25);
        for (int j = 0; j < 3; j++) { vc.sp[j] = sp; vc.lsp[j] = lsp; sp = sp * 0.25 * skip_prob; lsp = lsp + std::log(0.25) + skip_lik; }
        vc.stay_lik = std::log(stay_prob);
        vc.stay_prob = stay_prob;
    }
    double t_obs = 0;
    if (ctx->trace) { CU(cudaStreamSynchronize(ctx->stream)); t_obs = now(); }
    if (ctx->vit_cluster) k_vit_chain_cluster<<<VIT_CL, N_STATES / VIT_CL, 0, ctx->stream>>>(d_obs, d_eobs, n_pos, vc, d_fwd, d_bp, d_last);
    else k_vit_chain<<<1, N_STATES, 0, ctx->stream>>>(d_obs, d_eobs, n_pos, vc, d_fwd, d_bp, d_last);
    ctx->launches++;
    CU(cudaGetLastError());
    PinVec<double> last = ctx->pinned<double>("vit_last_h");
    if (!last.resize(N_STATES)) { ps_set_error(ctx, "out of host memory staging ViterbiMutate"); return PS_E_INTERNAL; }
    CU(cudaMemcpyAsync(last.data(), d_last, N_STATES * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU((cudaError_t)ps_stream_wait(ctx));
    const double t_chain = now();
    const int startst = (int)(std::max_element(last.begin(), last.end()) - last.begin());

    std::vector<int> path(n_pos);
    if (nkeep == 0)
    {
        PinVec<int> bp = ctx->pinned<int>("vit_bp_h");
        if (!bp.resize((size_t)n_pos * N_STATES)) { ps_set_error(ctx, "out of host memory staging ViterbiMutate"); return PS_E_INTERNAL; }
        CU(cudaMemcpyAsync(bp.data(), d_bp, bp.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU((cudaError_t)ps_stream_wait(ctx));
        int cur = startst;
        for (int t = n_pos - 1; t >= 0; t--) { path[t] = cur; cur = bp[(size_t)t * N_STATES + cur]; }
        out.push_back(states_to_sequence(path));
        return PS_OK;
    }
    // glibc rand() stream in the reference's call order: sample-major, positions from the end
    PinVec<double> rnd = ctx->pinned<double>("vit_rnd_h");
    PinVec<int> paths = ctx->pinned<int>("vit_paths_h");
    if (!rnd.resize((size_t)nkeep * n_pos) || !paths.resize((size_t)nkeep * n_pos))
    {
        ps_set_error(ctx, "out of host memory staging ViterbiMutate");
        return PS_E_INTERNAL;
    }
    if (R->own_rng) for (size_t q = 0; q < rnd.size(); q++) rnd[q] = R->next_uniform();
    else for (size_t q = 0; q < rnd.size(); q++) rnd[q] = rand() / (double(RAND_MAX) + 1);
    TRY(vroom(ctx, "vit_rnd", rnd.size(), &d_rnd));
    TRY(vroom(ctx, "vit_paths", rnd.size(), &d_paths));
    CU(cudaMemcpyAsync(d_rnd, rnd.data(), rnd.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_vit_sample<<<nkeep, N_STATES, 0, ctx->stream>>>(d_fwd, n_pos, startst, d_rnd, skip_prob, stay_prob, mut_min, mut_max, nkeep, d_paths);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(paths.data(), d_paths, paths.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU((cudaError_t)ps_stream_wait(ctx));
    if (ctx->trace)
        fprintf(stderr, "[ps] ViterbiMutate: %d positions, %d reads: host + emission pooling %.1f ms, chain %.1f ms, %d sampled walks %.1f ms\n",
                n_pos, E, t_obs - t_begin, t_chain - t_obs, nkeep, now() - t_chain);
    for (int k = 0; k < nkeep; k++)
    {
        // paths are stored end-first (as walked); flip to chain order
        for (int t = 0; t < n_pos; t++) path[t] = paths[(size_t)k * n_pos + (n_pos - 1 - t)];
        out.push_back(states_to_sequence(path));
    }
    return PS_OK;
}

extern "C" {

// Host only (no device work): the positions ViterbiMutate keeps and how many reads sit on each -- for the CPU tests of
// the skip / stop rule.  Returns PS_E_ARG like ps_viterbi_mutate when an event carries no alignment.
extern "C" int ps_viterbi_positions(ps_region* R, int cap, int* positions, int* reads_here, int* n_positions)
{
    if (!R || cap < 0 || !n_positions || (cap > 0 && (!positions || !reads_here))) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_viterbi_positions");
    if (R->events.empty()) { ps_set_error(R->ctx, "ViterbiMutate needs at least one event"); return PS_E_ARG; }
    for (HostEvent& ev : R->events) ev.ensure_refs();
    for (const HostEvent& ev : R->events)
        if (ev.ri_empty) { ps_set_error(R->ctx, "ViterbiMutate needs every event to carry an alignment"); return PS_E_ARG; }
    std::vector<VitSlot> slots;
    std::vector<int> slot_off, pos;
    int max_lik = 1;
    vit_positions(R, slots, slot_off, max_lik, &pos);
    *n_positions = (int)pos.size();
    if ((int)pos.size() > cap) { ps_set_error(R->ctx, "ps_viterbi_positions: %zu positions, room for %d", pos.size(), cap); return PS_E_CAPACITY; }
    for (size_t k = 0; k < pos.size(); k++) { positions[k] = pos[k]; reads_here[k] = slot_off[k + 1] - slot_off[k]; }
    return PS_OK;
}

int ps_viterbi_mutate(ps_region* R, int nkeep, double skip_prob, double stay_prob, double mut_min, double mut_max, int* n_seqs)
{
    if (!R || nkeep < 0) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_viterbi_mutate");
    TRY(ps_viterbi_list(R, nkeep, skip_prob, stay_prob, mut_min, mut_max, R->viterbi));
    if (n_seqs) *n_seqs = (int)R->viterbi.size();
    return PS_OK;
}

int ps_get_viterbi_sequence(ps_region* R, int i, char* out, int cap)
{
    if (!R || i < 0 || i >= (int)R->viterbi.size()) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_get_viterbi_sequence");
    const std::string& s = R->viterbi[i];
    if (!out) return (int)s.size();
    if ((int)s.size() + 1 > cap) return PS_E_CAPACITY;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

} // extern "C"
