// ps_lockstep.cu -- the consensus loop (poreseq/Mutate.py:47-99) of MANY regions in lockstep.
//
// One region's loop is ~35 dependent GPU jobs of 20-40 events each (a realignment, the seeds' realignments, a scoring
// pass, ... per repetition), and a job of 40 events leaves most of the GPU idle for the ~1 ms its fill wave takes.  The
// reference scales by running one process per region; regions in flight on separate streams (ps_consensus_batch's first
// form) overlap those small jobs only partly (~40 kb/s for 1 kb x 10x regions).  Here every step of the loop is done for
// ALL regions that are at that step, as ONE job: the fills, backtraces and mutation kernels see hundreds of events per
// launch, the way the batched ScorePoints of the benchmark does.  What stays per region -- the Smith-Waterman maps of
// FindMutations, the Viterbi chain (both small dependent kernels per region), the candidate picking and the accept
// loops on the host -- runs on `in_flight` lanes (a host thread + a context with its own stream each) side by side.
//
// Each region goes through exactly the steps ps_consensus takes for it alone, on its own data; regions drop out of a
// loop when the reference's loop would stop for them.  Results are identical to ps_consensus one region at a time.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "ps_internal.h"

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

namespace
{
typedef std::vector<std::string> Seeds;

struct Lanes                                  // ctx + its helper contexts: `n` host threads with a stream each
{
    std::vector<ps_ctx*> ctx;
    // fn(lane context, item) for item in [0, count), items dealt to the lanes dynamically; first error wins
    int run(int count, const std::function<int(ps_ctx*, int)>& fn)
    {
        if (count <= 0) return PS_OK;
        const int n = std::max(1, std::min((int)ctx.size(), count));
        std::atomic<int> next(0);
        std::vector<int> rcs(n, PS_OK);
        auto work = [&](int lane) {
            for (;;)
            {
                const int k = next.fetch_add(1);
                if (k >= count) return;
                const int rc = fn(ctx[lane], k);
                if (rc && !rcs[lane]) rcs[lane] = rc;
            }
        };
        std::vector<std::thread> th;
        for (int lane = 1; lane < n; lane++) th.emplace_back(work, lane);
        work(0);
        for (std::thread& t : th) t.join();
        for (int lane = 0; lane < n; lane++)
            if (rcs[lane]) { if (lane) ps_set_error(ctx[0], "%s", ctx[lane]->error.c_str()); return rcs[lane]; }
        return PS_OK;
    }
};

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// FindMutations (cpp/FindMutations.cpp:24-186) for every region of `regs` with its own seeds: see ps_find_mutation_list
// for the single-region form this follows step by step.
int find_mutations(Lanes& L, const std::vector<ps_region*>& regs, const std::vector<const Seeds*>& seeds,
                   std::vector<std::vector<HostMut>>& found)
{
    ps_ctx* ctx = L.ctx[0];
    const int n = (int)regs.size();
    found.assign(n, std::vector<HostMut>());
    if (n == 0) return PS_OK;
    const double t0 = now_ms();
    // 1. every region realigned to its current sequence, per-base likelihood profiles: one job
    std::vector<std::vector<double>> base;
    TRY(ps_run_alignments(ctx, regs, nullptr, &base));
    const double t1 = now_ms();
    // 2. Smith-Waterman maps of every region's sequence to its seeds: per region on the lanes
    std::vector<std::vector<SWResult>> sw(n);
    std::vector<char> sw_gpu(n, 0);
    TRY(L.run(n, [&](ps_ctx* lane, int r) {
        sw_gpu[r] = !lane->sw_host && psi_swfull_batch(lane, regs[r]->bases, *seeds[r], sw[r]) == PS_OK;
        return PS_OK;
    }));
    const double t2 = now_ms();
    // which (region, seed) pairs need a shadow region: first occurrence of a seed whose profile is not cached
    struct Pair { int r, s; };
    std::vector<Pair> pairs;                      // every (region, seed)
    std::vector<std::vector<char>> need(n);
    std::vector<std::vector<std::string>> shadow_key(n);
    for (int r = 0; r < n; r++)
    {
        const Seeds& sd = *seeds[r];
        need[r].assign(sd.size(), 0);
        for (size_t s = 0; s < sd.size(); s++)
        {
            const auto hit = regs[r]->seqlikes.find(sd[s]);
            const bool cached = hit != regs[r]->seqlikes.end() && !hit->second.empty();
            const bool queued = std::find(shadow_key[r].begin(), shadow_key[r].end(), sd[s]) != shadow_key[r].end();
            if (!cached && !queued && sd[s].size() >= 5) { need[r][s] = 1; shadow_key[r].push_back(sd[s]); }
            pairs.push_back(Pair{r, (int)s});
        }
    }
    ps_parallel_for(n, [&](int r) {
        if (!shadow_key[r].empty())
            for (HostEvent& he : regs[r]->events) he.ensure_levrec();
    });
    // (the profile caches stay out of the copies)
    std::vector<std::map<std::string, std::vector<double>>> cache(n);
    for (int r = 0; r < n; r++) cache[r].swap(regs[r]->seqlikes);
    std::vector<std::vector<SWResult>> als(n);
    std::vector<std::vector<ps_region*>> nds(n);
    for (int r = 0; r < n; r++) { als[r].resize(seeds[r]->size()); nds[r].assign(seeds[r]->size(), nullptr); }
    ps_parallel_for((int)pairs.size(), [&](int q) {
        const int r = pairs[q].r, s = pairs[q].s;
        ps_region* R = regs[r];
        const std::string& seed = (*seeds[r])[s];
        if (need[r][s])
        {
            ps_region* nd = ps_shadow_region(R);          // level data borrowed from R, not copied
            als[r][s] = sw_gpu[r] ? psi_map_alignments_with(nd, seed, sw[r][s]) : psi_map_alignments(nd, seed);
            nds[r][s] = nd;
        }
        else
        {
            als[r][s] = sw_gpu[r] ? sw[r][s] : psi_swfull(R->bases, seed);
            psi_fillinds(als[r][s]);
        }
    });
    for (int r = 0; r < n; r++) regs[r]->seqlikes.swap(cache[r]);
    // 3. all shadow regions of all regions realigned in one job
    std::vector<ps_region*> shadows;
    std::vector<Pair> shadow_of;                  // (region, index into shadow_key[region])
    for (int r = 0; r < n; r++)
    {
        int k = 0;
        for (size_t s = 0; s < seeds[r]->size(); s++)
            if (need[r][s]) { shadows.push_back(nds[r][s]); shadow_of.push_back(Pair{r, k++}); }
    }
    const double t3 = now_ms();
    if (!shadows.empty())
    {
        std::vector<std::vector<double>> likes;
        const int rc = ps_run_alignments(ctx, shadows, nullptr, &likes);
        for (size_t k = 0; k < shadows.size(); k++)
        {
            if (!rc) regs[shadow_of[k].r]->seqlikes[shadow_key[shadow_of[k].r][shadow_of[k].s]] = likes[k];
            delete shadows[k];
        }
        if (rc) return rc;
    }
    const double t4 = now_ms();
    // 4. CUSUM of the profile differences and greedy peak picking, per region on the host threads
    ps_parallel_for(n, [&](int r) {
        ps_region* R = regs[r];
        const Seeds& sd = *seeds[r];
        std::vector<const std::vector<double>*> profs(sd.size());
        for (size_t s = 0; s < sd.size(); s++)
        {
            std::vector<double>& prof = R->seqlikes[sd[s]];
            if (prof.empty()) prof.assign(sd[s].size(), 0.0);
            profs[s] = &prof;
        }
        psi_pick_candidates(R->bases, base[r], sd, profs, als[r], found[r]);
    });
    if (ctx->trace)
        fprintf(stderr, "[ps] lockstep FindMutations, %d regions: base realign %.1f ms, SW maps %.1f, shadow copies %.1f, %zu shadow regions realigned %.1f, picking %.1f\n",
                n, t1 - t0, t2 - t1, t3 - t2, shadows.size(), t4 - t3, now_ms() - t4);
    return PS_OK;
}

// MakeMutations (cpp/MakeMutations.cpp:74-146) for every region with its scored list; the reference's recursion on
// more than ten deferred mutations becomes rounds: the deferred lists of all regions that have one are scored together
int make_mutations(Lanes& L, std::vector<ps_region*> regs, std::vector<std::vector<HostMut>> lists, std::vector<int*> nbases)
{
    ps_ctx* ctx = L.ctx[0];
    for (int* p : nbases) *p = 0;
    while (!regs.empty())
    {
        const int n = (int)regs.size();
        std::vector<int> changed(n, 0);
        std::vector<std::vector<HostMut>> deferred(n);
        ps_parallel_for(n, [&](int r) { ps_make_mutation_pass(regs[r], std::move(lists[r]), &changed[r], &deferred[r]); });
        std::vector<ps_region*> nregs;
        std::vector<std::vector<HostMut>> nlists;
        std::vector<int*> nnb;
        for (int r = 0; r < n; r++)
        {
            *nbases[r] += changed[r];
            if (deferred[r].size() > 10) { nregs.push_back(regs[r]); nlists.push_back(std::move(deferred[r])); nnb.push_back(nbases[r]); }
        }
        if (nregs.empty()) break;
        std::vector<std::vector<HostMut>*> ptrs;
        for (auto& v : nlists) ptrs.push_back(&v);
        TRY(ps_score_mutation_lists(ctx, nregs, ptrs));
        regs.swap(nregs); lists.swap(nlists); nbases.swap(nnb);
    }
    return PS_OK;
}

// PSAlign.Mutate's loop (poreseq/_poreseqcpp.pyx:424-431) for every region with its own seeds
int mutate(Lanes& L, const std::vector<ps_region*>& regs, const std::vector<const Seeds*>& seeds, int reps, std::vector<int>& total)
{
    ps_ctx* ctx = L.ctx[0];
    const int n = (int)regs.size();
    total.assign(n, 0);
    std::vector<int> active(n);
    for (int r = 0; r < n; r++) { active[r] = r; regs[r]->seqlikes.clear(); }
    for (int rep = 0; rep < reps && !active.empty(); rep++)
    {
        const double t0 = now_ms();
        std::vector<ps_region*> ar;
        std::vector<const Seeds*> as;
        for (int r : active) { ar.push_back(regs[r]); as.push_back(seeds[r]); }
        std::vector<std::vector<HostMut>> cand;
        TRY(find_mutations(L, ar, as, cand));
        const double t1 = now_ms();
        std::vector<std::vector<HostMut>*> ptrs;
        for (auto& v : cand) ptrs.push_back(&v);
        TRY(ps_score_mutation_lists(ctx, ar, ptrs));
        const double t2 = now_ms();
        std::vector<int> nb(ar.size(), 0);
        std::vector<int*> nbp;
        for (int& v : nb) nbp.push_back(&v);
        TRY(make_mutations(L, ar, std::move(cand), nbp));
        if (ctx->trace)
            fprintf(stderr, "[ps] lockstep Mutate rep %d, %zu regions: FindMutations %.1f ms, ScoreMutations %.1f, MakeMutations %.1f\n",
                    rep, ar.size(), t1 - t0, t2 - t1, now_ms() - t2);
        std::vector<int> still;
        for (size_t k = 0; k < active.size(); k++)
            if (nb[k] != 0) { total[active[k]] += nb[k]; still.push_back(active[k]); }
        active.swap(still);
    }
    for (int r = 0; r < n; r++) regs[r]->seqlikes.clear();
    return PS_OK;
}

// PSAlign.Refine (pyx:437-472) for every region: all point mutations scored in one job, then the accept loops
int refine(Lanes& L, const std::vector<ps_region*>& regs, int point_width, std::vector<int>& nb)
{
    ps_ctx* ctx = L.ctx[0];
    const int n = (int)regs.size();
    nb.assign(n, 0);
    if (n == 0) return PS_OK;
    std::vector<int> width(n);
    for (int r = 0; r < n; r++) { width[r] = regs[r]->params.scoring_width; regs[r]->params.scoring_width = point_width; }
    std::vector<std::vector<HostMut>> lists(n);
    ps_parallel_for(n, [&](int r) { lists[r] = ps_point_mutations(regs[r]); });
    std::vector<std::vector<HostMut>*> ptrs;
    for (auto& v : lists) ptrs.push_back(&v);
    int rc = ps_score_mutation_lists(ctx, regs, ptrs);
    if (!rc)
    {
        std::vector<int*> nbp;
        for (int& v : nb) nbp.push_back(&v);
        rc = make_mutations(L, regs, std::move(lists), nbp);
    }
    for (int r = 0; r < n; r++) regs[r]->params.scoring_width = width[r];
    return rc;
}
} // namespace

int ps_consensus_lockstep(ps_ctx* ctx, ps_region* const* regions, int n_regions, int reps, int point_width, int in_flight)
{
    TRY(ctx->init());
    in_flight = std::max(1, std::min(std::min(in_flight, n_regions), 64));
    while ((int)ctx->helpers.size() < in_flight - 1)
    {
        ps_ctx* h = ps_create(ctx->device);
        if (!h) { ps_set_error(ctx, "ps_consensus_batch: out of memory"); return PS_E_INTERNAL; }
        ctx->helpers.push_back(h);
    }
    Lanes L;
    L.ctx.push_back(ctx);
    for (int k = 0; k + 1 < in_flight; k++) { ctx->helpers[k]->precision = ctx->precision; ctx->helpers[k]->blocking_wait = true; L.ctx.push_back(ctx->helpers[k]); }
    // regions the loop runs for (Mutate.py:50-53: fewer than 5 events -> untouched); all of them move to this context
    std::vector<ps_region*> regs;
    std::vector<ps_ctx*> home;
    for (int k = 0; k < n_regions; k++)
    {
        ps_region* R = regions[k];
        R->stage_log.clear(); R->stage_nbases.clear();
        if (R->events.size() < 5) continue;
        if (!R->own_rng) R->rng_seed(1);
        regs.push_back(R); home.push_back(R->ctx);
        R->ctx = ctx;
    }
    auto finish = [&](int rc) { for (size_t k = 0; k < regs.size(); k++) regs[k]->ctx = home[k]; return rc; };
    auto note = [&](const std::vector<ps_region*>& rs, const std::string& name, const std::vector<int>& nb) {
        for (size_t k = 0; k < rs.size(); k++) { rs[k]->stage_log.emplace_back(name, rs[k]->bases); rs[k]->stage_nbases.push_back(nb[k]); }
    };
    const int n = (int)regs.size();
    if (n == 0) return finish(PS_OK);
    // Mutate('self'): the 2D sequences of every other event (pyx:412-414)
    std::vector<Seeds> self(n);
    std::vector<const Seeds*> sp(n);
    for (int r = 0; r < n; r++)
    {
        for (size_t e = 0; e < regs[r]->events.size(); e += 2) self[r].push_back(regs[r]->events[e].seq2d);
        sp[r] = &self[r];
    }
    std::vector<int> nb;
    int rc = mutate(L, regs, sp, reps, nb);
    if (rc) return finish(rc);
    note(regs, "mutate_self", nb);
    std::vector<ps_region*> active(regs);
    for (int k = 0; k < reps && !active.empty(); k++)
    {
        const int m = (int)active.size();
        // ViterbiMutate per region on the lanes (a 1024-state chain over the positions: one CTA per region)
        std::vector<Seeds> vit(m);
        const double t0 = now_ms();
        rc = L.run(m, [&](ps_ctx* lane, int r) {
            ps_region* R = active[r];
            R->ctx = lane;
            const int q = ps_viterbi_list(R, 16, 0.05, 0.01, 0.33, 0.75, vit[r]);     // pyx:417
            R->ctx = ctx;
            return q;
        });
        if (rc) return finish(rc);
        if (ctx->trace) fprintf(stderr, "[ps] lockstep ViterbiMutate, %d regions: %.1f ms\n", m, now_ms() - t0);
        std::vector<const Seeds*> vp(m);
        for (int r = 0; r < m; r++) vp[r] = &vit[r];
        rc = mutate(L, active, vp, 4, nb);                                         // Mutate.py:76: reps defaults to 4
        if (rc) return finish(rc);
        note(active, "mutate_viterbi_" + std::to_string(k), nb);
        rc = refine(L, active, point_width, nb);
        if (rc) return finish(rc);
        note(active, "refine_" + std::to_string(k), nb);
        std::vector<ps_region*> still;
        for (int r = 0; r < m; r++) if (nb[r] != 0) still.push_back(active[r]);
        active.swap(still);
    }
    return finish(PS_OK);
}
