// ps_drivers.cu -- host drivers around the kernels: integer Smith-Waterman (swfull), alignment
// re-mapping (MapAlignments), seed-based candidate discovery (FindMutations) and the
// Find/Score/Make iteration of PSAlign.Mutate.  All DP over events runs on the GPU through
// ps_run_job(); what stays here is sequential glue the reference also runs on one core.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <atomic>

#include "ps_internal.h"

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

// fillinds (cpp/swlib.cpp:342-365): gaps take the previous aligned index of their own sequence
void psi_fillinds(SWResult& al)
{
    if (al.inds1.empty()) return;
    int a = al.inds1[0], b = al.inds2[0];
    for (size_t k = 0; k < al.inds1.size(); k++)
    {
        if (al.inds1[k] > 0) a = al.inds1[k]; else al.inds1[k] = a;
        if (al.inds2[k] > 0) b = al.inds2[k]; else al.inds2[k] = b;
    }
}

// MapAlignments (cpp/EventUtil.cpp:12-55): realign the region to `newseq` by integer SW and carry
// every level's ref_align across: value v -> inds2[lower_bound(inds1, v)], 0 outside the aligned span.
SWResult psi_map_alignments(ps_region* R, const std::string& newseq)
{
    return psi_map_alignments_with(R, newseq, psi_swfull(R->bases, newseq));
}

SWResult psi_map_alignments_with(ps_region* R, const std::string& newseq, SWResult al)
{
    psi_fillinds(al);
    R->set_sequence(newseq);
    const std::vector<int>& a = al.inds1;
    const std::vector<int>& b = al.inds2;
    // lower_bound(a, v) for every v in [a.front(), a.back()] in one sweep (a is non-decreasing once its gaps are
    // filled): the per-level search of the reference becomes a table look-up
    std::vector<int> first_ge;
    if (!a.empty() && a.back() >= a.front())
    {
        first_ge.resize((size_t)(a.back() - a.front()) + 1);
        size_t idx = 0;
        for (int v = a.front(); v <= a.back(); v++)
        {
            while (idx < a.size() && a[idx] < v) idx++;
            first_ge[(size_t)(v - a.front())] = (int)idx;
        }
    }
    for (HostEvent& ev : R->events)
    {
        for (int j = 0; j < ev.n0; j++)
        {
            const int v = (int)ev.ref_align[j];
            if (a.empty() || v < a.front() || v > a.back()) { ev.ref_align[j] = 0; continue; }
            const size_t at = (size_t)first_ge[(size_t)(v - a.front())];
            ev.ref_align[j] = at < b.size() ? b[at] : 0;
        }
        ev.update_refs();
    }
    return al;
}

// FindMutations, second half (cpp/FindMutations.cpp:51-186): given the per-base likelihood profile of the current
// sequence (`base`), of every seed (`profs[s]`, one value per seed base) and the SW alignment of the sequence to every
// seed with its gaps filled (`als[s]`, consumed), the CUSUM of the profile difference along each alignment and the
// greedy peak picking that turns its maxima into candidate edits.  Host only.
void psi_pick_candidates(const std::string& bases, const std::vector<double>& base, const std::vector<std::string>& seeds,
                         const std::vector<const std::vector<double>*>& profs, std::vector<SWResult>& als,
                         std::vector<HostMut>& found)
{
    found.clear();
    const size_t L = bases.size(), S = seeds.size();
    // 3. CUSUM of the profile difference along each SW alignment (:51-94)
    std::vector<std::vector<double>> dl(S);
    for (size_t s = 0; s < S; s++)
    {
        const std::vector<double>& prof = *profs[s];
        std::vector<int>& i1 = als[s].inds1;
        std::vector<int>& i2 = als[s].inds2;
        for (size_t j = 0; j < i1.size(); j++) { i1[j] -= 2; i2[j] -= 2; }
        size_t drop = 0;
        while (drop < i1.size() && (i1[drop] < 0 || i2[drop] < 0)) drop++;
        i1.erase(i1.begin(), i1.begin() + drop);
        i2.erase(i2.begin(), i2.begin() + drop);
        const size_t n = i1.size();
        std::vector<double> a1(n), a2(n);
        for (size_t j = 0; j < n; j++) { a1[j] = base[i1[j]]; a2[j] = prof[i2[j]]; }
        for (size_t j = n; j-- > 1;) { a1[j] -= a1[j - 1]; a2[j] -= a2[j - 1]; }
        if (n) { a1[0] = 0; a2[0] = 0; }
        dl[s].resize(n);
        double cus = 0;
        for (size_t j = 0; j < n; j++)
        {
            cus += a2[j] - a1[j];
            if (cus < 0) cus = 0;
            dl[s][j] = std::fabs(a1[j] - a2[j]) < 1e-5 ? 0.0 : cus;
        }
    }
    // 4. greedy peak picking (:111-183)
    std::vector<long> peak_at(S, -1);              // cached first argmax of every seed's curve, -1 = stale
    // The reference rescans every seed's curve per pick (std::max_element over ~L values, up to L/3 picks).  Only the
    // curve the previous pick zeroed can have a new first maximum, so the others keep their cached one; and that
    // curve's first maximum is found through per-block maxima (blocks of 64, refreshed where the pick zeroed):
    // the first block holding the largest block maximum contains the first occurrence of the largest value.  Same
    // picks, same tie order.
    constexpr int PB = 64;
    std::vector<std::vector<double>> bmax(S);
    auto refresh_block = [&](size_t s, int b) {
        const std::vector<double>& d = dl[s];
        const int lo = b * PB, hi = std::min((int)d.size(), lo + PB);
        double m = d[lo];
        for (int k = lo + 1; k < hi; k++) m = d[k] > m ? d[k] : m;
        bmax[s][b] = m;
    };
    auto first_argmax = [&](size_t s) -> long {
        const std::vector<double>& d = dl[s];
        const std::vector<double>& bm = bmax[s];
        int bb = 0;
        for (int b = 1; b < (int)bm.size(); b++) if (bm[b] > bm[bb]) bb = b;
        const int lo = bb * PB, hi = std::min((int)d.size(), lo + PB);
        int at = lo;
        for (int k = lo + 1; k < hi; k++) if (d[k] > d[at]) at = k;
        return at;
    };
    for (size_t s = 0; s < S; s++)
    {
        if (dl[s].empty()) continue;
        bmax[s].resize((dl[s].size() + PB - 1) / PB);
        for (int b = 0; b < (int)bmax[s].size(); b++) refresh_block(s, b);
    }
    while (found.size() < L / 3)
    {
        int smax = -1, ind = 0;
        double vmax = 0;
        for (size_t s = 0; s < S; s++)
        {
            if (dl[s].empty()) continue;
            if (peak_at[s] < 0) peak_at[s] = first_argmax(s);
            const size_t at = (size_t)peak_at[s];
            if (smax < 0 || dl[s][at] > vmax) { smax = (int)s; ind = (int)at; vmax = dl[s][at]; }
        }
        if (smax < 0 || vmax < 0.25) break;
        peak_at[smax] = -1;
        std::vector<double>& d = dl[smax];
        const int n = (int)d.size();
        int i1 = ind;
        while (i1 < n && d[i1] != 0) i1++;
        int i0 = ind;
        while (i0 >= 0 && d[i0] != 0) i0--;
        if (i0 < 0) i0 = 0;
        if (i1 >= n) i1 = n - 1;
        const int start1 = als[smax].inds1[i0], start2 = als[smax].inds2[i0];
        const int end1 = als[smax].inds1[ind], end2 = als[smax].inds2[ind];
        HostMut m;
        m.start = start1;
        m.orig = bases.substr(start1, end1 - start1);
        m.mut = seeds[smax].substr(start2, end2 - start2);
        while (!m.orig.empty() && !m.mut.empty() && m.orig.front() == m.mut.front())
        {
            m.orig.erase(0, 1); m.mut.erase(0, 1); m.start++;
        }
        while (!m.orig.empty() && !m.mut.empty() && m.orig.back() == m.mut.back())
        {
            m.orig.pop_back(); m.mut.pop_back();
        }
        if (!m.orig.empty() || !m.mut.empty()) found.push_back(m);
        std::fill(d.begin() + i0, d.begin() + i1 + 1, 0.0);
        for (int b = i0 / PB; b <= i1 / PB; b++) refresh_block((size_t)smax, b);
    }
}

// ------------------------------------------------------------------------------------------
// FindMutations (cpp/FindMutations.cpp:24-186).  The S seed realignments (S x E wide forward
// fills + backtraces) that dominate it are submitted as ONE batched GPU job: every distinct,
// not-yet-cached seed becomes a shadow region (copy of the events, alignments mapped through SW).
int ps_find_mutation_list(ps_region* R, const std::vector<std::string>& seeds, std::vector<HostMut>& found)
{
    found.clear();
    ps_ctx* ctx = R->ctx;
    const size_t L = R->bases.size();
    const bool trace = ctx->trace;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    // 1. realign to the current sequence, keep its per-base likelihood profile
    std::vector<double> base(L, 0.0);
    {
        std::vector<ps_region*> one(1, R);
        std::vector<std::vector<double>> likes;
        TRY(ps_run_alignments(ctx, one, nullptr, &likes));
        base = likes[0];
    }
    // 2. per seed: SW map, likelihood profile (cached by seed string for the life of the region)
    const size_t S = seeds.size();
    std::vector<SWResult> als(S);
    std::vector<ps_region*> shadows;
    std::vector<std::string> shadow_key;
    const double t_base = now();
    // the SW maps (O(L^2) each, the host-side cost of FindMutations) are independent per seed: worker threads
    // on the GPU (ps_sw.cu) for sequences up to 16 k bases, else on the host worker threads
    std::vector<ps_region*> nds(S, nullptr);
    std::vector<SWResult> sw;
    const bool sw_gpu = !ctx->sw_host && psi_swfull_batch(ctx, R->bases, seeds, sw) == PS_OK;
    // Only the first occurrence of a seed whose profile is not cached yet needs a shadow region (a copy of every
    // event, 36 MB for 10 kb x 30x); for the others the SW alignment with its gaps filled (what MapAlignments returns)
    // is all the CUSUM below reads.
    std::vector<char> need(S, 0);
    for (size_t s = 0; s < S; s++)
    {
        const auto hit = R->seqlikes.find(seeds[s]);
        const bool cached = hit != R->seqlikes.end() && !hit->second.empty();
        const bool queued = std::find(shadow_key.begin(), shadow_key.end(), seeds[s]) != shadow_key.end();
        if (!cached && !queued && seeds[s].size() >= 5) { need[s] = 1; shadow_key.push_back(seeds[s]); }
    }
    // the shadow regions inherit the level records (3 log(stdv) per level) instead of each computing its own
    if (!shadow_key.empty())
        ps_parallel_for((int)R->events.size(), [&](int e) { R->events[e].ensure_levrec(); });
    // (the profile cache stays out of the copies: up to one L-long profile per seed ever seen)
    std::map<std::string, std::vector<double>> cache;
    cache.swap(R->seqlikes);
    ps_parallel_for((int)S, [&](int s) {
        if (need[s])
        {
            ps_region* nd = ps_shadow_region(R);          // level data borrowed from R, not copied
            als[s] = sw_gpu ? psi_map_alignments_with(nd, seeds[s], sw[s]) : psi_map_alignments(nd, seeds[s]);
            nds[s] = nd;
        }
        else
        {
            als[s] = sw_gpu ? sw[s] : psi_swfull(R->bases, seeds[s]);
            psi_fillinds(als[s]);
        }
    });
    R->seqlikes.swap(cache);
    for (size_t s = 0; s < S; s++)
        if (need[s]) shadows.push_back(nds[s]);        // same order as shadow_key
    const double t_sw = now();
    if (!shadows.empty())
    {
        std::vector<std::vector<double>> likes;
        int rc = ps_run_alignments(ctx, shadows, nullptr, &likes);
        for (size_t k = 0; k < shadows.size(); k++)
        {
            if (!rc) R->seqlikes[shadow_key[k]] = likes[k];
            delete shadows[k];
        }
        if (rc) return rc;
    }
    if (trace) fprintf(stderr, "[ps] FindMutations: base realign %.1f ms, %zu SW maps %.1f ms, %zu shadow regions realigned %.1f ms\n",
                       t_base - t_start, S, t_sw - t_base, shadows.size(), now() - t_sw);
    // 3./4. CUSUM of the profile differences and greedy peak picking
    std::vector<const std::vector<double>*> profs(S);
    for (size_t s = 0; s < S; s++)
    {
        std::vector<double>& prof = R->seqlikes[seeds[s]];
        if (prof.empty()) prof.assign(seeds[s].size(), 0.0);
        profs[s] = &prof;
    }
    psi_pick_candidates(R->bases, base, seeds, profs, als, found);
    return PS_OK;
}

// PSAlign.Mutate loop body (poreseq/_poreseqcpp.pyx:424-431)
int ps_mutate_loop(ps_region* R, const std::vector<std::string>& seeds, int reps, int* totbases)
{
    int total = 0;
    for (int rep = 0; rep < reps; rep++)
    {
        const bool trace = R->ctx->trace;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = now();
        std::vector<HostMut> cand;
        TRY(ps_find_mutation_list(R, seeds, cand));
        const double t1 = now();
        TRY(ps_score_mutation_list(R, cand));
        const double t2 = now();
        int nb = 0;
        TRY(ps_make_mutation_list(R, cand, &nb));
        if (trace) fprintf(stderr, "[ps] Mutate rep %d: FindMutations %.1f ms (%zu candidates), ScoreMutations %.1f ms, MakeMutations %.1f ms (%d bases)\n",
                           rep, t1 - t0, cand.size(), t2 - t1, now() - t2, nb);
        if (nb == 0) break;
        total += nb;
    }
    *totbases = total;
    return PS_OK;
}

// ------------------------------------------------------------------------------------------
extern "C" {

int ps_swfull(const char* seq1, const char* seq2, int* inds1, int* inds2, int cap, int* n, int* score, double* accuracy)
{
    if (!seq1 || !seq2) return PS_E_ARG;
    SWResult r = psi_swfull(std::string(seq1), std::string(seq2));
    if (n) *n = (int)r.inds1.size();
    if (score) *score = r.score;
    if (accuracy) *accuracy = r.accuracy;
    if ((int)r.inds1.size() > cap) return PS_E_CAPACITY;
    for (size_t k = 0; k < r.inds1.size(); k++) { if (inds1) inds1[k] = r.inds1[k]; if (inds2) inds2[k] = r.inds2[k]; }
    return PS_OK;
}

int ps_swfull_device(ps_ctx* ctx, const char* seq1, const char* seq2, int* inds1, int* inds2, int cap, int* n, int* score, double* accuracy)
{
    if (!ctx || !seq1 || !seq2) return PS_BAD_ARGS(ctx, "ps_swfull_device");
    std::vector<SWResult> out;
    int rc = psi_swfull_batch(ctx, std::string(seq1), std::vector<std::string>(1, std::string(seq2)), out);
    if (rc == PS_E_ARG) { ps_set_error(ctx, "ps_swfull_device: sequences of 1..16384 bases only"); return rc; }
    if (rc) return rc;
    const SWResult& r = out[0];
    if (n) *n = (int)r.inds1.size();
    if (score) *score = r.score;
    if (accuracy) *accuracy = r.accuracy;
    if ((int)r.inds1.size() > cap) return PS_E_CAPACITY;
    for (size_t k = 0; k < r.inds1.size(); k++) { if (inds1) inds1[k] = r.inds1[k]; if (inds2) inds2[k] = r.inds2[k]; }
    return PS_OK;
}

int ps_map_alignments(ps_region* R, const char* newseq)
{
    if (!R || !newseq) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_map_alignments");
    psi_map_alignments(R, std::string(newseq));
    return PS_OK;
}

int ps_find_mutations(ps_region* R, int n_seeds, const char* const* seeds, int* n_found)
{
    if (!R || n_seeds < 0 || (n_seeds > 0 && !seeds)) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_find_mutations");
    std::vector<std::string> sv(n_seeds);
    for (int i = 0; i < n_seeds; i++) sv[i] = seeds[i] ? seeds[i] : "";
    R->seqlikes.clear();               // a stand-alone FindMutations starts from an empty profile cache (fresh AlignData)
    const int rc = ps_find_mutation_list(R, sv, R->found);
    R->seqlikes.clear();
    if (rc) return rc;
    if (n_found) *n_found = (int)R->found.size();
    return PS_OK;
}

int ps_pick_candidates(ps_region* R, int n_seeds, const char* const* seeds, const double* base_profile,
                       const double* const* seed_profiles, int* n_found)
{
    if (!R || n_seeds < 0 || (n_seeds > 0 && (!seeds || !seed_profiles)) || !base_profile)
        return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_pick_candidates");
    std::vector<std::string> sv(n_seeds);
    std::vector<std::vector<double>> pv(n_seeds);
    std::vector<const std::vector<double>*> profs(n_seeds);
    std::vector<SWResult> als(n_seeds);
    for (int s = 0; s < n_seeds; s++)
    {
        if (!seeds[s] || !seed_profiles[s]) return PS_BAD_ARGS(R->ctx, "ps_pick_candidates");
        sv[s] = seeds[s];
        pv[s].assign(seed_profiles[s], seed_profiles[s] + sv[s].size());
        profs[s] = &pv[s];
    }
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    ps_parallel_for(n_seeds, [&](int s) { als[s] = psi_swfull(R->bases, sv[s]); psi_fillinds(als[s]); });
    const std::vector<double> base(base_profile, base_profile + R->bases.size());
    const double t1 = now();
    psi_pick_candidates(R->bases, base, sv, profs, als, R->found);
    if (R->ctx && R->ctx->trace)
        fprintf(stderr, "[ps] pick_candidates: %d SW maps %.1f ms, CUSUM + peak picking %.1f ms (%zu candidates)\n", n_seeds, t1 - t0, now() - t1, R->found.size());
    if (n_found) *n_found = (int)R->found.size();
    return PS_OK;
}

int ps_get_found_mutation(ps_region* R, int i, int* start, char* orig, int orig_cap, char* mut, int mut_cap)
{
    if (!R || i < 0 || i >= (int)R->found.size()) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_get_found_mutation");
    const HostMut& m = R->found[i];
    if ((int)m.orig.size() + 1 > orig_cap || (int)m.mut.size() + 1 > mut_cap) return PS_E_CAPACITY;
    if (start) *start = m.start;
    memcpy(orig, m.orig.c_str(), m.orig.size() + 1);
    memcpy(mut, m.mut.c_str(), m.mut.size() + 1);
    return PS_OK;
}

int ps_found_mutation_sizes(ps_region* R, int i, int* n_orig, int* n_mut)
{
    if (!R || i < 0 || i >= (int)R->found.size()) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_found_mutation_sizes");
    if (n_orig) *n_orig = (int)R->found[i].orig.size();
    if (n_mut) *n_mut = (int)R->found[i].mut.size();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------
// The consensus loop of one region below the C-ABI (poreseq/Mutate.py:47-99): Mutate('self', reps) then up to `reps`
// rounds of (Mutate('viterbi'), Refine) until Refine changes nothing.  One region handle for the whole loop: nothing is
// re-marshalled between the steps (the Python policy builds a fresh native region for every PSAlign method, like the
// reference does, pyx:139-153) and the per-level log(stdv) records are computed once.  The handle's own state after each
// step is exactly what the reference writes back to Python and marshals again for the next one (sequence, ref_align,
// ref_like); ref_index is rebuilt from ref_align with the same arithmetic (HostEvent::update_refs).
static int consensus_one(ps_region* R, int reps, int point_width)
{
    R->stage_log.clear(); R->stage_nbases.clear();
    if (R->events.size() < 5) return PS_OK;                          // Mutate.py:50-53
    if (!R->own_rng) R->rng_seed(1);                                 // one process per region in the reference: rand() starts at seed 1
    auto note = [&](const std::string& name, int nb) { R->stage_log.emplace_back(name, R->bases); R->stage_nbases.push_back(nb); };
    const int scoring_width = R->params.scoring_width;
    std::vector<std::string> seeds;
    for (size_t e = 0; e < R->events.size(); e += 2) seeds.push_back(R->events[e].seq2d);     // pyx:412-414
    int nb = 0;
    R->seqlikes.clear();
    TRY(ps_mutate_loop(R, seeds, reps, &nb));
    R->seqlikes.clear();
    note("mutate_self", nb);
    for (int k = 0; k < reps; k++)
    {
        std::vector<std::string> vit;
        TRY(ps_viterbi_list(R, 16, 0.05, 0.01, 0.33, 0.75, vit));      // pyx:417
        R->seqlikes.clear();
        TRY(ps_mutate_loop(R, vit, 4, &nb));                          // Mutate.py:76: pa.Mutate(seqs='viterbi'), reps defaults to 4
        R->seqlikes.clear();
        note("mutate_viterbi_" + std::to_string(k), nb);
        R->params.scoring_width = point_width;                         // pyx:464-465
        const int rc = ps_refine_region(R, &nb);
        R->params.scoring_width = scoring_width;
        if (rc) return rc;
        note("refine_" + std::to_string(k), nb);
        if (nb == 0) break;
    }
    return PS_OK;
}

extern "C" int ps_consensus(ps_region* R, int reps, int point_width, int* n_stages)
{
    if (!R || reps < 0 || point_width < 0) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_consensus");
    TRY(consensus_one(R, reps, point_width));
    if (n_stages) *n_stages = (int)R->stage_log.size();
    return PS_OK;
}

// Many regions at once: `in_flight` regions are worked on side by side, each by a host thread of the library with a
// context of its own (stream, staging and device buffers) on ctx's device -- the reference's scaling model (one process
// per region, README.md:48-54) folded into one process, below the C-ABI: no interpreter, no marshalling between the
// steps.  The loop of one region is ~400 small dependent launches; regions in flight is what fills the GPU.
extern "C" int ps_consensus_batch(ps_ctx* ctx, ps_region* const* regions, int n_regions, int reps, int point_width, int in_flight)
{
    if (!ctx || n_regions < 0 || (n_regions > 0 && !regions) || reps < 0 || point_width < 0) return PS_BAD_ARGS(ctx, "ps_consensus_batch");
    for (int k = 0; k < n_regions; k++) if (!regions[k]) return PS_BAD_ARGS(ctx, "ps_consensus_batch");
    if (n_regions == 0) return PS_OK;
    TRY(ctx->init());
    {
        std::vector<ps_region*> seen(regions, regions + n_regions);
        std::sort(seen.begin(), seen.end());
        if (std::adjacent_find(seen.begin(), seen.end()) != seen.end())
        {
            ps_set_error(ctx, "ps_consensus_batch: the same region handle appears twice");
            return PS_E_ARG;
        }
    }
    // default: every step of the loop as one job over all regions that are at it (ps_lockstep.cu)
    // lockstep needs one set of band widths for all regions (they share jobs); otherwise regions in flight on threads
    bool uniform = true;
    for (int k = 1; k < n_regions; k++)
    {
        const ps_params &p = regions[k]->params, &q = regions[0]->params;
        if (p.lik_offset != q.lik_offset || p.realign_width != q.realign_width || p.scoring_width != q.scoring_width) uniform = false;
    }
    if (!ctx->threads_consensus && n_regions > 1 && uniform)
    {
        // A few lockstep groups side by side (a host thread + context each, with lanes of their own): while one group's
        // host step runs (staging a job, picking candidates, the accept loops) the GPU works on another group's job.
        // Measured on 64 regions of 1 kb x 10x: 1 group 27 kb/s, 2: 34, 4: 40, 8: 46-49, 16: 53 (regions in flight on
        // threads: 25-43, unstable).
        int groups = ctx->consensus_groups > 0 ? ctx->consensus_groups : std::max(1, std::min(16, n_regions / 4));
        groups = std::max(1, std::min(groups, n_regions));
        if (groups == 1) return ps_consensus_lockstep(ctx, regions, n_regions, reps, point_width, in_flight);
        while ((int)ctx->group_ctx.size() < groups - 1)
        {
            ps_ctx* h = ps_create(ctx->device);
            if (!h) { ps_set_error(ctx, "ps_consensus_batch: out of memory"); return PS_E_INTERNAL; }
            ctx->group_ctx.push_back(h);
        }
        std::vector<std::vector<ps_region*>> part(groups);
        for (int k = 0; k < n_regions; k++) part[k % groups].push_back(regions[k]);
        std::vector<int> rcs(groups, PS_OK);
        const int lanes = std::max(1, in_flight / groups);
        // every group's context grows band buffers of its own: share the device memory between them
        const double keep_budget = ctx->band_budget;
        const bool keep_wait = ctx->blocking_wait;
        const double group_budget = std::min(keep_budget > 0 ? keep_budget : 32e9, 0.6 * (double)ctx->total_mem / groups);
        auto run = [&](int g) {
            ps_ctx* c = g == 0 ? ctx : ctx->group_ctx[g - 1];
            c->precision = ctx->precision;
            c->band_budget = group_budget;
            c->blocking_wait = true;                                     // (restored for ctx below)
            rcs[g] = ps_consensus_lockstep(c, part[g].data(), (int)part[g].size(), reps, point_width, lanes);
        };
        std::vector<std::thread> th;
        for (int g = 1; g < groups; g++) th.emplace_back(run, g);
        run(0);
        for (std::thread& t : th) t.join();
        ctx->band_budget = keep_budget;
        ctx->blocking_wait = keep_wait;
        for (int g = 0; g < groups; g++)
            if (rcs[g]) { if (g) ps_set_error(ctx, "ps_consensus_batch: %s", ctx->group_ctx[g - 1]->error.c_str()); return rcs[g]; }
        return PS_OK;
    }
    in_flight = std::max(1, std::min(std::min(in_flight, n_regions), 64));
    while ((int)ctx->helpers.size() < in_flight - 1)
    {
        ps_ctx* h = ps_create(ctx->device);
        if (!h) { ps_set_error(ctx, "ps_consensus_batch: out of memory"); return PS_E_INTERNAL; }
        ctx->helpers.push_back(h);
    }
    std::vector<ps_ctx*> lanes(1, ctx);
    for (int k = 0; k + 1 < in_flight; k++) { ctx->helpers[k]->precision = ctx->precision; ctx->helpers[k]->blocking_wait = true; lanes.push_back(ctx->helpers[k]); }
    std::atomic<int> next(0);
    std::vector<int> rcs(in_flight, PS_OK);
    std::vector<std::string> errs(in_flight);
    auto work = [&](int lane) {
        ps_ctx* mine = lanes[lane];
        for (;;)
        {
            const int k = next.fetch_add(1);
            if (k >= n_regions) return;
            ps_region* R = regions[k];
            ps_ctx* home = R->ctx;
            R->ctx = mine;                                               // the region's jobs run on this lane's stream and buffers
            const int rc = consensus_one(R, reps, point_width);
            R->ctx = home;
            if (rc && !rcs[lane]) { rcs[lane] = rc; errs[lane] = mine->error; }
        }
    };
    std::vector<std::thread> th;
    for (int lane = 1; lane < in_flight; lane++) th.emplace_back(work, lane);
    work(0);
    for (std::thread& t : th) t.join();
    for (int lane = 0; lane < in_flight; lane++)
        if (rcs[lane]) { ps_set_error(ctx, "ps_consensus_batch: %s", errs[lane].c_str()); return rcs[lane]; }
    return PS_OK;
}

extern "C" int ps_region_num_stages(ps_region* R) { return R ? (int)R->stage_log.size() : PS_E_ARG; }

extern "C" int ps_region_get_stage(ps_region* R, int k, char* name, int name_cap, char* seq, int seq_cap, int* nbases)
{
    if (!R || k < 0 || k >= (int)R->stage_log.size()) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_region_get_stage");
    const auto& st = R->stage_log[k];
    if (nbases) *nbases = R->stage_nbases[k];
    if (name) { if ((int)st.first.size() + 1 > name_cap) return PS_E_CAPACITY; memcpy(name, st.first.c_str(), st.first.size() + 1); }
    if (seq) { if ((int)st.second.size() + 1 > seq_cap) return PS_E_CAPACITY; memcpy(seq, st.second.c_str(), st.second.size() + 1); }
    return (int)st.second.size();
}

int ps_mutate(ps_region* R, int n_seeds, const char* const* seeds, int reps, int* totbases)
{
    if (!R || n_seeds < 0 || (n_seeds > 0 && !seeds)) return PS_BAD_ARGS(R ? R->ctx : nullptr, "ps_mutate");
    std::vector<std::string> sv(n_seeds);
    for (int i = 0; i < n_seeds; i++) sv[i] = seeds[i] ? seeds[i] : "";
    int tot = 0;
    R->seqlikes.clear();               // the reference's profile cache lives for ONE Mutate call (fresh AlignData, pyx:407-408; A.3-14)
    TRY(ps_mutate_loop(R, sv, reps, &tot));
    R->seqlikes.clear();
    if (totbases) *totbases = tot;
    return PS_OK;
}

} // extern "C"
