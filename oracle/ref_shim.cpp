/* ref_shim.cpp -- wraps the UNMODIFIED reference C++ (compiled from /root/reference/cpp where it
 * lies; see oracle/Makefile) behind the flat C interface of oracle_api.h.
 *
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline).  This file contains no algorithm: it only
 * marshals flat arrays into the reference's own AlignData / EventData / MutInfo types
 * (cpp/AlignData.h, cpp/EventData.h, cpp/AlignUtil.h) the same way the reference's Cython layer
 * does (poreseq/_poreseqcpp.pyx:99-153) and calls the reference entry points
 * (cpp/Mutations.h:18-24, cpp/EventUtil.h:17, cpp/Viterbi.h:67, cpp/swlib.h:36).
 */
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "Mutations.h"
#include "EventUtil.h"
#include "Viterbi.h"
#include "swlib.h"

#include "oracle_api.h"

namespace {

AlignData to_align_data(const orc_region* r)
{
    AlignData data;
    data.sequence = Sequence(std::string(r->seq, r->seq_len));
    data.params.lik_offset = r->lik_offset;
    data.params.scoring_width = r->scoring_width;
    data.params.realign_width = r->realign_width;
    data.params.verbose = 0;
    for (int e = 0; e < r->n_events; e++)
    {
        int o = r->lev_off[e];
        int n = r->lev_off[e + 1] - o;
        EventData ev;
        ev.setData(n, const_cast<double*>(r->mean + o), const_cast<double*>(r->stdv + o),
                   r->ref_align + o, r->ref_like + o);
        const double* m = r->model + (size_t)e * 4 * N_STATES;
        ev.model.setData(const_cast<double*>(m), const_cast<double*>(m + N_STATES),
                         const_cast<double*>(m + 2 * N_STATES), const_cast<double*>(m + 3 * N_STATES),
                         false);
        const double* t = r->trans + (size_t)e * 4;
        ev.model.setParams(t[0], t[1], t[2], t[3]);
        if (r->ev_seq && r->ev_seq[e])
            ev.sequence = Sequence(std::string(r->ev_seq[e]));
        data.events.push_back(ev);
    }
    return data;
}

void write_back(orc_region* r, const AlignData& data)
{
    for (int e = 0; e < r->n_events; e++)
    {
        int o = r->lev_off[e];
        int n = r->lev_off[e + 1] - o;
        for (int i = 0; i < n; i++)
        {
            r->ref_align[o + i] = data.events[e].ref_align[i];
            r->ref_like[o + i] = data.events[e].ref_like[i];
        }
    }
}

std::vector<MutInfo> to_muts(int n, const int* start, const char* const* orig, const char* const* mut)
{
    std::vector<MutInfo> v(n);
    for (int i = 0; i < n; i++)
    {
        v[i].start = start[i];
        v[i].orig = orig[i];
        v[i].mut = mut[i];
    }
    return v;
}

std::string dot(const std::string& s) { return s.empty() ? std::string(".") : s; }

std::string text_of(const std::vector<MutScore>& v)
{
    std::string out;
    char buf[64];
    for (size_t i = 0; i < v.size(); i++)
    {
        snprintf(buf, sizeof buf, "%d", v[i].start);
        out += buf; out += '\t'; out += dot(v[i].orig); out += '\t'; out += dot(v[i].mut); out += '\t';
        snprintf(buf, sizeof buf, "%.17g", v[i].score);
        out += buf; out += '\n';
    }
    return out;
}

std::string text_of(const std::vector<MutInfo>& v)
{
    std::string out;
    char buf[64];
    for (size_t i = 0; i < v.size(); i++)
    {
        snprintf(buf, sizeof buf, "%d", v[i].start);
        out += buf; out += '\t'; out += dot(v[i].orig); out += '\t'; out += dot(v[i].mut); out += "\t0\n";
    }
    return out;
}

int put(const std::string& s, char* out, int cap)
{
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return 0;
}

std::vector<Sequence> to_seqs(int n, const char* const* seeds)
{
    std::vector<Sequence> v;
    for (int i = 0; i < n; i++) v.push_back(Sequence(std::string(seeds[i])));
    return v;
}

} // namespace

extern "C" {

const char* orc_name(void) { return "reference"; }

int orc_score_alignments(orc_region* r, double* scores, double* likes)
{
    AlignData data = to_align_data(r);
    std::vector<double> s = ScoreAlignments(data, likes);
    for (size_t i = 0; i < s.size(); i++) scores[i] = s[i];
    write_back(r, data);
    return 0;
}

int orc_score_mutations(orc_region* r, int n, const int* start, const char* const* orig,
                        const char* const* mut, double* scores)
{
    AlignData data = to_align_data(r);
    std::vector<MutScore> s = ScoreMutations(data, to_muts(n, start, orig, mut));
    for (int i = 0; i < n; i++) scores[i] = s[i].score;
    write_back(r, data);
    return 0;
}

int orc_score_points(orc_region* r, char* out, int cap)
{
    AlignData data = to_align_data(r);
    std::vector<MutInfo> m = FindPointMutations(data);
    std::vector<MutScore> s = ScoreMutations(data, m);
    write_back(r, data);
    return put(text_of(s), out, cap);
}

int orc_make_mutations(orc_region* r, int n, const int* start, const char* const* orig,
                       const char* const* mut, const double* scores, char* seq_out, int cap, int* nbases)
{
    AlignData data = to_align_data(r);
    std::vector<MutInfo> m = to_muts(n, start, orig, mut);
    std::vector<MutScore> s(m.begin(), m.end());
    for (int i = 0; i < n; i++) s[i].score = scores[i];
    *nbases = MakeMutations(data, s);
    write_back(r, data);
    return put(data.sequence.bases, seq_out, cap);
}

int orc_refine(orc_region* r, char* seq_out, int cap, int* nbases)
{
    AlignData data = to_align_data(r);
    std::vector<MutInfo> m = FindPointMutations(data);
    std::vector<MutScore> s = ScoreMutations(data, m);
    *nbases = MakeMutations(data, s);
    write_back(r, data);
    return put(data.sequence.bases, seq_out, cap);
}

int orc_find_mutations(orc_region* r, int n_seeds, const char* const* seeds, char* out, int cap)
{
    AlignData data = to_align_data(r);
    std::vector<MutInfo> m = FindMutations(data, to_seqs(n_seeds, seeds));
    write_back(r, data);
    return put(text_of(m), out, cap);
}

int orc_mutate(orc_region* r, int n_seeds, const char* const* seeds, int reps, char* seq_out, int cap,
               int* totbases)
{
    AlignData data = to_align_data(r);
    std::vector<Sequence> sequences = to_seqs(n_seeds, seeds);
    int tot = 0;
    for (int i = 0; i < reps; i++)
    {
        std::vector<MutInfo> m = FindMutations(data, sequences);
        std::vector<MutScore> s = ScoreMutations(data, m);
        int nb = MakeMutations(data, s);
        if (nb == 0) break;
        tot += nb;
    }
    *totbases = tot;
    write_back(r, data);
    return put(data.sequence.bases, seq_out, cap);
}

int orc_viterbi_mutate(orc_region* r, int nkeep, double skip_prob, double stay_prob, double mut_min,
                       double mut_max, char* out, int cap)
{
    AlignData data = to_align_data(r);
    std::vector<Sequence> seqs = ViterbiMutate(data.events, nkeep, skip_prob, stay_prob, mut_min, mut_max, false);
    std::string s;
    for (size_t i = 0; i < seqs.size(); i++) { s += seqs[i].bases; s += '\n'; }
    return put(s, out, cap);
}

int orc_swfull(const char* seq1, const char* seq2, int* inds1, int* inds2, int cap, int* n, int* score,
               double* accuracy)
{
    SWAlignment al = swfull(std::string(seq1), std::string(seq2));
    *n = (int)al.inds1.size();
    *score = al.score;
    *accuracy = al.accuracy;
    if (*n > cap) return -1;
    for (int i = 0; i < *n; i++) { inds1[i] = al.inds1[i]; inds2[i] = al.inds2[i]; }
    return 0;
}

int orc_map_alignments(orc_region* r, const char* newseq)
{
    AlignData data = to_align_data(r);
    MapAlignments(data, Sequence(std::string(newseq)));
    write_back(r, data);
    return 0;
}

int orc_seq_to_states(const char* seq, int len, int* states)
{
    Sequence s(std::string(seq, len));
    for (size_t i = 0; i < s.states.size(); i++) states[i] = s.states[i];
    return (int)s.states.size();
}

} // extern "C"
