/* ps_oracle.cpp -- independent CPU restatement of PoreSeq's event-to-sequence scoring path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared with in tests/ (and the
 * "port" CPU baseline of bench.py); the product (poreseq_b200/) never includes, links or loads it.
 *
 * PARITY PINNED: tests/test_oracle_vs_reference.py checks every entry point of this file against
 * oracle/_ref/libps_ref.so -- the reference's own C++ compiled from /root/reference/cpp -- and
 * against the committed fixtures in tests/golden/ generated from that library
 * (tests/golden/make_golden.py).  The reference itself ships no tests or golden vectors
 * (SURVEY.md section 4), so the compiled reference is the pin.
 *
 * Restated here: the DP path (fills, join, backtrace, scoreMutation, ScoreAlignments, ScoreMutations,
 * FindPointMutations, MakeMutations) and its drivers (swfull, fillinds, MapAlignments, FindMutations, the Mutate
 * loop, ViterbiMutate).
 *
 * The restatement is written dense-array style (one flat band buffer per event, one shared cell
 * routine for both directions) rather than the reference's column-object style.  Each routine
 * cites the reference lines whose behaviour it restates.  Arithmetic is IEEE double, evaluated
 * left to right exactly as the reference parenthesises it, no FMA contraction.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "oracle_api.h"

namespace {

const int NS = 1024;                       /* cpp/AlignUtil.h:19 */
const double NEG = -1e300;                 /* cpp/AlignUtil.h:20  "inf" is 1e300, not IEEE inf */
const double LOG2PI = std::log(2 * M_PI);  /* cpp/AlignUtil.h:24 */

/* step codes, cpp/Alignment.cpp:19-28 */
enum { SKIP = 0, MATCH = 1, INSERT = 2, IGNORE = 3, STAY = 4, EXTEND = 5, IMPLICIT = 255 };

/* ---------------------------------------------------------------- sequence (cpp/Sequence.h) */

/* cpp/Sequence.h:69-100: rolling 10-bit window; -1 iff the LEFTMOST base of the window is not
 * ACGT (only that one is tested); other characters pollute the code silently. */
std::vector<int> states_of(const std::string& b)
{
    std::vector<int> st;
    if (b.size() < 5) return st;
    std::string v = b;
    for (size_t i = 0; i < v.size(); i++)
    {
        char c = v[i];
        v[i] = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : c;
    }
    int cur = 0;
    for (int i = 0; i < 4; i++) cur = (cur << 2) + v[i];
    for (size_t i = 4; i < v.size(); i++)
    {
        if (v[i - 4] < 4) { cur = (NS - 1) & ((cur << 2) + v[i]); st.push_back(cur); }
        else { cur = 0; st.push_back(-1); }
    }
    return st;
}

struct Mut { int start; std::string orig, mut; double score; };

/* cpp/Sequence.h:38-59 */
std::string apply_mut(const std::string& b, const Mut& m)
{
    if ((size_t)m.start >= b.size()) return b;
    std::string out = b.substr(0, m.start) + m.mut;
    size_t rem = m.start + m.orig.size();
    if (rem < b.size()) out += b.substr(rem);
    return out;
}

/* ---------------------------------------------------------------- model + event (cpp/EventData.h) */

struct Model
{
    double lev_mean[NS], lev_stdv[NS], sd_mean[NS], log_lev[NS], sd_lambda[NS], log_lambda[NS];
    double lskip, lstay, lext, lins;
};

struct Event
{
    int n0;
    std::vector<double> mean, stdv, log_stdv, ref_align, ref_like, ref_index;
    int refstart, refend;
    Model model;

    /* cpp/EventData.h:110-169 */
    void updaterefs()
    {
        refstart = refend = -1;
        int a = 0, b = n0 - 1;
        while (a < n0 && !(ref_align[a] > 0)) a++;
        while (b >= 0 && !(ref_align[b] > 0)) b--;
        if (a == n0 || b < 0) { ref_index.clear(); return; }
        refstart = (int)ref_align[a];
        refend = (int)ref_align[b];
        ref_index = ref_align;
        double slope = (ref_align[b] - ref_align[a]) / (double)(b - a);
        double icpt = ref_align[a] - slope * a;
        int last = -1;
        for (int i = 0; i < n0; i++)
        {
            if (i < a || i > b) ref_index[i] = slope * i + icpt;
            else if (ref_align[i] > 0)
            {
                if (last > 0)    /* sic: a first aligned level at index 0 is never interpolated from */
                {
                    double m = (ref_align[i] - ref_align[last]) / (i - last);
                    for (int j = last + 1; j < i; j++) ref_index[j] = m * (j - last) + ref_align[last];
                }
                last = i;
            }
        }
    }

    /* cpp/EventData.h:172-183: std::lower_bound restated as the explicit halving search */
    int centre(int c) const
    {
        int first = 0, len = (int)ref_index.size();
        double v = (double)c;
        while (len > 0)
        {
            int half = len >> 1;
            if (ref_index[first + half] < v) { first += half + 1; len -= half + 1; }
            else len = half;
        }
        return first;
    }
};

/* cpp/EventData.h:48-73, 208-224 */
void load_event(const orc_region* r, int e, Event& ev)
{
    int o = r->lev_off[e];
    ev.n0 = r->lev_off[e + 1] - o;
    ev.mean.assign(r->mean + o, r->mean + o + ev.n0);
    ev.stdv.assign(r->stdv + o, r->stdv + o + ev.n0);
    ev.ref_align.assign(r->ref_align + o, r->ref_align + o + ev.n0);
    ev.ref_like.assign(r->ref_like + o, r->ref_like + o + ev.n0);
    ev.log_stdv.resize(ev.n0);
    for (int i = 0; i < ev.n0; i++) ev.log_stdv[i] = std::log(ev.stdv[i]);
    const double* m = r->model + (size_t)e * 4 * NS;
    for (int s = 0; s < NS; s++)
    {
        ev.model.lev_mean[s] = m[s];
        ev.model.lev_stdv[s] = m[NS + s];
        ev.model.sd_mean[s] = m[2 * NS + s];
        double sdsd = m[3 * NS + s];
        ev.model.log_lev[s] = std::log(m[NS + s]);
        ev.model.sd_lambda[s] = std::pow(m[2 * NS + s], 3) / std::pow(sdsd, 2);
        ev.model.log_lambda[s] = std::log(ev.model.sd_lambda[s]);
    }
    const double* t = r->trans + (size_t)e * 4;
    ev.model.lskip = std::log(t[0]);
    ev.model.lstay = std::log(t[1]);
    ev.model.lext = std::log(t[2]);
    ev.model.lins = std::log(t[3]);
    ev.updaterefs();
}

void store_event(orc_region* r, int e, const Event& ev)
{
    int o = r->lev_off[e];
    for (int i = 0; i < ev.n0; i++) { r->ref_align[o + i] = ev.ref_align[i]; r->ref_like[o + i] = ev.ref_like[i]; }
}

/* cpp/AlignUtil.h:34-53 plus the "+ lik_offset" of cpp/Alignment.cpp:169-173 / 345-349.
 * level = 0-based level whose mean/stdv are used; lsd_level = level whose log(stdv) is used
 * (the forward pass reads log_stdv[n0-i] with stdv[i-1]: quirk A.3-1, reproduced). */
double emission(const Event& ev, int s, int level, int lsd_level, double offset)
{
    const Model& m = ev.model;
    double d = (ev.mean[level] - m.lev_mean[s]) / m.lev_stdv[s];
    double l = -0.5 * (d * d + LOG2PI) - m.log_lev[s];
    double x = ev.stdv[level];
    double g = (x - m.sd_mean[s]) / m.sd_mean[s];
    l += 0.5 * (m.log_lambda[s] - 3 * ev.log_stdv[lsd_level] - LOG2PI - g * g * m.sd_lambda[s] / x);
    l += offset;
    return l;
}

/* ---------------------------------------------------------------- banded columns */

struct Best { double score; int i, j; Best() : score(0), i(0), j(0) {} };

/* One band column: rows i0 .. i0+len-1, two matrices (main C, stay S), emissions, step bytes. */
struct Column
{
    int i0, len, col;
    std::vector<double> C, S, E;
    std::vector<unsigned char> stepC, stepS;
    Best best;
    Column() : i0(0), len(0), col(0) {}
    void shape(int i0_, int len_, int col_)
    {
        i0 = i0_; len = len_; col = col_;
        C.assign(len, 0.0); S.assign(len, 0.0); E.assign(len, 0.0);
        stepC.assign(len, 0); stepS.assign(len, 0);
    }
    bool has(int i) const { return i >= i0 && i < i0 + len; }
};

/* Fill one column given the previous one (cpp/Alignment.cpp:111-274 forward, :280-444 reverse).
 *   forward : row i <-> level i-1, emission of the destination cell is added on match/stay/extend
 *   reverse : row i <-> level n0-i, transitions add the emission of the SOURCE cell
 *             (prev column's E[i-1] for match, this column's E[i-1] for stay/extend)
 * c is the 1-based state column of `states` this column represents; colid is what is stored in
 * Column::col (c for forward, -(k) for the k-th reverse column). */
void fill_column(const Event& ev, const std::vector<int>& states, int c, int colid, int width,
                 double offset, bool reverse, const Column& P, Column& Q)
{
    int n0 = ev.n0;
    int s = states[c - 1];
    int mid = 1;
    if (!ev.ref_index.empty()) mid = reverse ? n0 - ev.centre(c) + 1 : ev.centre(c);
    if (mid < 1) mid = 1;
    if (mid > n0) mid = n0;
    int i0 = std::max(1, mid - width), i1 = std::min(n0, mid + width);
    Q.shape(i0, i1 - i0 + 1, colid);
    Q.best = P.best;
    if (s < 0) return;                                   /* :162-163 */

    for (int i = i0; i <= i1; i++)
        Q.E[i - i0] = reverse ? emission(ev, s, n0 - i, n0 - i, offset) : emission(ev, s, i - 1, n0 - i, offset);

    const Model& m = ev.model;
    int p0 = P.i0, p1 = P.i0 + P.len - 1;
    double upC = 0, upS = 0, upE = 0;
    for (int i = i0; i <= i1; i++)
    {
        int r = i - i0;
        double e = Q.E[r];
        double cand[4] = {0, 0, 0, 0};
        unsigned char code[4] = {SKIP, MATCH, INSERT, IGNORE};
        if (i >= p0 && i <= p1) cand[0] = P.C[i - p0] + m.lskip;
        else { cand[0] = m.lskip; code[0] = IMPLICIT; }
        if (i > p0 && i <= p1)
        {
            cand[1] = reverse ? P.C[i - 1 - p0] + P.E[i - 1 - p0] : P.C[i - 1 - p0] + e;
            cand[3] = P.C[i - 1 - p0] + m.lins;
        }
        else { cand[1] = reverse ? 0.0 : e; code[1] = IMPLICIT; }
        double stay = NEG, ext = NEG;
        if (i > i0)
        {
            double src = reverse ? upE : e;
            stay = upC + src + m.lstay;
            cand[2] = upC + m.lins;
            ext = upS + src + m.lext;
        }
        double S = (i == i0) ? NEG : 0.0;
        unsigned char sS = 0;
        if (stay > S) { S = stay; sS = STAY; }
        if (ext > S) { S = ext; sS = EXTEND; }
        double Cv = 0;
        unsigned char sC = 0;
        for (int k = 0; k < 4; k++)
            if (cand[k] > Cv) { Cv = cand[k]; sC = code[k]; }
        if (S > Cv) { Cv = S; sC = STAY; }
        Q.C[r] = Cv; Q.S[r] = S; Q.stepC[r] = sC; Q.stepS[r] = sS;
        if (Cv > Q.best.score) { Q.best.score = Cv; Q.best.i = i; Q.best.j = c; }
        upC = Cv; upS = S; upE = e;
    }
}

/* cpp/Alignment.h:178-214: best forward+backward join between a forward and a reverse column. */
double column_join(const Event& ev, const Column& F, const Column& B)
{
    double sm = 0;
    for (int jf = 1; jf <= ev.n0; jf++)
    {
        int jb = ev.n0 - jf + 1;
        for (int k = 0; k < 2; k++)
        {
            double s = 0;
            if (F.has(jf)) s += k ? F.S[jf - F.i0] : F.C[jf - F.i0];
            if (B.has(jb)) s += k ? B.S[jb - B.i0] : B.C[jb - B.i0];
            sm = std::max(s, sm);
        }
        sm = std::max(sm, F.best.score);
        sm = std::max(sm, B.best.score);
    }
    return sm;
}

/* One event aligned against one sequence: forward columns 0..N and reverse columns 0..N. */
struct Aligner
{
    Event* ev;
    const std::vector<int>* states;
    double offset;
    int realign_width, scoring_width;
    bool usable;                      /* cpp/Alignment.cpp:51-59: decided once, at construction */
    std::vector<Column> F, B;

    Aligner(Event& e, const orc_region* r) : ev(&e), states(0), offset(r->lik_offset),
        realign_width(r->realign_width), scoring_width(r->scoring_width)
    {
        usable = !e.ref_index.empty() && realign_width != 0;
    }

    void blank(std::vector<Column>& v) { v.assign(1, Column()); v[0].shape(0, ev->n0 + 1, 0); }

    void forward(const std::vector<int>& st)       /* cpp/Alignment.cpp:84-91 */
    {
        states = &st;
        blank(F);
        if (!usable) return;
        int N = (int)st.size();
        F.resize(N + 1);
        for (int c = 1; c <= N; c++) fill_column(*ev, st, c, c, realign_width, offset, false, F[c - 1], F[c]);
    }

    void backward(const std::vector<int>& st)      /* cpp/Alignment.cpp:93-100 */
    {
        blank(B);
        if (!usable) return;
        int N = (int)st.size();
        B.resize(N + 1);
        for (int k = 1; k <= N; k++) fill_column(*ev, st, N - k + 1, -k, realign_width, offset, true, B[k - 1], B[k]);
    }

    /* cpp/Alignment.cpp:516-624 */
    void backtrace()
    {
        if (!usable) return;
        std::vector<int> li, lj;
        std::vector<double> ll;
        int i = F.back().best.i, j = F.back().best.j, arr = 0;
        while (i > 0)
        {
            const Column& q = F[j];
            int r = i - q.i0;
            unsigned char st = arr ? q.stepS[r] : q.stepC[r];
            double sc = arr ? q.S[r] : q.C[r];
            if (sc <= 0.0) break;
            if (st == SKIP) j--;
            else if (st == MATCH) { li.push_back(i); lj.push_back(j); ll.push_back(sc); i--; j--; }
            else if (st == IGNORE) { li.push_back(i); lj.push_back(-1); ll.push_back(sc); i--; j--; }
            else if (st == INSERT) { li.push_back(i); lj.push_back(-1); ll.push_back(sc); i--; }
            else if (st == STAY)
            {
                if (arr == 1) { li.push_back(i); lj.push_back(j); ll.push_back(sc); i--; }
                arr = 1 - arr;
            }
            else if (st == EXTEND) { li.push_back(i); lj.push_back(j); ll.push_back(sc); i--; }
            else i = 0;
        }
        std::fill(ev->ref_align.begin(), ev->ref_align.end(), 0.0);
        std::fill(ev->ref_like.begin(), ev->ref_like.end(), 0.0);
        for (size_t k = 0; k < li.size(); k++) { ev->ref_align[li[k] - 1] = lj[k]; ev->ref_like[li[k] - 1] = ll[k]; }
        ev->updaterefs();
    }

    double total() const { return std::max(F.back().best.score, B.back().best.score); }   /* Alignment.h:127 */

    double join_at(int raf, int rab) const            /* clamps of cpp/Alignment.h:181-185 */
    {
        if (raf >= (int)F.size()) raf = (int)F.size() - 1;
        if (rab >= (int)B.size()) rab = (int)B.size() - 1;
        if (raf < 0) raf = 0;
        if (rab < 0) rab = 0;
        return column_join(*ev, F[raf], B[rab]);
    }

    /* cpp/Alignment.cpp:447-512.  The reference appends phony columns to its forward vector and
     * pops them again; here they live in a local vector. */
    double score_mutation(const Mut& mu, const std::vector<int>& mst)
    {
        if (!usable) return 0;
        int N = (int)states->size(), Nm = (int)mst.size();
        double old = join_at(std::max(mu.start - 3, 1), N - std::max(mu.start - 3, 1) + 1);
        int startind = std::max(mu.start - 4, 0);
        std::vector<Column> T(1, F[startind]);                 /* T[0] is the shared seed column */
        int want = (int)mu.mut.size() + 6;
        for (int n = 0; n < want && scoring_width != 0; n++)
        {
            int c = T.back().col + 1;
            if (c > Nm) break;
            T.push_back(Column());
            fill_column(*ev, mst, c, c, scoring_width, offset, false, T[T.size() - 2], T.back());
        }
        int refind = mu.start + (int)mu.mut.size() + 1;
        int f = (int)T.size() - 1;
        while (f > 0 && T[f].col > refind) f--;
        refind = T[f].col;
        int rab = Nm - refind + 1;
        if (rab >= (int)B.size()) rab = (int)B.size() - 1;
        if (rab < 0) rab = 0;
        double neu = column_join(*ev, T[f], B[rab]);
        return neu - old;
    }
};

/* ---------------------------------------------------------------- drivers */

struct Region
{
    std::string bases;
    std::vector<int> states;
    std::vector<Event> events;
    const orc_region* raw;

    explicit Region(const orc_region* r) : bases(r->seq, r->seq_len), raw(r)
    {
        states = states_of(bases);
        events.resize(r->n_events);
        for (int e = 0; e < r->n_events; e++) load_event(r, e, events[e]);
    }
    void set_sequence(const std::string& b) { bases = b; states = states_of(b); }
    void store(orc_region* r) const { for (int e = 0; e < r->n_events; e++) store_event(r, e, events[e]); }
};

/* cpp/MakeMutations.cpp:148-195 */
std::vector<double> score_alignments(Region& R, double* likes)
{
    std::vector<double> out;
    std::vector<Aligner> al;
    for (size_t e = 0; e < R.events.size(); e++) al.push_back(Aligner(R.events[e], R.raw));
    for (size_t e = 0; e < al.size(); e++)
    {
        al[e].forward(R.states);
        al[e].blank(al[e].B);
        al[e].backtrace();
        out.push_back(al[e].total());
        if (likes)
        {
            const Event& ev = R.events[e];
            double last = 0;
            int at = 1;
            for (int j = 0; j < ev.n0; j++)
                if (ev.ref_align[j] > 0)
                {
                    for (int k = at; k < ev.ref_align[j]; k++) likes[k + 1] += last;
                    last = ev.ref_like[j];
                    at = (int)ev.ref_align[j];
                }
            for (int k = at; k < (int)R.states.size() + 3; k++) likes[k + 1] += last;
        }
        al[e].F.clear(); al[e].B.clear();
    }
    return out;
}

/* cpp/MakeMutations.cpp:23-69: events outer, mutations inner; score starts at -1e-6 */
void score_mutations(Region& R, std::vector<Mut>& muts)
{
    for (size_t i = 0; i < muts.size(); i++) muts[i].score = -1e-6;
    std::vector<Aligner> al;
    for (size_t e = 0; e < R.events.size(); e++) al.push_back(Aligner(R.events[e], R.raw));
    for (size_t e = 0; e < al.size(); e++)
    {
        al[e].forward(R.states);
        al[e].backward(R.states);
        al[e].backtrace();
        for (size_t i = 0; i < muts.size(); i++)
        {
            if ((size_t)muts[i].start > R.bases.size()) continue;
            std::vector<int> mst = states_of(apply_mut(R.bases, muts[i]));
            muts[i].score += al[e].score_mutation(muts[i], mst);
        }
        al[e].F.clear(); al[e].B.clear();
    }
}

/* cpp/FindMutations.cpp:191-234 */
std::vector<Mut> point_mutations(const Region& R)
{
    static const char* acgt = "ACGT";
    std::vector<Mut> v;
    for (int i = 0; i < (int)R.states.size(); i++)
    {
        Mut m; m.start = i; m.score = 0;
        m.orig = std::string(1, R.bases[i]); m.mut = "";
        v.push_back(m);
        for (int j = 0; j < 4; j++)
            if (R.bases[i] != acgt[j]) { m.mut = std::string(1, acgt[j]); v.push_back(m); }
        m.orig = "";
        for (int j = 0; j < 4; j++) { m.mut = std::string(1, acgt[j]); v.push_back(m); }
    }
    return v;
}

bool operator<(const Mut& a, const Mut& b) { return a.score > b.score; }   /* cpp/MakeMutations.cpp:16 */

/* cpp/MakeMutations.cpp:74-146 */
int make_mutations(Region& R, std::vector<Mut> muts)
{
    const int spacing = 10;
    int changed = 0;
    std::sort(muts.begin(), muts.end());
    while (!muts.empty() && muts.back().score < 0) muts.pop_back();
    if (muts.empty()) return 0;
    std::vector<Mut> later;
    for (size_t i = 0; i < muts.size(); i++)
    {
        if (muts[i].score < 0) { later.push_back(muts[i]); continue; }
        R.set_sequence(apply_mut(R.bases, muts[i]));
        changed += (int)std::max(muts[i].orig.size(), muts[i].mut.size());
        for (size_t j = i + 1; j < muts.size(); j++)
        {
            int lo = std::max(muts[i].start, muts[j].start);
            int hi = (int)std::min(muts[i].start + muts[i].mut.size(), muts[j].start + muts[j].mut.size());
            if (lo < hi + spacing && muts[j].score > 0) { muts[j].score = -1; continue; }
            if ((size_t)muts[j].start >= muts[i].start + muts[i].orig.size())
                muts[j].start += (int)(muts[i].mut.size() - muts[i].orig.size());
        }
    }
    if (later.size() > 10)
    {
        score_mutations(R, later);
        changed += make_mutations(R, later);
    }
    return changed;
}

/* ------------------------------------------------------------- sequence alignment (cpp/swlib.cpp) */

/* cpp/swlib.h:21-33 */
struct Pairing { int score; double accuracy; std::vector<int> a, b; Pairing() : score(0), accuracy(0) {} };

/* swfull, cpp/swlib.cpp:211-340, restated with two rolling score rows and one byte per cell that holds the
 * move (1 = gap in seq1, 2 = gap in seq2, 3 = pair) and whether the cell's score is 0 (where the reference's
 * traceback stops, :299-301; scores are never negative).  Compare order and ties as in :246-268: the gap moves
 * win only when strictly better, the pairing move wins ties (also the tie with 0); the first maximum in
 * (seq2 position, seq1 position) order is the end of the alignment (:273-278). */
Pairing sw_align(const std::string& s1, const std::string& s2)
{
    const int n1 = (int)s1.size(), n2 = (int)s2.size();
    const int match = 5, mismatch = -4, gap = -8;                      /* cpp/swlib.h:21-23 */
    std::vector<unsigned char> cell((size_t)(n1 + 1) * (n2 + 1), 4);   /* border: score 0 */
    std::vector<int> before(n1 + 1, 0), now(n1 + 1, 0);
    int top = 0, ti = 0, tj = 0;
    for (int j = 1; j <= n2; j++)
    {
        now[0] = 0;
        unsigned char* row = &cell[(size_t)j * (n1 + 1)];
        for (int i = 1; i <= n1; i++)
        {
            int best = 0, how = 0;
            const int from_j = before[i] + gap, from_i = now[i - 1] + gap;
            const int both = before[i - 1] + (s1[i - 1] == s2[j - 1] ? match : mismatch);
            if (from_j > best) { best = from_j; how = 1; }
            if (from_i > best) { best = from_i; how = 2; }
            if (both >= best) { best = both; how = 3; }
            now[i] = best;
            row[i] = (unsigned char)(how | (best <= 0 ? 4 : 0));
            if (best > top) { top = best; ti = i; tj = j; }
        }
        before.swap(now);
    }
    Pairing out;
    out.score = top;
    int i = ti, j = tj, same = 0;
    while (i > 0 && j > 0)
    {
        const unsigned char c = cell[(size_t)j * (n1 + 1) + i];
        if (c & 4) break;
        const int how = c & 3;
        if (how == 1) { out.a.push_back(0); out.b.push_back(j); j--; }
        else if (how == 2) { out.a.push_back(i); out.b.push_back(0); i--; }
        else { out.a.push_back(i); out.b.push_back(j); if (s1[i - 1] == s2[j - 1]) same++; i--; j--; }
    }
    std::reverse(out.a.begin(), out.a.end());
    std::reverse(out.b.begin(), out.b.end());
    out.accuracy = 100.0 * same / (double)out.a.size();
    return out;
}

/* fillinds, cpp/swlib.cpp:342-364: a gap repeats the last index seen on its side (the first entry as it is) */
void fill_gaps(Pairing& p)
{
    if (p.a.empty()) return;
    int ka = p.a[0], kb = p.b[0];
    for (size_t k = 0; k < p.a.size(); k++)
    {
        if (p.a[k] > 0) ka = p.a[k]; else p.a[k] = ka;
        if (p.b[k] > 0) kb = p.b[k]; else p.b[k] = kb;
    }
}

/* forward declarations of the region-level routines further down */
struct Region;
std::vector<double> score_alignments(Region& R, double* likes);

/* MapAlignments, cpp/EventUtil.cpp:12-55: every level's column is sent through the pairing of the old and the
 * new sequence -- the first pairing entry whose old-side index is not below it (std::lower_bound, :41) -- levels
 * outside the paired stretch lose their alignment; then updaterefs. */
Pairing map_alignments(Region& R, const std::string& newseq)
{
    Pairing p = sw_align(R.bases, newseq);
    fill_gaps(p);
    R.set_sequence(newseq);
    for (size_t e = 0; e < R.events.size(); e++)
    {
        Event& ev = R.events[e];
        for (size_t j = 0; j < ev.ref_align.size(); j++)
        {
            const int col = (int)ev.ref_align[j];
            if (p.a.empty() || col < p.a.front() || col > p.a.back()) { ev.ref_align[j] = 0; continue; }
            const size_t at = std::lower_bound(p.a.begin(), p.a.end(), col) - p.a.begin();
            ev.ref_align[j] = at < p.b.size() ? p.b[at] : 0;
        }
        ev.updaterefs();
    }
    return p;
}

typedef std::map<std::string, std::vector<double> > ProfileCache;     /* cpp/AlignData.h:34 */

/* FindMutations, cpp/FindMutations.cpp:24-186.  Per-base likelihood profiles of the current sequence and of every
 * seed (events remapped onto the seed, realigned; cached by seed string, :44-49), compared along the pairing of
 * the two sequences: differences of consecutive profile values, a CUSUM of (seed - current) clamped at 0 and
 * zeroed where the two differences agree to 1e-5 (:83-94); then the greedy peak picking of :111-183. */
std::vector<Mut> find_mutations(Region& R, const std::vector<std::string>& seeds, ProfileCache& cache)
{
    std::vector<double> mine(R.bases.size(), 0.0);
    score_alignments(R, mine.data());
    std::vector<std::vector<double> > gain;
    std::vector<Pairing> pairs;
    for (size_t s = 0; s < seeds.size(); s++)
    {
        Region other(R);
        Pairing p = map_alignments(other, seeds[s]);
        std::vector<double>& prof = cache[seeds[s]];
        if (prof.empty())
        {
            prof.assign(seeds[s].size(), 0.0);
            score_alignments(other, prof.data());
        }
        /* "because i did it in matlab code" (:51-63): both index lists minus 2, leading negatives dropped */
        for (size_t k = 0; k < p.a.size(); k++) { p.a[k] -= 2; p.b[k] -= 2; }
        size_t drop = 0;
        while (drop < p.a.size() && (p.a[drop] < 0 || p.b[drop] < 0)) drop++;
        p.a.erase(p.a.begin(), p.a.begin() + drop);
        p.b.erase(p.b.begin(), p.b.begin() + drop);
        const size_t n = p.a.size();
        std::vector<double> d1(n), d2(n), g(n);
        for (size_t k = 0; k < n; k++) { d1[k] = mine[p.a[k]]; d2[k] = prof[p.b[k]]; }
        for (size_t k = n; k-- > 1;) { d1[k] -= d1[k - 1]; d2[k] -= d2[k - 1]; }
        if (n) { d1[0] = 0; d2[0] = 0; }
        double run = 0;
        for (size_t k = 0; k < n; k++)
        {
            run += d2[k] - d1[k];
            if (run < 0) run = 0;
            g[k] = std::fabs(d1[k] - d2[k]) < 1e-5 ? 0.0 : run;
        }
        gain.push_back(g);
        pairs.push_back(p);
    }
    std::vector<Mut> found;
    while (found.size() < R.bases.size() / 3)
    {
        /* the seed with the highest peak (first on ties), the first position of that peak */
        size_t who = 0, where = 0;
        double peak = 0;
        bool any = false;
        for (size_t s = 0; s < gain.size(); s++)
        {
            if (gain[s].empty()) continue;
            const size_t at = std::max_element(gain[s].begin(), gain[s].end()) - gain[s].begin();
            if (!any || gain[s][at] > peak) { any = true; peak = gain[s][at]; who = s; where = at; }
        }
        if (!any) break;
        std::vector<double>& g = gain[who];
        if (g[where] < 0.25) break;
        /* the stretch between the exact zeros around the peak (:134-145) */
        long right = (long)where;
        while (right < (long)g.size() && g[right] != 0) right++;
        long left = (long)where;
        while (left >= 0 && g[left] != 0) left--;
        if (left < 0) left = 0;
        if (right >= (long)g.size()) right = (long)g.size() - 1;
        const Pairing& p = pairs[who];
        const int from1 = p.a[left], from2 = p.b[left], to1 = p.a[where], to2 = p.b[where];
        Mut m;
        m.start = from1;
        m.orig = R.bases.substr(from1, to1 - from1);
        m.mut = seeds[who].substr(from2, to2 - from2);
        m.score = -1e-6;
        while (!m.orig.empty() && !m.mut.empty() && m.orig[0] == m.mut[0]) { m.orig.erase(0, 1); m.mut.erase(0, 1); m.start++; }
        while (!m.orig.empty() && !m.mut.empty() && m.orig[m.orig.size() - 1] == m.mut[m.mut.size() - 1])
        {
            m.orig.erase(m.orig.size() - 1); m.mut.erase(m.mut.size() - 1);
        }
        if (!m.orig.empty() || !m.mut.empty()) found.push_back(m);
        std::fill(g.begin() + left, g.begin() + right + 1, 0.0);
    }
    return found;
}

/* ------------------------------------------------------------- ViterbiMutate (cpp/Viterbi.cpp) */

/* One position of the 1024-state chain: Viterbi scores with back pointers and normalised forward probabilities. */
struct Layer { std::vector<double> lik, fwd; std::vector<int> from; Layer() : lik(NS), fwd(NS), from(NS) {} };

/* V_LIK::V_LIK, cpp/Viterbi.cpp:39-102: a state is entered by an advance of 1, 2 or 3 bases -- predecessor
 * (d >> 2j) + (k << (10 - 2j)), cpp/Viterbi.h:28-29, weight .25^j skip^(j-1), in that order, the first best wins --
 * or by staying; the forward sum runs over the same terms, is multiplied by exp(obs) and normalised by the
 * reciprocal of the plain sum (cpp/Viterbi.h:56-64). */
void chain_step(const Layer& prev, const std::vector<double>& obs, double skip, double stay, Layer& out)
{
    const double lskip = std::log(skip), lstay = std::log(stay);
    for (int d = 0; d < NS; d++)
    {
        double top = NEG, mass = 0.0;
        int arg = -1;
        double w = 0.25, lw = std::log(0.25);
        for (int j = 1; j <= 3; j++)
        {
            for (int k = 0; k < (1 << (2 * j)); k++)
            {
                const int p = (d >> (2 * j)) + (k << (10 - 2 * j));
                double l = obs[d] + lw;
                l += prev.lik[p];
                mass += w * prev.fwd[p];
                if (l > top) { top = l; arg = p; }
            }
            w = w * 0.25 * skip;
            lw = lw + std::log(0.25) + lskip;
        }
        const double l = obs[d] + lstay + prev.lik[d];
        if (l > top) { top = l; arg = d; }
        mass += stay * prev.fwd[d];
        mass *= std::exp(obs[d]);
        out.lik[d] = top; out.from[d] = arg; out.fwd[d] = mass;
    }
    double tot = 0;
    for (int d = 0; d < NS; d++) tot += out.fwd[d];
    tot = 1.0 / tot;
    for (int d = 0; d < NS; d++) out.fwd[d] *= tot;
}

/* StatesToSequence, cpp/Viterbi.cpp:171-237: first base of the first state; a change of state emits the bases the
 * smallest advance (1..4, then smallest index) that explains it has shifted out, or, if none does, the first base
 * of the new state; repeats are stays; finally the last four bases of the last state. */
std::string path_to_bases(const std::vector<int>& path)
{
    /* base `at` (0 = leftmost) of a 5-mer state, cpp/Viterbi.h:35-39 */
    struct { char operator()(int st, int at) const { return "ACGT"[3 & (st >> (2 * (4 - at)))]; } } base;
    std::string seq;
    int cur = path[0];
    seq.push_back(base(cur, 0));
    for (size_t i = 1; i < path.size(); i++)
    {
        if (path[i] == cur) continue;
        int adv = 0;
        for (int n = 1; n <= 4 && !adv; n++)
            for (int k = 0; k < (1 << (2 * n)); k++)
                if ((((cur << (2 * n)) & (NS - 1)) + k) == path[i]) { adv = n; break; }
        if (adv) for (int j = 1; j <= adv; j++) seq.push_back(base(cur, j));
        else seq.push_back(base(path[i], 0));
        cur = path[i];
    }
    for (int j = 1; j <= 4; j++) seq.push_back(base(cur, j));
    return seq;
}

/* ViterbiMutate, cpp/Viterbi.cpp:239-426.  Positions run from the smallest refstart; per position
 * every read contributes the pdf of the MEAN of its levels aligned there (getrefstates, cpp/EventData.h:187-204: the
 * first level whose ref_index equals the position exactly, then the following levels while ref_align <= position,
 * keeping the aligned ones), without lik_offset; per state the lowest quarter of the reads is dropped and the rest
 * averaged (:327-343); positions with too few reads are skipped, the chain ends where no read is left (:310-325).
 * nkeep = 0: the best path; otherwise nkeep sampled paths, drawn backwards with rand() from
 * T[cur][i] * fwd_i^atten (:105-131, T with FOUR advance lengths and the diagonal overwritten by stay, :134-169). */
std::vector<std::string> viterbi_mutate(Region& R, int nkeep, double skip, double stay, double mut_min, double mut_max)
{
    const int E = (int)R.events.size();
    std::vector<Layer> layers(1);
    for (int d = 0; d < NS; d++) { layers[0].lik[d] = 0; layers[0].from[d] = -1; layers[0].fwd[d] = 1.0 / NS; }
    int pos = R.events[0].refstart;
    for (int e = 0; e < E; e++) pos = std::min(pos, R.events[e].refstart);
    std::vector<double> pool((size_t)NS * E), obs(NS);
    for (;;)
    {
        int used = 0;
        for (int e = 0; e < E; e++)
        {
            const Event& ev = R.events[e];
            std::vector<int> at;
            for (size_t i = 0; i < ev.ref_index.size(); i++)
                if (ev.ref_index[i] == pos)
                {
                    at.push_back((int)i);
                    for (int q = (int)i + 1; q < ev.n0 && ev.ref_align[q] <= pos; q++)
                        if (ev.ref_align[q] > 0) at.push_back(q);
                    break;
                }
            if (at.empty()) continue;
            used++;
            double lvl = 0, sd = 0;
            for (size_t q = 0; q < at.size(); q++) { lvl += ev.mean[at[q]]; sd += ev.stdv[at[q]]; }
            lvl = lvl / at.size();
            sd = sd / at.size();
            const Model& m = ev.model;
            const double lsd = std::log(sd);
            for (int st = 0; st < NS; st++)
            {
                const double d = (lvl - m.lev_mean[st]) / m.lev_stdv[st];
                double l = -0.5 * (d * d + LOG2PI) - m.log_lev[st];
                const double g = (sd - m.sd_mean[st]) / m.sd_mean[st];
                l += 0.5 * (m.log_lambda[st] - 3 * lsd - LOG2PI - g * g * m.sd_lambda[st] / sd);
                pool[(size_t)st * E + used - 1] = l;
            }
        }
        int covering = 0;
        for (int e = 0; e < E; e++)
            if (pos >= R.events[e].refstart && pos <= R.events[e].refend) covering++;
        if (used <= covering * 0.2)
        {
            if (covering == 0) break;
            pos++;
            continue;
        }
        if (used > 1)
        {
            int drop = (int)std::floor(used * 0.25);
            if (drop > used - 2) drop = 0;
            for (int st = 0; st < NS; st++)
            {
                double* v = &pool[(size_t)st * E];
                std::sort(v, v + used);
                double sum = 0.0;
                for (int q = drop; q < used; q++) sum += v[q];
                obs[st] = sum / (used - drop);
            }
        }
        else
            for (int st = 0; st < NS; st++) obs[st] = pool[(size_t)st * E];
        layers.push_back(Layer());
        chain_step(layers[layers.size() - 2], obs, skip, stay, layers.back());
        pos++;
    }
    std::vector<std::string> out;
    const Layer& last = layers.back();
    const int head = (int)(std::max_element(last.lik.begin(), last.lik.end()) - last.lik.begin());
    const int n = (int)layers.size() - 1;
    std::vector<int> path;
    if (nkeep == 0)
    {
        int cur = head;
        for (int i = n - 1; i >= 0; i--) { path.push_back(cur); cur = layers[i + 1].from[cur]; }
        std::reverse(path.begin(), path.end());
        out.push_back(path_to_bases(path));
        return out;
    }
    std::vector<double> T((size_t)NS * NS, 0.0);
    for (int d = 0; d < NS; d++)
    {
        double w = 0.25;
        for (int j = 1; j <= 4; j++)
        {
            for (int k = 0; k < (1 << (2 * j)); k++) T[(size_t)d * NS + (d >> (2 * j)) + (k << (10 - 2 * j))] += w;
            w = w * 0.25 * skip;
        }
    }
    for (int d = 0; d < NS; d++) T[(size_t)d * (NS + 1)] = stay;
    std::vector<double> pr(NS);
    for (int s = 0; s < nkeep; s++)
    {
        const double atten = mut_min + (mut_max - mut_min) * s / (double)nkeep;
        path.clear();
        int cur = head;
        for (int i = n - 1; i >= 0; i--)
        {
            path.push_back(cur);
            const Layer& L = layers[i + 1];
            const double r = rand() / (double(RAND_MAX) + 1);
            for (int q = 0; q < NS; q++) pr[q] = T[(size_t)cur * NS + q] * std::pow(L.fwd[q], atten);
            double tot = 0;
            for (int q = 0; q < NS; q++) tot += pr[q];
            tot = 1.0 / tot;
            for (int q = 0; q < NS; q++) pr[q] *= tot;
            double run = 0;
            int pick = NS - 1;
            for (int q = 0; q < NS; q++) { run += pr[q]; if (r < run) { pick = q; break; } }
            cur = pick;
        }
        std::reverse(path.begin(), path.end());
        out.push_back(path_to_bases(path));
    }
    return out;
}

std::string dot(const std::string& s) { return s.empty() ? std::string(".") : s; }

std::string text_of(const std::vector<Mut>& v, bool scored)
{
    std::string out;
    char buf[64];
    for (size_t i = 0; i < v.size(); i++)
    {
        snprintf(buf, sizeof buf, "%d", v[i].start);
        out += buf; out += '\t'; out += dot(v[i].orig); out += '\t'; out += dot(v[i].mut); out += '\t';
        snprintf(buf, sizeof buf, "%.17g", scored ? v[i].score : 0.0);
        out += buf; out += '\n';
    }
    return out;
}

int put(const std::string& s, char* out, int cap)
{
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return 0;
}

std::vector<Mut> to_muts(int n, const int* start, const char* const* orig, const char* const* mut)
{
    std::vector<Mut> v(n);
    for (int i = 0; i < n; i++) { v[i].start = start[i]; v[i].orig = orig[i]; v[i].mut = mut[i]; v[i].score = -1e-6; }
    return v;
}

} // namespace

extern "C" {

const char* orc_name(void) { return "restatement"; }

int orc_score_alignments(orc_region* r, double* scores, double* likes)
{
    Region R(r);
    std::vector<double> s = score_alignments(R, likes);
    for (size_t i = 0; i < s.size(); i++) scores[i] = s[i];
    R.store(r);
    return 0;
}

int orc_score_mutations(orc_region* r, int n, const int* start, const char* const* orig,
                        const char* const* mut, double* scores)
{
    Region R(r);
    std::vector<Mut> m = to_muts(n, start, orig, mut);
    score_mutations(R, m);
    for (int i = 0; i < n; i++) scores[i] = m[i].score;
    R.store(r);
    return 0;
}

int orc_score_points(orc_region* r, char* out, int cap)
{
    Region R(r);
    std::vector<Mut> m = point_mutations(R);
    score_mutations(R, m);
    R.store(r);
    return put(text_of(m, true), out, cap);
}

int orc_make_mutations(orc_region* r, int n, const int* start, const char* const* orig,
                       const char* const* mut, const double* scores, char* seq_out, int cap, int* nbases)
{
    Region R(r);
    std::vector<Mut> m = to_muts(n, start, orig, mut);
    for (int i = 0; i < n; i++) m[i].score = scores[i];
    *nbases = make_mutations(R, m);
    R.store(r);
    return put(R.bases, seq_out, cap);
}

int orc_refine(orc_region* r, char* seq_out, int cap, int* nbases)
{
    Region R(r);
    std::vector<Mut> m = point_mutations(R);
    score_mutations(R, m);
    *nbases = make_mutations(R, m);
    R.store(r);
    return put(R.bases, seq_out, cap);
}

/* FindMutations with a fresh profile cache (one call = one AlignData, _poreseqcpp.pyx:408) */
int orc_find_mutations(orc_region* r, int n_seeds, const char* const* seeds, char* out, int cap)
{
    Region R(r);
    ProfileCache cache;
    std::vector<std::string> sd(seeds, seeds + n_seeds);
    std::vector<Mut> m = find_mutations(R, sd, cache);
    R.store(r);
    return put(text_of(m, false), out, cap);
}

/* PSAlign.Mutate's loop (_poreseqcpp.pyx:424-431): the profile cache lives across the repetitions although the
 * sequence changes under it */
int orc_mutate(orc_region* r, int n_seeds, const char* const* seeds, int reps, char* seq_out, int cap, int* totbases)
{
    Region R(r);
    ProfileCache cache;
    std::vector<std::string> sd(seeds, seeds + n_seeds);
    int total = 0;
    for (int k = 0; k < reps; k++)
    {
        std::vector<Mut> m = find_mutations(R, sd, cache);
        score_mutations(R, m);
        const int nb = make_mutations(R, m);
        if (nb == 0) break;
        total += nb;
    }
    *totbases = total;
    R.store(r);
    return put(R.bases, seq_out, cap);
}

int orc_viterbi_mutate(orc_region* r, int nkeep, double skip_prob, double stay_prob, double mut_min, double mut_max,
                       char* out, int cap)
{
    Region R(r);
    std::vector<std::string> seqs = viterbi_mutate(R, nkeep, skip_prob, stay_prob, mut_min, mut_max);
    std::string s;
    for (size_t i = 0; i < seqs.size(); i++) { s += seqs[i]; s += '\n'; }
    return put(s, out, cap);
}

int orc_swfull(const char* seq1, const char* seq2, int* inds1, int* inds2, int cap, int* n, int* score, double* accuracy)
{
    Pairing p = sw_align(seq1, seq2);
    *n = (int)p.a.size();
    *score = p.score;
    *accuracy = p.accuracy;
    if (*n > cap) return -1;
    for (int k = 0; k < *n; k++) { inds1[k] = p.a[k]; inds2[k] = p.b[k]; }
    return 0;
}

int orc_map_alignments(orc_region* r, const char* newseq)
{
    Region R(r);
    map_alignments(R, newseq);
    R.store(r);
    return 0;
}

int orc_seq_to_states(const char* seq, int len, int* states)
{
    std::vector<int> st = states_of(std::string(seq, len));
    for (size_t i = 0; i < st.size(); i++) states[i] = st[i];
    return (int)st.size();
}

} // extern "C"
