// ps_swhost.cpp -- swfull on the host (the single-pair `swalign` of the Python surface, RealignTo, the accuracy
// line of the consensus driver): full-matrix local alignment, +5 match / -4 mismatch / -8 gap (cpp/swlib.h:21-23,
// cpp/swlib.cpp:211-340).  Tie rules of the reference: the two gap moves must beat the running best strictly, the
// pairing move wins ties; the best cell is the first maximum in (column of seq2, row of seq1) order.
//
// The reference walks the matrix column by column, one cell after the other (each cell waits for the one above it:
// ~11 ns per cell, 1.1 s for 10 kb x 10 kb).  Here the matrix is walked by anti-diagonals d = i + j, whose cells are
// independent of each other -- the inner loop is a plain element-wise loop over three score arrays that the compiler
// vectorises -- and the move bytes (the only thing the traceback reads) are stored diagonal-major, so they are written
// as one contiguous run per diagonal.  Same scores, same moves, same first-maximum cell, hence the same alignment.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ps_internal.h"

namespace
{
// One anti-diagonal: cells (i, d - i) for i = ilo..ihi.  h1 = diagonal d-1, h0 = diagonal d-2 (indexed by i),
// a = seq1 (a[i-1] is row i), br = seq2 reversed and shifted so that br[i] is the base of column d - i.
// Returns the largest score of the diagonal.
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx2", "default")))
#endif
int sw_diagonal(int ilo, int ihi, const int* __restrict h0, const int* __restrict h1, int* __restrict h2,
                const char* __restrict a, const char* __restrict br, uint8_t* __restrict mv)
{
    int dmax = 0;
    for (int i = ilo; i <= ihi; i++)
    {
        const int up = h1[i] - 8;                                    // from (i, j-1): move 1
        const int left = h1[i - 1] - 8;                              // from (i-1, j): move 2
        const int diag = h0[i - 1] + (a[i - 1] == br[i] ? 5 : -4);   // from (i-1, j-1): move 3, wins ties
        int sc = up > 0 ? up : 0;
        int m = up > 0 ? 1 : 0;
        m = left > sc ? 2 : m;
        sc = left > sc ? left : sc;
        m = diag >= sc ? 3 : m;
        sc = diag >= sc ? diag : sc;
        h2[i] = sc;
        mv[i - ilo] = (uint8_t)(m | (sc <= 0 ? 4 : 0));
        dmax = sc > dmax ? sc : dmax;
    }
    return dmax;
}

#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx2", "default")))
#endif
int sw_last_equal(int ilo, int ihi, const int* __restrict h, int value)
{
    int last = 0;
    for (int i = ilo; i <= ihi; i++) last = h[i] == value ? i : last;      // i ascends: the last hit is the largest
    return last;
}
}

SWResult psi_swfull(const std::string& s1, const std::string& s2)
{
    const int n1 = (int)s1.size(), n2 = (int)s2.size();
    SWResult r;
    r.score = 0;
    int best = 0, bi = 0, bj = 0;
    std::vector<size_t> off;                      // first move byte of diagonal d
    std::unique_ptr<uint8_t[]> move;
    auto ilo_of = [n2](int d) { return std::max(1, d - n2); };
    if (n1 > 0 && n2 > 0)
    {
        off.assign((size_t)n1 + n2 + 2, 0);
        for (int d = 2; d <= n1 + n2; d++)
            off[d + 1] = off[d] + (size_t)(std::min(n1, d - 1) - ilo_of(d) + 1);
        move.reset(new uint8_t[off[(size_t)n1 + n2 + 1] + 1]);
        // seq2 reversed: column j = d - i of diagonal d is rev[n2 - d + i]
        std::string rev(s2.rbegin(), s2.rend());
        std::vector<int> ha((size_t)n1 + 2, 0), hb((size_t)n1 + 2, 0), hc((size_t)n1 + 2, 0);
        int *h0 = ha.data(), *h1 = hb.data(), *h2 = hc.data();
        for (int d = 2; d <= n1 + n2; d++)
        {
            const int ilo = ilo_of(d), ihi = std::min(n1, d - 1);
            const int dmax = sw_diagonal(ilo, ihi, h0, h1, h2, s1.data(), rev.data() + (n2 - d), move.get() + off[d]);
            if (d <= n1) h2[d] = 0;               // cell (d, 0) of this diagonal: the blank column
            if (dmax > 0 && dmax >= best)
            {
                // first maximum in (column, row) order: on one diagonal the smallest column is the largest row
                const int i = sw_last_equal(ilo, ihi, h2, dmax), j = d - i;
                if (dmax > best || j < bj) { best = dmax; bi = i; bj = j; }
            }
            int* t = h0; h0 = h1; h1 = h2; h2 = t;
        }
    }
    r.score = best;
    int i = bi, j = bj, nmatch = 0;
    while (i > 0 && j > 0)
    {
        const int d = i + j;
        const uint8_t mv = move[off[d] + (size_t)(i - ilo_of(d))];
        if (mv & 4) break;
        const int m = mv & 3;
        if (m == 1) { r.inds1.push_back(0); r.inds2.push_back(j); j--; }
        else if (m == 2) { r.inds1.push_back(i); r.inds2.push_back(0); i--; }
        else if (m == 3)
        {
            r.inds1.push_back(i); r.inds2.push_back(j);
            if (s1[i - 1] == s2[j - 1]) nmatch++;
            i--; j--;
        }
        else break;      // cannot happen for a positive score
    }
    std::reverse(r.inds1.begin(), r.inds1.end());
    std::reverse(r.inds2.begin(), r.inds2.end());
    r.accuracy = 100.0 * nmatch / (double)r.inds1.size();
    return r;
}
