// ps_fast.cuh -- the FP32 throughput form of the per-(mutation, event) kernel.
//
// The exact FP64 kernels of ps_device.cuh stay the source of truth for everything a decision hangs
// on.  In PS_PRECISION_FAST mode the mutation scan runs in two passes:
//   1. k_mutscore_rows_f32 scores EVERY (mutation, event) pair in log-space FP32.  Path scores grow
//      with the region (~2.3 per level), so each task works on values rebased to its own seed
//      column: forward values carry x - a (a = seed value at the band centre), the reverse join
//      values B - (R - a) with R the task's old score, so the numbers the FP32 units see are O(10)
//      and the per-pair error stays ~1e-5 regardless of region length.  The 0 floor of the local
//      alignment becomes the constant -a.  Emissions use fused per-state coefficients
//      e = a_s (x-mu)^2 + c_s + f_s (y-mu2)^2 / y + e_y  (8 FP32 ops, no division, no log).
//   2. every mutation whose FP32 total is above -tau (tau >> the FP32 error) is re-scored by the
//      exact FP64 kernel over a compacted list, and its score is replaced.
// Accepted mutations (score >= 0), their order and their scores are therefore bit-identical to the
// reference; clearly negative scores carry the FP32 error (~1e-6 relative), far inside the 1e-4
// tolerance of BASELINE.json.
#pragma once
#include "ps_device.cuh"

namespace psdev {

constexpr float NEGF = -3.0e38f;

__device__ __forceinline__ float emission_f(const LevelRecF& l, const StateParamsF& p)
{
    const float d1 = l.x - p.mu, d2 = l.y - p.mu2;
    return __fmaf_rn(p.a_s, d1 * d1, p.c_s) + __fmaf_rn(p.f_s * l.ry, d2 * d2, l.ey);
}

// one cell in rebased FP32: `fl` is the rebased 0 floor (cpp/Alignment.cpp:194-271 with 0 -> fl)
__device__ __forceinline__ void dp_cell_f(bool first_row, bool skip_ok, bool diag_ok, float Pi, float Pi1, float e,
                                          float upC, float upS, float fl, const float4 tr, float& C, float& S)
{
    const float skip = (skip_ok ? Pi : fl) + tr.x;
    const float match = (diag_ok ? Pi1 : fl) + e;
    const float ignore = diag_ok ? Pi1 + tr.w : fl;
    float stay = NEGF, ext = NEGF, ins = fl;
    if (!first_row)
    {
        stay = upC + e + tr.y;
        ins = upC + tr.w;
        ext = upS + e + tr.z;
    }
    S = fmaxf(first_row ? NEGF : fl, fmaxf(stay, ext));
    C = fmaxf(fmaxf(fmaxf(fl, skip), fmaxf(match, ins)), fmaxf(ignore, S));
}

// ------------------------------------------------------------------------------------------
// k_mutscore_rows_f32: the FP32 scan, row-major.  A mutation that replaces at most one base
// re-fills at most NC = 6 narrow columns.  Instead of finishing one column before the next (which
// needs the previous column in a per-thread ring and re-reads the level records once per column),
// the thread sweeps the rows once and keeps one (main, stay) pair per column in registers: cell
// (i, c) reads (i, c-1) computed a moment ago, (i-1, c-1) and (i-1, c) from the registers of the
// previous row.  Per row it loads one 16-byte row record (mean, stdv, 1/stdv of level i-1 and the
// -1.5 log stdv of level n0-i the forward pass pairs with them, quirk A.3-1; built by the host), one seed
// value and (for the last column) one reverse cell; the six emissions of a row are independent.
// Mutations with longer replacement strings are left to the exact pass (k_flag sends them there).
constexpr int NC = 6;

struct ColState
{
    StateParamsF sp;
    int i0, i1;
    bool valid;
};

struct RowSweep               // everything the row loop carries, all in registers
{
    ColState col[NC];
    float C[NC], S[NC];       // row i-1 of every narrow column
    int ncol;
    int p0, p1;               // band of the seed column
    const double* seed;       // seed column (+ row_off), nullptr for the blank column 0
    double a;                 // rebasing offset
    const double* Bm; int b0, b1; double dRa; bool joined;   // reverse column of the last narrow column
    const LevelRecF* lev; int n0; long long rs;
    float fl; float4 tr;
    float best, joinmax;
    // of the row about to be computed: emission of every column, seed value (rebased) at rows i and i-1
    float em[NC];
    float sd, sd_prev;
};

// seed column value of row i as stored (the rebasing offset itself outside its band / for the blank
// column, so that it rebases to the floor)
__device__ __forceinline__ double seed_raw(const RowSweep& q, int i)
{
    if (i < q.p0 || i > q.p1 || !q.seed) return 0.0;
    return q.seed[row_off(q.rs, i)];
}

// rows [ia, ib].  FAST: every one of the first five columns (and the sixth, when there is one) is inside
// its band and past its first row, the previous column covers rows i-1 and i, every state is valid
// and (when joined) the reverse cell exists -- no predicate is left in the row body.  Otherwise the
// general form with all edge cases.
// Loads are issued at the top of a row and consumed at its bottom (the next row's emissions, the
// join), so no loaded value is carried around the loop.
// MASKED (edge rows, all states valid): the same predicate-free arithmetic for every cell, made right by
// sentinels instead of branches -- a cell outside its band leaves NEGF in the column registers, which
// (a) makes stay / extend / insert lose on the column's first row and puts the stay matrix's floor at
// NEGF there (min(upC, fl)), and (b) reads as the floor fl through max(., fl) when the next column
// looks left or diagonally; a column's last row is replaced by NEGF after use so that the diagonal
// move out of it is implicit (cpp/Alignment.cpp:213-225 tests i <= p1, not i-1 <= p1).
enum { ROWS_GENERAL = 0, ROWS_FAST = 1, ROWS_MASKED = 2 };

template <int MODE>
__device__ __forceinline__ void sweep_rows(RowSweep& q, int ia, int ib, int rhi)
{
    constexpr bool FAST = MODE == ROWS_FAST;
    // FAST rows lie strictly inside every band (and before the last row), so their loads need no range test
    // and walk with running pointers: the row record of row i+1, the seed cell of row i+1 (a row pair shares
    // a 2x2 tile: +2 inside the pair, +rs-2 to the next pair) and the reverse cell of row i (rows run backwards)
    const float4* plev = reinterpret_cast<const float4*>(q.lev + ia);
    const double* pseed = q.seed ? q.seed + row_off(q.rs, ia + 1) : nullptr;
    const double* pB = q.Bm + (q.joined ? row_off(q.rs, max(q.n0 - ia + 1, 1)) : 0);
    const long long hop = q.rs - 2;
    for (int i = ia; i <= ib; i++)
    {
        // requests: row record and seed value of row i+1, reverse cell of row i
        LevelRecF lrn; double sdn = 0.0, bmr = 0.0;
        const int jb = q.n0 - i + 1;
        const bool jin = jb >= q.b0 && jb <= q.b1;
        if (FAST)
        {
            const float4 v = *plev++;
            lrn.x = v.x; lrn.y = v.y; lrn.ry = v.z; lrn.ey = v.w;
            if (pseed) { sdn = *pseed; pseed += (i & 1) ? hop : 2; }            // row i+1 -> i+2: i+1 even ends a pair
            if (q.joined) { bmr = *pB; pB -= (jb & 1) ? hop : 2; }              // row jb -> jb-1: jb odd starts a pair
        }
        else
        {
            lrn.x = lrn.y = lrn.ry = lrn.ey = 0.f;
            if (i < rhi)
            {
                lrn = q.lev[i];
                sdn = seed_raw(q, i + 1);
            }
            if (q.joined && jin) bmr = q.Bm[row_off(q.rs, jb)];
        }
        float left = q.sd, diag = q.sd_prev;                 // (i, c-1) and (i-1, c-1)
        if (FAST)
        {
#pragma unroll
            for (int c = 0; c < 5; c++)
            {
                const float upC = q.C[c], upS = q.S[c];
                float Cn, Sn;
                dp_cell_f(false, true, true, left, diag, q.em[c], upC, upS, q.fl, q.tr, Cn, Sn);
                q.best = fmaxf(q.best, Cn);
                q.C[c] = Cn; q.S[c] = Sn;
                diag = upC; left = Cn;
            }
            if (q.ncol == 6)
            {
                // sixth column (every edit but a deletion), predicated instead of a second code path
                const float upC = q.C[5], upS = q.S[5];
                float Cn, Sn;
                dp_cell_f(false, true, true, left, diag, q.em[5], upC, upS, q.fl, q.tr, Cn, Sn);
                q.best = fmaxf(q.best, Cn);
                q.C[5] = Cn; q.S[5] = Sn;
            }
            if (q.joined)
            {
                const float Cl = q.ncol == 6 ? q.C[5] : q.C[4];
                q.joinmax = fmaxf(q.joinmax, Cl + (float)(bmr - q.dRa));
            }
        }
        else if (MODE == ROWS_MASKED)
        {
            const float bmf = jin ? (float)(bmr - q.dRa) : NEGF;
            float jv4 = NEGF, jv5 = NEGF;
#pragma unroll
            for (int c = 0; c < NC; c++)
            {
                const float upC = q.C[c], upS = q.S[c], em = q.em[c];
                const float skip = left + q.tr.x;
                const float match = diag + em;
                const float ignore = diag + q.tr.w;
                const float stay = upC + em + q.tr.y;
                const float ins = upC + q.tr.w;
                const float ext = upS + em + q.tr.z;
                const float Sn = fmaxf(fminf(upC, q.fl), fmaxf(stay, ext));
                const float Cn = fmaxf(fmaxf(fmaxf(q.fl, skip), fmaxf(match, ins)), fmaxf(ignore, Sn));
                const bool act = i >= q.col[c].i0 && i <= q.col[c].i1;
                const float Cm = act ? Cn : NEGF, Sm = act ? Sn : NEGF;
                q.best = fmaxf(q.best, Cm);
                if (c == 4) jv4 = Cm + bmf;
                if (c == 5) jv5 = Cm + bmf;
                left = fmaxf(Cm, q.fl);
                diag = fmaxf(upC, q.fl);                     // (i-1, c) as the next column's diagonal
                q.C[c] = i == q.col[c].i1 ? NEGF : Cm;
                q.S[c] = Sm;
            }
            if (q.joined) q.joinmax = fmaxf(q.joinmax, q.ncol == 6 ? jv5 : jv4);
        }
        else
        {
            int q0 = q.p0, q1 = q.p1;                        // band of column c-1
#pragma unroll
            for (int c = 0; c < NC; c++)
            {
                if (c < q.ncol)
                {
                    const bool act = i >= q.col[c].i0 && i <= q.col[c].i1;
                    const float upC = q.C[c], upS = q.S[c];
                    if (act)
                    {
                        float Cn = q.fl, Sn = q.fl;
                        if (q.col[c].valid)
                        {
                            const bool skip_ok = i >= q0 && i <= q1;
                            const bool diag_ok = i > q0 && i <= q1;
                            dp_cell_f(i == q.col[c].i0, skip_ok, diag_ok, left, diag, q.em[c], upC, upS, q.fl, q.tr, Cn, Sn);
                            q.best = fmaxf(q.best, Cn);
                        }
                        if (c == q.ncol - 1 && q.joined && jin)
                            q.joinmax = fmaxf(q.joinmax, Cn + (float)(bmr - q.dRa));
                        q.C[c] = Cn; q.S[c] = Sn;
                    }
                    diag = upC;                              // (i-1, c) is the next column's diagonal
                    left = q.C[c];
                    q0 = q.col[c].i0; q1 = q.col[c].i1;
                }
            }
        }
        // bottom of the row: what row i+1 needs
#pragma unroll
        for (int c = 0; c < NC; c++) q.em[c] = emission_f(lrn, q.col[c].sp);
        q.sd_prev = (MODE == ROWS_MASKED && i == q.p1) ? q.fl : q.sd;   // the diagonal out of the seed column's last row is implicit
                                                                          // (fast rows end before any column's last row)
        q.sd = (float)(sdn - q.a);
    }
}

__global__ void __launch_bounds__(128) k_mutscore_rows_f32(Batch b)
{
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int W = b.scoring_width;
    for (long long t = gtid; t < b.n_tasks; t += nthreads)
    {
        int e, m;
        if (!task_decode(b, t, e, m)) continue;
        const EvDesc ev = b.ev[e];
        const MutDev mu = b.muts[ev.mut_off + m];
        double result = 0.0;
        if (ev.usable && !((unsigned)mu.start > (unsigned)ev.L) && mu.n_mut <= 1)
        {
            const int N = ev.N, n0 = ev.n0, L = ev.L;
            MutView mv;
            mv.bases = b.bases + ev.base_off; mv.L = L; mv.mstr = b.mut_str + mu.str_off;
            mv.start = mu.start; mv.n_orig = mu.n_orig; mv.n_mut = mu.n_mut;
            mv.applied = mu.start < L;
            mv.Lm = mv.applied ? mu.start + mu.n_mut + max(0, L - mu.start - mu.n_orig) : L;
            const int Nm = mv.Lm >= 5 ? mv.Lm - 4 : 0;
            const int raf = max(mu.start - 3, 1);
            const double R = raf <= N ? b.old[ev.col_off + raf] : thread_join(b, ev, raf, N - raf + 1);
            const int startind = max(mu.start - 4, 0);
            const int refind = mu.start + mu.n_mut + 1;
            int last = min(min(refind, startind + mu.n_mut + 6), Nm);
            if (W == 0) last = startind;
            if (last <= startind)
                result = thread_join(b, ev, startind, Nm - startind + 1) - R;     // boundary case: exact
            else
            {
                RowSweep q;
                q.ncol = last - startind;                                         // 1..NC
                const StateParamsF* stf = b.stf + (size_t)ev.model * N_STATES;
                const bool ri_empty = b.ri_empty[e] != 0;
                q.rs = ev.rs; q.n0 = n0;
                q.tr = b.trf[ev.model];
                bool all_valid = true;
                // 5-mer states of the narrow columns: one rolling window over the mutated bases when the region
                // and the replacement base are plain ACGT (cpp/Sequence.h:79-98 without its reset cases)
                const bool plain = !ev.inv && (mu.n_mut == 0 || base_code(mv.mstr[0]) < 4);
                int win = 0;
                if (plain)
                {
#pragma unroll
                    for (int t = 0; t < 4; t++) win = (win << 2) | mut_base(mv, startind + t);
                }
#pragma unroll
                for (int c = 0; c < NC; c++)
                {
                    q.col[c].i0 = 1; q.col[c].i1 = 0; q.col[c].valid = false;
                    q.col[c].sp = stf[0];
                    if (c < q.ncol)
                    {
                        band_of(ri_empty ? 1 : b.cen_new[ev.cen_off + startind + 1 + c], n0, W, q.col[c].i0, q.col[c].i1);
                        int s;
                        if (plain) { win = ((win << 2) | mut_base(mv, startind + c + 4)) & (N_STATES - 1); s = win; }
                        else s = mut_state(mv, startind + c);
                        q.col[c].valid = s >= 0;
                        all_valid = all_valid && s >= 0;
                        q.col[c].sp = stf[max(s, 0)];
                    }
                }
                // seed column (forward column startind, or the blank column 0) and the rebasing offset
                q.p0 = 0; q.p1 = n0; q.seed = nullptr; q.a = 0.0;
                double best_d = 0.0;
                if (startind > 0)
                {
                    const long long gs = ev.col_off + startind;
                    q.p0 = b.Fi0[gs]; q.p1 = q.p0 + b.Flen[gs] - 1;
                    q.seed = b.Fm + col_base(ev, startind);
                    best_d = b.Fbest[gs];
                    const int mid = min(max((q.col[0].i0 + q.col[0].i1) >> 1, q.p0), q.p1);
                    q.a = q.seed[row_off(q.rs, mid)];
                }
                q.fl = (float)(-q.a);
                q.best = (float)(best_d - q.a);
                // reverse column the last narrow column is joined with; values relative to R
                const int rab = min(max(Nm - last + 1, 0), N);
                q.dRa = R - q.a;
                double mb = 0.0;
                q.b0 = 1; q.b1 = 0; q.joined = rab > 0;
                q.Bm = b.Bm;
                if (rab > 0)
                {
                    const long long gb = ev.col_off + rab;
                    q.b0 = b.Bi0[gb]; q.b1 = q.b0 + b.Blen[gb] - 1;
                    mb = b.Bbest[gb];
                    const long long bbase = col_base(ev, rab);
                    q.Bm = b.Bm + bbase;
                }
                q.joinmax = NEGF;
                q.lev = b.levf + ev.lev_off;
                // rows swept: from the first row of the first column to the last row of the last one; the
                // predicate-free rows are those every column treats as interior (see sweep_rows)
                int rlo = q.col[0].i0, rhi = q.col[0].i1;
                int flo = max(q.col[0].i0, q.p0) + 1, fhi = min(q.col[0].i1, q.p1) - 1;   // (last rows: sentinel handling)
#pragma unroll
                for (int c = 1; c < NC; c++)
                    if (c < q.ncol)
                    {
                        rlo = min(rlo, q.col[c].i0); rhi = max(rhi, q.col[c].i1);
                        flo = max(flo, max(q.col[c].i0, q.col[c - 1].i0) + 1);
                        fhi = min(fhi, min(q.col[c].i1, q.col[c - 1].i1) - 1);
                    }
                if (q.joined) { flo = max(flo, n0 + 1 - q.b1); fhi = min(fhi, n0 + 1 - q.b0); }
                if (flo > fhi) { flo = rhi + 1; fhi = rhi; }
#pragma unroll
                for (int c = 0; c < NC; c++) { q.C[c] = q.fl; q.S[c] = q.fl; }
                // (rlo - 1, seed) is the first row's diagonal source; the move exists only while rlo itself is inside the
                // seed column's band (cpp/Alignment.cpp:213: i <= p1) -- a seed band that ends right above the first
                // narrow row leaves the floor (found by scripts/gpu_sweep.py; the general rows test the predicate)
                q.sd_prev = rlo <= q.p1 ? (float)(seed_raw(q, rlo - 1) - q.a) : q.fl;
                q.sd = (float)(seed_raw(q, rlo) - q.a);
                {
                    const LevelRecF lr = q.lev[rlo - 1];
#pragma unroll
                    for (int c = 0; c < NC; c++) q.em[c] = emission_f(lr, q.col[c].sp);
                }
#ifdef PS_SCAN_GENERAL_ONLY
                if (false)
#else
                if (all_valid && q.ncol >= 5)
#endif
                {
                    // edge rows with sentinels, interior rows with nothing: the column registers start as "outside"
#pragma unroll
                    for (int c = 0; c < NC; c++) { q.C[c] = NEGF; q.S[c] = NEGF; }
                    sweep_rows<ROWS_MASKED>(q, rlo, flo - 1, rhi);
                    sweep_rows<ROWS_FAST>(q, flo, fhi, rhi);
                    sweep_rows<ROWS_MASKED>(q, fhi + 1, rhi, rhi);
                }
                else
                    sweep_rows<ROWS_GENERAL>(q, rlo, rhi, rhi);
                float joinmax = q.joinmax;
                // a blank reverse column (all zeros) adds nothing beyond the running best of the forward cells
                if (rab == 0) joinmax = q.best - (float)q.dRa;
                // new - old = max(join, best, reverse best, 0) - R, all relative to R
                const float rel = fmaxf(fmaxf(joinmax, q.best - (float)q.dRa), fmaxf((float)(mb - R), (float)(-R)));
                result = (double)rel;
            }
        }
        b.delta[t] = result;
    }
}

// mutations whose FP32 total is not clearly negative go to the exact pass
// (and so do the mutations the FP32 scan does not handle: replacement strings longer than max_mut bases)
__global__ void k_flag(Batch b, long long n_muts, int max_mut)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_muts) return;
    // tau is per event of the mutation's region: an FP32 total over E events carries at most E times the per-pair error
    int lo = 0, hi = b.n_regs - 1;
    while (lo < hi)
    {
        const int mid = (lo + hi + 1) >> 1;
        if (b.regs[mid].mut_off <= g) lo = mid; else hi = mid - 1;
    }
    const int nev = b.tau_events > 0 ? b.tau_events : b.regs[lo].nev;
    if (b.scores[g] > -b.tau * (double)nev || b.muts[g].n_mut > max_mut)
    {
        const int q = atomicAdd(b.flag_count, 1);
        b.flag_list[q] = (int)g;
    }
}

} // namespace psdev
