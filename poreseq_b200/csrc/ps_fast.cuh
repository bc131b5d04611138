// ps_fast.cuh -- the FP32 throughput form of the per-(mutation, event) kernel.
//
// The exact FP64 kernels of ps_device.cuh stay the source of truth for everything a decision hangs
// on.  In PS_PRECISION_FAST mode the mutation scan runs in two passes:
//   1. k_mutscore_f32 scores EVERY (mutation, event) pair in log-space FP32.  Path scores grow
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

__device__ __forceinline__ float emission_f(const LevelRecF& l, float ey, const StateParamsF& p)
{
    const float d1 = l.x - p.mu, d2 = l.y - p.mu2;
    return __fmaf_rn(p.a_s, d1 * d1, p.c_s) + __fmaf_rn(p.f_s * l.ry, d2 * d2, ey);
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

// Per-thread strip in shared memory, indexed by (row & mask) so no wrap logic is needed: the
// main-matrix values of the previous column, updated in place (see ps_device.cuh).  The seed
// column is staged into it (rebased, FP32) before the first narrow column, so the column routine
// has one form for every column but the last, whose rows are joined with the reverse column on the
// fly.  Interior rows carry no band predicate; the level record (and the reverse cells) of row i+1
// are requested while row i is computed.
struct ColF
{
    float* ringC;                // + (row & mask) * 128
    const LevelRecF* lev;        // event levels
    const double* Bm; const double* Bs;   // reverse column (last narrow column only): + row_off(ts, jb)
    long long ts;
    double dRa;
    int n0, mask;
    int p0, p1;                  // previous column's band
    int b0, b1;                  // reverse column's band (rows jb, inclusive); empty when there is none
    float fl;
    float4 tr;
    float best, joinmax;
};

template <bool EDGE, bool LAST>
__device__ __forceinline__ void row_f(ColF& q, const StateParamsF& sp, int i, int i0, int i1,
                                      const LevelRecF*& lv, const LevelRecF*& lq, LevelRecF& lr, float& ey,
                                      float& bm, float& bs, float& diag, float& upC, float& upS)
{
    const LevelRecF lr_c = lr;
    const float ey_c = ey, bm_c = bm, bs_c = bs;
    if (!EDGE || i < i1)
    {
        // next row: level i (mean/stdv) and level n0-i-1 (its -1.5 log stdv, quirk A.3-1), reverse row jb-1
        lv++; lq--;
        lr = *lv; ey = lq->ey;
        if (LAST)
        {
            const int jn = q.n0 - i;
            if (!EDGE || (jn >= q.b0 && jn <= q.b1))
            {
                const long long ro = row_off(q.ts, jn);
                bm = (float)(q.Bm[ro] - q.dRa); bs = (float)(q.Bs[ro] - q.dRa);
            }
        }
    }
    const float e = emission_f(lr_c, ey_c, sp);
    const bool skip_ok = EDGE ? (i >= q.p0 && i <= q.p1) : true;
    const bool diag_ok = EDGE ? (i > q.p0 && i <= q.p1) : true;
    const bool first = EDGE ? (i == i0) : false;
    float* slot = q.ringC + (i & q.mask) * 128;
    float Pi = q.fl;
    if (skip_ok) Pi = *slot;
    float C, Sv;
    dp_cell_f(first, skip_ok, diag_ok, Pi, diag, e, upC, upS, q.fl, q.tr, C, Sv);
    q.best = fmaxf(q.best, C);
    if (LAST)
    {
        const int jb = q.n0 - i + 1;
        if (!EDGE || (jb >= q.b0 && jb <= q.b1)) q.joinmax = fmaxf(q.joinmax, fmaxf(C + bm_c, Sv + bs_c));
    }
    else *slot = C;
    diag = Pi;
    upC = C; upS = Sv;
}

template <bool LAST>
__device__ __forceinline__ void column_f(ColF& q, const StateParamsF& sp, int i0, int i1)
{
    float diag = q.fl;
    if (i0 > q.p0 && i0 <= q.p1) diag = q.ringC[((i0 - 1) & q.mask) * 128];
    float upC = q.fl, upS = q.fl;
    // interior: rows i-1 and i of the previous column exist and (last column) so does the reverse cell,
    // also for the row after (its values are requested one row ahead)
    int lo = max(i0 + 1, q.p0 + 1), hi = min(i1 - 1, q.p1);
    if (LAST) { lo = max(lo, q.n0 + 1 - q.b1); hi = min(hi, q.n0 - q.b0); }
    if (lo > hi) { lo = i1 + 1; hi = i1; }
    const LevelRecF* lv = q.lev + (i0 - 1);
    const LevelRecF* lq = q.lev + (q.n0 - i0);
    LevelRecF lr = *lv;
    float ey = lq->ey;
    float bm = 0.f, bs = 0.f;
    if (LAST)
    {
        const int jb = q.n0 - i0 + 1;
        if (jb >= q.b0 && jb <= q.b1) { const long long ro = row_off(q.ts, jb); bm = (float)(q.Bm[ro] - q.dRa); bs = (float)(q.Bs[ro] - q.dRa); }
    }
    int i = i0;
    for (; i < lo; i++) row_f<true, LAST>(q, sp, i, i0, i1, lv, lq, lr, ey, bm, bs, diag, upC, upS);
    for (; i <= hi; i++) row_f<false, LAST>(q, sp, i, i0, i1, lv, lq, lr, ey, bm, bs, diag, upC, upS);
    for (; i <= i1; i++) row_f<true, LAST>(q, sp, i, i0, i1, lv, lq, lr, ey, bm, bs, diag, upC, upS);
}

__global__ void __launch_bounds__(128) k_mutscore_f32(Batch b, int mask)
{
    extern __shared__ float ringf[];
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
        if (ev.usable && !((unsigned)mu.start > (unsigned)ev.L))
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
                const StateParamsF* stf = b.stf + (size_t)ev.model * N_STATES;
                const bool ri_empty = b.ri_empty[e] != 0;
                const long long ts = ev.rs;
                ColF q;
                q.ringC = ringf + threadIdx.x;
                q.lev = b.levf + ev.lev_off;
                q.n0 = n0; q.mask = mask; q.ts = ts;
                q.tr = b.trf[ev.model];
                q.p0 = 0; q.p1 = n0;
                double best_d = 0.0, a = 0.0;
                // band of the first narrow column: the seed rows it can touch are [i0-1, i1]
                int f0, f1;
                band_of(ri_empty ? 1 : b.cen_new[ev.cen_off + startind + 1], n0, W, f0, f1);
                if (startind > 0)
                {
                    const long long gs = ev.col_off + startind;
                    q.p0 = b.Fi0[gs]; q.p1 = q.p0 + b.Flen[gs] - 1;
                    const double* seed = b.Fm + col_base(ev, startind);
                    best_d = b.Fbest[gs];
                    // rebase to the seed value on the band centre of the first narrow column
                    const int mid = min(max((f0 + f1) >> 1, q.p0), q.p1);
                    a = seed[row_off(ts, mid)];
                    // stage the seed rows the first column reads, rebased, into the ring
                    const int s0 = max(f0 - 1, q.p0), s1 = min(f1, q.p1);
                    for (int i = s0; i <= s1; i++) q.ringC[(i & mask) * 128] = (float)(seed[row_off(ts, i)] - a);
                }
                q.fl = (float)(-a);
                q.best = (float)(best_d - a);
                if (startind == 0)
                {
                    // blank column 0: rows 0..n0, all zeros
                    for (int i = max(f0 - 1, 0); i <= f1; i++) q.ringC[(i & mask) * 128] = q.fl;
                }
                // reverse column the last narrow column is joined with; values relative to R
                const int rab = min(max(Nm - last + 1, 0), N);
                const double dRa = R - a;
                double mb = 0.0;
                q.dRa = dRa;
                q.joinmax = NEGF;
                q.b0 = 1; q.b1 = 0;
                q.Bm = b.Bm; q.Bs = b.Bs;
                if (rab > 0)
                {
                    const long long gb = ev.col_off + rab;
                    q.b0 = b.Bi0[gb]; q.b1 = q.b0 + b.Blen[gb] - 1;
                    mb = b.Bbest[gb];
                    const long long bbase = col_base(ev, rab);
                    q.Bm = b.Bm + bbase; q.Bs = b.Bs + bbase;
                }
                int i0 = f0, i1 = f1;
                for (int c = startind + 1; c <= last; c++)
                {
                    if (c > startind + 1) band_of(ri_empty ? 1 : b.cen_new[ev.cen_off + c], n0, W, i0, i1);
                    const int s = mut_state(mv, c - 1);
                    if (s >= 0)
                    {
                        const StateParamsF sp = stf[s];
                        if (c == last && rab > 0) column_f<true>(q, sp, i0, i1); else column_f<false>(q, sp, i0, i1);
                    }
                    else
                    {
                        // invalid state (cpp/Alignment.cpp:162-163): all-zero column inheriting the running best
                        for (int i = i0; i <= i1; i++) q.ringC[(i & mask) * 128] = q.fl;
                        if (c == last && rab > 0)
                            for (int i = max(i0, n0 + 1 - q.b1); i <= min(i1, n0 + 1 - q.b0); i++)
                            {
                                const long long jb = n0 - i + 1;
                                q.joinmax = fmaxf(q.joinmax, q.fl + fmaxf((float)(q.Bm[row_off(ts, (int)jb)] - dRa), (float)(q.Bs[row_off(ts, (int)jb)] - dRa)));
                            }
                    }
                    q.p0 = i0; q.p1 = i1;
                }
                float joinmax = q.joinmax;
                // a blank reverse column (all zeros) adds nothing beyond the running best of the forward cells
                if (rab == 0) joinmax = q.best - (float)dRa;
                // new - old = max(join, best, reverse best, 0) - R, all relative to R
                const float rel = fmaxf(fmaxf(joinmax, q.best - (float)dRa), fmaxf((float)(mb - R), (float)(-R)));
                result = (double)rel;
            }
        }
        b.delta[t] = result;
    }
}

// mutations whose FP32 total is not clearly negative go to the exact pass
__global__ void k_flag(Batch b, long long n_muts)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_muts) return;
    if (b.scores[g] > -b.tau)
    {
        const int q = atomicAdd(b.flag_count, 1);
        b.flag_list[q] = (int)g;
    }
}

} // namespace psdev
