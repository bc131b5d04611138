// ps_device.cuh -- device-side data layout and the exact (FP64) kernels of the PoreSeq scoring
// path, written for sm_100a.  Compiled with -fmad=false: the recurrence has no transcendentals
// (every log() is hoisted to the host, cpp/EventData.h:57-62,220), so IEEE double add/mul/div
// evaluated in the reference's order reproduces the reference's x86-64 results bit for bit.
//
// Kernels (SURVEY.md 2.1 numbering):
//   k_centres      band-centre table  imid[e][c] = lower_bound(ref_index, c)   (cpp/EventData.h:172-183)
//   k_fill         K1/K2  wide-band forward + reverse fill, anti-diagonal wavefront,
//                  one CTA per (event, direction), one thread per column   (cpp/Alignment.cpp:111-444)
//   k_backtrace    K3     best-path pointer chase + updaterefs               (cpp/Alignment.cpp:516-624,
//                                                                             cpp/EventData.h:110-169)
//   k_join         K4a    old[e][c] = columnMax(c)                           (cpp/Alignment.h:169-214)
//   k_mutscore     K4     one thread per (mutation, event): narrow re-fill + join (cpp/Alignment.cpp:447-512),
//                         previous column kept in an in-place shared-memory ring
//   k_reduce       K5     score[m] = -1e-6 + sum over events in event order  (cpp/MakeMutations.cpp:38-52)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ps_types.cuh"

namespace psdev {

struct EvDesc                 // one event of one region in the batch
{
    int       region;
    int       n0;             // levels
    int       N;              // states of the region sequence
    int       L;              // bases of the region sequence
    int       usable;         // cpp/Alignment.cpp:51-59 (ref_index non-empty at call start)
    int       model;          // index into the model table
    int       n_muts;         // mutations of this event's region
    int       inv;            // the region sequence has a non-ACGT base (some state is -1)
    long long lev_off;        // into the per-level arrays
    long long col_off;        // band column g = col_off + c,  c = 1..N
    long long state_off;      // into states[]
    long long base_off;       // into bases[]
    long long cen_off;        // into centre tables: index cen_off + c, c = 0..N+cen_pad
    long long mut_off;        // first mutation of the region in the mutation arrays
    long long task_off;       // first (event, mutation) task of this event; delta[task_off + m]
    long long band_off;       // first element of this event's band storage (same for F and B arrays)
    long long strip_off;      // first StripRec of this event: forward strips 0..J (J = sentinel), then reverse
    int       ts;             // strips that can be live on one wavefront step (band storage slot count)
    int       rs;             // row stride of the band storage = ts * CW
    int       lazy;           // P - 1: every thread idles > P steps between two strips, strips are switched on every P-th step (P = 1, 2, 4, 8)
};

struct MutDev
{
    int start, n_orig, n_mut, str_off;   // str_off: offset of the mut string in the char pool
};

struct Batch                  // everything the kernels need, passed by value
{
    const EvDesc*     ev;
    int               n_events;
    const ModelDev*   models;
    const int*        states;
    const char*       bases;
    // per level
    LevelRec*         lev;
    const LevIn*       lev_in;                  // staged by the host; k_rows builds lev / rowF / rowB / levf from it
    double*           ref_align;
    double*           ref_like;
    double*           ref_index;
    int*              bt_src;        // scratch for the backtrace gather
    // per event
    int*              ri_empty;      // ref_index empty after the last updaterefs
    int*              refstart;
    int*              refend;
    int*              mono;          // band centres nondecreasing (wavefront schedule is valid)
    const int*        fill_list;     // event indices grouped by wavefront width class (k_fill grid.x indexes it)
    struct StripRec*  strips;        // per (event, direction, strip): everything the fill needs to take the strip up
    LevelRec*         rowF;          // per level: forward row record (row i = level i-1, with 3 log stdv of level n0-i)
    LevelRec*         rowB;          // per level: reverse row record (row i = level n0-i)
    // centre tables (forward lower_bound index, 0..n0), old = before backtrace, new = after
    int*              cen_old;
    int*              cen_new;
    int               cen_pad;
    long long         warp_limit;               // exact pass: up to this many tasks go to k_mutscore_warp (0: never)
    // band storage, wavefront-major: the cells of one anti-diagonal d = k+i of an event are
    // contiguous (slot k % ts), so the fill's per-step stores and the join's loads coalesce
    int               RS;            // 2*realign_width+1 rounded up (serial-fallback smem strips)
    double*           Fm; double* Fs; double* Bm;    // no reverse stay matrix: main >= stay, joins read B's main only
    uint8_t*          Fstep;
    int*              Fi0; int* Flen; int* Bi0; int* Blen;
    double*           Fcb; int* Fcbi;          // per-column best (score, row) before the running max
    double*           Bcb; int* Bcbi;
    double*           Fbest; int* Fbi; int* Fbj; // running best up to and including the column
    double*           Bbest;
    double*           old;                      // columnMax(c) per band column
    // mutations
    const MutDev*     muts;
    const char*       mut_str;
    double*           delta;                    // per task
    double*           scores;                   // per mutation
    long long         n_tasks;
    // narrow-fill scratch
    double*           scratch;
    long long         scratch_slots;
    // parameters
    double            lik_offset;
    double            log2pi;
    int               realign_width;
    int               scoring_width;
    // FP32 fast pass + exact re-score of the candidates that matter (ps_fast.cuh)
    const StateParamsF* stf;                    // [models][1024]
    LevelRecF*        levf;                     // per level
    const float4*     trf;                      // per model: log skip, stay, extend, insert
    const RegTabDev*  regs;                     // per region
    int               n_regs;
    int               max_ev;                   // most events in one region
    int*              flag_list;                // global indices of the mutations to re-score exactly
    int*              flag_count;
    double            tau;                      // re-score when the FP32 total is above -tau * (events of the region)
    int               tau_events;               // > 0: the region's events over ALL ranks of an event-sharded job
};

// ------------------------------------------------------------------------------------------
// Division by a divisor whose correctly rounded reciprocal r = RN(1/b) is already known
// (per state / per level, computed once with an IEEE division): q0 = a*r followed by two
// Markstein corrections q <- fma(fma(-q, b, a), r, q).  After the first correction q is within an
// ulp of a/b, so the second one returns RN(a/b) -- the same bits the reference's `a / b` gives --
// for 5 FP64 instructions instead of the ~25 of a generic double division.
__device__ __forceinline__ double div_by(double a, double b, double r)
{
    double q = a * r;
    q = __fma_rn(__fma_rn(-q, b, a), r, q);
    q = __fma_rn(__fma_rn(-q, b, a), r, q);
    return q;
}

// c = x, code = k when x > c -- as setp + selp.  Written in C++ with c == 0 on entry, NVVM turns the first compare of a
// chain into max.f64, which sm_100 (no FP64 min/max instruction) expands into an 11-instruction NaN-aware sequence;
// the explicit form is DSETP + 2 FSEL + SEL, and it is the reference's own compare-and-assign (cpp/Alignment.cpp:243-263).
__device__ __forceinline__ void take_gt(double x, int k, double& c, int& code)
{
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %2, %0;\n\tselp.f64 %0, %2, %0, p;\n\tselp.s32 %1, %3, %1, p;\n\t}"
        : "+d"(c), "+r"(code) : "d"(x), "r"(k));
}

// x > c ? x : c  (fmax() without the NaN handling: the same three instructions)
__device__ __forceinline__ double max_gt(double x, double c)
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(x), "d"(c));
    return r;
}

// emission: lognormpdf + logigpdf + lik_offset   (cpp/AlignUtil.h:34-53, cpp/Alignment.cpp:169-173)
// x = level mean, y = level stdv (ry its reciprocal), lsd3 = 3*log(stdv) of the level the
// reference indexes (quirk A.3-1: the forward pass reads log_stdv[n0-i] beside stdv[i-1])
__device__ __forceinline__ double emission(double x, double y, double ry, double lsd3, const StateParams& p,
                                           double log2pi, double offset)
{
#ifdef PS_ABL_NOEMIS
    return (x - p.lev_mean) + (y - p.sd_mean) + lsd3 + offset;          // timing ablation: 4 ops instead of 29
#endif
    double d = div_by(x - p.lev_mean, p.lev_stdv, p.r_lev_stdv);
    double l = -0.5 * (d * d + log2pi) - p.log_lev;
    double g = div_by(y - p.sd_mean, p.sd_mean, p.r_sd_mean);
    l += 0.5 * (p.log_lambda - lsd3 - log2pi - div_by(g * g * p.sd_lambda, y, ry));
    l += offset;
    return l;
}

// emission of the cell (row i) of a forward (level i-1) or reverse (level n0-i) column
template <bool REV>
__device__ __forceinline__ double cell_emission(const LevelRec* lev, int n0, int i, const StateParams& p,
                                                double log2pi, double offset)
{
    const LevelRec a = lev[REV ? n0 - i : i - 1];
    const double lsd3 = REV ? a.lsd3 : lev[n0 - i].lsd3;
    return emission(a.mean, a.stdv, a.rstdv, lsd3, p, log2pi, offset);
}

struct Trans { double lskip, lstay, lext, lins; };

// Band storage, wavefront-major.  The fill gives every thread a strip of CW = 2 consecutive columns
// (strip j = columns 2j+1, 2j+2) and computes one row pair r (rows 2r+1, 2r+2) of its strip per step
// d = j + r.  The 2x2 tile of (strip j, row pair r) is stored where its step puts it: 4 contiguous
// cells [row 2r+1: col 0, col 1; row 2r+2: col 0, col 1], tiles computed in the same step are
// neighbours (slot j % ts).  So every step of the fill writes one contiguous run per warp, and a
// reader that walks down a column alternates between +2 and the row-pair stride rs = 4 ts:
//   cell (k, i)  ->  col_base(k) + row_off(rs, i).
constexpr int CW = 2;

__device__ __forceinline__ long long col_base(const EvDesc& ev, int k)
{
    const int j = (k - 1) >> 1;
    return ev.band_off + ((long long)j * ev.ts + (j % ev.ts)) * 4 + ((k - 1) & 1);
}

__device__ __forceinline__ long long row_off(long long rs, int i)      // i >= 1
{
    return (long long)((i - 1) >> 1) * rs + (((i - 1) & 1) << 1);
}

__device__ __forceinline__ long long cell_at(const EvDesc& ev, int k, int i)
{
    return col_base(ev, k) + row_off(ev.rs, i);
}

// One cell of the coupled (main C, stay S) recurrence, cpp/Alignment.cpp:194-271 (forward) and
// :370-441 (reverse).  eM = emission added on the diagonal move (forward: this cell's; reverse:
// the source cell's, 0 when implicit), eU = emission added on stay/extend.
__device__ __forceinline__ void dp_cell(bool first_row, bool skip_ok, bool diag_ok, double Pi, double Pi1,
                                        double eM, double eU, double upC, double upS, const Trans& t,
                                        double& C, double& S, int& step)
{
    double skip = skip_ok ? Pi + t.lskip : t.lskip;
    double match = diag_ok ? Pi1 + eM : eM;
    double ignore = diag_ok ? Pi1 + t.lins : 0.0;
    double stay = NEG, ext = NEG, ins = 0.0;
    if (!first_row)
    {
        stay = upC + eU + t.lstay;
        ins = upC + t.lins;
        ext = upS + eU + t.lext;
    }
    double s = first_row ? NEG : 0.0;
    int ss = 0;
    take_gt(stay, 1, s, ss);
    take_gt(ext, 2, s, ss);
    double c = 0.0;
    int sc = ST_STOP;
    take_gt(skip, skip_ok ? ST_SKIP : ST_IMPLICIT, c, sc);
    take_gt(match, diag_ok ? ST_MATCH : ST_IMPLICIT, c, sc);
    take_gt(ins, ST_INSERT, c, sc);
    take_gt(ignore, ST_IGNORE, c, sc);
    take_gt(s, ST_STAY, c, sc);
    C = c; S = s; step = sc | (ss << 3);
}

// std::lower_bound over ref_index restated as libstdc++'s halving search (cpp/EventData.h:172-183)
__device__ __forceinline__ int lower_bound_index(const double* a, int n, double v)
{
    int first = 0, len = n;
    while (len > 0)
    {
        int half = len >> 1;
        if (a[first + half] < v) { first += half + 1; len -= half + 1; }
        else len = half;
    }
    return first;
}

__device__ __forceinline__ void band_of(int mid, int n0, int w, int& i0, int& i1)
{
    if (mid < 1) mid = 1;
    if (mid > n0) mid = n0;
    i0 = max(1, mid - w);
    i1 = min(n0, mid + w);
}

// ------------------------------------------------------------------------------------------
// k_centres: cen[cen_off + c] = lower_bound(ref_index, c) for c = 0..N+pad (1 when ref_index is
// empty, cpp/Alignment.cpp:129-132); optionally flags non-monotone centres per event.
__global__ void k_centres(Batch b, int* cen, int check_mono)
{
    const EvDesc ev = b.ev[blockIdx.y];
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int cmax = ev.N + b.cen_pad;
    if (c > cmax) return;
    int v = 1;
    if (!b.ri_empty[blockIdx.y])
    {
        const double* ri = b.ref_index + ev.lev_off;
        v = lower_bound_index(ri, ev.n0, (double)c);
        if (check_mono && c >= 1)
        {
            int u = lower_bound_index(ri, ev.n0, (double)(c - 1));
            if (u > v) b.mono[blockIdx.y] = 0;
        }
    }
    cen[ev.cen_off + c] = v;
}

// ------------------------------------------------------------------------------------------
// k_fill: wide-band fill of one (event, direction) per CTA.
//
// Wavefront over (strip, row pair): thread t owns the strips j = t, t+T, ... (2 consecutive columns
// each, in processing order k = c forward, k = N-c+1 reverse) and at step d computes the 2x2 tile of
// row pair r = d - j of its current strip.  Inside the tile and down the strip every dependency is a
// register; only the strip's first column looks outside: rows 2r, 2r+1, 2r+2 of the left
// neighbour's second column were produced at steps d-2 and d-1 and are read from a 4-deep
// shared-memory ring.  One barrier separates steps.
//
// The step body is branch-free so that the four emissions (3 reciprocal divisions each) and the four
// cells' candidate scores interleave: every term of a cell except the horizontal move is folded into
// (M1, code) first, and the only serial dependency between horizontally adjacent cells is
//     C = (skip_ok && Pc + lskip >= M1) ? Pc + lskip : M1
// which returns the value and step code of cpp/Alignment.cpp:252-267 (ties: skip precedes every other
// main-matrix move and the stay matrix only wins when strictly greater) as long as lskip <= 0;
// events with prob_skip > 1 take the serial schedule.
//
// With nondecreasing band centres and T >= the number of strips that can be live on one step
// (planned on the host, Job::plan_event) a thread never has two live strips; events whose centres
// go backwards are filled serially by thread 0.  The 5-mer parameters of a thread's next strip are
// fetched into shared memory with cp.async one step after the current strip was taken up, so the
// switch costs a shared-memory read instead of two dependent global round trips.
struct StripRec               // 192 bytes: the two columns of a strip, ready to be copied into registers
{
    StateParams p[2];
    int i0[2], i1[2], s[2];   // band and state of each column (empty band, s = -1 past the last column)
    int pp0, pp1;             // band of the column just before the strip
    int rlo, rhi;             // row pairs covered by the union of the two bands (1, 0 for the sentinel strip J)
    int lean_lo, lean_hi;     // row pairs whose four cells are interior ones (see fill_wave); empty = (1 << 30, -1)
    int slot4;                // (j % ts) * 4: offset of the strip's tile inside a step's run
    int pad[3];
};
static_assert(sizeof(StripRec) == 192, "StripRec is copied as 16-byte chunks (the first eleven carry data)");

__device__ __forceinline__ void col_band(const Batch& b, const EvDesc& ev, bool rev, int k, int& i0, int& i1)
{
    const int c = rev ? ev.N - k + 1 : k;
    const int cen = b.cen_old[ev.cen_off + c];
    band_of(rev ? ev.n0 - cen + 1 : cen, ev.n0, b.realign_width, i0, i1);
}

// k_strips: one thread per (event, direction, strip) builds the strip's record and publishes the band
// shape of its columns (Fi0/Flen, Bi0/Blen); grid (ceil((Jmax+1)/128), events, directions)
__global__ void k_strips(Batch b)
{
    const EvDesc ev = b.ev[blockIdx.y];
    if (!ev.usable || ev.N <= 0) return;
    const bool rev = blockIdx.z != 0;
    const int J = (ev.N + CW - 1) / CW;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > J) return;
    StripRec r;
    memset(&r, 0, sizeof r);
    r.pp0 = 0; r.pp1 = ev.n0;
    r.rlo = 1; r.rhi = 0;
    r.slot4 = (j % ev.ts) * 4;
    const ModelDev& md = b.models[ev.model];
    int lo = 1 << 30, hi = 0;
    for (int c = 0; c < CW; c++)
    {
        r.s[c] = -1; r.i0[c] = 1; r.i1[c] = 0;
        const int k = CW * j + 1 + c;
        if (j < J && k <= ev.N)
        {
            col_band(b, ev, rev, k, r.i0[c], r.i1[c]);
            r.s[c] = b.states[ev.state_off + (rev ? ev.N - k + 1 : k) - 1];
            lo = min(lo, r.i0[c]); hi = max(hi, r.i1[c]);
            const long long g = ev.col_off + k;
            (rev ? b.Bi0 : b.Fi0)[g] = r.i0[c];
            (rev ? b.Blen : b.Flen)[g] = r.i1[c] - r.i0[c] + 1;
        }
        r.p[c] = md.st[max(r.s[c], 0)];
    }
    r.lean_lo = 1 << 30; r.lean_hi = -1;
    if (j < J)
    {
        r.rlo = (lo - 1) >> 1; r.rhi = (hi - 1) >> 1;
        if (j > 0)
        {
            col_band(b, ev, rev, CW * j, r.pp0, r.pp1);
            // interior row pairs: rows 2r+1 and 2r+2 inside both columns' bands and the band of the column before,
            // 2r+1 not a first row, both states valid:  2r+1 > max(i0a, i0b, pp0)  and  2r+2 <= min(i1a, i1b, pp1)
            if (r.s[0] >= 0 && r.s[1] >= 0)
            {
                r.lean_lo = (max(max(r.i0[0], r.i0[1]), r.pp0) + 1) >> 1;
                r.lean_hi = (min(min(r.i1[0], r.i1[1]), r.pp1) - 2) >> 1;
            }
        }
    }
    b.strips[ev.strip_off + (rev ? J + 1 : 0) + j] = r;
}

// k_rows: expands the staged (mean, stdv, 3 log stdv) of every level into the records the kernels read: the level
// record with RN(1/stdv), the record each fill row reads per direction (quirk A.3-1: the forward pass pairs
// stdv[i-1] with log_stdv[n0-i], cpp/Alignment.cpp:171-172) and the FP32 row record of the mutation scan
__global__ void k_rows(Batch b)
{
    const EvDesc ev = b.ev[blockIdx.y];
    if (!ev.usable) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;       // row 1..n0
    if (i > ev.n0) return;
    const LevIn* in = b.lev_in + ev.lev_off;
    const LevIn a = in[i - 1], q = in[ev.n0 - i];
    LevelRec f;
    f.mean = a.mean; f.stdv = a.stdv; f.rstdv = 1.0 / a.stdv; f.lsd3 = a.lsd3;     // IEEE division: RN(1 / stdv)
    b.lev[ev.lev_off + i - 1] = f;
    LevelRec r;
    r.mean = q.mean; r.stdv = q.stdv; r.rstdv = 1.0 / q.stdv; r.lsd3 = q.lsd3;
    b.rowB[ev.lev_off + i - 1] = r;
    f.lsd3 = q.lsd3;
    b.rowF[ev.lev_off + i - 1] = f;
    if (b.levf)
    {
        // FP32 row record of the scan: -1.5 log(stdv) = -0.5 * (3 log stdv), exact halving
        LevelRecF g;
        g.x = (float)f.mean; g.y = (float)f.stdv; g.ry = (float)f.rstdv; g.ey = (float)(-0.5 * q.lsd3);
        b.levf[ev.lev_off + i - 1] = g;
    }
}

struct FillOut               // where one direction's band columns go
{
    double* Mm; double* Ms; int* Mi0; int* Mlen; double* Mcb; int* Mcbi;
};

// Everything of one cell except the horizontal (skip) move: M1 = max(0, match, insert, ignore, S) with
// the step code the reference's compare order gives, and the stay-matrix value S.
// Rows outside a column's band are computed too (the body is branch-free) and deliver NEG in both
// matrices; nobody reads them except the column's own first row, for which NEG is exactly what makes
// stay / extend / insert lose (cpp/Alignment.cpp:229-237: a first row has no cell above).  INV: the
// column's state may be invalid (non-ACGT base), its in-band cells are all zero (cpp/Alignment.cpp:162).
// SH: the step codes come out shifted to the byte of the cell inside the tile's packed word (main code << SH, stay
// code << SH + 3), so that packing four cells is three ORs.
template <bool INV, int SH = 0>
__device__ __forceinline__ void cell_pre(bool inb, bool valid, bool first, bool diag_ok, double Pd, double eM, double eU,
                                         double upC, double upS, const Trans& t,
                                         double& M1, int& m1, double& S, int& ss)
{
    const double Pe = diag_ok ? Pd : 0.0;
    const double match = Pe + eM;
    const double ignore = Pe + t.lins;                    // lins <= 0: an implicit diagonal never yields a positive ignore
    const double stay = (upC + eU) + t.lstay;             // upC = upS = NEG on a first row
    const double ins = upC + t.lins;
    const double ext = (upS + eU) + t.lext;
    double s = first ? NEG : 0.0;
    int q = 0;
    take_gt(stay, 1 << (SH + 3), s, q);
    take_gt(ext, 2 << (SH + 3), s, q);
    double c = 0.0;
    int sc = ST_STOP << SH;
    take_gt(match, (diag_ok ? ST_MATCH : ST_IMPLICIT) << SH, c, sc);
    take_gt(ins, ST_INSERT << SH, c, sc);
    take_gt(ignore, ST_IGNORE << SH, c, sc);
    take_gt(s, ST_STAY << SH, c, sc);
    if (INV && !valid) { c = 0.0; s = 0.0; sc = ST_STOP << SH; q = 0; }
    M1 = inb ? c : NEG; m1 = inb ? sc : (ST_STOP << SH); S = inb ? s : NEG; ss = inb ? q : 0;
}

__device__ __forceinline__ void cell_fin(bool skip_pred, double Pc, double M1, int m1, const Trans& t, double& C, int& code)
{
    const double skip = Pc + t.lskip;
    const bool w = skip_pred && (skip >= M1);
    C = w ? skip : M1;
    code = (w && skip > 0.0) ? ST_SKIP : m1;              // ST_SKIP == 0 in every byte
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}

// The strip a thread works on, in registers
struct StripRegs
{
    int j;                    // strip index, 1<<29 when past the end
    int rlo, rhi, pp0, pp1;
    int i0a, i1a, sa, i0b, i1b, sb;
    int slot4;                // k_fill2 only
    int lean_lo, lean_hi;
};

template <int MAXT>
__device__ __forceinline__ void strip_from_smem(const double2* nxt, StripRegs& c, StateParams& p0, StateParams& p1)
{
    double2 v[11];
#pragma unroll
    for (int q = 0; q < 11; q++) v[q] = nxt[q * MAXT];
    p0.lev_mean = v[0].x; p0.lev_stdv = v[0].y; p0.log_lev = v[1].x; p0.sd_mean = v[1].y;
    p0.sd_lambda = v[2].x; p0.log_lambda = v[2].y; p0.r_lev_stdv = v[3].x; p0.r_sd_mean = v[3].y;
    p1.lev_mean = v[4].x; p1.lev_stdv = v[4].y; p1.log_lev = v[5].x; p1.sd_mean = v[5].y;
    p1.sd_lambda = v[6].x; p1.log_lambda = v[6].y; p1.r_lev_stdv = v[7].x; p1.r_sd_mean = v[7].y;
    c.i0a = __double2loint(v[8].x); c.i0b = __double2hiint(v[8].x);
    c.i1a = __double2loint(v[8].y); c.i1b = __double2hiint(v[8].y);
    c.sa = __double2loint(v[9].x); c.sb = __double2hiint(v[9].x);
    c.pp0 = __double2loint(v[9].y); c.pp1 = __double2hiint(v[9].y);
    c.rlo = __double2loint(v[10].x); c.rhi = __double2hiint(v[10].x);
    c.lean_lo = __double2loint(v[10].y); c.lean_hi = __double2hiint(v[10].y);
}

template <int MAXT>
__device__ __forceinline__ void strip_request(double2* nxt, const StripRec* rec)
{
    const double2* src = reinterpret_cast<const double2*>(rec);
#pragma unroll
    for (int q = 0; q < 11; q++) cp_async16(&nxt[q * MAXT], src + q);
}

struct RowRecs { LevelRec a, b; };                        // row records of the two rows of a tile

__device__ __forceinline__ void load_rows(const LevelRec* rows, int n0, int r, RowRecs& o)
{
    // rows ia = 2r+1 clamped into the event and ia + 1: two adjacent records, one address.  A clamped row is outside
    // every band, and so is row n0 + 1 (the record after the event's last one: the next event's first, or the spare
    // record at the end of the array) -- their emissions are computed for nobody.
    const int ia = min(max(2 * r + 1, 1), n0);
    const LevelRec* p = rows + (ia - 1);
    o.a = p[0];
    o.b = p[1];
}

// split-phase CTA barrier on an mbarrier object: a warp signals that its step is written, prepares its
// next step (strip switch, emissions) and only then waits for the others
__device__ __forceinline__ void mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// A failed try_wait comes back after ~10 cycles on sm_100a (ptxas drops the suspend-time hint), so a waiting warp spins:
// try_wait, branch and yield were 24 % of the instructions the fill executed.  PS_WAIT_SLEEP_NS > 0 puts a nanosleep
// into the loop (A/B knob; the warps that wait have 4-5 steps of slack around the ring, see fill_wave).
#ifndef PS_WAIT_SLEEP_NS
#define PS_WAIT_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
#if PS_WAIT_SLEEP_NS > 0
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "WAIT_LOOP:\n"
        "nanosleep.u32 %2;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "n"(PS_WAIT_SLEEP_NS) : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
#endif
}

template <bool REV, int MAXT, bool INV>
__device__ __forceinline__ void fill_wave(const Batch& b, const EvDesc& ev, const FillOut& o, double* smem)
{
    const int T = blockDim.x, tid = threadIdx.x;
    const int n0 = ev.n0, N = ev.N;
    const int J = (N + CW - 1) / CW;                      // strips
    // [RD][MAXT] rings of the last RD steps: second-column main values (and, reverse pass, emissions)
    // of both rows of the tile; then the [11][MAXT] 16-byte chunks of the thread's next strip record.
    //
    // Step synchronisation.  A tile only looks at the left neighbour's previous two steps, so a warp only
    // has to wait for the warp left of it (warp 0: the last warp, whose last thread owns the strip before
    // thread 0's next one).  P2P (classes of at most 6 warps): warp w signals "step q written" on its own
    // mbarrier [w][q & 7] (one arrival per phase) and warp w+1 waits for it before its step q+1 -- no
    // CTA-wide barrier in the sweep; the warps run as a pipeline skewed by their own pace.  Around the ring
    // of nw warps a warp can get nw-1 steps ahead of the warp that reads it, hence 8-deep data rings (the
    // reader still needs up to nw+1 <= 7 older slots) and 8 barriers per warp (the second arrival on a slot
    // comes 8 steps later, never before the reader has passed, so the parity wait is unambiguous).
    // Wider classes keep the split-phase CTA barrier with 4-deep rings.
    constexpr bool P2P = MAXT <= 192;
    constexpr int RD = P2P ? 8 : 4;
    double2* myC = reinterpret_cast<double2*>(smem) + tid;
    double2* myE = myC + RD * MAXT;
    const int left = tid == 0 ? T - 1 : tid - 1;
    const double2* lfC = reinterpret_cast<const double2*>(smem) + left;
    const double2* lfE = lfC + RD * MAXT;
    double2* nxt = reinterpret_cast<double2*>(smem) + 2 * RD * MAXT + tid;    // chunk q at nxt[q * MAXT]
    __shared__ unsigned long long step_bar[P2P ? 64 : 1];
    const int wrp = tid >> 5, nwarps = T >> 5;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&step_bar[0]);
    const unsigned bar_mine = bar0 + (P2P ? wrp * 64 : 0);                              // + 8 * (q & 7)
    const unsigned bar_left = bar0 + (P2P ? (wrp == 0 ? nwarps - 1 : wrp - 1) * 64 : 0);
    if (P2P) { if (tid < 8 * nwarps) mbar_init(bar0 + 8 * tid, 1); }
    else if (tid == 0) mbar_init(bar0, nwarps);           // one arrival per warp and step
    const LevelRec* rows = (REV ? b.rowB : b.rowF) + ev.lev_off;
    const StripRec* srec = b.strips + ev.strip_off + (REV ? J + 1 : 0);  // strips 0..J-1, sentinel J
    const ModelDev& md = b.models[ev.model];
    const Trans tr = {md.lskip, md.lstay, md.lext, md.lins};
    const double off = b.lik_offset, l2p = b.log2pi;

    StripRegs cur;
    StateParams p0, p1;
    strip_request<MAXT>(nxt, srec + min(tid, J));
    const int dstart = srec[0].rlo, dend = (J - 1) + srec[J - 1].rhi;
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    strip_from_smem<MAXT>(nxt, cur, p0, p1);
    cur.j = tid < J ? tid : 1 << 29;
    strip_request<MAXT>(nxt, srec + min(tid + T, J));
    double upC0 = NEG, upS0 = NEG, upE0 = 0, upC1 = NEG, upS1 = NEG, upE1 = 0;
    double best0 = NEG, best1 = NEG;
    int besti0 = 0, besti1 = 0;
    int ph = dstart & (RD - 1);
    unsigned parity = 0;
    RowRecs rr;
    double eA = 0, eB = 0, eC = 0, eD = 0;                // emissions of the tile of the coming step
    load_rows(rows, n0, dstart - cur.j, rr);
    if (dstart - cur.j >= cur.rlo && dstart - cur.j <= cur.rhi)
    {
        eA = emission(rr.a.mean, rr.a.stdv, rr.a.rstdv, rr.a.lsd3, p0, l2p, off);
        eB = emission(rr.a.mean, rr.a.stdv, rr.a.rstdv, rr.a.lsd3, p1, l2p, off);
        eC = emission(rr.b.mean, rr.b.stdv, rr.b.rstdv, rr.b.lsd3, p0, l2p, off);
        eD = emission(rr.b.mean, rr.b.stdv, rr.b.rstdv, rr.b.lsd3, p1, l2p, off);
    }
    load_rows(rows, n0, dstart + 1 - ((dstart + 1 > cur.j + cur.rhi) ? (cur.j + T < J ? cur.j + T : 1 << 29) : cur.j), rr);
    __syncthreads();                                      // the barrier object is initialised
    // band storage of step d: the run starts at band_off + d * rs, this thread's tile is at + 4 * tid (a wave class has
    // ts == blockDim.x slots per step, Job::build, so strip j's slot j % ts is the thread's own index)
    long long a = ev.band_off + (long long)dstart * ev.rs + 4 * tid;
    unsigned slot = 0, par = 0;                           // P2P: barrier slot 8 * (q & 7) and parity (q >> 3) & 1 of q = d - dstart
    for (int d = dstart; d <= dend; d++)
    {
#ifndef PS_ABL_NOSYNC
        if (d > dstart)
#else
        if (false)
#endif
        {
            if (P2P)                                                     // the left warp has written step d-1
                mbar_wait(bar_left + ((slot + 56u) & 63u), slot == 0 ? par ^ 1u : par);
            else { mbar_wait(bar0, parity); parity ^= 1u; }              // every warp has written step d-1
        }
        const int r = d - cur.j;
        const int w0 = ph, w1 = (ph + RD - 1) & (RD - 1), w2 = (ph + RD - 2) & (RD - 1);   // steps d, d-1, d-2
#ifndef PS_NO_LEAN
        // Is every tile of this warp's step an interior one (all four cells inside their bands and past their first
        // rows, every predecessor inside the band of the column before, valid states: the strip's lean_lo..lean_hi,
        // k_strips)?  Then the warp runs the same arithmetic with every mask a compile-time `true`: the selects that
        // only apply masks fold away.
        const bool warp_lean = __all_sync(0xffffffffu, !(r >= cur.rlo && r <= cur.rhi) || (r >= cur.lean_lo && r <= cur.lean_hi));
#endif
        if (r >= cur.rlo && r <= cur.rhi)
        {
            const int ia = 2 * r + 1, ib = ia + 1;
            const bool v0 = !INV || cur.sa >= 0, v1 = !INV || cur.sb >= 0;
            if (INV) { if (!v0) { eA = 0.0; eC = 0.0; } if (!v1) { eB = 0.0; eD = 0.0; } }
            // the column left of the strip: left neighbour's second column, or the blank column 0
            double Lm = 0, La = 0, Lb = 0, LEm = 0, LEa = 0;             // rows ia-1, ia, ib
            if (cur.j > 0)
            {
                const double2 u = lfC[w1 * MAXT];
                La = u.x; Lb = u.y;
                Lm = lfC[w2 * MAXT].y;
                if (REV) { LEa = lfE[w1 * MAXT].x; LEm = lfE[w2 * MAXT].y; }
            }
            const bool inA = ia >= cur.i0a && ia <= cur.i1a, inB = ia >= cur.i0b && ia <= cur.i1b;
            const bool inC = ib >= cur.i0a && ib <= cur.i1a, inD = ib >= cur.i0b && ib <= cur.i1b;
            // band predicates of the horizontal / diagonal moves (previous column's band)
            const bool skA = ia >= cur.pp0 && ia <= cur.pp1, dgA = ia > cur.pp0 && ia <= cur.pp1;
            const bool skC = ib >= cur.pp0 && ib <= cur.pp1, dgC = ib > cur.pp0 && ib <= cur.pp1;
            const bool skB = inA, dgB = ia > cur.i0a && ia <= cur.i1a;
            const bool skD = inC, dgD = ib > cur.i0a && ib <= cur.i1a;
            double CA, CB, CC, CD, SA, SB, SC, SD, M;
            int kA, kB, kC, kD, qA, qB, qC, qD, m;
#ifndef PS_NO_LEAN
            const bool lean = warp_lean;
            if (lean)
            {
                cell_pre<false, 0>(true, true, false, true, Lm, REV ? LEm : eA, REV ? upE0 : eA, upC0, upS0, tr, M, m, SA, qA);
                cell_fin(true, La, M, m, tr, CA, kA);
                cell_pre<false, 8>(true, true, false, true, upC0, REV ? upE0 : eB, REV ? upE1 : eB, upC1, upS1, tr, M, m, SB, qB);
                cell_fin(true, CA, M, m, tr, CB, kB);
                cell_pre<false, 16>(true, true, false, true, La, REV ? LEa : eC, REV ? eA : eC, CA, SA, tr, M, m, SC, qC);
                cell_fin(true, Lb, M, m, tr, CC, kC);
                cell_pre<false, 24>(true, true, false, true, CA, REV ? eA : eD, REV ? eB : eD, CB, SB, tr, M, m, SD, qD);
                cell_fin(true, CC, M, m, tr, CD, kD);
            }
            else
#endif
            {
            // row ia
            cell_pre<INV, 0>(inA, v0, ia == cur.i0a, dgA, Lm, REV ? (dgA ? LEm : 0.0) : eA, REV ? upE0 : eA, upC0, upS0, tr, M, m, SA, qA);
            cell_fin(skA && inA && v0, La, M, m, tr, CA, kA);
            cell_pre<INV, 8>(inB, v1, ia == cur.i0b, dgB, upC0, REV ? (dgB ? upE0 : 0.0) : eB, REV ? upE1 : eB, upC1, upS1, tr, M, m, SB, qB);
            cell_fin(skB && inB && v1, CA, M, m, tr, CB, kB);
            // row ib
            cell_pre<INV, 16>(inC, v0, ib == cur.i0a, dgC, La, REV ? (dgC ? LEa : 0.0) : eC, REV ? eA : eC, CA, SA, tr, M, m, SC, qC);
            cell_fin(skC && inC && v0, Lb, M, m, tr, CC, kC);
            cell_pre<INV, 24>(inD, v1, ib == cur.i0b, dgD, CA, REV ? (dgD ? eA : 0.0) : eD, REV ? eB : eD, CB, SB, tr, M, m, SD, qD);
            cell_fin(skD && inD && v1, CC, M, m, tr, CD, kD);
            }
            // the second column of the strip is what the right neighbour reads
            myC[w0 * MAXT] = make_double2(CB, CD);
            if (REV) myE[w0 * MAXT] = make_double2(eB, eD);
            if (v0 && CA > best0) { best0 = CA; besti0 = ia; }
            if (v0 && CC > best0) { best0 = CC; besti0 = ib; }
            if (v1 && CB > best1) { best1 = CB; besti1 = ia; }
            if (v1 && CD > best1) { best1 = CD; besti1 = ib; }
            upC0 = CC; upS0 = SC; upE0 = eC; upC1 = CD; upS1 = SD; upE1 = eD;
            double2* pm = reinterpret_cast<double2*>(o.Mm + a);
            double2* ps = reinterpret_cast<double2*>(o.Ms + a);
            pm[0] = make_double2(CA, CB); pm[1] = make_double2(CC, CD);
            // the reverse stay matrix is never read (main >= stay in every cell: the joins need B's main only)
            if (!REV) { ps[0] = make_double2(SA, SB); ps[1] = make_double2(SC, SD); }
            if (!REV)
                *reinterpret_cast<unsigned*>(b.Fstep + a) = (unsigned)((kA | qA | kB) | (qB | kC | qC) | (kD | qD));
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(P2P ? bar_mine + slot : bar0);   // this warp's part of step d is in shared memory
        slot = (slot + 8u) & 63u; par ^= (slot == 0);
        a += ev.rs;
        ph = (ph + 1) & (RD - 1);
        // ---- preparation of step d+1: overlaps the other warps' step d ----
        // Lanes finish their strips on different steps (about one lane every other step in the warp at the lower band
        // edge); when the event leaves every thread enough idle steps between two strips (ev.lazy, Job::plan_event), the
        // warp switches all its finished lanes together on every 2nd, 4th or 8th step instead of paying the switch every time.
        if (d + 1 > cur.j + cur.rhi && ((d & ev.lazy) == ev.lazy || d == dend))
        {
            // strip finished: publish the best cell of its columns, take the next strip from shared memory
            // and request the one after it
            if (cur.j < J)
            {
                const int k = CW * cur.j + 1;
                const long long g = ev.col_off + k;
                o.Mcb[g] = best0; o.Mcbi[g] = besti0;
                if (k + 1 <= N) { o.Mcb[g + 1] = best1; o.Mcbi[g + 1] = besti1; }
            }
            best0 = NEG; best1 = NEG; besti0 = 0; besti1 = 0;
            upC0 = NEG; upS0 = NEG; upC1 = NEG; upS1 = NEG;
            const int jn = cur.j + T;
            asm volatile("cp.async.wait_all;\n" ::: "memory");
            strip_from_smem<MAXT>(nxt, cur, p0, p1);
            cur.j = jn < J ? jn : 1 << 29;
            strip_request<MAXT>(nxt, srec + min(jn + T, J));
        }
        {
            const int rn = d + 1 - cur.j;
            if (rn >= cur.rlo && rn <= cur.rhi)
            {
                eA = emission(rr.a.mean, rr.a.stdv, rr.a.rstdv, rr.a.lsd3, p0, l2p, off);
                eB = emission(rr.a.mean, rr.a.stdv, rr.a.rstdv, rr.a.lsd3, p1, l2p, off);
                eC = emission(rr.b.mean, rr.b.stdv, rr.b.rstdv, rr.b.lsd3, p0, l2p, off);
                eD = emission(rr.b.mean, rr.b.stdv, rr.b.rstdv, rr.b.lsd3, p1, l2p, off);
            }
            // row records of step d+2, in flight across the wait
            const int rnext = d + 2 - ((d + 2 > cur.j + cur.rhi) ? (cur.j + T < J ? cur.j + T : 1 << 29) : cur.j);
            load_rows(rows, n0, rnext, rr);
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// Serial schedule (arbitrary band layout): the same cells, column by column, by one thread; the
// previous column's emissions (reverse pass only) alternate between two RS-long strips of smem.
template <bool REV>
__device__ void fill_serial(const Batch& b, const EvDesc& ev, const FillOut& o, double* smem)
{
    const int n0 = ev.n0, N = ev.N, RS = b.RS;
    const LevelRec* lev = b.lev + ev.lev_off;
    const ModelDev& md = b.models[ev.model];
    const Trans tr = {md.lskip, md.lstay, md.lext, md.lins};
    int p0 = 0, p1 = n0;
    for (int k = 1; k <= N; k++)
    {
        int i0, i1;
        col_band(b, ev, REV, k, i0, i1);
        const int s = b.states[ev.state_off + (REV ? N - k + 1 : k) - 1];
        StateParams sp;
        if (s >= 0) sp = md.st[s];
        double upC = 0, upS = 0, upE = 0, best = NEG;
        int besti = 0;
        double* Ecur = smem + (k & 1) * RS;
        const double* Eprev = smem + ((k - 1) & 1) * RS;
        for (int i = i0; i <= i1; i++)
        {
            double C = 0, S = 0, e = 0;
            int step = ST_STOP;
            if (s >= 0)
            {
                e = cell_emission<REV>(lev, n0, i, sp, b.log2pi, b.lik_offset);
                const bool skip_ok = i >= p0 && i <= p1;
                const bool diag_ok = i > p0 && i <= p1;
                double Pi = 0, Pi1 = 0, PE = 0;
                if (k > 1)
                {
                    if (skip_ok) Pi = o.Mm[cell_at(ev, k - 1, i)];
                    if (diag_ok) { Pi1 = o.Mm[cell_at(ev, k - 1, i - 1)]; PE = Eprev[i - 1 - p0]; }
                }
                const double eM = REV ? (diag_ok ? PE : 0.0) : e;
                const double eU = REV ? upE : e;
                dp_cell(i == i0, skip_ok, diag_ok, Pi, Pi1, eM, eU, upC, upS, tr, C, S, step);
                if (C > best) { best = C; besti = i; }
            }
            const long long a = cell_at(ev, k, i);
            o.Mm[a] = C;
            if (!REV) o.Ms[a] = S;
            if (!REV) b.Fstep[a] = (uint8_t)step;
            Ecur[i - i0] = e;
            upC = C; upS = S; upE = e;
        }
        const long long g = ev.col_off + k;
        o.Mi0[g] = i0; o.Mlen[g] = i1 - i0 + 1;
        o.Mcb[g] = best; o.Mcbi[g] = besti;
        p0 = i0; p1 = i1;
    }
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_fill(Batch b, int list_off)
{
    extern __shared__ double smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const int e = b.fill_list[list_off + blockIdx.x];
    const EvDesc ev = b.ev[e];
    const bool rev = blockIdx.y != 0;
    if (!ev.usable || ev.N <= 0) return;
    const int N = ev.N;
    FillOut o;
    o.Mm = rev ? b.Bm : b.Fm; o.Ms = rev ? nullptr : b.Fs;
    o.Mi0 = rev ? b.Bi0 : b.Fi0; o.Mlen = rev ? b.Blen : b.Flen;
    o.Mcb = rev ? b.Bcb : b.Fcb; o.Mcbi = rev ? b.Bcbi : b.Fcbi;
    double* Mcb = o.Mcb; int* Mcbi = o.Mcbi;
    if (b.mono[e])
    {
        if (ev.inv)
        {
            if (rev) fill_wave<true, MAXT, true>(b, ev, o, smem); else fill_wave<false, MAXT, true>(b, ev, o, smem);
        }
        else
        {
            if (rev) fill_wave<true, MAXT, false>(b, ev, o, smem); else fill_wave<false, MAXT, false>(b, ev, o, smem);
        }
    }
    else if (tid == 0)
    {
        if (rev) fill_serial<true>(b, ev, o, smem); else fill_serial<false>(b, ev, o, smem);
    }
    __syncthreads();

    // running best over columns: first maximum in (column, row) order wins (cpp/Alignment.cpp:31-36,
    // :158, :270).  Columns whose best is not > the carried one leave it untouched.
    double* Mbest = rev ? b.Bbest : b.Fbest;
    __shared__ double sh_s[32];
    __shared__ int sh_i[32], sh_j[32];
    __shared__ double car_s;
    __shared__ int car_i, car_j;
    if (tid == 0) { car_s = 0.0; car_i = 0; car_j = 0; }
    __syncthreads();
    const int lane = tid & 31, wid = tid >> 5, nw = (T + 31) >> 5;
    for (int base = 0; base < N; base += T)
    {
        int k = base + tid + 1;
        double s = NEG; int bi = 0, bj = 0;
        if (k <= N) { s = Mcb[ev.col_off + k]; bi = Mcbi[ev.col_off + k]; bj = rev ? N - k + 1 : k; }
        // inclusive scan, earlier element wins ties
        for (int o = 1; o < 32; o <<= 1)
        {
            double s2 = __shfl_up_sync(0xffffffffu, s, o);
            int i2 = __shfl_up_sync(0xffffffffu, bi, o), j2 = __shfl_up_sync(0xffffffffu, bj, o);
            if (lane >= o && !(s > s2)) { s = s2; bi = i2; bj = j2; }
        }
        if (lane == 31) { sh_s[wid] = s; sh_i[wid] = bi; sh_j[wid] = bj; }
        __syncthreads();
        if (wid == 0)
        {
            double ws = lane < nw ? sh_s[lane] : NEG;
            int wi = lane < nw ? sh_i[lane] : 0, wj = lane < nw ? sh_j[lane] : 0;
            for (int o = 1; o < 32; o <<= 1)
            {
                double s2 = __shfl_up_sync(0xffffffffu, ws, o);
                int i2 = __shfl_up_sync(0xffffffffu, wi, o), j2 = __shfl_up_sync(0xffffffffu, wj, o);
                if (lane >= o && !(ws > s2)) { ws = s2; wi = i2; wj = j2; }
            }
            if (lane < nw) { sh_s[lane] = ws; sh_i[lane] = wi; sh_j[lane] = wj; }
        }
        __syncthreads();
        if (wid > 0 && !(s > sh_s[wid - 1])) { s = sh_s[wid - 1]; bi = sh_i[wid - 1]; bj = sh_j[wid - 1]; }
        if (!(s > car_s)) { s = car_s; bi = car_i; bj = car_j; }
        if (k <= N)
        {
            Mbest[ev.col_off + k] = s;
            if (!rev) { b.Fbi[ev.col_off + k] = bi; b.Fbj[ev.col_off + k] = bj; }
        }
        __syncthreads();
        if (tid == T - 1) { car_s = s; car_i = bi; car_j = bj; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// updaterefs (cpp/EventData.h:110-169), sequential restatement used by k_backtrace's thread 0.
__device__ void dev_updaterefs(const double* ra, double* ri, int n0, int& empty, int& refstart, int& refend)
{
    int a = 0, z = n0 - 1;
    while (a < n0 && !(ra[a] > 0)) a++;
    while (z >= 0 && !(ra[z] > 0)) z--;
    if (a == n0 || z < 0) { empty = 1; refstart = -1; refend = -1; return; }
    empty = 0;
    refstart = (int)ra[a];
    refend = (int)ra[z];
    double slope = (ra[z] - ra[a]) / (double)(z - a);
    double icpt = ra[a] - slope * a;
    int last = -1;
    for (int i = 0; i < n0; i++)
    {
        double v = ra[i];
        if (i < a || i > z) ri[i] = slope * i + icpt;
        else
        {
            ri[i] = v;
            if (v > 0)
            {
                if (last > 0)
                {
                    double m = (v - ra[last]) / (i - last);
                    for (int j = last + 1; j < i; j++) ri[j] = m * (j - last) + ra[last];
                }
                last = i;
            }
        }
    }
}

// The same by a whole warp (k_backtrace): every level between the first and the last aligned one that is not aligned
// itself is interpolated between its nearest aligned neighbours p < i < n with the sequential loop's own expression
// (m = (ra[n] - ra[p]) / (n - p); ri[i] = m * (i - p) + ra[p]) -- unless p == 0: the loop above tests `last > 0`, so
// the gap behind an aligned level 0 keeps its raw values; the ends are extrapolated.  p comes from a forward max-scan
// (kept in `prevpos`, n0 ints of scratch), n from a backward min-scan.  `ri` may alias `ra`: aligned entries are never
// rewritten and every other entry is read by the lane that rewrites it.
__device__ void warp_updaterefs(const double* ra, double* ri, int* prevpos, int n0, int lane, int& empty, int& refstart, int& refend)
{
    int a = 1 << 30, z = -1;
    for (int i = lane; i < n0; i += 32)
        if (ra[i] > 0) { a = min(a, i); z = i; }
    for (int o = 16; o; o >>= 1)
    {
        a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
        z = max(z, __shfl_xor_sync(0xffffffffu, z, o));
    }
    if (z < 0) { empty = 1; refstart = -1; refend = -1; return; }
    empty = 0;
    refstart = (int)ra[a];
    refend = (int)ra[z];
    const double slope = (ra[z] - ra[a]) / (double)(z - a);
    const double icpt = ra[a] - slope * a;
    int carry = -1;
    for (int c0 = 0; c0 < n0; c0 += 32)
    {
        const int i = c0 + lane;
        int v = (i < n0 && ra[i] > 0) ? i : -1;
        for (int o = 1; o < 32; o <<= 1)
        {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v = max(v, u);
        }
        v = max(v, carry);
        if (i < n0) prevpos[i] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    carry = 1 << 30;
    for (int c0 = ((n0 - 1) >> 5) << 5; c0 >= 0; c0 -= 32)
    {
        const int i = c0 + lane;
        const double x = i < n0 ? ra[i] : 0.0;
        const bool pos = i < n0 && x > 0;
        int v = pos ? i : 1 << 30;
        for (int o = 1; o < 32; o <<= 1)
        {
            const int u = __shfl_down_sync(0xffffffffu, v, o);
            if (lane + o < 32) v = min(v, u);
        }
        v = min(v, carry);
        if (i < n0)
        {
            double r = x;
            if (i < a || i > z) r = slope * i + icpt;
            else if (!pos)
            {
                const int p = prevpos[i];
                if (p > 0)
                {
                    const double m = (ra[v] - ra[p]) / (double)(v - p);
                    r = m * (double)(i - p) + ra[p];
                }
            }
            if (!(pos && ri == ra)) ri[i] = r;            // in place, an aligned entry stays as it is: other lanes read it
        }
        carry = __shfl_sync(0xffffffffu, v, 0);
        __syncwarp();
    }
}

// k_backtrace: one WARP per event.  The best path is followed from the best cell through the packed step
// bytes (cpp/Alignment.cpp:516-605).  A pointer chase through global memory costs one memory round trip
// per move, so the warp fetches a block of 32 columns x 32 rows below and left of the current cell in one
// go -- in the wavefront-major layout a 2x2 tile is one 32-bit word of step bytes, lane l takes eight tiles of
// strip s_hi - (l & 15) -- into shared memory, walks inside the block (the walk itself is uniform across the
// warp; runs of matches are taken by the lanes together), and fetches the next block where the walk
// leaves this one: one round trip per ~32 moves.  Every visited level records its column and matrix; the
// warp then gathers ref_like in parallel and lane 0 rebuilds ref_index.  BT_WARPS events per CTA, their
// per-level scratch in shared memory when the events fit.
constexpr int BT_WARPS = 4;
constexpr int BT_BS = 16;           // block side in tiles (strips x row pairs): 32 x 32 cells per fetch

__global__ void __launch_bounds__(32 * BT_WARPS) k_backtrace(Batch b, int smem_levels, int warps_per_cta)
{
    extern __shared__ double bt_dyn[];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int e = blockIdx.x * warps_per_cta + wrp;
    if (wrp >= warps_per_cta || e >= b.n_events) return;
    const EvDesc ev = b.ev[e];
    if (!ev.usable) return;
    const int n0 = ev.n0, N = ev.N;
    double* ra = b.ref_align + ev.lev_off;
    double* rl = b.ref_like + ev.lev_off;
    double* ri = b.ref_index + ev.lev_off;
    // the per-level scratch of the walk lives in shared memory when the event fits
    const bool in_smem = n0 <= smem_levels;
    double* val = in_smem ? bt_dyn + (size_t)wrp * smem_levels : ra;
    int* src = in_smem ? (int*)(bt_dyn + (size_t)warps_per_cta * smem_levels) + (size_t)wrp * smem_levels : b.bt_src + ev.lev_off;
    for (int i = lane; i < n0; i += 32) { val[i] = 0.0; src[i] = 0; }
    __syncwarp();
    int i = N > 0 ? b.Fbi[ev.col_off + N] : 0, j = N > 0 ? b.Fbj[ev.col_off + N] : 0, arr = 0;
    bool go = i > 0 && j > 0;
    const int ts = ev.ts;
    // the block of step words around the walk: BT_BS strips x BT_BS row pairs (one 32-bit word per 2x2 tile) and the
    // bands of its 2 BT_BS columns, staged in shared memory
    __shared__ unsigned bt_words[BT_WARPS][BT_BS * BT_BS];
    __shared__ int2 bt_band[BT_WARPS][2 * BT_BS];
    unsigned* W = bt_words[wrp];
    int2* BB = bt_band[wrp];
    while (go)
    {
        // block of strips s_hi-15 .. s_hi and row pairs r_hi-15 .. r_hi, the current cell in its top right tile:
        // 32 columns x 32 rows.  Lane l fetches the words of strip s_hi - (l & 15), row pairs r_hi - (l >> 4) - 2q.
        const int s_hi = (j - 1) >> 1, r_hi = (i - 1) >> 1;
        const int kcol0 = 2 * (s_hi - (BT_BS - 1)) + 1;
        __syncwarp();                                            // the walk through the previous block is over
        {
            const int ds = lane & (BT_BS - 1), sl = s_hi - ds;
            int slot = sl >= 0 ? sl % ts : 0;
            unsigned w[BT_BS / 2];
#pragma unroll
            for (int q = 0; q < BT_BS / 2; q++)
            {
                const int r = r_hi - (lane >> 4) - 2 * q;
                w[q] = (unsigned)ST_STOP * 0x01010101u;
                if (sl >= 0 && r >= 0)
                    w[q] = *reinterpret_cast<const unsigned*>(b.Fstep + ev.band_off + ((long long)(sl + r) * ts + slot) * 4);
            }
            int c0 = 1, c1 = 0;
            {
                const int k = kcol0 + lane;
                if (k >= 1 && k <= N) { const long long g = ev.col_off + k; c0 = b.Fi0[g]; c1 = c0 + b.Flen[g] - 1; }
            }
#pragma unroll
            for (int q = 0; q < BT_BS / 2; q++) W[((lane >> 4) + 2 * q) * BT_BS + ds] = w[q];
            BB[lane] = make_int2(c0, c1);
        }
        __syncwarp();
        const int jlo = kcol0, ilo = 2 * (r_hi - (BT_BS - 1)) + 1;   // the block covers columns >= jlo and rows >= ilo
        while (true)
        {
            if (!(i > 0 && j > 0)) { go = false; break; }
            if (j < jlo || i < ilo) break;                       // left the block: fetch the next one
            if (arr == 0)
            {
                // Runs of matches (most of a path) in one go: lane l looks at the main-matrix cell l steps down the
                // diagonal, (i - l, j - l); the lanes before the first one that is not an in-band match inside the
                // block record their level and the walk jumps over them.  What ends the run is handled by the
                // one-move code below, exactly as if the moves had been taken one by one.
                const int il = i - lane, jl = j - lane;
                bool is_match = false;
                if (il >= ilo && jl >= jlo && il > 0 && jl > 0)
                {
                    const int2 bd = BB[jl - kcol0];
                    if (il >= bd.x && il <= bd.y)
                    {
                        const unsigned wl = W[(r_hi - ((il - 1) >> 1)) * BT_BS + (s_hi - ((jl - 1) >> 1))];
                        is_match = ((wl >> (8 * ((((il - 1) & 1) << 1) + ((jl - 1) & 1)))) & 7u) == (unsigned)ST_MATCH;
                    }
                }
                const int run = __ffs(~__ballot_sync(0xffffffffu, is_match)) - 1;    // 32 matches: ~0 has no bit, ffs = 0
                if (run != 0)
                {
                    const int n = run < 0 ? 32 : run;
                    if (lane < n) { val[il - 1] = (double)jl; src[il - 1] = 2 * jl; }
                    i -= n; j -= n;
                    continue;
                }
            }
            const int2 bd = BB[j - kcol0];
            int st = ST_STOP;
            if (i >= bd.x && i <= bd.y)
            {
                const unsigned w = W[(r_hi - ((i - 1) >> 1)) * BT_BS + (s_hi - ((j - 1) >> 1))];
                st = (int)((w >> (8 * ((((i - 1) & 1) << 1) + ((j - 1) & 1)))) & 0xffu);
            }
            const int mv = arr ? ((st >> 3) & 3) : (st & 7);
            if (arr == 0)
            {
                if (mv == ST_STOP || mv == ST_IMPLICIT || mv == 5) { go = false; break; }
                if (mv == ST_SKIP) { j--; }
                else if (mv == ST_MATCH) { if (lane == 0) { val[i - 1] = (double)j; src[i - 1] = 2 * j; } i--; j--; }
                else if (mv == ST_IGNORE) { if (lane == 0) { val[i - 1] = -1.0; src[i - 1] = 2 * j; } i--; j--; }
                else if (mv == ST_INSERT) { if (lane == 0) { val[i - 1] = -1.0; src[i - 1] = 2 * j; } i--; }
                else /* ST_STAY: hop to the stay matrix, same cell */ arr = 1;
            }
            else
            {
                if (mv == 0) { go = false; break; }
                if (lane == 0) { val[i - 1] = (double)j; src[i - 1] = 2 * j + 1; }
                i--;
                if (mv == 1) arr = 0;                             // stay: back to the main matrix one row up
            }
        }
    }
    __syncwarp();
    // ref_align out, ref_like gathered from the recorded (column, matrix) of every aligned level
    for (int q = lane; q < n0; q += 32)
    {
        if (in_smem) ra[q] = val[q];
        const int sc = src[q];
        double like = 0.0;
        if (sc)
        {
            const long long a = cell_at(ev, sc >> 1, q + 1);
            like = (sc & 1) ? b.Fs[a] : b.Fm[a];
        }
        rl[q] = like;
    }
    __syncwarp();
    int empty = 0;
    {
        int rs, re;
        // in shared memory the interpolation runs in place; the walk's `src` is free again and serves as scratch
        warp_updaterefs(val, in_smem ? val : ri, src, n0, lane, empty, rs, re);
        if (lane == 0)
        {
            b.ri_empty[e] = empty;
            b.refstart[e] = rs;
            b.refend[e] = re;
        }
    }
    __syncwarp();
    if (in_smem && !empty)
        for (int q = lane; q < n0; q += 32) ri[q] = val[q];
}

// ------------------------------------------------------------------------------------------
// Join of a forward column with a reverse column (cpp/Alignment.h:178-214).  Every cell has
// main >= stay and every column's running best >= each of its main cells, so rows present in only
// one of the two bands can never beat the two running bests: the maximum over all rows reduces to
// the rows present in BOTH bands plus the two running bests (and the floor 0).
//
// k_join: old[g] = columnMax(c) = join(F[c], B[N-c+1]) for every band column.  A block covers 32
// consecutive columns = 16 forward strips (lane = column) with 8 warps that take the wavefront steps
// d = strip + row pair round robin; in the wavefront-major layout the forward tiles a warp reads in
// one step are one contiguous 512-byte run, and the matching reverse cells (column N-c+1, rows
// n0-jf+1) belong to one reverse step as well.
__global__ void __launch_bounds__(256) k_join(Batch b)
{
    const EvDesc ev = b.ev[blockIdx.y];
    if (!ev.usable) return;
    const int N = ev.N, n0 = ev.n0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane + 1;
    const bool have = c <= N;
    const int cb = N - c + 1;                           // reverse column joined with forward column c
    const int j = (c - 1) >> 1;                         // forward strip of this lane
    int lo = 1, hi = 0;
    double m = 0.0;
    long long gf = 0;
    if (have)
    {
        gf = ev.col_off + c;
        const long long gb = ev.col_off + cb;
        const int f0 = b.Fi0[gf], flen = b.Flen[gf], b0 = b.Bi0[gb], blen = b.Blen[gb];
        // jf in [f0, f0+flen-1] and jb = n0-jf+1 in [b0, b0+blen-1]
        lo = max(max(f0, n0 + 1 - (b0 + blen - 1)), 1);
        hi = min(min(f0 + flen - 1, n0 + 1 - b0), n0);
        m = fmax(b.Fbest[gf], b.Bbest[gb]);
    }
    // common step range of the 32 columns
    int dlo = have && lo <= hi ? j + ((lo - 1) >> 1) : 1 << 30, dhi = have && lo <= hi ? j + ((hi - 1) >> 1) : -1;
    for (int o = 16; o; o >>= 1)
    {
        dlo = min(dlo, __shfl_xor_sync(0xffffffffu, dlo, o));
        dhi = max(dhi, __shfl_xor_sync(0xffffffffu, dhi, o));
    }
    const long long fb = have ? col_base(ev, c) : 0, bb = have ? col_base(ev, cb) : 0;
    const long long rs = ev.rs;
    for (int d = dlo + w; d <= dhi; d += 8)
    {
        const int r = d - j;
#pragma unroll
        for (int h = 0; h < 2; h++)
        {
            const int jf = 2 * r + 1 + h;
            if (have && jf >= lo && jf <= hi)
            {
                const long long af = fb + (long long)r * rs + 2 * h;
                const long long ab = bb + row_off(rs, n0 - jf + 1);
                m = fmax(m, b.Fm[af] + b.Bm[ab]);
            }
        }
    }
    __shared__ double part[8][32];
    part[w][lane] = m;
    __syncthreads();
    if (w == 0 && have)
    {
#pragma unroll
        for (int q = 1; q < 8; q++) m = fmax(m, part[q][lane]);
        b.old[gf] = fmax(m, 0.0);
    }
}

// ------------------------------------------------------------------------------------------
// Mutated-sequence helpers (cpp/Sequence.h:38-100).
struct MutView
{
    const char* bases; int L;        // region sequence
    const char* mstr;                // mutation's replacement bases
    int start, n_orig, n_mut;
    bool applied;                    // start < L (cpp/Sequence.h:41)
    int Lm;                          // mutated length
};

__device__ __forceinline__ int base_code(char ch)
{
    return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : (int)ch;
}

__device__ __forceinline__ int mut_base(const MutView& v, int q)
{
    char ch;
    if (!v.applied || q < v.start) ch = v.bases[q];
    else if (q < v.start + v.n_mut) ch = v.mstr[q - v.start];
    else ch = v.bases[q - v.n_mut + v.n_orig];
    return base_code(ch);
}

// state of mutated position k (0-based): -1 iff the leftmost base of the window is not ACGT; bases
// added before the most recent such reset are dropped (cpp/Sequence.h:79-98)
__device__ __forceinline__ int mut_state(const MutView& v, int k)
{
    int b0 = mut_base(v, k);
    if (b0 >= 4) return -1;                          // cpp/Sequence.h:87 tests `< 4` only
    int from = k;                                   // first base index still contributing
    for (int kk = k - 1; kk >= max(0, k - 4); kk--)
    {
        int bb = mut_base(v, kk);
        if (bb >= 4) { from = max(from, kk + 5); break; }
    }
    int cur = 0;
    for (int q = max(k, from); q <= k + 4; q++) cur += mut_base(v, q) << (2 * (k + 4 - q));
    return cur & (N_STATES - 1);
}

// ------------------------------------------------------------------------------------------
// k_mutscore (generic form): one thread per (event, mutation) task, columns one after another,
// the previous column's main-matrix values kept in a per-thread ring (see below).  Handles any
// mutation length and every boundary case of cpp/Alignment.cpp:447-512.
__device__ __forceinline__ bool task_decode(const Batch& b, long long t, int& e, int& m)
{
    // events are laid out with nondecreasing task_off; find the event owning task t
    int lo = 0, hi = b.n_events - 1;
    while (lo < hi)
    {
        int mid = (lo + hi + 1) >> 1;
        if (b.ev[mid].task_off <= t) lo = mid; else hi = mid - 1;
    }
    e = lo;
    m = (int)(t - b.ev[lo].task_off);
    return m < b.ev[lo].n_muts;
}

// single-thread join used for the boundary cases (seed column joined directly, blank columns)
__device__ double thread_join(const Batch& b, const EvDesc& ev, int raf, int rab)
{
    const int N = ev.N, n0 = ev.n0;
    raf = min(max(raf, 0), N); rab = min(max(rab, 0), N);
    const long long gf = ev.col_off + raf, gb = ev.col_off + rab;
    int f0 = 0, flen = n0 + 1, b0 = 0, blen = n0 + 1;
    double mf = 0.0, mb = 0.0;
    if (raf > 0) { f0 = b.Fi0[gf]; flen = b.Flen[gf]; mf = b.Fbest[gf]; }
    if (rab > 0) { b0 = b.Bi0[gb]; blen = b.Blen[gb]; mb = b.Bbest[gb]; }
    int lo = max(max(f0, n0 + 1 - (b0 + blen - 1)), 1), hi = min(min(f0 + flen - 1, n0 + 1 - b0), n0);
    double m = fmax(0.0, fmax(mf, mb));
    for (int jf = lo; jf <= hi; jf++)
    {
        int jb = n0 - jf + 1;
        const long long af = cell_at(ev, raf, jf), ab = cell_at(ev, rab, jb);
        const double fm = raf > 0 ? b.Fm[af] : 0.0, bm = rab > 0 ? b.Bm[ab] : 0.0;
        m = fmax(m, fm + bm);
    }
    return m;
}

// Strip storage of the previous narrow column's main-matrix values.  SMEM: a ring of S = 2W+2
// slots per thread in shared memory, updated in place (slot = row % S; the cell (c, i) reads the
// old value of its own slot = (c-1, i) and keeps it one more iteration as (c-1, i-1)).
// GLOBAL: the same ring in a per-thread strip of global scratch, for widths whose ring does not
// fit in shared memory.
// LIST: the tasks are the (flagged mutation, event of its region) pairs of b.flag_list instead of all pairs
__device__ __forceinline__ bool list_task(const Batch& b, long long t, int& e, int& m)
{
    const long long q = t / b.max_ev;
    const int el = (int)(t % b.max_ev);
    const long long g = b.flag_list[q];
    int lo = 0, hi = b.n_regs - 1;
    while (lo < hi)
    {
        int mid = (lo + hi + 1) >> 1;
        if (b.regs[mid].mut_off <= g) lo = mid; else hi = mid - 1;
    }
    if (el >= b.regs[lo].nev) return false;
    e = b.regs[lo].ev0 + el;
    m = (int)(g - b.regs[lo].mut_off);
    return true;
}

template <bool SMEM, bool LIST>
__global__ void __launch_bounds__(128) k_mutscore(Batch b)
{
    extern __shared__ double ring_smem[];
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int W = b.scoring_width;
    const int S = 2 * W + 2;
    double* ring = SMEM ? ring_smem + threadIdx.x : b.scratch + gtid;
    const long long rstride = SMEM ? 128 : nthreads;              // element r at ring[r * rstride]
    const long long n_total = LIST ? (long long)(*b.flag_count) * b.max_ev : b.n_tasks;
    if (LIST && n_total <= b.warp_limit) return;                   // few flagged pairs: k_mutscore_warp does them
    for (long long t = gtid; t < n_total; t += nthreads)
    {
        int e, m;
        if (LIST ? !list_task(b, t, e, m) : !task_decode(b, t, e, m)) continue;
        const EvDesc ev = b.ev[e];
        const MutDev mu = b.muts[ev.mut_off + m];
        double result = 0.0;
        if (ev.usable && !((unsigned)mu.start > (unsigned)ev.L))        // cpp/MakeMutations.cpp:46
        {
            const int N = ev.N, n0 = ev.n0, L = ev.L;
            MutView mv;
            mv.bases = b.bases + ev.base_off; mv.L = L; mv.mstr = b.mut_str + mu.str_off;
            mv.start = mu.start; mv.n_orig = mu.n_orig; mv.n_mut = mu.n_mut;
            mv.applied = mu.start < L;
            mv.Lm = mv.applied ? mu.start + mu.n_mut + max(0, L - mu.start - mu.n_orig) : L;
            const int Nm = mv.Lm >= 5 ? mv.Lm - 4 : 0;
            // old score: columnMax(max(start-3,1)) with the reference's clamps
            const int raf = max(mu.start - 3, 1);
            double old;
            if (raf <= N) old = b.old[ev.col_off + raf];
            else old = thread_join(b, ev, raf, N - raf + 1);
            const int startind = max(mu.start - 4, 0);
            const int refind = mu.start + mu.n_mut + 1;
            int last = min(min(refind, startind + mu.n_mut + 6), Nm);      // last column that gets filled
            if (W == 0) last = startind;
            double neu;
            if (last <= startind)
            {
                // nothing appended: the seed column itself is joined (cpp/Alignment.cpp:483-499)
                neu = thread_join(b, ev, startind, Nm - startind + 1);
            }
            else
            {
                const LevelRec* lev = b.lev + ev.lev_off;
                const ModelDev& md = b.models[ev.model];
                const Trans tr = {md.lskip, md.lstay, md.lext, md.lins};
                const bool ri_empty = b.ri_empty[e] != 0;
                // seed column
                int p0 = 0, p1 = n0;
                const double* seed = nullptr;
                double best = 0.0;
                if (startind > 0)
                {
                    const long long gs = ev.col_off + startind;
                    p0 = b.Fi0[gs]; p1 = p0 + b.Flen[gs] - 1;
                    seed = b.Fm + col_base(ev, startind);          // + row * rs
                    best = b.Fbest[gs];
                }
                // reverse column to join with
                const int rab = min(max(Nm - last + 1, 0), N);
                const long long gb = ev.col_off + rab;
                int b0 = 0, blen = n0 + 1;
                double mb = 0.0;
                if (rab > 0) { b0 = b.Bi0[gb]; blen = b.Blen[gb]; mb = b.Bbest[gb]; }
                const long long bbase = col_base(ev, rab);            // + reverse row * rs
                const double* Bm = b.Bm + bbase;
                const long long ts = ev.rs;
                double joinmax = 0.0;
                for (int c = startind + 1; c <= last; c++)
                {
                    const int mid = ri_empty ? 1 : b.cen_new[ev.cen_off + c];
                    int i0, i1;
                    band_of(mid, n0, W, i0, i1);
                    const int s = mut_state(mv, c - 1);
                    const bool first_col = (c == startind + 1), last_col = (c == last);
                    int slot = i0 % S;
                    if (s >= 0)
                    {
                        const StateParams sp = md.st[s];
                        double upC = 0, upS = 0;
                        // (c-1, i0-1): still in its slot, nothing of this column overwrites it
                        double diag = 0.0;
                        if (i0 > p0 && i0 <= p1)
                            diag = first_col ? (seed ? seed[row_off(ts, i0 - 1)] : 0.0) : ring[(long long)((i0 - 1) % S) * rstride];
                        // software pipeline: level record, seed value and join values of row i+1 are
                        // requested while row i is computed (the loads are the latency that matters here)
                        const LevelRec* lv = lev + (i0 - 1);                 // row i reads level i-1 ...
                        const LevelRec* lq = lev + (n0 - i0);                // ... and 3 log stdv of level n0-i
                        const double* sd = (first_col && seed) ? seed : nullptr;                // + row_off(ts, row)
                        const bool joinB = last_col && rab > 0;
                        // reverse row jb = n0-i+1 lives at + row_off(ts, jb)
                        LevelRec lr = *lv;
                        double lsd3 = lq->lsd3;
                        double sv = (sd && i0 >= p0 && i0 <= p1) ? sd[row_off(ts, i0)] : 0.0;
                        double bmv = 0.0;
                        {
                            const int jb = n0 - i0 + 1;
                            if (joinB && jb >= b0 && jb < b0 + blen) bmv = Bm[row_off(ts, jb)];
                        }
                        for (int i = i0; i <= i1; i++)
                        {
                            const LevelRec lr_c = lr;
                            const double lsd3_c = lsd3, sv_c = sv, bm_c = bmv;
                            if (i < i1)
                            {
                                lv++; lq--;
                                lr = *lv; lsd3 = lq->lsd3;
                                if (sd) sv = (i + 1 >= p0 && i + 1 <= p1) ? sd[row_off(ts, i + 1)] : 0.0;
                                if (joinB)
                                {
                                    const int jn = n0 - i;
                                    bmv = 0.0;
                                    if (jn >= b0 && jn < b0 + blen) bmv = Bm[row_off(ts, jn)];
                                }
                            }
                            const double e_i = emission(lr_c.mean, lr_c.stdv, lr_c.rstdv, lsd3_c, sp, b.log2pi, b.lik_offset);
                            const bool skip_ok = i >= p0 && i <= p1;
                            const bool diag_ok = i > p0 && i <= p1;
                            double Pi = 0.0;
                            if (skip_ok) Pi = first_col ? sv_c : ring[(long long)slot * rstride];
                            double C, Sv; int step;
                            dp_cell(i == i0, skip_ok, diag_ok, Pi, diag, e_i, e_i, upC, upS, tr, C, Sv, step);
                            best = max_gt(C, best);
                            if (last_col)
                            {
                                const int jb = n0 - i + 1;
                                if (jb >= b0 && jb < b0 + blen) joinmax = max_gt(C + bm_c, joinmax);
                            }
                            else ring[(long long)slot * rstride] = C;
                            diag = Pi;
                            upC = C; upS = Sv;
                            slot = slot + 1 == S ? 0 : slot + 1;
                        }
                    }
                    else
                    {
                        // invalid state: an all-zero column that inherits the running best
                        for (int i = i0; i <= i1; i++)
                        {
                            if (last_col)
                            {
                                const int jb = n0 - i + 1;
                                if (jb >= b0 && jb < b0 + blen)
                                {
                                    joinmax = fmax(joinmax, rab > 0 ? Bm[row_off(ts, jb)] : 0.0);
                                }
                            }
                            else ring[(long long)slot * rstride] = 0.0;
                            slot = slot + 1 == S ? 0 : slot + 1;
                        }
                    }
                    p0 = i0; p1 = i1;
                }
                neu = fmax(fmax(joinmax, 0.0), fmax(best, mb));
            }
            result = neu - old;
        }
        b.delta[LIST ? ev.task_off + m : t] = result;
    }
}

// k_mutscore_warp: the exact FP64 task of k_mutscore with ONE WARP per (mutation, event) pair, for jobs
// with few pairs (FindMutations' candidate lists at scoring_width, the flagged re-scores of the FAST mode):
// a thread-per-pair launch leaves the machine empty there and its run time is the latency of one thread
// walking (|mut| + 6) x (2W + 1) cells.  Lane c owns narrow column startind + 1 + c and the warp runs the
// anti-diagonal wavefront: at step s lane c computes row rmin + s - c, its left neighbour (row i of column
// c-1) is what lane c-1 computed one step earlier and arrives by __shfl_up_sync, the diagonal (row i-1) is
// the value received the step before, the cell above is the lane's own previous result.  Lane 0 reads the
// seed column, the last lane joins with the reverse column; level records, seed and reverse cells are
// requested one step ahead.  Same cell arithmetic (dp_cell, emission) as the thread form: bit-identical.
// Longer replacement strings run as chunks of 32 columns.  With more than b.warp_limit tasks the kernel is a
// no-op and the thread form launched beside it does the work, and vice versa (the flagged count only exists
// on the device).
template <bool LIST>
__global__ void __launch_bounds__(128) k_mutscore_warp(Batch b)
{
    extern __shared__ double wbuf[];                               // [warps][2][S] hand-over strips (see below)
    const int S = 2 * b.scoring_width + 2;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int W = b.scoring_width;
    const long long n_total = LIST ? (long long)(*b.flag_count) * b.max_ev : b.n_tasks;
    if (n_total > b.warp_limit) return;
    for (long long t = gw; t < n_total; t += nwarps)
    {
        int e, m;
        if (LIST ? !list_task(b, t, e, m) : !task_decode(b, t, e, m)) continue;
        const EvDesc ev = b.ev[e];
        const MutDev mu = b.muts[ev.mut_off + m];
        double result = 0.0;
        if (ev.usable && !((unsigned)mu.start > (unsigned)ev.L))        // cpp/MakeMutations.cpp:46
        {
            const int N = ev.N, n0 = ev.n0, L = ev.L;
            MutView mv;
            mv.bases = b.bases + ev.base_off; mv.L = L; mv.mstr = b.mut_str + mu.str_off;
            mv.start = mu.start; mv.n_orig = mu.n_orig; mv.n_mut = mu.n_mut;
            mv.applied = mu.start < L;
            mv.Lm = mv.applied ? mu.start + mu.n_mut + max(0, L - mu.start - mu.n_orig) : L;
            const int Nm = mv.Lm >= 5 ? mv.Lm - 4 : 0;
            const int raf = max(mu.start - 3, 1);
            double old;
            if (raf <= N) old = b.old[ev.col_off + raf];
            else old = thread_join(b, ev, raf, N - raf + 1);
            const int startind = max(mu.start - 4, 0);
            const int refind = mu.start + mu.n_mut + 1;
            int last = min(min(refind, startind + mu.n_mut + 6), Nm);
            if (W == 0) last = startind;
            const int ncol = last - startind;
            double neu;
            if (ncol <= 0) neu = thread_join(b, ev, startind, Nm - startind + 1);     // cpp/Alignment.cpp:483-499
            else
            {
                const LevelRec* lev = b.lev + ev.lev_off;
                const ModelDev& md = b.models[ev.model];
                const Trans tr = {md.lskip, md.lstay, md.lext, md.lins};
                const bool ri_empty = b.ri_empty[e] != 0;
                const long long ts = ev.rs;
                // reverse column joined with the last narrow column
                const int rab = min(max(Nm - last + 1, 0), N);
                const long long gb = ev.col_off + rab;
                int b0 = 0, blen = n0 + 1;
                double mb = 0.0;
                if (rab > 0) { b0 = b.Bi0[gb]; blen = b.Blen[gb]; mb = b.Bbest[gb]; }
                const double* Bm = b.Bm + col_base(ev, rab);
                // the column left of the first chunk: the seed column, or the blank column 0
                int pb0 = 0, pb1 = n0;
                const double* seed = nullptr;
                double best = lane == 0 ? 0.0 : NEG, joinmax = 0.0;
                if (startind > 0)
                {
                    const long long gs = ev.col_off + startind;
                    pb0 = b.Fi0[gs]; pb1 = pb0 + b.Flen[gs] - 1;
                    seed = b.Fm + col_base(ev, startind);
                    if (lane == 0) best = b.Fbest[gs];
                }
                // more than 32 narrow columns: chunks of 32, the last column of a chunk is handed to the next chunk's
                // lane 0 through a per-warp strip of shared memory (two strips, alternating)
                double* bin = wbuf + (size_t)(threadIdx.x >> 5) * 2 * S;
                double* bout = bin + S;
                for (int c0 = 0; c0 < ncol; c0 += 32)
                {
                    const int ncq = min(32, ncol - c0);
                    const bool first_chunk = c0 == 0, final_chunk = c0 + 32 >= ncol;
                    const bool mine = lane < ncq, lastl = final_chunk && lane == ncq - 1, handl = !final_chunk && lane == 31;
                    // this lane's column
                    int i0 = 1 << 29, i1 = -(1 << 29), st = -1;
                    StateParams sp;
                    if (mine)
                    {
                        const int k = startind + 1 + c0 + lane;
                        band_of(ri_empty ? 1 : b.cen_new[ev.cen_off + k], n0, W, i0, i1);
                        st = mut_state(mv, k - 1);
                        sp = md.st[max(st, 0)];
                    }
                    else sp = md.st[0];
                    const bool valid = st >= 0;
                    // band of the column before: the left lane's, or (lane 0) the column left of the chunk
                    int p0 = __shfl_up_sync(0xffffffffu, i0, 1), p1 = __shfl_up_sync(0xffffffffu, i1, 1);
                    if (lane == 0) { p0 = pb0; p1 = pb1; }
                    const bool have_left = !first_chunk || seed != nullptr;      // lane 0: a stored column exists
                    int rmin = mine ? i0 : 1 << 29, rmax = mine ? i1 : -(1 << 29);
                    for (int o = 16; o; o >>= 1)
                    {
                        rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
                        rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
                    }
                    const int nsteps = rmax - rmin + ncq;
                    double upC = 0.0, upS = 0.0, recv = 0.0, recv_prev = 0.0, Cpub = 0.0;
                    // requests of the first step; lane 0 follows the column left of the chunk on every step (its
                    // diagonal is that column's cell of the row before)
                    int i = rmin - lane;
                    bool act = mine && i >= i0 && i <= i1;
                    LevelRec lr = lev[0]; double lsd3 = 0.0, sv = 0.0, bmv = 0.0;
                    if (lane == 0 && have_left)
                    {
                        if (i >= p0 && i <= p1) sv = first_chunk ? seed[row_off(ts, i)] : bin[i - p0];
                        if (i - 1 >= p0 && i - 1 <= p1) recv_prev = first_chunk ? seed[row_off(ts, i - 1)] : bin[i - 1 - p0];
                    }
                    if (act)
                    {
                        lr = lev[i - 1]; lsd3 = lev[n0 - i].lsd3;
                        const int jb = n0 - i + 1;
                        if (lastl && rab > 0 && jb >= b0 && jb < b0 + blen) bmv = Bm[row_off(ts, jb)];
                    }
                    for (int s = 0; s < nsteps; s++)
                    {
                        const bool act_c = act;
                        const int i_c = i;
                        const LevelRec lr_c = lr;
                        const double lsd3_c = lsd3, sv_c = sv, bm_c = bmv;
                        // next step's row: requested now, consumed after this step's cell
                        i = i_c + 1;
                        act = mine && i >= i0 && i <= i1;
                        sv = 0.0; bmv = 0.0;
                        if (lane == 0 && have_left && i >= p0 && i <= p1) sv = first_chunk ? seed[row_off(ts, i)] : bin[i - p0];
                        if (act)
                        {
                            lr = lev[i - 1]; lsd3 = lev[n0 - i].lsd3;
                            const int jb = n0 - i + 1;
                            if (lastl && rab > 0 && jb >= b0 && jb < b0 + blen) bmv = Bm[row_off(ts, jb)];
                        }
                        if (act_c)
                        {
                            double C = 0.0, Sv = 0.0;
                            if (valid)
                            {
                                const double e_i = emission(lr_c.mean, lr_c.stdv, lr_c.rstdv, lsd3_c, sp, b.log2pi, b.lik_offset);
                                const bool skip_ok = i_c >= p0 && i_c <= p1;
                                const bool diag_ok = i_c > p0 && i_c <= p1;
                                const double Pi = skip_ok ? (lane == 0 ? sv_c : recv) : 0.0;
                                const double Pd = diag_ok ? recv_prev : 0.0;
                                int step;
                                dp_cell(i_c == i0, skip_ok, diag_ok, Pi, Pd, e_i, e_i, upC, upS, tr, C, Sv, step);
                                best = max_gt(C, best);
                            }
                            if (lastl)
                            {
                                const int jb = n0 - i_c + 1;
                                if (jb >= b0 && jb < b0 + blen) joinmax = max_gt(C + bm_c, joinmax);
                            }
                            if (handl) bout[i_c - i0] = C;
                            upC = C; upS = Sv;
                            Cpub = C;
                        }
                        // hand this step's cell to the right neighbour; what it held becomes its diagonal
                        recv_prev = lane == 0 ? sv_c : recv;
                        const double got = __shfl_up_sync(0xffffffffu, Cpub, 1);
                        if (lane > 0) recv = got;
                    }
                    // the chunk's last column becomes the column left of the next chunk
                    pb0 = __shfl_sync(0xffffffffu, i0, 31); pb1 = __shfl_sync(0xffffffffu, i1, 31);
                    __syncwarp();
                    double* tmp = bin; bin = bout; bout = tmp;
                }
                // the warp's best cell and the last lane's join
                for (int o = 16; o; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
                joinmax = __shfl_sync(0xffffffffu, joinmax, (ncol - 1) & 31);
                neu = fmax(fmax(joinmax, 0.0), fmax(best, mb));
            }
            result = neu - old;
        }
        if (lane == 0) b.delta[LIST ? ev.task_off + m : t] = result;
    }
}

// k_points: the implicit point-mutation table of a region whose sequence is plain ACGT, in FindPointMutations order
// (cpp/FindMutations.cpp:191-234): per state one deletion, the 3 substitutions by the other bases, 4 insertions.
// Regions with other characters (a fourth substitution where the base is not ACGT) get their table from the host.
__global__ void k_points(Batch b)
{
    const RegTabDev rt = b.regs[blockIdx.y];
    if (!rt.plain_points || rt.nev == 0) return;
    const EvDesc ev = b.ev[rt.ev0];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ev.N) return;
    const char here = b.bases[ev.base_off + i];
    MutDev* out = const_cast<MutDev*>(b.muts) + rt.mut_off + 8LL * i;
    MutDev d;
    d.start = i; d.n_orig = 1; d.n_mut = 0; d.str_off = 0;
    *out++ = d;
    d.n_mut = 1;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if ("ACGT"[j] != here) { d.str_off = j; *out++ = d; }
    d.n_orig = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { d.str_off = j; *out++ = d; }
}

// k_reduce: score[m] = -1e-6 + sum_e delta(e, m), events in order (cpp/MakeMutations.cpp:38-52,
// cpp/AlignUtil.h:84-90).  One thread per mutation; the region table gives its events.
__global__ void k_reduce(Batch b, const RegTabDev* regs, int n_regs, long long n_muts, double start, int from_list,
                         const double* start_arr, double* out)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (from_list)
    {
        if (g >= *b.flag_count) return;
        g = b.flag_list[g];
    }
    else if (g >= n_muts) return;
    int lo = 0, hi = n_regs - 1;
    while (lo < hi)
    {
        int mid = (lo + hi + 1) >> 1;
        if (regs[mid].mut_off <= g) lo = mid; else hi = mid - 1;
    }
    const int m = (int)(g - regs[lo].mut_off);
    // -1e-6 (cpp/AlignUtil.h:86); 0 for a partial sum over an event shard; the running sum of the ranks before this one
    // when the events of the region are split across GPUs in order (ps_comm.cu)
    double s = start_arr ? start_arr[g] : start;
    for (int e = regs[lo].ev0; e < regs[lo].ev0 + regs[lo].nev; e++) s += b.delta[b.ev[e].task_off + m];
    out[g] = s;
}

// event-sharded FAST mode: the exactly re-scored totals of the flagged mutations replace the FP32 ones
__global__ void k_merge_flagged(Batch b, const double* exact, double add)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *b.flag_count) return;
    const int g = b.flag_list[k];
    b.scores[g] = exact[g] + add;
}

__global__ void k_add_scalar(double* v, long long n, double add)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) v[g] += add;
}

// per-event alignment score = running best of the last forward column (cpp/Alignment.h:127-130)
__global__ void k_event_scores(Batch b, double* out)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= b.n_events) return;
    const EvDesc ev = b.ev[e];
    out[e] = (ev.usable && ev.N > 0) ? b.Fbest[ev.col_off + ev.N] : 0.0;
}

} // namespace psdev
