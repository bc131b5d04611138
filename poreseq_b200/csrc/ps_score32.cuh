// ps_score32.cuh -- k_score_f32: the score-only forward fill of PSAlign.ScoreEvents in log-space FP32
// (cpp/MakeMutations.cpp:148-195 called with likes == NULL, poreseq/_poreseqcpp.pyx:263-276: E doubles out, the
// realignment is NOT propagated -- the one entry point of the path whose whole contract is "scores within 1e-4").
//
// Nothing is stored: no band matrices, no step bytes, no backtrace.  The recurrence is cpp/Alignment.cpp:111-274
//  cell for cell -- same bands (centre = lower_bound(ref_index, c), +- realign_width), same implicit
// zeros outside the previous column's band, same first-row rule, same floor -- with the fused 8-op FP32 emission of
// ps_fast.cuh.  What comes out is max over all main-matrix cells (= the last column's running best, Alignment.h:127).
//
// Mapping (the shape BASELINE.json's north_star describes: a warp runs an anti-diagonal wavefront over the band and
// hands scores along by __shfl_sync):
//   * a warp owns a BLOCK of 32 consecutive columns, lane = column; at step s lane c computes row R0 + s - c.  The cell
//     left of it (same row, column c-1) is what lane c-1 computed one step earlier and arrives by __shfl_up_sync; the
//     diagonal cell is the value received the step before; the cell above is the lane's own previous result.  One
//     shuffle per cell, everything else is registers.
//   * the W warps of a CTA take the blocks of ONE event round robin (warp w: blocks w, w+W, ...).  Block b+1 needs the
//     last column of block b: lane 31 of the producing warp writes it to a strip of shared memory (slot = row mod
//     STRIP), lane 0 of the consuming warp reads it ~62 steps later (32 columns + ~30 rows of band drift further down
//     the anti-diagonal); a progress word per strip carries "rows final up to", updated every 8 rows.
//   * the level records of the event (16 B per level: mean, stdv, 1/stdv, -1.5 log stdv) are read once per cell.
//     STAGE = true: the CTA brings ALL of them into shared memory with ONE cp.async.bulk (TMA 1-D bulk copy, completion
//     on an mbarrier with expect_tx) before the sweep -- events up to PS_SCORE32_STAGE_LEVELS levels; a lane's read is
//     then a conflict-free LDS.128 (lanes read consecutive records).  STAGE = false: LDG.128 through L1 (long events).
//   * steps in which all 32 lanes are strictly inside their own band and their left neighbour's (82 % of the steps of
//     a 601-row band) run a predicate-free body; the edges run the same arithmetic under band masks.
//
// Per cell in the interior body: 8 emission ops, 8 adds, 5 max (FMNMX3 folds two), one shuffle, one 16-byte load.
#pragma once
#include "ps_fast.cuh"

namespace psdev {

constexpr int   PS_SCORE32_WARPS = 4;                 // warps per CTA = blocks of one event in flight
constexpr int   PS_SCORE32_STAGE_LEVELS = 2048;       // events up to this many levels are staged whole (32 KB)
constexpr float S32_BIG = 1.0e30f;                    // "never wins" (cpp/AlignUtil.h:20 uses 1e300)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}

struct Score32Args
{
    const int* list;              // event indices (usable ones first), grid.x indexes it
    double*    out;               // per event: best main-matrix cell, floor 0
    int        strip;             // slots per hand-over strip (power of two >= 2 * realign_width + 64)
};

// one cell of the recurrence in FP32; fl = floor (0), out-of-band predecessors already replaced by the floor
__device__ __forceinline__ void cell32(float left, float diag, float e, float uC, float uS, float s0, const float4 tr,
                                       float& C, float& S)
{
    const float skip = left + tr.x;
    const float match = diag + e;
    const float ign = diag + tr.w;
    const float stay = (uC + e) + tr.y;
    const float ins = uC + tr.w;
    const float ext = (uS + e) + tr.z;
    S = fmaxf(s0, fmaxf(stay, ext));
    C = fmaxf(fmaxf(fmaxf(0.f, skip), match), fmaxf(fmaxf(ins, ign), S));
}

template <bool STAGE, bool INV>
__global__ void __launch_bounds__(32 * PS_SCORE32_WARPS) k_score_f32(Batch b, Score32Args a)
{
    extern __shared__ __align__(16) unsigned char s32_smem[];
    constexpr int W = PS_SCORE32_WARPS;
    const int e = a.list[blockIdx.x];
    const EvDesc ev = b.ev[e];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int N = ev.N, n0 = ev.n0;
    __shared__ float warp_best[W];
    // strip q (the last column of blocks q, q+W, ...): (block << 32) | "rows final up to" -- one 8-byte word, so a reader
    // can never pair a new block number with an old row count; rdone[q]: last block of strip q that was read to its end
    __shared__ unsigned long long prog[W];
    __shared__ int rdone[W];
    __shared__ unsigned long long stage_bar;
    if (!ev.usable || N <= 0)
    {
        if (threadIdx.x == 0) a.out[e] = 0.0;
        return;
    }
    const int SM = a.strip - 1;
    float* strips = reinterpret_cast<float*>(s32_smem);                   // [W][strip]
    const LevelRecF* glev = b.levf + ev.lev_off;
    const float4* lev = reinterpret_cast<const float4*>(glev);           // row i at lev[i - 1]
    if (STAGE)
    {
        float4* slev = reinterpret_cast<float4*>(s32_smem + (size_t)W * a.strip * sizeof(float));
        const unsigned bar = smem_u32(&stage_bar);
        if (threadIdx.x == 0)
        {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const unsigned bytes = (unsigned)n0 * (unsigned)sizeof(LevelRecF);
            mbar_expect_tx(bar, bytes);
            bulk_load(slev, glev, bytes, bar);
        }
        lev = slev;
    }
    if (threadIdx.x < W) { prog[threadIdx.x] = 0xffffffff00000000ull; rdone[threadIdx.x] = -1; }
    __syncthreads();
    if (STAGE) mbar_wait(smem_u32(&stage_bar), 0);
    const StateParamsF* stf = b.stf + (size_t)ev.model * N_STATES;
    const float4 tr = b.trf[ev.model];
    const int* cen = b.cen_old + ev.cen_off;
    const int* states = b.states + ev.state_off;
    const int rw = b.realign_width;
    const int nblocks = (N + 31) >> 5;
    float best = 0.f;

    for (int blk = wrp; blk < nblocks; blk += W)
    {
        // this lane's column
        const int k = (blk << 5) + lane + 1;
        const bool mine = k <= N;
        int i0 = 1 << 28, i1 = -(1 << 28), st = 0;
        if (mine) { band_of(cen[k], n0, rw, i0, i1); st = states[k - 1]; }
        const bool valid = !INV || st >= 0;
        const StateParamsF sp = stf[max(st, 0)];
        // band of the column before: the left lane's; lane 0: the last column of the previous block (its strip), or the
        // blank column 0 (rows 0..n0, all zeros, cpp/Alignment.cpp:42)
        int p0 = __shfl_up_sync(0xffffffffu, i0, 1), p1 = __shfl_up_sync(0xffffffffu, i1, 1);
        if (lane == 0)
        {
            if (blk == 0) { p0 = 0; p1 = n0; }
            else band_of(cen[k - 1], n0, rw, p0, p1);
        }
        const bool from_strip = blk > 0;
        const float* sin = strips + (size_t)((blk + W - 1) % W) * a.strip;
        float* sout = strips + (size_t)(blk % W) * a.strip;
        volatile unsigned long long* pin = &prog[(blk + W - 1) % W];
        volatile unsigned long long* pout = &prog[blk % W];
        const unsigned long long tag = (unsigned long long)(unsigned)blk << 32;
        const int lastl = min(31, N - (blk << 5) - 1);                   // lane of the block's last column
        const bool handl = lane == lastl && blk + 1 < nblocks;
        // steps: lane c computes row R0 + s - c
        int R0 = mine ? i0 + lane : 1 << 28, Rend = mine ? i1 + lane : -(1 << 28);
        // interior steps: every lane strictly inside its band and the previous column's (and not on a first row)
        int s_lo = mine ? max(i0, p0) + 1 + lane : 1 << 28, s_hi = mine ? min(i1, p1) + lane : -(1 << 28);
        for (int o = 16; o; o >>= 1)
        {
            R0 = min(R0, __shfl_xor_sync(0xffffffffu, R0, o));
            Rend = max(Rend, __shfl_xor_sync(0xffffffffu, Rend, o));
            s_lo = max(s_lo, __shfl_xor_sync(0xffffffffu, s_lo, o));
            s_hi = min(s_hi, __shfl_xor_sync(0xffffffffu, s_hi, o));
        }
        const int nsteps = Rend - R0 + 1;
        s_lo -= R0; s_hi -= R0;                                          // interior steps [s_lo, s_hi]
        if (lastl < 31 || s_lo > s_hi) { s_lo = nsteps; s_hi = nsteps - 1; }   // partial block: general body throughout
        // the output strip was last used by block blk - W; its reader (the warp of block blk - W + 1) finished a block's
        // worth of steps ago in any regular schedule -- wait for its word anyway, then open the strip for this block
        if (blk >= W)
            while (*(volatile int*)&rdone[blk % W] < blk - W) __nanosleep(64);
        __syncwarp();
        if (handl) *pout = tag | (unsigned)(i0 - 1);                     // rows above the band are final (never read)
        float upC = 0.f, upS = 0.f, recv = 0.f, recv_prev = 0.f, Cpub = 0.f;
        int known = from_strip ? -1 : 1 << 30;                            // progress of the input strip as last seen
        int i = R0 - lane;                                               // row of step 0
        // level record of the row of the coming step (clamped while the lane is outside the event)
        float4 lr = lev[min(max(i, 1), n0) - 1];
        // strip value of lane 0 for the coming step
        auto strip_wait = [&](int row) {
            // lane 0 needs rows <= min(row, p1) of the input strip final; uniform loop (every lane sees lane 0's need)
            const int need = __shfl_sync(0xffffffffu, min(row, p1), 0);
            while (known < need)
            {
                const unsigned long long v = *pin;
                known = (int)(v >> 32) == blk - 1 ? (int)(unsigned)v : -1;
                if (known < need) __nanosleep(32);
            }
            __threadfence_block();
        };

        auto general_steps = [&](int sa, int sb) {
            for (int s = sa; s <= sb; s++)
            {
                if (from_strip && (s & 7) == 0) strip_wait(i + 7);
                const float4 lr_c = lr;
                const int i_c = i;
                i = i_c + 1;
                lr = lev[min(max(i, 1), n0) - 1];
                // left neighbour: previous step's result of lane c-1; lane 0 reads the strip (or the blank column)
                float left = recv;
                if (lane == 0) left = (from_strip && i_c >= p0 && i_c <= p1) ? sin[i_c & SM] : 0.f;
                const bool act = mine && i_c >= i0 && i_c <= i1;
                const bool skip_ok = i_c >= p0 && i_c <= p1, diag_ok = i_c > p0 && i_c <= p1;
                const bool first = i_c == i0;
                LevelRecF l4; l4.x = lr_c.x; l4.y = lr_c.y; l4.ry = lr_c.z; l4.ey = lr_c.w;
                const float em = emission_f(l4, sp);
                float C, S;
                cell32(skip_ok ? left : 0.f, diag_ok ? recv_prev : 0.f, em, first ? -S32_BIG : upC, first ? -S32_BIG : upS,
                       first ? -S32_BIG : 0.f, tr, C, S);
                if (INV && !valid) { C = 0.f; S = 0.f; }                  // cpp/Alignment.cpp:162: the column stays all zero
                if (act)
                {
                    best = fmaxf(best, C);
                    upC = C; upS = S; Cpub = C;
                    if (handl) sout[i_c & SM] = C;
                }
                if (handl && ((s & 7) == 7) && i_c >= i0)
                {
                    __threadfence_block();
                    *pout = tag | (unsigned)min(i_c, i1);
                }
                recv_prev = left;
                const float got = __shfl_up_sync(0xffffffffu, Cpub, 1);
                recv = got;
            }
        };

        general_steps(0, min(s_lo, nsteps) - 1);
        // interior: no masks, no first rows, every lane active
        for (int s = s_lo; s <= s_hi; s++)
        {
            if (from_strip && (s & 7) == 0) strip_wait(i + 7);
            const float4 lr_c = lr;
            const int i_c = i;
            i = i_c + 1;
            lr = lev[i - 1];                                             // (row n0 + 1 at most: the records are padded by one)
            float left = recv;
            if (lane == 0) left = from_strip ? sin[i_c & SM] : 0.f;
            LevelRecF l4; l4.x = lr_c.x; l4.y = lr_c.y; l4.ry = lr_c.z; l4.ey = lr_c.w;
            const float em = emission_f(l4, sp);
            float C, S;
            cell32(left, recv_prev, em, upC, upS, 0.f, tr, C, S);
            if (INV && !valid) { C = 0.f; S = 0.f; }
            best = fmaxf(best, C);
            upC = C; upS = S;
            if (handl) sout[i_c & SM] = C;
            if (handl && ((s & 7) == 7))
            {
                __threadfence_block();
                *pout = tag | (unsigned)i_c;
            }
            recv_prev = left;
            recv = __shfl_up_sync(0xffffffffu, C, 1);
            Cpub = C;
        }
        if (s_lo <= s_hi) general_steps(s_hi + 1, nsteps - 1);
        if (handl)
        {
            __threadfence_block();
            *pout = tag | 0x7fffffffu;                                   // the whole column is final
        }
        if (from_strip && lane == 0) *(volatile int*)&rdone[(blk + W - 1) % W] = blk - 1;   // the input strip is free again
        __syncwarp();
    }
    for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) warp_best[wrp] = best;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        float m = 0.f;
        for (int q = 0; q < W; q++) m = fmaxf(m, warp_best[q]);
        a.out[e] = (double)m;
    }
}

} // namespace psdev
