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
//   * a warp owns a BLOCK of 64 consecutive columns, lane c = columns 2c, 2c+1 of the block; at step s it computes row
//     R0 + s - c of both.  The cell left of its first column (same row, last column of lane c-1) is what lane c-1
//     computed one step earlier and arrives by __shfl_up_sync; the diagonal cell is the value received the step before;
//     its second column reads the first one's registers.  One shuffle and one 16-byte level record per TWO cells.
//   * the W warps of a CTA take the blocks of ONE event round robin (warp w: blocks w, w+W, ...).  Block b+1 needs the
//     last column of block b: lane 31 of the producing warp writes it to a strip of shared memory (slot = row -
//     first row of the column's band), lane 0 of the consuming warp reads it ~90 steps later (64 columns + the band's drift further down the
//     anti-diagonal).  The value IS the flag: main-matrix cells are >= 0, a slot holds -1 until it is written and the
//     reader spins on the slot itself -- no progress words, no fences in the sweep.  At the end of its block the reader
//     puts the -1 back into the whole strip and says so in one word, which the warp that re-uses the strip (W blocks
//     later) checks first.
//   * the level records of the event (16 B per level: mean, stdv, 1/stdv, -1.5 log stdv) are read once per step.
//     STAGE = true: the CTA brings ALL of them into shared memory with ONE cp.async.bulk (TMA 1-D bulk copy, completion
//     on an mbarrier with expect_tx) before the sweep -- events up to PS_SCORE32_STAGE_LEVELS levels; a warp's read is
//     then a conflict-free LDS.128 (lanes read consecutive records).  STAGE = false: LDG.128 through L1 (long events).
//   * steps in which all 32 lanes are strictly inside the bands of both their columns and of the column before (3/4 of
//     the steps of a 601-row band) run a predicate-free body; the edges run the same arithmetic under band masks.
//
// Per cell in the interior body: 8 emission ops, 8 adds, 4.5 max (FMNMX3 folds two), half a shuffle, half a 16-byte load.
#pragma once
#include <type_traits>

#include "ps_fast.cuh"

namespace psdev {

constexpr int   PS_SCORE32_MAX_WARPS = 16;            // warps per CTA = blocks of one event in flight (4 for big batches);
                                                      // 32 warps were slower on long events (64 registers, spills): 190 vs 215 GCUPS
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
__device__ __forceinline__ float lds_volatile(unsigned addr)
{
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_f4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_volatile(unsigned addr, float v)
{
    asm volatile("st.volatile.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}

struct Score32Args
{
    const int* list;              // event indices of this launch
    int        count;
    float*     out32;             // per event (indexed by event): best main-matrix cell, floor 0; zeroed before the launch
    int        strip;             // slots per hand-over strip: 2 * realign_width + 1 rows of a band, padded
};

// one cell of the recurrence in FP32 (floor 0); out-of-band predecessors already replaced by 0
__device__ __forceinline__ void cell32(float left, float diag, float e, float uC, float uS, float s0, const float4 tr,
                                       float& C, float& S)
{
    const float skip = left + tr.x;
    const float match = diag + e;
    const float insign = fmaxf(uC, diag) + tr.w;          // max(insert, ignore): both add log p_insert
    const float se = fmaxf(uC + tr.y, uS + tr.z) + e;     // max(stay, extend): both add the emission
    S = fmaxf(s0, se);
    C = fmaxf(fmaxf(fmaxf(0.f, skip), match), fmaxf(insign, S));
}

__device__ __forceinline__ float emis32(const float4 l, const StateParamsF& p)
{
    const float d1 = l.x - p.mu, d2 = l.y - p.mu2;
    return __fmaf_rn(p.a_s, d1 * d1, p.c_s) + __fmaf_rn(p.f_s * l.z, d2 * d2, l.w);
}

template <bool STAGE, bool INV, int MAXT>
__global__ void __launch_bounds__(MAXT) k_score_f32(Batch b, Score32Args a)
{
    extern __shared__ __align__(16) unsigned char s32_smem[];
    const int W = blockDim.x >> 5;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    __shared__ int rdone[PS_SCORE32_MAX_WARPS];            // strip q: last block whose hand-over through it was read to the end
    __shared__ unsigned long long stage_bar;
    float* strips = reinterpret_cast<float*>(s32_smem);                   // [W + 1][strip]; strip W = the blank column 0
    for (int q = threadIdx.x; q < (W + 1) * a.strip; q += blockDim.x) strips[q] = q < W * a.strip ? -1.f : 0.f;
    if (threadIdx.x < W) rdone[threadIdx.x] = -1;
    if (STAGE && threadIdx.x == 0)
    {
        mbar_init(smem_u32(&stage_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int rw = b.realign_width;
    // A CTA takes the events list[blockIdx.x], list[blockIdx.x + gridDim.x], ... one after the other WITHOUT a barrier
    // between them: its warps share out the blocks of the whole sequence round robin (block G of the sequence goes to
    // warp G mod W), so a warp that has finished its last block of one event starts on the next event while the others
    // are still busy -- the pipeline of blocks is filled once per CTA, not once per event.  (STAGE: one event per CTA,
    // the grid is the list.)
    int gbase = 0;                                                       // blocks of the events this CTA has been through
    for (int q = blockIdx.x; q < a.count; q += gridDim.x)
    {
    const int e = a.list[q];
    const EvDesc ev = b.ev[e];
    const int N = ev.N, n0 = ev.n0;
    if (!ev.usable || N <= 0) continue;                                  // its score stays 0
    const LevelRecF* glev = b.levf + ev.lev_off;
    const float4* lev = reinterpret_cast<const float4*>(glev);           // row i at lev[i - 1]
    if (STAGE)
    {
        float4* slev = reinterpret_cast<float4*>(s32_smem + (size_t)(W + 1) * a.strip * sizeof(float));
        const unsigned bar = smem_u32(&stage_bar);
        if (threadIdx.x == 0)
        {
            const unsigned bytes = (unsigned)n0 * (unsigned)sizeof(LevelRecF);
            mbar_expect_tx(bar, bytes);
            bulk_load(slev, glev, bytes, bar);
        }
        lev = slev;
        mbar_wait(bar, 0);
    }
    const StateParamsF* stf = b.stf + (size_t)ev.model * N_STATES;
    const float4 tr = b.trf[ev.model];
    const int* cen = b.cen_old + ev.cen_off;
    const int* states = b.states + ev.state_off;
    const int nblocks = (N + 63) >> 6;
    float best = 0.f;
    bool any = false;

    for (int blk = ((wrp - gbase) % W + W) % W; blk < nblocks; blk += W)
    {
        const int G = gbase + blk;                                       // the block's number in the CTA's sequence
        any = true;

        // this lane's two columns
        const int kA = (blk << 6) + 2 * lane + 1, kB = kA + 1;
        const bool mineA = kA <= N, mineB = kB <= N;
        int i0a = 1 << 28, i1a = -(1 << 28), i0b = 1 << 28, i1b = -(1 << 28), sa = 0, sb = 0;
        if (mineA) { band_of(cen[kA], n0, rw, i0a, i1a); sa = states[kA - 1]; }
        if (mineB) { band_of(cen[kB], n0, rw, i0b, i1b); sb = states[kB - 1]; }
        const bool validA = !INV || sa >= 0, validB = !INV || sb >= 0;
        const StateParamsF spA = stf[max(sa, 0)], spB = stf[max(sb, 0)];
        // band of the column before the lane's first one: the left lane's second column; lane 0: the last column of the
        // previous block (its strip), or the blank column 0 (rows 0..n0, all zeros, cpp/Alignment.cpp:42)
        int p0 = __shfl_up_sync(0xffffffffu, i0b, 1), p1 = __shfl_up_sync(0xffffffffu, i1b, 1);
        const bool from_strip = blk > 0;
        if (lane == 0)
        {
            if (!from_strip) { p0 = 0; p1 = n0; }
            else band_of(cen[kA - 1], n0, rw, p0, p1);
        }
        const int l0p0 = __shfl_sync(0xffffffffu, p0, 0), l0p1 = __shfl_sync(0xffffffffu, p1, 0);
        // strips are indexed by (row - first band row of their column); the blank column 0 is one zero that never moves
        const unsigned sin = smem_u32(strips + (size_t)(from_strip ? (G + W - 1) % W : W) * a.strip);
        const unsigned sout = smem_u32(strips + (size_t)(G % W) * a.strip);
        const unsigned sin_step = from_strip ? 4u : 0u;
        const int nslot = a.strip - 1;
        const bool hand = blk + 1 < nblocks;                             // (then the block is full: lane 31's second column is its last)
        const bool handl = hand && lane == 31;
        // steps: lane c computes row R0 + s - c
        int R0 = mineA ? min(i0a, i0b) + lane : 1 << 28, Rend = mineA ? max(i1a, i1b) + lane : -(1 << 28);
        // interior steps: the lane strictly inside both bands and the previous column's, past every first row
        int s_lo = max(max(i0a, i0b), p0) + 1 + lane, s_hi = min(min(i1a, i1b), p1) + lane;
        for (int o = 16; o; o >>= 1)
        {
            R0 = min(R0, __shfl_xor_sync(0xffffffffu, R0, o));
            Rend = max(Rend, __shfl_xor_sync(0xffffffffu, Rend, o));
            s_lo = max(s_lo, __shfl_xor_sync(0xffffffffu, s_lo, o));
            s_hi = min(s_hi, __shfl_xor_sync(0xffffffffu, s_hi, o));
        }
        const int nsteps = Rend - R0 + 1;
        s_lo -= R0; s_hi -= R0;                                          // interior steps [s_lo, s_hi]
        if (s_lo > s_hi) { s_lo = nsteps; s_hi = nsteps - 1; }           // (a partial last block has none: some lane owns no column)
        // the output strip was last used by block blk - W; wait for its reader (the warp of block blk - W + 1, which
        // finished about a block's worth of steps ago in any regular schedule) to say so
        if (hand && G >= W)
            while (*(volatile int*)&rdone[G % W] < G - W) __nanosleep(64);
        float upCA = 0.f, upSA = 0.f, upCB = 0.f, upSB = 0.f, recv = 0.f, recv_prev = 0.f, pub = 0.f;
        int i = R0 - lane;                                               // this lane's row of step 0
        int il0 = R0;                                                    // lane 0's row
        auto in_addr = [&](int row) { return sin + (from_strip ? (unsigned)min(max(row - l0p0, 0), nslot) << 2 : 0u); };
        float4 lr = lev[min(max(i, 1), n0) - 1];                         // level record of the coming step
        float sv_next = lds_volatile(in_addr(il0));                      // strip value of lane 0's coming row (-1: not there yet)

        // Edge steps.  PH_PRE (before the interior steps of a regular block): lanes enter their bands, nobody has reached
        // the end of one -- only the lower band limits and the first rows are tested.  PH_POST (after them): lanes leave
        // their bands -- only the upper limits.  PH_ANY: every test (blocks without interior steps).
        enum { PH_ANY = 0, PH_PRE = 1, PH_POST = 2 };
        auto edge_steps = [&](auto phase, int s_from, int s_to) {
            constexpr int PH = decltype(phase)::value;
            constexpr bool LO = PH != PH_POST, HI = PH != PH_PRE;
            for (int s = s_from; s <= s_to; s++)
            {
                const float4 lr_c = lr;
                const int i_c = i;
                i = i_c + 1;
                lr = lev[min(max(i, 1), n0) - 1];
                // lane 0's left neighbour comes from the strip: wait until the producer has written the row
                float sv = sv_next;
                if ((!LO || il0 >= l0p0) && (!HI || il0 <= l0p1))
                    while (__any_sync(0xffffffffu, sv < 0.f)) { __nanosleep(20); sv = lds_volatile(in_addr(il0)); }
                il0++;
                sv_next = lds_volatile(in_addr(il0));
                const float left = lane == 0 ? sv : recv;
                // first column
                const bool actA = (!LO || i_c >= i0a) && (!HI || i_c <= i1a);
                const bool firstA = LO && i_c == i0a;
                const bool skA = (!LO || i_c >= p0) && (!HI || i_c <= p1), dgA = (!LO || i_c > p0) && (!HI || i_c <= p1);
                const float emA = emis32(lr_c, spA), emB = emis32(lr_c, spB);
                float CA, SA;
                cell32(skA ? left : 0.f, dgA ? recv_prev : 0.f, emA,
                       firstA ? -S32_BIG : upCA, firstA ? -S32_BIG : upSA, firstA ? -S32_BIG : 0.f, tr, CA, SA);
                if (INV && !validA) { CA = 0.f; SA = 0.f; }               // cpp/Alignment.cpp:162: the column stays all zero
                const float prevCA = upCA;
                if (actA) { best = fmaxf(best, CA); upCA = CA; upSA = SA; }
                // second column: its left neighbour is the first one
                const bool actB = (!LO || i_c >= i0b) && (!HI || i_c <= i1b);
                const bool firstB = LO && i_c == i0b;
                const bool dgB = (!LO || i_c > i0a) && (!HI || i_c <= i1a);
                float CB, SB;
                cell32(actA ? CA : 0.f, dgB ? prevCA : 0.f, emB,
                       firstB ? -S32_BIG : upCB, firstB ? -S32_BIG : upSB, firstB ? -S32_BIG : 0.f, tr, CB, SB);
                if (INV && !validB) { CB = 0.f; SB = 0.f; }
                if (actB)
                {
                    best = fmaxf(best, CB); upCB = CB; upSB = SB; pub = CB;
                    if (handl) sts_volatile(sout + ((unsigned)(i_c - i0b) << 2), CB);
                }
                recv_prev = left;
                recv = __shfl_up_sync(0xffffffffu, pub, 1);
            }
        };
        const bool regular = s_lo <= s_hi;
        if (regular) edge_steps(std::integral_constant<int, PH_PRE>(), 0, s_lo - 1);
        else edge_steps(std::integral_constant<int, PH_ANY>(), 0, nsteps - 1);
        // interior: no masks, no first rows, every lane active in both columns.  Running addresses: the level record of
        // the coming step (16 bytes further per step), lane 0's strip slot and lane 31's output slot (4 bytes further)
        if (regular)
        {
            unsigned ia = in_addr(il0);                                  // slot of lane 0's row of the coming step (sv_next)
            unsigned oa = sout + ((unsigned)(handl ? i - i0b : 0) << 2); // slot of this lane's row of the coming step
            unsigned lev_sh = STAGE ? smem_u32(lev) + ((unsigned)i << 4) : 0u;   // record of row i + 1
            const float4* lev_gl = lev + i;
            // one step: `use` holds the level record of this step, `load` receives the next one (two copies of the body
            // with the roles swapped instead of a register rotation)
            auto step = [&](const float4& use, float4& load, const float sv_in, float& sv_out) {
                if (STAGE) { load = lds_f4(lev_sh); lev_sh += 16; }      // (row n0 + 1 at most: the records are padded by one)
                else load = *lev_gl++;
                float sv = sv_in;
                // the wait is uniform by construction (every lane looks at lane 0's slot): vote, so that the branch is too
                while (__any_sync(0xffffffffu, sv < 0.f)) { __nanosleep(20); sv = lds_volatile(ia); }
                ia += sin_step;
                sv_out = lds_volatile(ia);
                const float left = lane == 0 ? sv : recv;
                const float emA = emis32(use, spA), emB = emis32(use, spB);
                float CA, SA, CB, SB;
                cell32(left, recv_prev, emA, upCA, upSA, 0.f, tr, CA, SA);
                if (INV && !validA) { CA = 0.f; SA = 0.f; }
                cell32(CA, upCA, emB, upCB, upSB, 0.f, tr, CB, SB);
                if (INV && !validB) { CB = 0.f; SB = 0.f; }
                best = fmaxf(best, fmaxf(CA, CB));
                upCA = CA; upSA = SA; upCB = CB; upSB = SB;
                if (handl) sts_volatile(oa, CB);
                oa += 4;
                recv_prev = left;
                recv = __shfl_up_sync(0xffffffffu, CB, 1);
                pub = CB;
            };
            float4 lr2;
            float sv2;
            int s = s_lo;
            for (; s + 1 <= s_hi; s += 2)
            {
                step(lr, lr2, sv_next, sv2);
                step(lr2, lr, sv2, sv_next);
            }
            if (s <= s_hi)
            {
                step(lr, lr2, sv_next, sv2);
                lr = lr2; sv_next = sv2;
            }
            const int done = s_hi - s_lo + 1;
            i += done; il0 += done;
        }
        if (regular) edge_steps(std::integral_constant<int, PH_POST>(), s_hi + 1, nsteps - 1);
        if (from_strip)
        {
            // the input strip goes back to "nothing written" (every row of its column's band, read or not), then it is
            // free for the block W further on
            for (int q = lane; q <= l0p1 - l0p0; q += 32) sts_volatile(sin + ((unsigned)q << 2), -1.f);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *(volatile int*)&rdone[(G + W - 1) % W] = G - 1;
        }
        if (!hand && lane == 0) *(volatile int*)&rdone[G % W] = G;       // nobody reads the last block of an event
        __syncwarp();
    }
    if (any)
    {
        for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
        // scores are >= 0: their order as floats is their order as ints
        if (lane == 0) atomicMax(reinterpret_cast<int*>(a.out32) + e, __float_as_int(best));
    }
    gbase += nblocks;
    }
}

// the per-event maxima as doubles (the C-ABI's type)
__global__ void k_score_f32_out(const float* in, double* out, int n)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = (double)in[e];
}

} // namespace psdev
