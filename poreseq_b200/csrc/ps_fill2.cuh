// ps_fill2.cuh -- k_fill2: the exact FP64 wide-band fill (cpp/Alignment.cpp:111-444) on the warp-block schedule of
// k_score_f32 instead of the CTA-wide wavefront of k_fill.  Same cells, same arithmetic (cell_pre / cell_fin /
// emission of ps_device.cuh), same band storage (wavefront-major 2x2 tiles, so every reader is unchanged); what differs
// is how the cells are handed between threads:
//   * a warp owns a BLOCK of 32 strips (64 columns), lane c = strip 32 blk + c; at step s it computes the 2x2 tile of row
//     pair P0 + s - c.  The left neighbour's second column (rows of the pair and the row above) arrives by
//     __shfl_up_sync; no shared-memory ring, no mbarrier, no waiting for a neighbouring warp inside a block.
//   * the W warps of a CTA take the blocks of one (event, direction) round robin; block b+1 gets the last column of block
//     b through a strip of shared memory whose main-matrix values are their own flags (cells are >= 0, -1 = not written).
//   * steps in which every lane's tile is interior run the body with all masks compile-time true.
//   * a CTA walks through its (event, direction) items without a barrier; the running best over columns (Fbest / Fbi /
//     Fbj) is a separate small kernel (k_fill_best).
// The tile of (strip j, pair r) is stored at band_off + (j + r) * rs + (j mod ts) * 4: j + r is the same for all lanes of
// a step, so a warp's stores are one contiguous run, as in k_fill.
#pragma once
#include <type_traits>

#include "ps_device.cuh"

namespace psdev {

constexpr int F2_MAX_WARPS = 4;

struct Fill2Args
{
    const int* list;              // events of this launch (wavefront-capable ones)
    int        count;             // number of events; items = count * dirs
    int        dirs;              // 1: forward only, 2: forward + reverse
    int        slots;             // rows per hand-over strip (2 * realign_width + 1, padded)
};

__device__ __forceinline__ double2 lds_d2_volatile(unsigned addr)
{
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_d2_volatile(unsigned addr, double a, double b)
{
    asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ unsigned f2_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <bool REV, bool INV>
__device__ __forceinline__ void fill2_block(const Batch& b, const EvDesc& ev, const FillOut& o, int blk, int nblocks, int G, int W,
                                            double2* strips, int slots, volatile int* rdone)
{
    const int lane = threadIdx.x & 31;
    const int n0 = ev.n0, N = ev.N;
    const int J = (N + CW - 1) / CW;
    const StripRec* srec = b.strips + ev.strip_off + (REV ? J + 1 : 0);
    const LevelRec* rows = (REV ? b.rowB : b.rowF) + ev.lev_off;
    const ModelDev& md = b.models[ev.model];
    const Trans tr = {md.lskip, md.lstay, md.lext, md.lins};
    const double off = b.lik_offset, l2p = b.log2pi;
    // this lane's strip
    const int j = (blk << 5) + lane;
    const bool mine = j < J;
    const StripRec& rec = srec[min(j, J)];                   // strip J is the sentinel (empty bands)
    StateParams p0 = rec.p[0], p1 = rec.p[1];
    StripRegs cur;
    cur.j = j; cur.rlo = rec.rlo; cur.rhi = rec.rhi; cur.pp0 = rec.pp0; cur.pp1 = rec.pp1;
    cur.i0a = rec.i0[0]; cur.i1a = rec.i1[0]; cur.sa = rec.s[0];
    cur.i0b = rec.i0[1]; cur.i1b = rec.i1[1]; cur.sb = rec.s[1];
    cur.slot4 = rec.slot4;
    const bool from_strip = blk > 0;
    const bool hand = blk + 1 < nblocks;
    const bool handl = hand && lane == 31;
    const int l0pp0 = __shfl_sync(0xffffffffu, cur.pp0, 0), l0pp1 = __shfl_sync(0xffffffffu, cur.pp1, 0);
    // strips: slot = row - first band row of the strip's column; entry = (main value, emission of that cell [reverse])
    const unsigned sin = f2_smem_u32(strips + (size_t)(from_strip ? (G + W - 1) % W : W) * slots);
    const unsigned sout = f2_smem_u32(strips + (size_t)(G % W) * slots);
    // steps: lane c computes row pair P0 + s - c
    int P0 = mine && cur.rlo <= cur.rhi ? cur.rlo + lane : 1 << 28, Pend = mine && cur.rlo <= cur.rhi ? cur.rhi + lane : -(1 << 28);
    // interior pairs of this lane: 2r+1 > max(i0a, i0b, pp0) and 2r+2 <= min(i1a, i1b, pp1), valid states
    int s_lo = 1 << 28, s_hi = -(1 << 28);
    if (mine && (!INV || (cur.sa >= 0 && cur.sb >= 0)) && cur.i0a <= cur.i1a && cur.i0b <= cur.i1b)
    {
        const int X = max(max(cur.i0a, cur.i0b), cur.pp0), Y = min(min(cur.i1a, cur.i1b), cur.pp1);
        s_lo = ((X + 1) >> 1) + lane; s_hi = (Y >> 1) - 1 + lane;
    }
    for (int q = 16; q; q >>= 1)
    {
        P0 = min(P0, __shfl_xor_sync(0xffffffffu, P0, q));
        Pend = max(Pend, __shfl_xor_sync(0xffffffffu, Pend, q));
        s_lo = max(s_lo, __shfl_xor_sync(0xffffffffu, s_lo, q));
        s_hi = min(s_hi, __shfl_xor_sync(0xffffffffu, s_hi, q));
    }
    const int nsteps = Pend - P0 + 1;
    s_lo -= P0; s_hi -= P0;
    const bool regular = s_lo <= s_hi && s_lo >= 0 && s_hi < nsteps;
    if (!regular) { s_lo = nsteps; s_hi = nsteps - 1; }
    // the output strip was last used by block G - W: wait for its reader to have let go of it
    if (hand && G >= W)
        while (rdone[G % W] < G - W) __nanosleep(64);
    double upC0 = NEG, upS0 = NEG, upE0 = 0, upC1 = NEG, upS1 = NEG, upE1 = 0;
    double best0 = NEG, best1 = NEG;
    int besti0 = 0, besti1 = 0;
    // what the left lane produced: second-column main values (and emissions) of its last two rows
    double rLb = 0, rLEb = 0;                                // row 2r (the row above this step's pair): diagonal of row ia
    double pubC_a = 0, pubC_b = 0, pubE_a = 0, pubE_b = 0;   // this lane's second column of the step just computed
    int r = P0 - lane;
    RowRecs rr;
    load_rows(rows, n0, r, rr);

    auto in_slot = [&](int row) { return sin + (from_strip ? (unsigned)min(max(row - l0pp0, 0), slots - 1) << 4 : 0u); };
    // lane 0's diagonal of its very first row: the strip's row above the first pair (later ones come with the steps)
    if (from_strip)
    {
        const int row = 2 * P0;
        double2 v = lds_d2_volatile(in_slot(row));
        if (row >= l0pp0 && row <= l0pp1)
            while (__any_sync(0xffffffffu, v.x < 0.0)) { __nanosleep(20); v = lds_d2_volatile(in_slot(row)); }
        if (lane == 0) { rLb = v.x; rLEb = v.y; }
    }

    auto step = [&](auto lean_tag) {
        constexpr bool LEAN = decltype(lean_tag)::value;
        const RowRecs rc = rr;
        const int r_c = r;
        r = r_c + 1;
        load_rows(rows, n0, r, rr);                          // row records of the next step
        const int ia = 2 * r_c + 1, ib = ia + 1;
        const bool act = mine && r_c >= cur.rlo && r_c <= cur.rhi;
        // lane 0's left neighbour: the strip of the previous block (rows ia, ib; the row above came with the step before)
        const int l0a = 2 * (r_c + lane) + 1;                // lane 0's row ia (the same number in every lane)
        double2 sa = lds_d2_volatile(in_slot(l0a)), sb = lds_d2_volatile(in_slot(l0a + 1));
        if (from_strip)
        {
            const bool needA = LEAN || (l0a >= l0pp0 && l0a <= l0pp1), needB = LEAN || (l0a + 1 >= l0pp0 && l0a + 1 <= l0pp1);
            while (__any_sync(0xffffffffu, (needA && sa.x < 0.0) || (needB && sb.x < 0.0)))
            {
                __nanosleep(20);
                sa = lds_d2_volatile(in_slot(l0a)); sb = lds_d2_volatile(in_slot(l0a + 1));
            }
        }
        // left neighbour's second column: rows ia, ib (this step's shuffle) and ia - 1 (kept from the step before)
        double La = __shfl_up_sync(0xffffffffu, pubC_a, 1), Lb = __shfl_up_sync(0xffffffffu, pubC_b, 1);
        double LEa = 0, LEb = 0;
        if (REV) { LEa = __shfl_up_sync(0xffffffffu, pubE_a, 1); LEb = __shfl_up_sync(0xffffffffu, pubE_b, 1); }
        if (lane == 0) { La = sa.x; Lb = sb.x; LEa = sa.y; LEb = sb.y; }
        const double Lm = rLb, LEm = rLEb;
        rLb = Lb; rLEb = LEb;
        if (!act && !LEAN) return;
        double eA = emission(rc.a.mean, rc.a.stdv, rc.a.rstdv, rc.a.lsd3, p0, l2p, off);
        double eB = emission(rc.a.mean, rc.a.stdv, rc.a.rstdv, rc.a.lsd3, p1, l2p, off);
        double eC = emission(rc.b.mean, rc.b.stdv, rc.b.rstdv, rc.b.lsd3, p0, l2p, off);
        double eD = emission(rc.b.mean, rc.b.stdv, rc.b.rstdv, rc.b.lsd3, p1, l2p, off);
        const bool v0 = !INV || cur.sa >= 0, v1 = !INV || cur.sb >= 0;
        if (INV && !LEAN) { if (!v0) { eA = 0.0; eC = 0.0; } if (!v1) { eB = 0.0; eD = 0.0; } }
        double CA, CB, CC, CD, SA, SB, SC, SD, M;
        int kA, kB, kC, kD, qA, qB, qC, qD, m;
        if (LEAN)
        {
            cell_pre<false>(true, true, false, true, Lm, REV ? LEm : eA, REV ? upE0 : eA, upC0, upS0, tr, M, m, SA, qA);
            cell_fin(true, La, M, m, tr, CA, kA);
            cell_pre<false>(true, true, false, true, upC0, REV ? upE0 : eB, REV ? upE1 : eB, upC1, upS1, tr, M, m, SB, qB);
            cell_fin(true, CA, M, m, tr, CB, kB);
            cell_pre<false>(true, true, false, true, La, REV ? LEa : eC, REV ? eA : eC, CA, SA, tr, M, m, SC, qC);
            cell_fin(true, Lb, M, m, tr, CC, kC);
            cell_pre<false>(true, true, false, true, CA, REV ? eA : eD, REV ? eB : eD, CB, SB, tr, M, m, SD, qD);
            cell_fin(true, CC, M, m, tr, CD, kD);
        }
        else
        {
            // the column left of the strip is the blank column 0 for the first strip (all zeros, every row)
            const double zLm = cur.j > 0 ? Lm : 0.0, zLa = cur.j > 0 ? La : 0.0, zLb = cur.j > 0 ? Lb : 0.0;
            const double zLEm = cur.j > 0 ? LEm : 0.0, zLEa = cur.j > 0 ? LEa : 0.0;
            const bool inA = ia >= cur.i0a && ia <= cur.i1a, inB = ia >= cur.i0b && ia <= cur.i1b;
            const bool inC = ib >= cur.i0a && ib <= cur.i1a, inD = ib >= cur.i0b && ib <= cur.i1b;
            const bool skA = ia >= cur.pp0 && ia <= cur.pp1, dgA = ia > cur.pp0 && ia <= cur.pp1;
            const bool skC = ib >= cur.pp0 && ib <= cur.pp1, dgC = ib > cur.pp0 && ib <= cur.pp1;
            const bool skB = inA, dgB = ia > cur.i0a && ia <= cur.i1a;
            const bool skD = inC, dgD = ib > cur.i0a && ib <= cur.i1a;
            cell_pre<INV>(inA, v0, ia == cur.i0a, dgA, zLm, REV ? (dgA ? zLEm : 0.0) : eA, REV ? upE0 : eA, upC0, upS0, tr, M, m, SA, qA);
            cell_fin(skA && inA && v0, zLa, M, m, tr, CA, kA);
            cell_pre<INV>(inB, v1, ia == cur.i0b, dgB, upC0, REV ? (dgB ? upE0 : 0.0) : eB, REV ? upE1 : eB, upC1, upS1, tr, M, m, SB, qB);
            cell_fin(skB && inB && v1, CA, M, m, tr, CB, kB);
            cell_pre<INV>(inC, v0, ib == cur.i0a, dgC, zLa, REV ? (dgC ? zLEa : 0.0) : eC, REV ? eA : eC, CA, SA, tr, M, m, SC, qC);
            cell_fin(skC && inC && v0, zLb, M, m, tr, CC, kC);
            cell_pre<INV>(inD, v1, ib == cur.i0b, dgD, CA, REV ? (dgD ? eA : 0.0) : eD, REV ? eB : eD, CB, SB, tr, M, m, SD, qD);
            cell_fin(skD && inD && v1, CC, M, m, tr, CD, kD);
        }
        pubC_a = CB; pubC_b = CD; pubE_a = eB; pubE_b = eD;
        if (v0 && CA > best0) { best0 = CA; besti0 = ia; }
        if (v0 && CC > best0) { best0 = CC; besti0 = ib; }
        if (v1 && CB > best1) { best1 = CB; besti1 = ia; }
        if (v1 && CD > best1) { best1 = CD; besti1 = ib; }
        upC0 = CC; upS0 = SC; upE0 = eC; upC1 = CD; upS1 = SD; upE1 = eD;
        // hand the second column's rows to the next block (in-band rows only: their values are >= 0)
        if (handl)
        {
            if (LEAN || (ia >= cur.i0b && ia <= cur.i1b)) sts_d2_volatile(sout + ((unsigned)(ia - cur.i0b) << 4), CB, eB);
            if (LEAN || (ib >= cur.i0b && ib <= cur.i1b)) sts_d2_volatile(sout + ((unsigned)(ib - cur.i0b) << 4), CD, eD);
        }
        const long long a = ev.band_off + (long long)(cur.j + r_c) * ev.rs + cur.slot4;
        double2* pm = reinterpret_cast<double2*>(o.Mm + a);
        pm[0] = make_double2(CA, CB); pm[1] = make_double2(CC, CD);
        if (!REV)
        {
            double2* ps = reinterpret_cast<double2*>(o.Ms + a);
            ps[0] = make_double2(SA, SB); ps[1] = make_double2(SC, SD);
            *reinterpret_cast<unsigned*>(b.Fstep + a) = (unsigned)(kA | (qA << 3)) | ((unsigned)(kB | (qB << 3)) << 8) |
                                                       ((unsigned)(kC | (qC << 3)) << 16) | ((unsigned)(kD | (qD << 3)) << 24);
        }
    };

    for (int s = 0; s < s_lo; s++) step(std::false_type());
    for (int s = s_lo; s <= s_hi; s++) step(std::true_type());
    for (int s = s_hi + 1; s < nsteps; s++) step(std::false_type());

    // best cell of the strip's columns
    if (mine)
    {
        const int k = CW * j + 1;
        const long long g = ev.col_off + k;
        o.Mcb[g] = best0; o.Mcbi[g] = besti0;
        if (k + 1 <= N) { o.Mcb[g + 1] = best1; o.Mcbi[g + 1] = besti1; }
    }
    if (from_strip)
    {
        // the input strip goes back to "nothing written", then it is free for the block W further on
        for (int q = lane; q <= l0pp1 - l0pp0; q += 32) sts_d2_volatile(sin + ((unsigned)q << 4), -1.0, 0.0);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) rdone[(G + W - 1) % W] = G - 1;
    }
    if (!hand && lane == 0) rdone[G % W] = G;
    __syncwarp();
}

template <bool INV>
__global__ void __launch_bounds__(128, 4) k_fill2(Batch b, Fill2Args a)
{
    extern __shared__ __align__(16) unsigned char f2_smem[];
    const int W = blockDim.x >> 5, wrp = threadIdx.x >> 5;
    __shared__ int rdone[F2_MAX_WARPS];
    double2* strips = reinterpret_cast<double2*>(f2_smem);               // [W + 1][slots]; strip W: the blank column 0
    for (int q = threadIdx.x; q < (W + 1) * a.slots; q += blockDim.x) strips[q] = make_double2(q < W * a.slots ? -1.0 : 0.0, 0.0);
    if (threadIdx.x < W) rdone[threadIdx.x] = -1;
    __syncthreads();
    int gbase = 0;
    const int items = a.count * a.dirs;
    for (int it = blockIdx.x; it < items; it += gridDim.x)
    {
        const int e = a.list[it / a.dirs];
        const bool rev = (it % a.dirs) != 0;
        const EvDesc ev = b.ev[e];
        if (!ev.usable || ev.N <= 0) continue;
        FillOut o;
        o.Mm = rev ? b.Bm : b.Fm; o.Ms = rev ? nullptr : b.Fs;
        o.Mi0 = rev ? b.Bi0 : b.Fi0; o.Mlen = rev ? b.Blen : b.Flen;
        o.Mcb = rev ? b.Bcb : b.Fcb; o.Mcbi = rev ? b.Bcbi : b.Fcbi;
        const int J = (ev.N + CW - 1) / CW;
        const int nblocks = (J + 31) >> 5;
        for (int blk = ((wrp - gbase) % W + W) % W; blk < nblocks; blk += W)
        {
            const int G = gbase + blk;
            if (rev) fill2_block<true, INV>(b, ev, o, blk, nblocks, G, W, strips, a.slots, rdone);
            else fill2_block<false, INV>(b, ev, o, blk, nblocks, G, W, strips, a.slots, rdone);
        }
        gbase += nblocks;
    }
}

// running best over columns: first maximum in (column, row) order wins (cpp/Alignment.cpp:31-36, :158, :270); one CTA per
// (event, direction).  The same scan k_fill ends with.
__global__ void __launch_bounds__(256) k_fill_best(Batch b, Fill2Args a)
{
    const int T = blockDim.x, tid = threadIdx.x;
    const int e = a.list[blockIdx.x];
    const bool rev = blockIdx.y != 0;
    const EvDesc ev = b.ev[e];
    if (!ev.usable || ev.N <= 0) return;
    const int N = ev.N;
    const double* Mcb = rev ? b.Bcb : b.Fcb; const int* Mcbi = rev ? b.Bcbi : b.Fcbi;
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

} // namespace psdev
