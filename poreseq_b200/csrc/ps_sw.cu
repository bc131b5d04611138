// ps_sw.cu -- integer Smith-Waterman with full traceback on the GPU (cpp/swlib.cpp:211-340, swfull),
// batched over the seed sequences FindMutations realigns the region to (cpp/EventUtil.cpp:16,
// cpp/FindMutations.cpp:40-41).  SURVEY.md 8f rank 1: the O(L^2) Amdahl term of the consensus loop
// once the HMM scoring runs on the GPU (0.67 s per 10 kb x 10 kb pair on a host core).
//
// One CTA per (region sequence, seed) pair.  Thread t owns K consecutive columns i of the score
// matrix and walks down the rows j, one row per step, skewed by one step per thread (row j of thread t
// at step j + t - 1): the left neighbour's last column of rows j and j-1 comes through a 3-deep
// shared-memory ring, everything else is registers.  Scores are int32 (+5 / -4 / -8, cpp/swlib.cpp:24-27);
// the move byte per cell (1 = from (j-1, i), 2 = from (j, i-1), 3 = diagonal, +4 = score <= 0) has the
// reference's tie order: vertical, then horizontal, then the diagonal which also wins ties (>=).
// The best cell is the first maximum in the reference's (j outer, i inner) order.  A second kernel
// walks the traceback (one thread per pair) and returns the aligned index pairs.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "ps_internal.h"

namespace {

constexpr int SW_T = 1024;        // threads per pair
constexpr int SW_KMAX = 16;       // columns per thread: sequences up to 16 k bases

struct SwPair
{
    int n2;
    long long s2_off;             // into the concatenated seed bases
    long long mv_off;             // into the move records (16 bytes each): (n2 + T) x T per pair, record (step s, thread t) at s*T + t
    long long out_off;            // into the index outputs (n1 + n2 entries reserved per pair)
};

struct SwBest { int score, i, j, n, nmatch, pad; };

__global__ void __launch_bounds__(SW_T) k_sw_fill(const char* __restrict__ s1, int n1, const char* __restrict__ s2all,
                                                  const SwPair* __restrict__ pairs, uint8_t* __restrict__ move,
                                                  SwBest* __restrict__ best, int K)
{
    const SwPair pr = pairs[blockIdx.x];
    const int n2 = pr.n2, t = threadIdx.x, T = blockDim.x;
    const char* s2 = s2all + pr.s2_off;
    // Move bytes, step-major: the K moves a thread produces in one step are ONE 16-byte record, and the records of one
    // step lie next to each other (thread t at step s: record s * T + t), so a warp's stores are one contiguous 512-byte
    // run.  (Row-major bytes -- row j at j * (n1 + 1) -- made every thread store K single bytes into a row of its own per
    // step: 10 240 scattered byte stores per step at 10 kb, 104 ms per batch of 10 kb x 10 kb pairs however many pairs.)
    uint4* mv = reinterpret_cast<uint4*>(move) + pr.mv_off;
    const int i0 = t * K + 1;                                    // first column (1-based) of this thread
    const int ncols = min(K, n1 - t * K);                        // <= 0: nothing to do
    __shared__ int ring[3][SW_T];
    __shared__ int red_s[SW_T / 32], red_i[SW_T / 32], red_j[SW_T / 32];
    int H[SW_KMAX];
    char c1[SW_KMAX];
#pragma unroll
    for (int c = 0; c < SW_KMAX; c++) { H[c] = 0; c1[c] = (c < ncols) ? s1[i0 - 1 + c] : 0; }
    int bs = 0, bi = 0, bj = 0;
    const int steps = n2 + T - 1;
    for (int s = 0; s < steps; s++)
    {
        const int j = s - t + 1;
        const int w0 = s % 3, w1 = (s + 2) % 3, w2 = (s + 1) % 3;    // steps s, s-1, s-2
        if (j >= 1 && j <= n2 && ncols > 0)
        {
            int left = 0, diag = 0;                               // H(j, i0-1), H(j-1, i0-1)
            if (t > 0)
            {
                left = ring[w1][t - 1];
                diag = j > 1 ? ring[w2][t - 1] : 0;
            }
            const char cj = s2[j - 1];
            unsigned pk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int c = 0; c < SW_KMAX; c++)
            {
                if (c < ncols)
                {
                    const int up = H[c];
                    int sc = 0, m = 0;
                    int v = up - 8;
                    if (v > sc) { sc = v; m = 1; }
                    v = left - 8;
                    if (v > sc) { sc = v; m = 2; }
                    v = diag + (c1[c] == cj ? 5 : -4);
                    if (v >= sc) { sc = v; m = 3; }
                    pk[c >> 2] |= (unsigned)(m | (sc <= 0 ? 4 : 0)) << (8 * (c & 3));
                    if (sc > bs) { bs = sc; bi = i0 + c; bj = j; }
                    diag = up; left = sc; H[c] = sc;
                }
            }
            mv[(long long)s * T + t] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            ring[w0][t] = left;
        }
        __syncthreads();
    }
    // first maximum in (j, i) order: larger score, then smaller j, then smaller i
    auto better = [](int s_a, int j_a, int i_a, int s_b, int j_b, int i_b) {
        return s_a > s_b || (s_a == s_b && (j_a < j_b || (j_a == j_b && i_a < i_b)));
    };
    if (bs == 0) { bj = 1 << 30; bi = 1 << 30; }
    for (int o = 16; o; o >>= 1)
    {
        const int s2_ = __shfl_xor_sync(0xffffffffu, bs, o), i2 = __shfl_xor_sync(0xffffffffu, bi, o), j2 = __shfl_xor_sync(0xffffffffu, bj, o);
        if (better(s2_, j2, i2, bs, bj, bi)) { bs = s2_; bi = i2; bj = j2; }
    }
    if ((t & 31) == 0) { red_s[t >> 5] = bs; red_i[t >> 5] = bi; red_j[t >> 5] = bj; }
    __syncthreads();
    if (t == 0)
    {
        for (int w = 1; w < (T + 31) / 32; w++)
            if (better(red_s[w], red_j[w], red_i[w], bs, bj, bi)) { bs = red_s[w]; bi = red_i[w]; bj = red_j[w]; }
        SwBest b;
        b.score = bs; b.i = bs > 0 ? bi : 0; b.j = bs > 0 ? bj : 0; b.n = 0; b.nmatch = 0; b.pad = 0;
        best[blockIdx.x] = b;
    }
}

// traceback (cpp/swlib.cpp:296-333): index pairs from the best cell back to the first cell with score <= 0,
// written in walk order (the host reverses them)
__global__ void k_sw_trace(const char* __restrict__ s1, int n1, const char* __restrict__ s2all, const SwPair* __restrict__ pairs,
                           const uint8_t* __restrict__ move, SwBest* __restrict__ best, int* __restrict__ o1, int* __restrict__ o2,
                           int K, int T)
{
    if (threadIdx.x != 0) return;
    const SwPair pr = pairs[blockIdx.x];
    const char* s2 = s2all + pr.s2_off;
    const uint8_t* mv = move + pr.mv_off * 16;
    SwBest b = best[blockIdx.x];
    int i = b.i, j = b.j, n = 0, nmatch = 0;
    int* a1 = o1 + pr.out_off;
    int* a2 = o2 + pr.out_off;
    while (i > 0 && j > 0)
    {
        const int t = (i - 1) / K, c = (i - 1) - t * K;             // cell (j, i): thread t, its column c, step j + t - 1
        const uint8_t x = mv[((long long)(j + t - 1) * T + t) * 16 + c];
        if (x & 4) break;
        const int m = x & 3;
        if (m == 1) { a1[n] = 0; a2[n] = j; j--; }
        else if (m == 2) { a1[n] = i; a2[n] = 0; i--; }
        else if (m == 3) { a1[n] = i; a2[n] = j; if (s1[i - 1] == s2[j - 1]) nmatch++; i--; j--; }
        else break;
        n++;
    }
    b.n = n; b.nmatch = nmatch;
    best[blockIdx.x] = b;
}

template <class T>
int dev_room(ps_ctx* ctx, const char* name, size_t count, T** out)
{
    DevBuf& buf = ctx->bufs[name];
    int rc = ctx->ensure(buf, std::max<size_t>(count, 1) * sizeof(T));
    if (rc) return rc;
    *out = (T*)buf.p;
    return PS_OK;
}

}  // namespace

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess)                                                                 \
        {                                                                                         \
            ps_set_error(ctx, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(err__), __FILE__, \
                         __LINE__, #call);                                                        \
            return PS_E_CUDA;                                                                     \
        }                                                                                         \
    } while (0)
#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

// swfull of s1 against every sequence of `others` on the GPU.  Returns PS_E_ARG (without touching
// `out`) when a sequence is too long for the kernel's register strips; the caller keeps the host form.
int psi_swfull_batch(ps_ctx* ctx, const std::string& s1, const std::vector<std::string>& others, std::vector<SWResult>& out)
{
    const int n1 = (int)s1.size();
    const size_t P = others.size();
    out.assign(P, SWResult());
    if (P == 0) return PS_OK;
    if (n1 == 0 || n1 > SW_T * SW_KMAX) return PS_E_ARG;
    TRY(ctx->init());
    CU(cudaSetDevice(ctx->device));
    // columns per thread and threads per pair: at least 8 columns per thread (short sequences take fewer threads: a
    // cheaper barrier, a shorter skew), at most SW_T threads
    const int K = std::max(8, (n1 + SW_T - 1) / SW_T);
    const int T = std::min(SW_T, (((n1 + K - 1) / K) + 31) & ~31);
    // pairs in sub-batches whose move matrices fit a quarter of the device memory
    const double budget = 0.25 * (double)ctx->total_mem;
    size_t a = 0;
    while (a < P)
    {
        size_t b = a;
        double bytes = 0;
        std::vector<SwPair> pairs;
        std::string cat;
        long long mv_off = 0, out_off = 0;
        while (b < P)
        {
            const double need = ((double)others[b].size() + T) * (double)T * 16.0;
            if (b > a && bytes + need > budget) break;
            SwPair p;
            p.n2 = (int)others[b].size();
            p.s2_off = (long long)cat.size();
            p.mv_off = mv_off;
            p.out_off = out_off;
            cat += others[b];
            mv_off += ((long long)p.n2 + T) * (long long)T;        // 16-byte records
            out_off += (long long)n1 + p.n2 + 2;
            bytes += need;
            pairs.push_back(p);
            b++;
        }
        const size_t np = pairs.size();
        char* d_s1; char* d_s2; SwPair* d_pairs; uint8_t* d_mv; SwBest* d_best; int* d_o1; int* d_o2;
        TRY(dev_room(ctx, "sw_s1", (size_t)n1, &d_s1));
        TRY(dev_room(ctx, "sw_s2", cat.size(), &d_s2));
        TRY(dev_room(ctx, "sw_pairs", np, &d_pairs));
        TRY(dev_room(ctx, "sw_move", (size_t)mv_off * 16, &d_mv));
        TRY(dev_room(ctx, "sw_best", np, &d_best));
        TRY(dev_room(ctx, "sw_o1", (size_t)out_off, &d_o1));
        TRY(dev_room(ctx, "sw_o2", (size_t)out_off, &d_o2));
        // host sides of all copies in pinned memory of the context (a pageable copy waits for the stream inside
        // the driver call and holds up the CUDA calls of other host threads meanwhile)
        PinVec<char> h_s1 = ctx->pinned<char>("sw_s1_h"), h_s2 = ctx->pinned<char>("sw_s2_h");
        PinVec<SwPair> h_pairs = ctx->pinned<SwPair>("sw_pairs_h");
        PinVec<SwBest> best = ctx->pinned<SwBest>("sw_best_h");
        PinVec<int> o1 = ctx->pinned<int>("sw_o1_h"), o2 = ctx->pinned<int>("sw_o2_h");
        if (!h_s1.resize((size_t)n1) || !h_s2.resize(std::max<size_t>(cat.size(), 1)) || !h_pairs.resize(np) || !best.resize(np) ||
            !o1.resize((size_t)out_off) || !o2.resize((size_t)out_off))
        {
            ps_set_error(ctx, "out of host memory staging the alignments");
            return PS_E_INTERNAL;
        }
        memcpy(h_s1.data(), s1.data(), (size_t)n1);
        memcpy(h_s2.data(), cat.data(), cat.size());
        std::copy(pairs.begin(), pairs.end(), h_pairs.data());
        CU(cudaMemcpyAsync(d_s1, h_s1.data(), (size_t)n1, cudaMemcpyHostToDevice, ctx->stream));
        if (!cat.empty()) CU(cudaMemcpyAsync(d_s2, h_s2.data(), cat.size(), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_pairs, h_pairs.data(), np * sizeof(SwPair), cudaMemcpyHostToDevice, ctx->stream));
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        double t0 = 0, t1 = 0;
        if (ctx->trace) { CU(cudaStreamSynchronize(ctx->stream)); t0 = now(); }
        k_sw_fill<<<(unsigned)np, T, 0, ctx->stream>>>(d_s1, n1, d_s2, d_pairs, d_mv, d_best, K);
        ctx->launches++;
        CU(cudaGetLastError());
        if (ctx->trace) { CU(cudaStreamSynchronize(ctx->stream)); t1 = now(); }
        k_sw_trace<<<(unsigned)np, 32, 0, ctx->stream>>>(d_s1, n1, d_s2, d_pairs, d_mv, d_best, d_o1, d_o2, K, T);
        ctx->launches++;
        CU(cudaGetLastError());
        if (ctx->trace)
        {
            CU((cudaError_t)ps_stream_wait(ctx));
            fprintf(stderr, "[ps] swfull batch: %zu pairs of %d x ~%zu: fill %.2f ms, traceback %.2f ms\n", np, n1, cat.size() / np, t1 - t0, now() - t1);
        }
        CU(cudaMemcpyAsync(best.data(), d_best, np * sizeof(SwBest), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(o1.data(), d_o1, (size_t)out_off * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(o2.data(), d_o2, (size_t)out_off * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU((cudaError_t)ps_stream_wait(ctx));
        for (size_t k = 0; k < np; k++)
        {
            SWResult& r = out[a + k];
            const SwBest& bb = best[k];
            r.score = bb.score;
            r.inds1.assign(o1.begin() + pairs[k].out_off, o1.begin() + pairs[k].out_off + bb.n);
            r.inds2.assign(o2.begin() + pairs[k].out_off, o2.begin() + pairs[k].out_off + bb.n);
            std::reverse(r.inds1.begin(), r.inds1.end());
            std::reverse(r.inds2.begin(), r.inds2.end());
            r.accuracy = 100.0 * bb.nmatch / (double)r.inds1.size();
        }
        a = b;
    }
    return PS_OK;
}
