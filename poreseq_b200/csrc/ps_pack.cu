// Event-pack files below the C-ABI.
//
// The reference loads every region through h5py (fast5 events, poreseq/EventData.py:100-175) and pysam (BAM
// mapping, poreseq/LoadData.py:67-153) into Python objects.  A pack holds the same fields of many regions in one
// flat file (layout: poreseq_b200/eventpack.py, which also writes it); here it is memory-mapped and its regions are
// handed to ps_regions_create as descriptors whose pointers are views into the mapping: no Python object, no
// intermediate copy between the page cache and the library's region objects.  Host code only (no CUDA call).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "ps_internal.h"

namespace
{
const char PACK_MAGIC[8] = {'P', 'S', 'E', 'P', '0', '0', '0', '1'};

inline uint64_t pad8(uint64_t n) { return (n + 7) & ~(uint64_t)7; }

struct PackParam { char name[17]; double value; };

// Where the blocks of one region lie in the mapping (validated against the region's size once, at open)
struct PackRegion
{
    uint32_t seq_len = 0, n_events = 0, n_models = 0, n_params = 0, n_levels = 0, seq2d_bytes = 0;
    const unsigned char* params = nullptr;       // n_params x (16 B name | f64)
    const char* sequence = nullptr;
    const int32_t *n0 = nullptr, *model_index = nullptr, *complement = nullptr, *seq2d_len = nullptr;
    const double *mean = nullptr, *stdv = nullptr, *ref_align = nullptr, *ref_like = nullptr, *models = nullptr, *probs = nullptr;
    const char* seq2d = nullptr;
    std::vector<uint64_t> seq2d_off;             // n_events + 1 offsets into seq2d
};
}

struct ps_pack
{
    int fd = -1;
    const unsigned char* map = nullptr;
    size_t size = 0;
    std::vector<PackRegion> regions;
    std::string error;
};

namespace
{
bool pack_fail(ps_pack* p, const char* path, const char* why)
{
    ps_set_error(nullptr, "ps_pack_open(%s): %s", path, why);
    if (p)
    {
        if (p->map) munmap(const_cast<unsigned char*>(p->map), p->size);
        if (p->fd >= 0) close(p->fd);
        delete p;
    }
    return false;
}

// Lays out region `r` over [base, base + size); false when a block would leave the region
bool pack_layout(const unsigned char* base, uint64_t size, PackRegion& r)
{
    if (size < 32) return false;
    uint32_t h[8];
    memcpy(h, base, 32);
    r.seq_len = h[0]; r.n_events = h[1]; r.n_models = h[2]; r.n_params = h[3]; r.n_levels = h[4]; r.seq2d_bytes = h[5];
    uint64_t at = 32;
    auto take = [&](uint64_t bytes, bool padded) -> const unsigned char* {
        const uint64_t step = padded ? pad8(bytes) : bytes;
        if (bytes > size || at > size - bytes) return nullptr;
        const unsigned char* p = base + at;
        at += step;
        return p;
    };
    const unsigned char* q;
    if (!(q = take((uint64_t)r.n_params * 24, false))) return false;
    r.params = q;
    if (!(q = take(r.seq_len, true))) return false;
    r.sequence = reinterpret_cast<const char*>(q);
    const int32_t** ints[3] = {&r.n0, &r.model_index, &r.complement};
    for (auto ip : ints)
    {
        if (!(q = take((uint64_t)r.n_events * 4, true))) return false;
        *ip = reinterpret_cast<const int32_t*>(q);
    }
    const double** lev[4] = {&r.mean, &r.stdv, &r.ref_align, &r.ref_like};
    for (auto dp : lev)
    {
        if (!(q = take((uint64_t)r.n_levels * 8, true))) return false;
        *dp = reinterpret_cast<const double*>(q);
    }
    if (!(q = take((uint64_t)r.n_models * 4 * PS_N_STATES * 8, true))) return false;
    r.models = reinterpret_cast<const double*>(q);
    if (!(q = take((uint64_t)r.n_models * 4 * 8, true))) return false;
    r.probs = reinterpret_cast<const double*>(q);
    if (!(q = take((uint64_t)r.n_events * 4, true))) return false;
    r.seq2d_len = reinterpret_cast<const int32_t*>(q);
    if (!(q = take(r.seq2d_bytes, true))) return false;
    r.seq2d = reinterpret_cast<const char*>(q);
    if (at > pad8(size)) return false;
    // consistency of the per-event tables with the block sizes
    uint64_t levels = 0, s2 = 0;
    r.seq2d_off.assign((size_t)r.n_events + 1, 0);
    for (uint32_t e = 0; e < r.n_events; e++)
    {
        if (r.n0[e] < 0 || r.seq2d_len[e] < 0 || r.model_index[e] < 0 || (uint32_t)r.model_index[e] >= r.n_models) return false;
        levels += (uint64_t)r.n0[e];
        s2 += (uint64_t)r.seq2d_len[e];
        r.seq2d_off[e + 1] = s2;
    }
    return levels == r.n_levels && s2 == r.seq2d_bytes;
}

bool pack_param(const PackRegion& r, const char* name, double* value)
{
    const size_t ln = strlen(name);
    if (ln > 16) return false;
    for (uint32_t k = 0; k < r.n_params; k++)
    {
        const unsigned char* q = r.params + (size_t)k * 24;
        size_t have = 0;
        while (have < 16 && q[have]) have++;
        if (have == ln && memcmp(q, name, ln) == 0)
        {
            memcpy(value, q + 16, 8);
            return true;
        }
    }
    return false;
}

// ps_params of a region the way the Python surface derives them (poreseq/_poreseqcpp.pyx:144-151; point_width
// override :293,361,465), AlignParams defaults (cpp/AlignUtil.h:57-66) for absent keys
ps_params pack_params(const PackRegion& r, const char* width_key)
{
    ps_params p;
    p.lik_offset = 4.5; p.scoring_width = 150; p.realign_width = 300; p.verbose = 0;
    double v;
    if (pack_param(r, "verbose", &v)) p.verbose = (int)v;
    if (pack_param(r, "lik_offset", &v)) p.lik_offset = v;
    if (pack_param(r, "realign_width", &v)) p.realign_width = (int)v;
    if (pack_param(r, "scoring_width", &v)) p.scoring_width = (int)v;
    if (width_key && pack_param(r, width_key, &v)) p.scoring_width = (int)v;
    return p;
}

void pack_desc(const PackRegion& r, const char* width_key, ps_region_desc* d)
{
    d->bases = r.sequence; d->len = (int)r.seq_len;
    d->params = pack_params(r, width_key);
    d->n_events = (int)r.n_events; d->n0 = r.n0;
    d->mean = r.mean; d->stdv = r.stdv; d->ref_align = r.ref_align; d->ref_like = r.ref_like;
    d->model_index = r.model_index; d->n_models = (int)r.n_models;
    d->models = r.models; d->probs = r.probs;
    d->complement = r.complement;
    d->seq2d = nullptr;                          // the 2D sequences stay in the file: ps_pack_event_sequence
}
}

extern "C" {

ps_pack* ps_pack_open(const char* path)
{
    if (!path) { ps_set_error(nullptr, "ps_pack_open: bad arguments"); return nullptr; }
    ps_pack* p = new ps_pack();
    p->fd = open(path, O_RDONLY);
    if (p->fd < 0) { pack_fail(p, path, "cannot open the file"); return nullptr; }
    struct stat st;
    if (fstat(p->fd, &st) != 0 || st.st_size < 24) { pack_fail(p, path, "too short for an event pack"); return nullptr; }
    p->size = (size_t)st.st_size;
    void* m = mmap(nullptr, p->size, PROT_READ, MAP_PRIVATE, p->fd, 0);
    if (m == MAP_FAILED) { p->map = nullptr; pack_fail(p, path, "mmap failed"); return nullptr; }
    p->map = static_cast<const unsigned char*>(m);
    if (memcmp(p->map, PACK_MAGIC, 8) != 0) { pack_fail(p, path, "not an event pack (bad magic)"); return nullptr; }
    uint64_t n = 0, at = 0;
    memcpy(&n, p->map + 8, 8);
    memcpy(&at, p->map + 16, 8);
    if (at > p->size || n > (p->size - at) / 16 || (at & 7)) { pack_fail(p, path, "index outside the file"); return nullptr; }
    p->regions.resize((size_t)n);
    for (uint64_t k = 0; k < n; k++)
    {
        uint64_t off = 0, size = 0;
        memcpy(&off, p->map + at + 16 * k, 8);
        memcpy(&size, p->map + at + 16 * k + 8, 8);
        if ((off & 7) || off > p->size || size > p->size - off || !pack_layout(p->map + off, size, p->regions[(size_t)k]))
        {
            char why[96];
            snprintf(why, sizeof why, "region %llu is inconsistent", (unsigned long long)k);
            pack_fail(p, path, why);
            return nullptr;
        }
    }
    return p;
}

void ps_pack_close(ps_pack* p)
{
    if (!p) return;
    if (p->map) munmap(const_cast<unsigned char*>(p->map), p->size);
    if (p->fd >= 0) close(p->fd);
    delete p;
}

int ps_pack_num_regions(ps_pack* p) { return p ? (int)p->regions.size() : PS_E_ARG; }

int ps_pack_region_desc(ps_pack* p, int k, const char* width_key, ps_region_desc* out)
{
    if (!p || !out || k < 0 || k >= (int)p->regions.size()) return PS_BAD_ARGS(nullptr, "ps_pack_region_desc");
    pack_desc(p->regions[k], width_key, out);
    return PS_OK;
}

int ps_pack_region_param(ps_pack* p, int k, const char* name, double* value)
{
    if (!p || !name || !value || k < 0 || k >= (int)p->regions.size()) return PS_BAD_ARGS(nullptr, "ps_pack_region_param");
    return pack_param(p->regions[k], name, value) ? PS_OK : PS_E_ARG;
}

int ps_pack_event_sequence(ps_pack* p, int k, int e, const char** seq, int* len)
{
    if (!p || !seq || !len || k < 0 || k >= (int)p->regions.size()) return PS_BAD_ARGS(nullptr, "ps_pack_event_sequence");
    const PackRegion& r = p->regions[k];
    if (e < 0 || e >= (int)r.n_events) return PS_BAD_ARGS(nullptr, "ps_pack_event_sequence");
    *seq = r.seq2d + r.seq2d_off[e];
    *len = (int)(r.seq2d_off[e + 1] - r.seq2d_off[e]);
    return PS_OK;
}

int ps_pack_regions_create(ps_ctx* ctx, ps_pack* p, int first, int count, const char* width_key, ps_region** out)
{
    if (!ctx || !p || first < 0 || count < 0 || first > (int)p->regions.size() || count > (int)p->regions.size() - first ||
        (count > 0 && !out))
        return PS_BAD_ARGS(ctx, "ps_pack_regions_create");
    std::vector<ps_region_desc> desc((size_t)count);
    for (int k = 0; k < count; k++) pack_desc(p->regions[(size_t)first + k], width_key, &desc[k]);
    return ps_regions_create(ctx, count, desc.data(), out);
}

} // extern "C"
