"""Event-pack files: many regions (sequence, params, per-event level arrays, pore models, seed alignments) in one
flat binary file that is memory-mapped and handed to the library without per-event Python objects.

SURVEY.md section 8f rank 3: the reference loads every region through h5py (fast5 events, poreseq/EventData.py:100-175)
and pysam (BAM mapping, poreseq/LoadData.py:67-153) into Python objects; at assembly scale (BASELINE.json configs[4],
~512 regions x 100 reads) that loader, not the scoring, sets the pace.  A pack is written once (by whatever front end
has the events) and read with zero copies: `read_pack(path)[k]` is a `poreseqcpp.PackedRegion` whose arrays are views
into the mapping, which is exactly what `ps_regions_create` takes.

Layout (little endian, every block starts on an 8-byte boundary):

    header   8 B magic "PSEP0001" | u64 n_regions | u64 index_offset
    region   u32 seq_len, n_events, n_models, n_params, n_levels, seq2d_bytes, 0, 0
             params      n_params x (16 B name, zero padded | f64 value)
             sequence    seq_len bytes
             n0, model_index, complement          int32[n_events] each
             mean, stdv, ref_align, ref_like      float64[n_levels] each, events concatenated
             models      float64[n_models][4][1024]   level_mean, level_stdv, sd_mean, sd_stdv
             probs       float64[n_models][4]         prob_skip, prob_stay, prob_extend, prob_insert
             seq2d_len   int32[n_events]; seq2d bytes concatenated
    index    n_regions x (u64 offset, u64 size)
"""
import struct

import numpy as np

from . import poreseqcpp

MAGIC = b"PSEP0001"


def _pad8(n):
    return (n + 7) & ~7


def _block(f, data):
    f.write(data)
    f.write(b"\0" * (_pad8(len(data)) - len(data)))


def write_pack(path, regions):
    """Writes PSAlign-like objects (`.sequence`, `.events`, `.params`) or PackedRegion objects to `path`."""
    index = []
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<QQ", len(regions), 0))
        for reg in regions:
            p = reg if isinstance(reg, poreseqcpp.PackedRegion) else poreseqcpp.PackedRegion(reg.sequence, reg.events, reg.params)
            start = f.tell()
            seq2d = [s.encode("ascii") for s in p.seq2d]
            names = sorted(p.params)
            f.write(struct.pack("<8I", len(p.sequence), len(p.n0), len(p.models), len(names), int(p.n0.sum()),
                                sum(len(s) for s in seq2d), 0, 0))
            for k in names:
                kb = k.encode("ascii")
                if len(kb) > 16:
                    raise ValueError("parameter name %r is longer than 16 bytes" % k)
                f.write(kb.ljust(16, b"\0") + struct.pack("<d", float(p.params[k])))
            _block(f, p.sequence)
            for a in (p.n0, p.model_index, p.complement):
                _block(f, np.ascontiguousarray(a, dtype="<i4").tobytes())
            for a in (p.mean, p.stdv, p.ref_align, p.ref_like, p.models, p.probs):
                _block(f, np.ascontiguousarray(a, dtype="<f8").tobytes())
            _block(f, np.array([len(s) for s in seq2d], dtype="<i4").tobytes())
            _block(f, b"".join(seq2d))
            index.append((start, f.tell() - start))
        at = f.tell()
        for off, size in index:
            f.write(struct.pack("<QQ", off, size))
        f.seek(len(MAGIC))
        f.write(struct.pack("<QQ", len(regions), at))


class Pack(object):
    """A memory-mapped event-pack file; `pack[k]` is region k as a PackedRegion over views into the mapping."""

    def __init__(self, path):
        self.map = np.memmap(path, dtype=np.uint8, mode="r")
        if bytes(self.map[:8]) != MAGIC:
            raise ValueError("%s is not an event pack (bad magic)" % path)
        n, at = struct.unpack("<QQ", bytes(self.map[8:24]))
        self.index = np.frombuffer(self.map, dtype="<u8", count=2 * n, offset=at).reshape(n, 2)

    def __len__(self):
        return len(self.index)

    def __getitem__(self, k):
        if k < 0:
            k += len(self)
        off = int(self.index[k, 0])
        seq_len, n_ev, n_mod, n_par, n_lev, n_2d, _, _ = struct.unpack("<8I", bytes(self.map[off:off + 32]))
        at = [off + 32]

        def take(dtype, count, itemsize):
            a = np.frombuffer(self.map, dtype=dtype, count=count, offset=at[0])
            at[0] += _pad8(count * itemsize)
            return a

        params = {}
        for _ in range(n_par):
            raw = bytes(self.map[at[0]:at[0] + 24])
            v = struct.unpack("<d", raw[16:])[0]
            name = raw[:16].rstrip(b"\0").decode("ascii")
            params[name] = int(v) if v.is_integer() and name != "lik_offset" else v
            at[0] += 24
        sequence = bytes(take(np.uint8, seq_len, 1))
        n0, model_index, complement = (take("<i4", n_ev, 4) for _ in range(3))
        mean, stdv, ref_align, ref_like = (take("<f8", n_lev, 8) for _ in range(4))
        models = take("<f8", n_mod * 4 * 1024, 8)
        probs = take("<f8", n_mod * 4, 8)
        lens = take("<i4", n_ev, 4)
        blob = bytes(take(np.uint8, n_2d, 1))
        seq2d, c = [], 0
        for ln in lens.tolist():
            seq2d.append(blob[c:c + ln].decode("ascii"))
            c += ln
        return poreseqcpp.PackedRegion.from_arrays(sequence, params, n0, mean, stdv, ref_align, ref_like, model_index,
                                                   models, probs, complement, seq2d)

    def __iter__(self):
        return (self[k] for k in range(len(self)))


def read_pack(path):
    return Pack(path)


class NativePack(object):
    """The same file opened below the C-ABI (`ps_pack_open`, csrc/ps_pack.cu): the library maps it and builds its
    regions straight from the mapping -- `regions(ctx, first, count)` is ONE call, with no per-region Python object
    and no numpy view in between.  This is the form a C5-sized job uses (hundreds of regions per pack)."""

    def __init__(self, path):
        import ctypes as C
        self.lib = poreseqcpp.lib()
        self.handle = self.lib.ps_pack_open(str(path).encode())
        if not self.handle:
            raise ValueError(self.lib.ps_last_error(None).decode())
        self._C = C

    def close(self):
        if self.handle:
            self.lib.ps_pack_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self.lib.ps_pack_num_regions(self.handle)

    def desc(self, k, width_key=None):
        """ps_region_desc of region k (pointers into the mapping)."""
        d = poreseqcpp.PSRegionDesc()
        rc = self.lib.ps_pack_region_desc(self.handle, int(k), width_key.encode() if width_key else None, self._C.byref(d))
        if rc:
            raise IndexError("region %d: %s" % (k, self.lib.ps_last_error(None).decode()))
        return d

    def param(self, k, name, default=None):
        v = self._C.c_double(0)
        if self.lib.ps_pack_region_param(self.handle, int(k), name.encode(), self._C.byref(v)):
            return default
        return v.value

    def event_sequence(self, k, e):
        ptr, ln = self._C.c_void_p(), self._C.c_int(0)
        if self.lib.ps_pack_event_sequence(self.handle, int(k), int(e), self._C.byref(ptr), self._C.byref(ln)):
            raise IndexError("region %d event %d" % (k, e))
        return self._C.string_at(ptr.value, ln.value).decode("ascii") if ln.value else ""

    def regions(self, ctx, first=0, count=None, width_key=None):
        """ps_pack_regions_create: NativeRegion objects for regions [first, first + count)."""
        C = self._C
        n = len(self) - first if count is None else count
        out = (C.c_void_p * max(n, 1))()
        ctx.check(self.lib.ps_pack_regions_create(ctx.handle, self.handle, int(first), int(n),
                                                  width_key.encode() if width_key else None, out))
        regs = []
        for k in range(n):
            d = self.desc(first + k)
            r = poreseqcpp.NativeRegion.__new__(poreseqcpp.NativeRegion)
            r.ctx, r.handle = ctx, out[k]
            r.n_levels = np.ctypeslib.as_array(C.cast(d.n0, C.POINTER(C.c_int)), shape=(d.n_events,)).tolist() if d.n_events else []
            regs.append(r)
        return regs
