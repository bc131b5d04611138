"""ctypes loader for the two CPU checkers behind oracle_api.h.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by poreseq_b200.

    ref = load("ref")        # oracle/_ref/libps_ref.so  -- the reference's own C++ (needs `make ref`)
    orc = load("oracle")     # oracle/_build/libps_oracle.so -- the independent restatement

Both expose the same methods; each takes a "region" = any object with .sequence (str),
.events (list of PSEvent-like) and .params (dict), i.e. what PSAlign holds
(poreseq/_poreseqcpp.pyx:225-229).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBS = {"ref": os.path.join(HERE, "_ref", "libps_ref.so"),
        "oracle": os.path.join(HERE, "_build", "libps_oracle.so")}
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class OrcRegion(C.Structure):
    _fields_ = [("seq", C.c_char_p), ("seq_len", C.c_int), ("n_events", C.c_int),
                ("lev_off", c_int_p), ("mean", c_double_p), ("stdv", c_double_p),
                ("ref_align", c_double_p), ("ref_like", c_double_p), ("model", c_double_p),
                ("trans", c_double_p), ("ev_seq", C.POINTER(C.c_char_p)),
                ("lik_offset", C.c_double), ("scoring_width", C.c_int), ("realign_width", C.c_int)]


def build(which="oracle", quiet=True):
    """Compile a checker with oracle/Makefile (building the checker is not using it)."""
    target = "ref" if which == "ref" else "oracle"
    subprocess.check_call(["make", "-C", HERE, target],
                          stdout=subprocess.DEVNULL if quiet else None)


def available(which):
    return os.path.exists(LIBS[which])


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _strs(items):
    arr = (C.c_char_p * max(len(items), 1))()
    for i, s in enumerate(items):
        arr[i] = s.encode("ascii") if isinstance(s, str) else s
    return arr


def parse_mutations(text):
    out = []
    for line in text.splitlines():
        if not line:
            continue
        s, o, m, sc = line.split("\t")
        out.append((int(s), "" if o == "." else o, "" if m == "." else m, float(sc)))
    return out


class Packed(object):
    """Flat arrays for one region (keeps the numpy buffers alive)."""

    def __init__(self, region, width_key="scoring_width"):
        evs = region.events
        self.n = [len(ev.mean) for ev in evs]
        self.lev_off = np.zeros(len(evs) + 1, dtype=np.int32)
        self.lev_off[1:] = np.cumsum(self.n)
        cat = lambda name: (np.ascontiguousarray(np.concatenate([np.asarray(getattr(ev, name), dtype="f8") for ev in evs]))
                            if evs else np.zeros(0))
        self.mean, self.stdv = cat("mean"), cat("stdv")
        self.ref_align, self.ref_like = cat("ref_align").copy(), cat("ref_like").copy()
        self.model = np.zeros((len(evs), 4, 1024))
        self.trans = np.zeros((len(evs), 4))
        for i, ev in enumerate(evs):
            m = ev.model
            self.model[i, 0], self.model[i, 1] = m.level_mean, m.level_stdv
            self.model[i, 2], self.model[i, 3] = m.sd_mean, m.sd_stdv
            self.trans[i] = (m.prob_skip, m.prob_stay, m.prob_extend, m.prob_insert)
        self.ev_seq = _strs([getattr(ev, "sequence", "") or "" for ev in evs])
        self.seq = region.sequence.encode("ascii")
        p = region.params
        self.c = OrcRegion(self.seq, len(self.seq), len(evs), self.lev_off.ctypes.data_as(c_int_p),
                           _dp(self.mean), _dp(self.stdv), _dp(self.ref_align), _dp(self.ref_like),
                           _dp(self.model), _dp(self.trans), self.ev_seq,
                           float(p.get("lik_offset", 4.5)), int(p.get(width_key, p.get("scoring_width", 150))),
                           int(p.get("realign_width", 300)))

    def aligns(self):
        """Per-event (ref_align, ref_like) after the call."""
        return [(self.ref_align[a:b].copy(), self.ref_like[a:b].copy())
                for a, b in zip(self.lev_off[:-1], self.lev_off[1:])]


class Checker(object):
    def __init__(self, which):
        self.which = which
        if not os.path.exists(LIBS[which]):
            raise RuntimeError("%s not built; run `make -C oracle %s`" % (LIBS[which], "ref" if which == "ref" else "oracle"))
        self.lib = C.CDLL(LIBS[which])
        self.lib.orc_name.restype = C.c_char_p
        self.libc = C.CDLL("libc.so.6")

    def srand(self, seed=1):
        self.libc.srand(C.c_uint(seed))

    def score_alignments(self, region, want_likes=False):
        pk = Packed(region)
        scores = np.zeros(len(region.events))
        likes = np.zeros(len(region.sequence)) if want_likes else None
        rc = self.lib.orc_score_alignments(C.byref(pk.c), _dp(scores), _dp(likes) if want_likes else None)
        assert rc == 0
        return scores, likes, pk.aligns()

    def score_mutations(self, region, starts, origs, muts, width_key="scoring_width"):
        pk = Packed(region, width_key)
        n = len(starts)
        st = np.asarray(starts, dtype=np.int32)
        scores = np.zeros(n)
        rc = self.lib.orc_score_mutations(C.byref(pk.c), n, st.ctypes.data_as(c_int_p), _strs(origs), _strs(muts), _dp(scores))
        assert rc == 0
        return scores, pk.aligns()

    def score_points(self, region):
        pk = Packed(region, "point_width")
        cap = 64 * 8 * (len(region.sequence) + 8)
        buf = C.create_string_buffer(cap)
        rc = self.lib.orc_score_points(C.byref(pk.c), buf, cap)
        assert rc == 0
        return parse_mutations(buf.value.decode()), pk.aligns()

    def make_mutations(self, region, starts, origs, muts, scores, width_key="point_width"):
        pk = Packed(region, width_key)
        n = len(starts)
        st = np.asarray(starts, dtype=np.int32)
        sc = np.asarray(scores, dtype="f8")
        cap = 2 * len(region.sequence) + sum(len(m) for m in muts) + 64
        buf = C.create_string_buffer(cap)
        nb = C.c_int(0)
        rc = self.lib.orc_make_mutations(C.byref(pk.c), n, st.ctypes.data_as(c_int_p), _strs(origs), _strs(muts),
                                         _dp(sc), buf, cap, C.byref(nb))
        assert rc == 0
        return buf.value.decode(), nb.value, pk.aligns()

    def refine(self, region):
        pk = Packed(region, "point_width")
        cap = 16 * len(region.sequence) + 1024           # (insertion probabilities above 1 make every insertion a gain)
        buf = C.create_string_buffer(cap)
        nb = C.c_int(0)
        rc = self.lib.orc_refine(C.byref(pk.c), buf, cap, C.byref(nb))
        assert rc == 0
        return buf.value.decode(), nb.value, pk.aligns()

    def find_mutations(self, region, seeds):
        pk = Packed(region)
        cap = 64 * (len(region.sequence) + 8) + 4 * sum(len(s) for s in seeds)
        buf = C.create_string_buffer(cap)
        rc = self.lib.orc_find_mutations(C.byref(pk.c), len(seeds), _strs(seeds), buf, cap)
        assert rc == 0
        return [m[:3] for m in parse_mutations(buf.value.decode())], pk.aligns()

    def mutate(self, region, seeds, reps=4):
        pk = Packed(region)
        cap = 4 * len(region.sequence) + 2 * max([len(s) for s in seeds] + [0]) + 64
        buf = C.create_string_buffer(cap)
        nb = C.c_int(0)
        rc = self.lib.orc_mutate(C.byref(pk.c), len(seeds), _strs(seeds), reps, buf, cap, C.byref(nb))
        assert rc == 0
        return buf.value.decode(), nb.value, pk.aligns()

    def viterbi_mutate(self, region, nkeep=16, skip=0.05, stay=0.01, mut_min=0.33, mut_max=0.75, seed=1):
        pk = Packed(region)
        cap = (max(nkeep, 1) + 1) * (2 * len(region.sequence) + 4096)
        buf = C.create_string_buffer(cap)
        if seed is not None:
            self.srand(seed)
        rc = self.lib.orc_viterbi_mutate(C.byref(pk.c), nkeep, C.c_double(skip), C.c_double(stay),
                                         C.c_double(mut_min), C.c_double(mut_max), buf, cap)
        assert rc == 0
        return buf.value.decode().split("\n")[:-1]

    def swfull(self, seq1, seq2):
        cap = len(seq1) + len(seq2) + 8
        i1 = np.zeros(cap, dtype=np.int32)
        i2 = np.zeros(cap, dtype=np.int32)
        n, score, acc = C.c_int(0), C.c_int(0), C.c_double(0)
        rc = self.lib.orc_swfull(seq1.encode(), seq2.encode(), i1.ctypes.data_as(c_int_p), i2.ctypes.data_as(c_int_p),
                                 cap, C.byref(n), C.byref(score), C.byref(acc))
        assert rc == 0
        return acc.value, score.value, list(zip(i1[:n.value].tolist(), i2[:n.value].tolist()))

    def map_alignments(self, region, newseq):
        pk = Packed(region)
        rc = self.lib.orc_map_alignments(C.byref(pk.c), newseq.encode())
        assert rc == 0
        return pk.aligns()

    def seq_to_states(self, seq):
        out = np.zeros(max(len(seq), 1), dtype=np.int32)
        n = self.lib.orc_seq_to_states(seq.encode(), len(seq), out.ctypes.data_as(c_int_p))
        return out[:n].copy()


_cache = {}


def load(which="oracle"):
    if which not in _cache:
        _cache[which] = Checker(which)
    return _cache[which]
