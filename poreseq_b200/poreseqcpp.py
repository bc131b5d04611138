"""Drop-in mirror of the reference's Cython module `poreseq.poreseqcpp`
(poreseq/_poreseqcpp.pyx:155-473): `PSAlign`, `swalign`, `seqtostates`, with every native call going
through the C-ABI of include/poreseq_b200.h into the sm_100a library.

Like the reference, every PSAlign method re-marshals the Python-side state into a fresh native
region (PythonToAlignData, pyx:139-153), and only Mutate / Refine / ApplyMuts write the realigned
events back (pyx:131-137, 374, 433, 470).

There is no CPU fallback: importing works anywhere, but any compute call raises RuntimeError when
the CUDA library is missing or no sm_100 device is present.
"""
import copy
import ctypes as C
import os

import numpy as np

from .Util import MutationInfo, MutationScore  # noqa: F401  (re-exported like the reference)

_HERE = os.path.dirname(os.path.abspath(__file__))
# PORESEQ_B200_LIB: an experiment build of the same library (python -m poreseq_b200.build -D... --out=...), never a fallback
_LIB_PATH = os.environ.get("PORESEQ_B200_LIB") or os.path.join(_HERE, "libporeseq_b200.so")
_c_double_p = C.POINTER(C.c_double)
_c_int_p = C.POINTER(C.c_int)

PS_T_NAMES = ["h2d", "centres", "forward", "backward", "backtrace", "join", "mutscore", "reduce", "d2h", "total"]


class PSParams(C.Structure):
    _fields_ = [("lik_offset", C.c_double), ("scoring_width", C.c_int), ("realign_width", C.c_int), ("verbose", C.c_int)]


class PSRegionDesc(C.Structure):
    _fields_ = [("bases", C.c_char_p), ("len", C.c_int), ("params", PSParams), ("n_events", C.c_int), ("n0", C.c_void_p),
                ("mean", C.c_void_p), ("stdv", C.c_void_p), ("ref_align", C.c_void_p), ("ref_like", C.c_void_p),
                ("model_index", C.c_void_p), ("n_models", C.c_int), ("models", C.c_void_p), ("probs", C.c_void_p),
                ("complement", C.c_void_p), ("seq2d", C.c_void_p)]


_lib = None


def lib():
    """The native library; raises loudly if it has not been built (python -m poreseq_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError("poreseq_b200: %s is missing -- build it with `python -m poreseq_b200.build`; "
                           "there is no CPU fallback" % _LIB_PATH)
    L = C.CDLL(_LIB_PATH)
    L.ps_create.restype = C.c_void_p
    L.ps_create.argtypes = [C.c_int]
    L.ps_destroy.argtypes = [C.c_void_p]
    L.ps_last_error.restype = C.c_char_p
    L.ps_last_error.argtypes = [C.c_void_p]
    L.ps_version.restype = C.c_char_p
    L.ps_launch_count.restype = C.c_longlong
    L.ps_launch_count.argtypes = [C.c_void_p]
    L.ps_set_precision.argtypes = [C.c_void_p, C.c_int]
    L.ps_last_timing.argtypes = [C.c_void_p, _c_double_p]
    L.ps_last_cells.argtypes = [C.c_void_p, _c_double_p, _c_double_p]
    L.ps_last_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.ps_region_create.restype = C.c_void_p
    L.ps_region_create.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(PSParams)]
    L.ps_region_destroy.argtypes = [C.c_void_p]
    L.ps_region_add_event.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8 + [C.c_int] + [C.c_double] * 4 + [C.c_char_p]
    L.ps_region_add_events.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                                                   C.c_void_p, C.c_void_p]
    L.ps_regions_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(PSRegionDesc), C.POINTER(C.c_void_p)]
    L.ps_regions_destroy.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.ps_region_set_params.argtypes = [C.c_void_p, C.POINTER(PSParams)]
    L.ps_region_num_events.argtypes = [C.c_void_p]
    L.ps_region_sequence_length.argtypes = [C.c_void_p]
    L.ps_region_get_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.ps_region_get_event_align.argtypes = [C.c_void_p, C.c_int, _c_double_p, _c_double_p]
    L.ps_score_alignments.argtypes = [C.c_void_p, _c_double_p, _c_double_p]
    L.ps_score_events.argtypes = [C.c_void_p, _c_double_p]
    L.ps_score_events_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, _c_double_p]
    L.ps_score_mutations.argtypes = [C.c_void_p, C.c_int, _c_int_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _c_double_p]
    L.ps_score_mutations_partial.argtypes = L.ps_score_mutations.argtypes
    L.ps_score_mutations_sharded.argtypes = L.ps_score_mutations.argtypes
    L.ps_consensus.argtypes = [C.c_void_p, C.c_int, C.c_int, _c_int_p]
    L.ps_consensus_batch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]
    L.ps_region_num_stages.argtypes = [C.c_void_p]
    L.ps_region_get_stage.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, _c_int_p]
    L.ps_comm_unique_id.argtypes = [C.c_char_p, C.c_int]
    L.ps_comm_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ps_comm_destroy.argtypes = [C.c_void_p]
    L.ps_comm_rank.argtypes = [C.c_void_p, _c_int_p, _c_int_p]
    L.ps_find_point_mutations.argtypes = [C.c_void_p, C.c_int, _c_int_p, _c_int_p, C.c_char_p, C.c_char_p]
    L.ps_score_points.argtypes = [C.c_void_p, C.c_int, _c_int_p, _c_int_p, C.c_char_p, C.c_char_p, _c_double_p]
    L.ps_score_points_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, _c_int_p, C.POINTER(C.c_longlong),
                                        _c_int_p, C.c_char_p, C.c_char_p, _c_double_p]
    L.ps_score_points_batch_begin.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, _c_int_p, C.POINTER(C.c_longlong),
                                              _c_int_p, C.c_char_p, C.c_char_p]
    L.ps_score_points_batch_end.argtypes = [C.c_void_p, _c_double_p]
    L.ps_score_points_direct_begin.argtypes = [C.c_void_p, C.c_int, C.POINTER(PSRegionDesc), C.c_int, _c_int_p, C.POINTER(C.c_longlong),
                                               _c_int_p, C.c_char_p, C.c_char_p]
    L.ps_score_points_direct_end.argtypes = [C.c_void_p, _c_double_p]
    L.ps_make_mutations.argtypes = [C.c_void_p, C.c_int, _c_int_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _c_double_p, _c_int_p]
    L.ps_refine.argtypes = [C.c_void_p, _c_int_p]
    L.ps_seq_to_states.argtypes = [C.c_char_p, C.c_int, _c_int_p]
    L.ps_swfull.argtypes = [C.c_char_p, C.c_char_p, _c_int_p, _c_int_p, C.c_int, _c_int_p, _c_int_p, _c_double_p]
    L.ps_swfull_device.argtypes = [C.c_void_p] + L.ps_swfull.argtypes
    L.ps_map_alignments.argtypes = [C.c_void_p, C.c_char_p]
    L.ps_find_mutations.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), _c_int_p]
    L.ps_found_mutation_sizes.argtypes = [C.c_void_p, C.c_int, _c_int_p, _c_int_p]
    L.ps_get_found_mutation.argtypes = [C.c_void_p, C.c_int, _c_int_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    L.ps_viterbi_positions.argtypes = [C.c_void_p, C.c_int, _c_int_p, _c_int_p, _c_int_p]
    L.ps_band_centres.argtypes = [C.c_void_p, C.c_int, C.c_int, _c_int_p, _c_double_p, _c_int_p, _c_int_p]
    L.ps_pick_candidates.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), _c_double_p, C.POINTER(_c_double_p), _c_int_p]
    L.ps_mutate.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, _c_int_p]
    L.ps_viterbi_mutate.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _c_int_p]
    L.ps_get_viterbi_sequence.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    L.ps_pack_open.restype = C.c_void_p
    L.ps_pack_open.argtypes = [C.c_char_p]
    L.ps_pack_close.argtypes = [C.c_void_p]
    L.ps_pack_num_regions.argtypes = [C.c_void_p]
    L.ps_pack_region_desc.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(PSRegionDesc)]
    L.ps_pack_region_param.argtypes = [C.c_void_p, C.c_int, C.c_char_p, _c_double_p]
    L.ps_pack_event_sequence.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), _c_int_p]
    L.ps_pack_regions_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    _lib = L
    return L


class Context(object):
    """One CUDA device + stream + scratch memory (ps_ctx)."""

    def __init__(self, device=0):
        self.lib = lib()
        self.handle = self.lib.ps_create(int(device))
        if not self.handle:
            raise RuntimeError("ps_create failed: %s" % self.lib.ps_last_error(None).decode())
        self.device = device

    def close(self):
        if self.handle:
            self.lib.ps_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("poreseq_b200 error %d: %s" % (rc, self.lib.ps_last_error(self.handle).decode()))

    def set_precision(self, mode):
        """'exact' (default, bit-identical FP64) or 'fast' (FP32 scan + exact re-score of candidates)."""
        self.check(self.lib.ps_set_precision(self.handle, {"exact": 0, "fast": 1}[mode]))

    def comm_init(self, unique_id, rank, n_ranks, ordered=True):
        """Joins the event-shard communicator of `n_ranks` contexts (one per GPU): ps_comm_init.  `unique_id` is what
        rank 0 got from comm_unique_id(), carried to the other ranks by the caller."""
        self.check(self.lib.ps_comm_init(self.handle, unique_id, len(unique_id), int(rank), int(n_ranks), 1 if ordered else 0))

    def comm_destroy(self):
        self.check(self.lib.ps_comm_destroy(self.handle))

    def launch_count(self):
        return int(self.lib.ps_launch_count(self.handle))

    def last_timing(self):
        ms = (C.c_double * len(PS_T_NAMES))()
        self.check(self.lib.ps_last_timing(self.handle, ms))
        return dict(zip(PS_T_NAMES, list(ms)))

    def last_bytes(self):
        """(host->device, device->host) bytes copied by the last batch on this context."""
        h, d = C.c_longlong(0), C.c_longlong(0)
        self.check(self.lib.ps_last_bytes(self.handle, C.byref(h), C.byref(d)))
        return h.value, d.value

    def last_cells(self):
        w, n = C.c_double(0), C.c_double(0)
        self.check(self.lib.ps_last_cells(self.handle, C.byref(w), C.byref(n)))
        return w.value, n.value


_default_ctx = {}


def comm_unique_id():
    """The 128-byte NCCL id rank 0 creates for ps_comm_init (ps_comm_unique_id)."""
    buf = C.create_string_buffer(128)
    rc = lib().ps_comm_unique_id(buf, 128)
    if rc != 0:
        raise RuntimeError("poreseq_b200 error %d: %s" % (rc, lib().ps_last_error(None).decode()))
    return buf.raw


def default_context(device=None):
    """Per-process context for a device (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("PORESEQ_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


_F8 = np.dtype("f8")


def _f8(a):
    if type(a) is np.ndarray and a.dtype == _F8 and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dtype="f8")


def _addr(a):
    return a.__array_interface__["data"][0]


def _cstrs(items):
    arr = (C.c_char_p * max(len(items), 1))()
    for i, s in enumerate(items):
        arr[i] = s.encode("ascii") if isinstance(s, str) else bytes(s)
    return arr


class PackedRegion(object):
    """Host buffers of one region laid out for ps_region_add_events: the level arrays of all events
    concatenated, one table of distinct pore models.  Built once from a PSAlign-like object; the
    marshalling of a call is then one C-ABI call per region instead of one per event."""

    def __init__(self, sequence, events, params):
        self.sequence = sequence.encode("ascii") if isinstance(sequence, str) else bytes(sequence)
        self.params = dict(params)
        self.n0 = np.array([len(ev.mean) for ev in events], dtype=np.int32)
        cat = lambda name: np.ascontiguousarray(np.concatenate([_f8(getattr(ev, name)) for ev in events]) if len(events) else np.zeros(0))
        self.mean, self.stdv, self.ref_align, self.ref_like = cat("mean"), cat("stdv"), cat("ref_align"), cat("ref_like")
        seen, tables, probs, index = {}, [], [], []
        for ev in events:
            m = ev.model
            if id(m) not in seen:
                seen[id(m)] = len(tables)
                tables.append(np.stack([_f8(m.level_mean)[:1024], _f8(m.level_stdv)[:1024], _f8(m.sd_mean)[:1024], _f8(m.sd_stdv)[:1024]]))
                probs.append([float(m.prob_skip), float(m.prob_stay), float(m.prob_extend), float(m.prob_insert)])
            index.append(seen[id(m)])
        self.model_index = np.array(index, dtype=np.int32)
        self.models = np.ascontiguousarray(np.stack(tables)) if tables else np.zeros((0, 4, 1024))
        self.probs = np.ascontiguousarray(np.array(probs, dtype="f8").reshape(-1, 4))
        self.complement = np.array([int(bool(ev.model.complement)) for ev in events], dtype=np.int32)
        self.seq2d = [(getattr(ev, "sequence", "") or "") for ev in events]
        self._seq2d_c = _cstrs(self.seq2d)

    @classmethod
    def from_arrays(cls, sequence, params, n0, mean, stdv, ref_align, ref_like, model_index, models, probs, complement, seq2d):
        """A PackedRegion over arrays that already have the packed layout (e.g. views into an event-pack file,
        poreseq_b200/eventpack.py): nothing is copied, the arrays must stay alive as long as the object."""
        self = cls.__new__(cls)
        self.sequence = sequence.encode("ascii") if isinstance(sequence, str) else bytes(sequence)
        self.params = dict(params)
        self.n0 = np.ascontiguousarray(n0, dtype=np.int32)
        self.mean, self.stdv, self.ref_align, self.ref_like = (np.ascontiguousarray(a, dtype="f8") for a in (mean, stdv, ref_align, ref_like))
        self.model_index = np.ascontiguousarray(model_index, dtype=np.int32)
        self.models = np.ascontiguousarray(models, dtype="f8").reshape(-1, 4, 1024)
        self.probs = np.ascontiguousarray(probs, dtype="f8").reshape(-1, 4)
        self.complement = np.ascontiguousarray(complement, dtype=np.int32)
        self.seq2d = list(seq2d)
        self._seq2d_c = _cstrs(self.seq2d)
        return self

    def nbytes(self):
        return (self.mean.nbytes + self.stdv.nbytes + self.ref_align.nbytes + self.ref_like.nbytes + self.models.nbytes +
                self.probs.nbytes + self.n0.nbytes + self.model_index.nbytes + self.complement.nbytes + len(self.sequence))


def _ps_params(params, width_key):
    p = PSParams(4.5, 150, 300, 0)          # cpp/AlignUtil.h:64 defaults
    if "verbose" in params:
        p.verbose = int(params["verbose"])
    if "lik_offset" in params:
        p.lik_offset = float(params["lik_offset"])
    if "realign_width" in params:
        p.realign_width = int(params["realign_width"])
    if "scoring_width" in params:
        p.scoring_width = int(params["scoring_width"])
    if width_key is not None and width_key in params:   # point_width override, pyx:293,361,465
        p.scoring_width = int(params[width_key])
    return p


class NativeRegion(object):
    """ps_region built from a PSAlign-like object (sequence, events, params)."""

    @classmethod
    def from_packed(cls, ctx, pack, width_key=None):
        """One ps_region_create + one ps_region_add_events from the host buffers of a PackedRegion."""
        self = cls.__new__(cls)
        self.ctx = ctx
        L = ctx.lib
        p = _ps_params(pack.params, width_key)
        self.handle = L.ps_region_create(ctx.handle, pack.sequence, len(pack.sequence), C.byref(p))
        if not self.handle:
            raise RuntimeError("ps_region_create failed: %s" % L.ps_last_error(ctx.handle).decode())
        self.n_levels = pack.n0.tolist()
        ctx.check(L.ps_region_add_events(self.handle, len(pack.n0), pack.n0.ctypes.data, pack.mean.ctypes.data,
                                         pack.stdv.ctypes.data, pack.ref_align.ctypes.data, pack.ref_like.ctypes.data,
                                         pack.model_index.ctypes.data, len(pack.models), pack.models.ctypes.data,
                                         pack.probs.ctypes.data, pack.complement.ctypes.data,
                                         C.cast(pack._seq2d_c, C.c_void_p)))
        return self

    def __init__(self, ctx, sequence, events, params, width_key=None):
        self.ctx = ctx
        L = ctx.lib
        p = PSParams(4.5, 150, 300, 0)          # cpp/AlignUtil.h:64 defaults
        if "verbose" in params:
            p.verbose = int(params["verbose"])
        if "lik_offset" in params:
            p.lik_offset = float(params["lik_offset"])
        if "realign_width" in params:
            p.realign_width = int(params["realign_width"])
        if "scoring_width" in params:
            p.scoring_width = int(params["scoring_width"])
        if width_key is not None and width_key in params:   # point_width override, pyx:293,361,465
            p.scoring_width = int(params[width_key])
        seq = sequence.encode("ascii") if isinstance(sequence, str) else bytes(sequence)
        self.handle = L.ps_region_create(ctx.handle, seq, len(seq), C.byref(p))
        if not self.handle:
            raise RuntimeError("ps_region_create failed: %s" % L.ps_last_error(ctx.handle).decode())
        self.n_levels = []
        for ev in events:
            mean, stdv, ra, rl = _f8(ev.mean), _f8(ev.stdv), _f8(ev.ref_align), _f8(ev.ref_like)
            m = ev.model
            lm, ls, sm, ss = _f8(m.level_mean), _f8(m.level_stdv), _f8(m.sd_mean), _f8(m.sd_stdv)
            if not (len(stdv) == len(ra) == len(rl) == len(mean)) or min(len(lm), len(ls), len(sm), len(ss)) < 1024:
                raise ValueError("event arrays must have equal length and models 1024 states")
            seq2d = getattr(ev, "sequence", "") or ""
            ctx.check(L.ps_region_add_event(self.handle, len(mean), _addr(mean), _addr(stdv), _addr(ra), _addr(rl),
                                            _addr(lm), _addr(ls), _addr(sm), _addr(ss), int(bool(m.complement)),
                                            float(m.prob_skip), float(m.prob_stay), float(m.prob_extend),
                                            float(m.prob_insert), seq2d.encode("ascii")))
            self.n_levels.append(len(mean))

    def close(self):
        if self.handle:
            self.ctx.lib.ps_region_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sequence(self):
        n = self.ctx.lib.ps_region_sequence_length(self.handle)
        buf = C.create_string_buffer(n + 1)
        self.ctx.check(self.ctx.lib.ps_region_get_sequence(self.handle, buf, n + 1))
        return buf.value.decode("ascii")

    def event_align(self, e):
        ra = np.zeros(self.n_levels[e])
        rl = np.zeros(self.n_levels[e])
        self.ctx.check(self.ctx.lib.ps_region_get_event_align(self.handle, e, _dp(ra), _dp(rl)))
        return ra, rl

    def write_back(self, events):
        """UpdatePythonEvents (pyx:131-137): in-place ref_align / ref_like."""
        for e, ev in enumerate(events):
            ra, rl = self.event_align(e)
            ev.ref_align[:] = ra
            ev.ref_like[:] = rl

    # -- compute ---------------------------------------------------------------------------
    def score_alignments(self, want_likes=False):
        n = len(self.n_levels)
        scores = np.zeros(n)
        likes = np.zeros(self.ctx.lib.ps_region_sequence_length(self.handle)) if want_likes else None
        self.ctx.check(self.ctx.lib.ps_score_alignments(self.handle, _dp(scores), _dp(likes) if want_likes else None))
        return scores, likes

    def score_events(self):
        """PSAlign.ScoreEvents: one score per event, alignments untouched (ps_score_events)."""
        n = self.ctx.lib.ps_region_num_events(self.handle)
        out = np.zeros(max(n, 1))
        self.ctx.check(self.ctx.lib.ps_score_events(self.handle, _dp(out)))
        return out[:n]

    def score_mutations(self, starts, origs, muts):
        n = len(starts)
        st = np.ascontiguousarray(starts, dtype=np.int32)
        scores = np.zeros(n)
        self.ctx.check(self.ctx.lib.ps_score_mutations(self.handle, n, st.ctypes.data_as(_c_int_p), _cstrs(origs),
                                                       _cstrs(muts), _dp(scores)))
        return scores

    def score_mutations_partial(self, starts, origs, muts):
        """Per-mutation sums over this region's events only, starting at 0 (for event-sharded scoring)."""
        n = len(starts)
        st = np.ascontiguousarray(starts, dtype=np.int32)
        out = np.zeros(n)
        self.ctx.check(self.ctx.lib.ps_score_mutations_partial(self.handle, n, st.ctypes.data_as(_c_int_p), _cstrs(origs),
                                                               _cstrs(muts), _dp(out)))
        return out

    def consensus(self, reps=4, point_width=20):
        """The whole Mutate.py loop on this handle (ps_consensus); returns the stages [(name, sequence, bases changed)]."""
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.ps_consensus(self.handle, int(reps), int(point_width), C.byref(n)))
        return self.stages()

    def stages(self):
        out = []
        for k in range(self.ctx.lib.ps_region_num_stages(self.handle)):
            ln = self.ctx.lib.ps_region_get_stage(self.handle, k, None, 0, None, 0, None)
            name, seq, nb = C.create_string_buffer(64), C.create_string_buffer(ln + 1), C.c_int(0)
            self.ctx.lib.ps_region_get_stage(self.handle, k, name, 64, seq, ln + 1, C.byref(nb))
            out.append((name.value.decode(), seq.value.decode(), nb.value))
        return out

    def score_mutations_sharded(self, starts, origs, muts):
        """This handle holds this rank's block of the region's events; complete scores come back on every rank
        (ps_score_mutations_sharded, sums combined over NCCL inside the library)."""
        n = len(starts)
        st = np.ascontiguousarray(starts, dtype=np.int32)
        out = np.zeros(n)
        self.ctx.check(self.ctx.lib.ps_score_mutations_sharded(self.handle, n, st.ctypes.data_as(_c_int_p), _cstrs(origs),
                                                               _cstrs(muts), _dp(out)))
        return out

    def score_points(self):
        cap = 9 * max(self.ctx.lib.ps_region_sequence_length(self.handle), 1)   # 8 per state, 9 where the base is not ACGT
        n = C.c_int(0)
        st = np.zeros(cap, dtype=np.int32)
        og = C.create_string_buffer(cap)
        mu = C.create_string_buffer(cap)
        sc = np.zeros(cap)
        self.ctx.check(self.ctx.lib.ps_score_points(self.handle, cap, C.byref(n), st.ctypes.data_as(_c_int_p), og, mu, _dp(sc)))
        k = n.value
        return st[:k], og.raw[:k], mu.raw[:k], sc[:k]

    def make_mutations(self, starts, origs, muts, scores):
        n = len(starts)
        st = np.ascontiguousarray(starts, dtype=np.int32)
        sc = np.ascontiguousarray(scores, dtype="f8")
        nb = C.c_int(0)
        self.ctx.check(self.ctx.lib.ps_make_mutations(self.handle, n, st.ctypes.data_as(_c_int_p), _cstrs(origs),
                                                      _cstrs(muts), _dp(sc), C.byref(nb)))
        return nb.value

    def refine(self):
        nb = C.c_int(0)
        self.ctx.check(self.ctx.lib.ps_refine(self.handle, C.byref(nb)))
        return nb.value

    def map_alignments(self, newseq):
        self.ctx.check(self.ctx.lib.ps_map_alignments(self.handle, newseq.encode("ascii")))

    def band_centres(self, event, n_cols):
        """ps_band_centres: (centres[n_cols], ref_index or None when the event carries no alignment, monotone)."""
        cen = np.zeros(max(n_cols, 1), dtype=np.int32)
        ra, _ = self.event_align(event)
        ri = np.zeros(max(len(ra), 1), dtype="f8")
        empty, mono = C.c_int(0), C.c_int(0)
        self.ctx.check(self.ctx.lib.ps_band_centres(self.handle, int(event), int(n_cols), cen.ctypes.data_as(_c_int_p), _dp(ri),
                                                    C.byref(empty), C.byref(mono)))
        return cen[:n_cols], (None if empty.value else ri[:len(ra)]), bool(mono.value)

    def viterbi_positions(self):
        """ps_viterbi_positions: [(position, reads sitting on it)] as ViterbiMutate's host half keeps them."""
        n = C.c_int(0)
        cap = 16
        while True:
            pos = np.zeros(cap, dtype=np.int32)
            cnt = np.zeros(cap, dtype=np.int32)
            rc = self.ctx.lib.ps_viterbi_positions(self.handle, cap, pos.ctypes.data_as(_c_int_p), cnt.ctypes.data_as(_c_int_p), C.byref(n))
            if rc == -3 and n.value > cap:                  # PS_E_CAPACITY
                cap = n.value
                continue
            self.ctx.check(rc)
            return list(zip(pos[:n.value].tolist(), cnt[:n.value].tolist()))

    def pick_candidates(self, seeds, base_profile, seed_profiles):
        """ps_pick_candidates: the host-only second half of FindMutations over profiles the caller already holds."""
        n = C.c_int(0)
        base = np.ascontiguousarray(base_profile, dtype="f8")
        profs = [np.ascontiguousarray(p, dtype="f8") for p in seed_profiles]
        arr = (_c_double_p * max(len(profs), 1))(*[_dp(p) for p in profs])
        self.ctx.check(self.ctx.lib.ps_pick_candidates(self.handle, len(seeds), _cstrs(seeds), _dp(base), arr, C.byref(n)))
        return self._found(n.value)

    def find_mutations(self, seeds):
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.ps_find_mutations(self.handle, len(seeds), _cstrs(seeds), C.byref(n)))
        return self._found(n.value)

    def _found(self, count):
        n = C.c_int(count)
        out = []
        for i in range(n.value):
            no, nm, st = C.c_int(0), C.c_int(0), C.c_int(0)
            self.ctx.check(self.ctx.lib.ps_found_mutation_sizes(self.handle, i, C.byref(no), C.byref(nm)))
            ob, mb = C.create_string_buffer(no.value + 1), C.create_string_buffer(nm.value + 1)
            self.ctx.check(self.ctx.lib.ps_get_found_mutation(self.handle, i, C.byref(st), ob, no.value + 1, mb, nm.value + 1))
            out.append((st.value, ob.value.decode("ascii"), mb.value.decode("ascii")))
        return out

    def viterbi_mutate(self, nkeep=16, skip=0.05, stay=0.01, mut_min=0.33, mut_max=0.75):
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.ps_viterbi_mutate(self.handle, int(nkeep), skip, stay, mut_min, mut_max, C.byref(n)))
        out = []
        for i in range(n.value):
            ln = self.ctx.lib.ps_get_viterbi_sequence(self.handle, i, None, 0)
            buf = C.create_string_buffer(ln + 1)
            self.ctx.lib.ps_get_viterbi_sequence(self.handle, i, buf, ln + 1)
            out.append(buf.value.decode("ascii"))
        return out

    def mutate(self, seeds, reps=4):
        tot = C.c_int(0)
        self.ctx.check(self.ctx.lib.ps_mutate(self.handle, len(seeds), _cstrs(seeds), int(reps), C.byref(tot)))
        return tot.value


def _region_descs(packs, width_key=None):
    n = len(packs)
    desc = (PSRegionDesc * n)()
    for k, p in enumerate(packs):
        d = desc[k]
        d.bases, d.len, d.params = p.sequence, len(p.sequence), _ps_params(p.params, width_key)
        d.n_events, d.n0 = len(p.n0), p.n0.ctypes.data
        d.mean, d.stdv, d.ref_align, d.ref_like = p.mean.ctypes.data, p.stdv.ctypes.data, p.ref_align.ctypes.data, p.ref_like.ctypes.data
        d.model_index, d.n_models, d.models, d.probs = p.model_index.ctypes.data, len(p.models), p.models.ctypes.data, p.probs.ctypes.data
        d.complement, d.seq2d = p.complement.ctypes.data, C.cast(p._seq2d_c, C.c_void_p)
    return desc


def native_regions_from_packed(ctx, packs, width_key=None):
    """ps_regions_create: fresh native regions for a batch of PackedRegion host buffers in ONE C-ABI call
    (the library copies them in on its worker threads)."""
    n = len(packs)
    desc = _region_descs(packs, width_key)
    out = (C.c_void_p * n)()
    ctx.check(ctx.lib.ps_regions_create(ctx.handle, n, desc, out))
    regs = []
    for k, p in enumerate(packs):
        r = NativeRegion.__new__(NativeRegion)
        r.ctx, r.handle, r.n_levels = ctx, out[k], p.n0.tolist()
        regs.append(r)
    return regs


class PendingDirect(object):
    """PSAlign.ScorePoints of many regions straight from host buffers (ps_score_points_direct_begin / _end): no native
    region objects, scores only (the reference drops the realignment of ScorePoints too, pyx:278-308)."""

    def __init__(self, ctx, packs, width_key="point_width"):
        self.ctx, self.packs = ctx, packs                         # (the host buffers must outlive the call)
        n = len(packs)
        self.desc = _region_descs(packs, width_key)
        cap = sum(9 * max(len(p.sequence), 1) for p in packs)
        self.n_out = (C.c_int * n)()
        self.off = (C.c_longlong * n)()
        self.st = np.zeros(cap, dtype=np.int32)
        self.og = C.create_string_buffer(cap)
        self.mu = C.create_string_buffer(cap)
        self.sc = np.zeros(cap)
        ctx.check(ctx.lib.ps_score_points_direct_begin(ctx.handle, n, self.desc, cap, self.n_out, self.off,
                                                       self.st.ctypes.data_as(_c_int_p), self.og, self.mu))

    def end(self):
        self.ctx.check(self.ctx.lib.ps_score_points_direct_end(self.ctx.handle, _dp(self.sc)))
        out = []
        og, mu = self.og.raw, self.mu.raw
        for k in range(len(self.packs)):
            a, b = self.off[k], self.off[k] + self.n_out[k]
            out.append((self.st[a:b], og[a:b], mu[a:b], self.sc[a:b]))
        return out


def score_points_direct(ctx, packs, width_key="point_width"):
    return PendingDirect(ctx, packs, width_key).end()


def close_regions(regions):
    """ps_regions_destroy: release a batch of NativeRegion objects in one C-ABI call."""
    live = [r for r in regions if r.handle]
    if not live:
        return
    handles = (C.c_void_p * len(live))(*[r.handle for r in live])
    live[0].ctx.lib.ps_regions_destroy(handles, len(live))
    for r in live:
        r.handle = None


def score_points_batch(ctx, regions):
    """ps_score_points_batch over NativeRegion objects: one launch sequence for all of them.
    Returns a list of (start, orig bytes, mut bytes, score) tuples of arrays, one per region."""
    n = len(regions)
    handles = (C.c_void_p * n)(*[r.handle for r in regions])
    cap = sum(9 * max(ctx.lib.ps_region_sequence_length(r.handle), 1) for r in regions)
    n_out = (C.c_int * n)()
    off = (C.c_longlong * n)()
    st = np.zeros(cap, dtype=np.int32)
    og = C.create_string_buffer(cap)
    mu = C.create_string_buffer(cap)
    sc = np.zeros(cap)
    ctx.check(ctx.lib.ps_score_points_batch(handles, n, cap, n_out, off, st.ctypes.data_as(_c_int_p), og, mu, _dp(sc)))
    out = []
    ogr, mur = og.raw, mu.raw
    for k in range(n):
        a, b = off[k], off[k] + n_out[k]
        out.append((st[a:b], ogr[a:b], mur[a:b], sc[a:b]))
    return out


def score_events_batch(ctx, regions):
    """ps_score_events_batch over NativeRegion objects: [scores per event] per region, one launch sequence."""
    n = len(regions)
    handles = (C.c_void_p * n)(*[r.handle for r in regions])
    counts = [ctx.lib.ps_region_num_events(r.handle) for r in regions]
    out = np.zeros(max(sum(counts), 1))
    ctx.check(ctx.lib.ps_score_events_batch(handles, n, _dp(out)))
    res, at = [], 0
    for c in counts:
        res.append(out[at:at + c])
        at += c
    return res


def consensus_batch(ctx, regions, reps=4, point_width=20, in_flight=16):
    """ps_consensus_batch over NativeRegion objects of `ctx`: the Mutate.py loop of every region, `in_flight` regions side
    by side on the GPU.  Afterwards every region's .sequence() / .event_align() / .stages() hold its result."""
    n = len(regions)
    handles = (C.c_void_p * n)(*[r.handle for r in regions])
    ctx.check(ctx.lib.ps_consensus_batch(ctx.handle, handles, n, int(reps), int(point_width), int(in_flight)))


class PendingBatch(object):
    """A ps_score_points_batch in flight (ps_score_points_batch_begin / _end)."""

    def __init__(self, ctx, regions):
        self.ctx, self.regions = ctx, regions
        n = len(regions)
        handles = (C.c_void_p * n)(*[r.handle for r in regions])
        cap = sum(9 * max(ctx.lib.ps_region_sequence_length(r.handle), 1) for r in regions)
        self.n_out = (C.c_int * n)()
        self.off = (C.c_longlong * n)()
        self.st = np.zeros(cap, dtype=np.int32)
        self.og = C.create_string_buffer(cap)
        self.mu = C.create_string_buffer(cap)
        self.sc = np.zeros(cap)
        ctx.check(ctx.lib.ps_score_points_batch_begin(handles, n, cap, self.n_out, self.off,
                                                      self.st.ctypes.data_as(_c_int_p), self.og, self.mu))

    def end(self):
        """Wait for the batch; returns [(start, orig bytes, mut bytes, score)] per region."""
        self.ctx.check(self.ctx.lib.ps_score_points_batch_end(self.ctx.handle, _dp(self.sc)))
        out = []
        og, mu = self.og.raw, self.mu.raw                 # one copy of each buffer, not one per region
        for k in range(len(self.regions)):
            a, b = self.off[k], self.off[k] + self.n_out[k]
            out.append((self.st[a:b], og[a:b], mu[a:b], self.sc[a:b]))
        return out


def score_points_batch_begin(ctx, regions):
    """Asynchronous ps_score_points_batch: stage + enqueue now, `.end()` later.  With two contexts on
    one device the host can stage batch k+1 while the GPU works on batch k."""
    return PendingBatch(ctx, regions)


def _score_objects(starts, origs, muts, scores):
    out = []
    for s, o, m, sc in zip(starts, origs, muts, scores):
        ms = MutationScore()
        ms.start = int(s)
        ms.orig = o
        ms.mut = m
        ms.score = float(sc)
        out.append(ms)
    return out


def _ch(b):
    return "" if b == 0 else chr(b)


def swalign(seq1, seq2):
    """Smith-Waterman align two sequences (pyx:155-174, cpp/swlib.cpp:211-340).

    Returns (accuracy in %, list of pairwise aligned 1-based indices, 0 = gap)."""
    a = seq1.encode("ascii") if isinstance(seq1, str) else bytes(seq1)
    b = seq2.encode("ascii") if isinstance(seq2, str) else bytes(seq2)
    cap = len(a) + len(b) + 8
    i1 = np.zeros(cap, dtype=np.int32)
    i2 = np.zeros(cap, dtype=np.int32)
    n, score, acc = C.c_int(0), C.c_int(0), C.c_double(0)
    rc = lib().ps_swfull(a, b, i1.ctypes.data_as(_c_int_p), i2.ctypes.data_as(_c_int_p), cap, C.byref(n), C.byref(score), C.byref(acc))
    if rc != 0:
        raise RuntimeError("ps_swfull failed (%d)" % rc)
    return (acc.value, list(zip(i1[:n.value].tolist(), i2[:n.value].tolist())))


def swalign_device(ctx, seq1, seq2):
    """swalign computed on the GPU (ps_swfull_device); returns (score, accuracy, index pairs)."""
    a = seq1.encode("ascii") if isinstance(seq1, str) else bytes(seq1)
    b = seq2.encode("ascii") if isinstance(seq2, str) else bytes(seq2)
    cap = len(a) + len(b) + 8
    i1 = np.zeros(cap, dtype=np.int32)
    i2 = np.zeros(cap, dtype=np.int32)
    n, score, acc = C.c_int(0), C.c_int(0), C.c_double(0)
    ctx.check(ctx.lib.ps_swfull_device(ctx.handle, a, b, i1.ctypes.data_as(_c_int_p), i2.ctypes.data_as(_c_int_p), cap,
                                       C.byref(n), C.byref(score), C.byref(acc)))
    return score.value, acc.value, list(zip(i1[:n.value].tolist(), i2[:n.value].tolist()))


def seqtostates(seq):
    """5-mer states [0,1023] of a nucleotide sequence (pyx:176-187, cpp/Sequence.h:69-100)."""
    s = seq.encode("ascii") if isinstance(seq, str) else bytes(seq)
    out = np.zeros(max(len(s), 1), dtype=np.int32)
    n = lib().ps_seq_to_states(s, len(s), out.ctypes.data_as(_c_int_p))
    return out[:max(n, 0)].tolist()


class PSAlign(object):
    """All data associated with reads aligned to a reference (pyx:189-472).

    Attributes: sequence (str), events (list of PSEvent-like), params (dict)."""

    def __init__(self):
        self.sequence = ""
        self.events = []
        self.params = {}
        self.ctx = None          # optional Context of its own (one per host thread); default: the process-wide one

    # -- plumbing --------------------------------------------------------------------------
    def _native(self, width_key=None):
        return NativeRegion(self.ctx or default_context(), self.sequence, self.events, self.params, width_key)

    def __deepcopy__(self, memo):
        other = PSAlign()
        other.sequence, other.params, other.ctx = self.sequence, dict(self.params), self.ctx
        other.events = copy.deepcopy(self.events, memo)
        return other

    def Copy(self):
        return copy.deepcopy(self)

    def Coverage(self):
        """Depth of coverage across the reference (pyx:233-247)."""
        cov = np.zeros(len(self.sequence))
        for ev in self.events:
            nzs = ev.ref_align[ev.ref_align > 0]
            if len(nzs) == 0:
                continue
            lo = int(nzs[0])
            hi = int(np.minimum(nzs[-1], len(cov) - 1))
            cov[lo:hi] += 1
        return cov

    def RealignTo(self, newseq):
        """Realign all events to a new reference sequence using swalign (pyx:249-261)."""
        align = swalign(self.sequence, newseq)
        if align[0] < 0.6:
            raise Exception('Error rate too large for realignment!')
        for ev in self.events:
            ev.mapaligns(np.array(align[1]))
        self.sequence = newseq

    # -- scoring ---------------------------------------------------------------------------
    def ScoreEvents(self):
        """Likelihood score of every event (pyx:263-276).  Like the reference, the realignment is
        not propagated back to the Python events."""
        reg = self._native()
        try:
            scores = reg.score_events()
        finally:
            reg.close()
        return scores.tolist()

    def ScorePoints(self):
        """Scores of all single-base mutations, at point_width (pyx:278-308)."""
        reg = self._native("point_width")
        try:
            st, og, mu, sc = reg.score_points()
        finally:
            reg.close()
        return _score_objects(st, [_ch(b) for b in og], [_ch(b) for b in mu], sc)

    def ScoreMutations(self, muts):
        """Scores of the given MutationInfo list, at scoring_width (pyx:310-345)."""
        reg = self._native()
        try:
            starts = [int(m.start) for m in muts]
            origs = [m.orig for m in muts]
            mutss = [m.mut for m in muts]
            sc = reg.score_mutations(starts, origs, mutss)
        finally:
            reg.close()
        return _score_objects(starts, origs, mutss, sc)

    def ApplyMuts(self, pymuts):
        """MakeMutations on an already scored list, at point_width (pyx:347-375)."""
        reg = self._native("point_width")
        try:
            reg.make_mutations([int(m.start) for m in pymuts], [m.orig for m in pymuts],
                               [m.mut for m in pymuts], [float(m.score) for m in pymuts])
            self.sequence = reg.sequence()
            reg.write_back(self.events)
        finally:
            reg.close()

    def Mutate(self, seqs='self', reps=4):
        """Use similar sequences as seeds to mutate the consensus (pyx:378-435).

        seqs: 'self' (2D sequences of every other event), 'viterbi', or a list of sequences."""
        reg = self._native()
        try:
            if isinstance(seqs, str) and seqs == 'self':
                seeds = [ev.sequence for ev in self.events[::2]]
            elif isinstance(seqs, str) and seqs == 'viterbi':
                seeds = reg.viterbi_mutate(16, 0.05, 0.01, 0.33, 0.75)
            else:
                seeds = list(seqs)
            tot = reg.mutate(seeds, reps)
            self.sequence = reg.sequence()
            reg.write_back(self.events)
        finally:
            reg.close()
        return tot

    def Refine(self):
        """Test all single-base mutations and make the improving ones (pyx:437-472)."""
        reg = self._native("point_width")
        try:
            nbases = reg.refine()
            self.sequence = reg.sequence()
            reg.write_back(self.events)
        finally:
            reg.close()
        return nbases
