# cython: language_level=3
# distutils: language = c
"""The binding a PoreSeq maintainer would add (INTEGRATION.md section 2), as a compilable module.

It replaces the `cdef extern from "cpp/..."` blocks of poreseq/_poreseqcpp.pyx:15-83 with ONE block over
include/poreseq_b200.h and keeps the Python surface of that module: `swalign`, `seqtostates`, and `PSAlign`
(shown here with ScoreEvents and Refine; the other methods follow the same three steps: build the region, make one
ps_* call, read the results back).  Built and exercised by
tests/test_host_and_abi.py::test_cython_stub_of_integration_md_builds_and_binds."""
from libc.stdlib cimport malloc, free
import numpy as np
cimport numpy as cnp

cdef extern from "poreseq_b200.h":
    ctypedef struct ps_ctx
    ctypedef struct ps_region
    ctypedef struct ps_params:
        double lik_offset
        int scoring_width
        int realign_width
        int verbose
    ps_ctx* ps_create(int device)
    void ps_destroy(ps_ctx*)
    const char* ps_last_error(ps_ctx*)
    ps_region* ps_region_create(ps_ctx*, const char* bases, int len, const ps_params*)
    void ps_region_destroy(ps_region*)
    int ps_region_add_event(ps_region*, int n0, const double* mean, const double* stdv,
                            const double* ref_align, const double* ref_like,
                            const double* level_mean, const double* level_stdv,
                            const double* sd_mean, const double* sd_stdv, int complement,
                            double prob_skip, double prob_stay, double prob_extend, double prob_insert,
                            const char* seq2d)
    int ps_region_num_events(ps_region*)
    int ps_region_get_sequence(ps_region*, char* out, int cap)
    int ps_region_sequence_length(ps_region*)
    int ps_region_get_event_align(ps_region*, int e, double* ref_align, double* ref_like)
    int ps_score_alignments(ps_region*, double* scores, double* likes)
    int ps_score_events(ps_region*, double* scores)
    int ps_refine(ps_region*, int* nbases)
    int ps_swfull(const char* s1, const char* s2, int* inds1, int* inds2, int cap, int* n, int* score, double* acc)
    int ps_seq_to_states(const char* seq, int len, int* states)

cdef ps_ctx* _ctx = ps_create(0)          # one context per process; CUDA starts on the first compute call


cdef double* getPr(cnp.ndarray arr):      # as poreseq/_poreseqcpp.pyx:86-88
    return <double*>arr.data


def _error():
    return ps_last_error(_ctx).decode()


def swalign(seq1, seq2):
    """(accuracy %, [(i, j) ...]) -- poreseq/_poreseqcpp.pyx:155-174"""
    a, b = seq1.encode('ascii'), seq2.encode('ascii')
    cdef int cap = len(a) + len(b) + 8, n = 0, score = 0
    cdef double acc = 0
    cdef int* i1 = <int*>malloc(cap * sizeof(int))
    cdef int* i2 = <int*>malloc(cap * sizeof(int))
    try:
        if ps_swfull(a, b, i1, i2, cap, &n, &score, &acc) != 0:
            raise RuntimeError("ps_swfull failed")
        return (acc, [(i1[k], i2[k]) for k in range(n)])
    finally:
        free(i1)
        free(i2)


def seqtostates(seq):
    """5-mer states of a sequence -- poreseq/_poreseqcpp.pyx:176-187"""
    s = seq.encode('ascii')
    cdef int n = len(s)
    cdef int* st = <int*>malloc(max(n, 1) * sizeof(int))
    try:
        n = ps_seq_to_states(s, n, st)
        return [st[k] for k in range(max(n, 0))]
    finally:
        free(st)


cdef ps_region* PythonToRegion(obj, width_key=None) except NULL:      # was PythonToAlignData, pyx:139-153
    cdef ps_params p
    p.lik_offset = obj.params.get('lik_offset', 4.5)
    p.scoring_width = int(obj.params.get('scoring_width', 150))
    if width_key is not None and width_key in obj.params:
        p.scoring_width = int(obj.params[width_key])                   # pyx:293,361,465
    p.realign_width = int(obj.params.get('realign_width', 300))
    p.verbose = int(obj.params.get('verbose', 0))
    seq = obj.sequence.encode('ascii')
    cdef ps_region* r = ps_region_create(_ctx, seq, len(seq), &p)
    if r == NULL:
        raise RuntimeError(_error())
    for pyev in obj.events:
        pyev.makecontiguous()
        m = pyev.model
        if ps_region_add_event(r, pyev.mean.size, getPr(pyev.mean), getPr(pyev.stdv),
                               getPr(pyev.ref_align), getPr(pyev.ref_like),
                               getPr(m.level_mean), getPr(m.level_stdv), getPr(m.sd_mean), getPr(m.sd_stdv),
                               m.complement, m.prob_skip, m.prob_stay, m.prob_extend, m.prob_insert,
                               pyev.sequence.encode('ascii')) != 0:
            ps_region_destroy(r)
            raise RuntimeError(_error())
    return r


cdef _region_sequence(ps_region* r):
    cdef int n = ps_region_sequence_length(r)
    buf = bytearray(n + 1)
    cdef char* out = buf
    if ps_region_get_sequence(r, out, n + 1) != 0:
        raise RuntimeError(_error())
    return bytes(buf[:n]).decode('ascii')


cdef _update_python_events(events, ps_region* r):                     # as pyx:131-137
    for e, ev in enumerate(events):
        if ps_region_get_event_align(r, e, getPr(ev.ref_align), getPr(ev.ref_like)) != 0:
            raise RuntimeError(_error())


class PSAlign(object):
    def __init__(self):
        self.sequence = ""
        self.events = []
        self.params = {}

    def NumEvents(self):
        """Host only: marshals the object and asks the library how many events arrived."""
        cdef ps_region* r = PythonToRegion(self)
        try:
            return ps_region_num_events(r)
        finally:
            ps_region_destroy(r)

    def ScoreEvents(self):                                             # was pyx:263-276
        cdef ps_region* r = PythonToRegion(self)
        cdef cnp.ndarray scores = np.zeros(len(self.events))
        try:
            if ps_score_events(r, getPr(scores)) != 0:
                raise RuntimeError(_error())
            return scores.tolist()
        finally:
            ps_region_destroy(r)

    def Refine(self):                                                  # was pyx:437-472
        cdef ps_region* r = PythonToRegion(self, 'point_width')
        cdef int nbases = 0
        try:
            if ps_refine(r, &nbases) != 0:
                raise RuntimeError(_error())
            self.sequence = _region_sequence(r)
            _update_python_events(self.events, r)
        finally:
            ps_region_destroy(r)
        return nbases
