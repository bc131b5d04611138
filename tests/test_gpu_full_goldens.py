"""Full-size parity against outputs of the REFERENCE's own C++ (tests/golden/f_*.npz, made by
tests/golden/make_golden_full.py from oracle/_ref on the CPU): BASELINE.json configs[2] (the whole consensus loop on a
10 kb region at 30x), one region of configs[4] (the same loop at 50x = 100 events), one region of configs[3]
(`poreseq variant`: 1200 multi-base edits at scoring_width 100 against 200 events), plus the same loop at 2 kb.

The inputs are regenerated here by the seeded synthetic generator; the fixture carries a checksum of every input array,
so a drifted generator fails loudly instead of comparing different problems.  Both precision modes run: decisions
(sequences after every stage, bases changed, every event's alignment) must be identical in both; scores are bit-exact
in EXACT mode and within 1e-4 RELATIVE (scores >= 0: bit-exact) in FAST mode.
"""
import ctypes
import hashlib
import os

import numpy as np
import pytest

from poreseq_b200 import drivers, poreseqcpp, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REL_TOL = 1e-4


def load(name):
    path = os.path.join(HERE, "golden", "f_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated yet (tests/golden/make_golden_full.py %s)" % (os.path.basename(path), name))
    z = np.load(path)
    kw = {k: (float(v) if k == "draft_error" else int(v)) for k, v in zip(z["kw_keys"].tolist(), z["kw_vals"].tolist())}
    reg = synth.make_region(**kw)
    h = hashlib.sha256()
    h.update(reg.sequence.encode())
    for ev in reg.events:
        for a in (ev.mean, ev.stdv, ev.ref_align):
            h.update(np.ascontiguousarray(a, dtype="f8").tobytes())
        h.update(ev.sequence.encode())
        m = ev.model
        for a in (m.level_mean, m.level_stdv, m.sd_mean, m.sd_stdv):
            h.update(np.ascontiguousarray(a, dtype="f8").tobytes())
    assert h.hexdigest() == str(z["input_sha256"]), "the synthetic generator no longer reproduces the fixture's input"
    return z, reg


def align_digest(aligns):
    h = hashlib.sha256()
    for ra in aligns:
        h.update(np.ascontiguousarray(ra, dtype="f8").tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("precision", ["exact", "fast"])
@pytest.mark.parametrize("name", ["c2s", "c2", "c4"])
def test_consensus_loop_matches_reference(name, precision):
    """poreseq/Mutate.py:47-99 stage by stage: the sequence after Mutate('self'), after every Mutate('viterbi') and
    every Refine, the bases each stage changed and all events' alignments equal the reference's."""
    z, reg = load(name)
    c = poreseqcpp.Context(0)
    try:
        c.set_precision(precision)
        pa = drivers.make_psalign(reg)
        pa.ctx = c
        ctypes.CDLL("libc.so.6").srand(1)        # ViterbiMutate draws from the process-global rand() stream (A.3b-6)
        stages = []
        drivers.consensus(pa, reps=4, stages=stages)
        want_names = z["stage_names"].tolist()
        assert [s[0] for s in stages] == want_names
        for k, (nm, seq, nb, al) in enumerate(stages):
            assert nb == int(z["stage_nbases"][k]), (nm, nb, int(z["stage_nbases"][k]))
            assert seq == str(z["stage_seqs"][k]), "sequence differs after %s" % nm
            assert align_digest(al) == str(z["stage_aligns"][k]), "alignments differ after %s" % nm
    finally:
        c.close()


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_variant_region_matches_reference(precision):
    """configs[3], one region: PSAlign.ScoreMutations (poreseq/Variant.py:71-76) of 1200 random single/multi-base edits
    at scoring_width 100 against 200 events of a 10 kb region, and ScoreEvents of the same region."""
    z, reg = load("c3")
    rng = np.random.default_rng(4242)
    st, og, mu = synth.random_mutations(reg.sequence, len(z["scores"]), rng, max_len=4)
    want = z["scores"]
    c = poreseqcpp.Context(0)
    try:
        c.set_precision(precision)
        nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params)
        got = nr.score_mutations(st, og, mu)
        al = [nr.event_align(e) for e in range(len(reg.events))]
        nr.close()
        if precision == "exact":
            assert np.array_equal(got, want), int(np.sum(got != want))
        else:
            assert np.array_equal(got[want >= 0], want[want >= 0]) and np.array_equal(got >= 0, want >= 0)
            assert np.all(np.abs(got - want) <= REL_TOL * np.abs(want)), float(np.max(np.abs(got - want) / np.abs(want)))
        assert align_digest([a[0] for a in al]) == str(z["aligns_sha256"])
        assert np.array_equal(np.array([float(np.sum(a[1])) for a in al]), z["ref_like_sum"])
        nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params)
        ev_scores, _ = nr.score_alignments()
        nr.close()
        assert np.array_equal(ev_scores, z["event_scores"])
    finally:
        c.close()


@pytest.mark.parametrize("precision", ["exact", "fast"])
@pytest.mark.parametrize("name", ["c2s", "c2", "c4"])
def test_native_consensus_matches_reference(name, precision):
    """The same loop below the C-ABI (ps_consensus: one region handle for the whole loop, the region's own rand()
    stream): every stage's sequence and bases changed equal the reference's, and so do the final alignments."""
    z, reg = load(name)
    c = poreseqcpp.Context(0)
    try:
        c.set_precision(precision)
        nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params)
        stages = nr.consensus(reps=4, point_width=int(reg.params["point_width"]))
        assert [s[0] for s in stages] == z["stage_names"].tolist()
        for k, (nm, seq, nb) in enumerate(stages):
            assert nb == int(z["stage_nbases"][k]), (nm, nb, int(z["stage_nbases"][k]))
            assert seq == str(z["stage_seqs"][k]), "sequence differs after %s" % nm
        al = [nr.event_align(e)[0] for e in range(len(reg.events))]
        assert align_digest(al) == str(z["stage_aligns"][-1])
        nr.close()
    finally:
        c.close()
