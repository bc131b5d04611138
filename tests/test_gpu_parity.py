"""CUDA path vs the CPU checkers, through the C-ABI.  Bit-exact: the FP64 kernels evaluate the
reference's recurrence in the reference's order (no FMA), so scores, accepted mutations, final
sequences and per-level alignments must all be identical, not merely within 1e-4."""
import numpy as np
import pytest

from poreseq_b200 import poreseqcpp, synth
from util import CASES, edge_mutations, region, same_aligns

pytestmark = pytest.mark.gpu

# BASELINE.json north_star: per-mutation log-likelihoods within 1e-4 RELATIVE (pure: no absolute term)
REL_TOL = 1e-4


def native(ctx, reg, width_key=None):
    return poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params, width_key)


def native_aligns(nr, reg):
    return [nr.event_align(e) for e in range(len(reg.events))]


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_score_alignments(ctx, orc, name):
    reg = region(name)
    want_s, want_l, want_a = orc.score_alignments(reg, True)
    nr = native(ctx, reg)
    got_s, got_l = nr.score_alignments(True)
    assert np.array_equal(got_s, want_s)
    assert np.array_equal(got_l, want_l)
    assert same_aligns(native_aligns(nr, reg), want_a)


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_score_points(ctx, orc, name):
    reg = region(name)
    want, want_a = orc.score_points(reg)
    nr = native(ctx, reg, "point_width")
    st, og, mu, sc = nr.score_points()
    assert len(st) == len(want)
    assert [int(s) for s in st] == [w[0] for w in want]
    assert [chr(b) if b else "" for b in og] == [w[1] for w in want]
    assert [chr(b) if b else "" for b in mu] == [w[2] for w in want]
    assert np.array_equal(sc, np.array([w[3] for w in want]))
    assert same_aligns(native_aligns(nr, reg), want_a)


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_score_mutations_edges(ctx, orc, name):
    reg = region(name)
    st, og, mu = edge_mutations(reg.sequence, 11)
    want, want_a = orc.score_mutations(reg, st, og, mu)
    nr = native(ctx, reg)
    got = nr.score_mutations(st, og, mu)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, [(int(i), st[i], og[i], mu[i], got[i], want[i]) for i in bad[:8]]
    assert same_aligns(native_aligns(nr, reg), want_a)


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_refine(ctx, orc, name):
    reg = region(name)
    want_seq, want_nb, want_a = orc.refine(reg)
    nr = native(ctx, reg, "point_width")
    nb = nr.refine()
    assert nb == want_nb
    assert nr.sequence() == want_seq
    assert same_aligns(native_aligns(nr, reg), want_a)


def test_psalign_surface(ctx, orc):
    """The PSAlign mirror: same call sequence a poreseq driver makes (Variant.py:48-78, Mutate.py:76-85)."""
    reg = region("draft_partial")
    pa = poreseqcpp.PSAlign()
    pa.sequence, pa.events, pa.params = reg.sequence, [e.copy() for e in reg.events], dict(reg.params)
    scores = pa.ScoreEvents()
    assert np.array_equal(np.array(scores), orc.score_alignments(reg)[0])
    pts = pa.ScorePoints()
    want, _ = orc.score_points(reg)
    assert [(p.start, p.orig, p.mut, p.score) for p in pts] == want
    nb = pa.Refine()
    want_seq, want_nb, want_a = orc.refine(reg)
    assert (nb, pa.sequence) == (want_nb, want_seq)
    assert same_aligns([(e.ref_align, e.ref_like) for e in pa.events], want_a)


def test_batch_matches_single(ctx, orc):
    regs = [region("clean"), region("draft_partial"), region("ragged")]
    nrs = [native(ctx, r, "point_width") for r in regs]
    out = poreseqcpp.score_points_batch(ctx, nrs)
    for r, (st, og, mu, sc) in zip(regs, out):
        want, _ = orc.score_points(r)
        assert np.array_equal(sc, np.array([w[3] for w in want]))


def test_unusable_and_invalid_bases(ctx, orc):
    reg = region("clean")
    reg.events[1].ref_align[:] = 0            # no alignment -> event unusable (cpp/Alignment.cpp:51-59)
    seq = list(reg.sequence)
    seq[50] = "N"; seq[51] = "N"; seq[200] = "-"
    reg.sequence = "".join(seq)
    want_s, _, want_a = orc.score_alignments(reg)
    nr = native(ctx, reg)
    got_s, _ = nr.score_alignments()
    assert np.array_equal(got_s, want_s)
    st, og, mu = edge_mutations(reg.sequence, 5, count=150)
    st += [48, 49, 50, 51, 52, 196, 199, 200]; og += ["", "A", "N", "N", "", "", "", "-"]; mu += ["A", "", "C", "", "G", "T", "A", "A"]
    want, want_a = orc.score_mutations(reg, st, og, mu)
    nr = native(ctx, reg)
    got = nr.score_mutations(st, og, mu)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, [(int(i), st[i], og[i], mu[i], got[i], want[i]) for i in bad[:8]]
    assert same_aligns(native_aligns(nr, reg), want_a)


@pytest.mark.parametrize("name", ["draft_partial", "ragged"])
def test_find_mutations(ctx, drv, name):
    """Seed-based candidate discovery (cpp/FindMutations.cpp:24-186): same candidates, same order."""
    reg = region(name)
    seeds = [ev.sequence for ev in reg.events[::2]]
    want, want_a = drv.find_mutations(reg, seeds)
    nr = native(ctx, reg)
    got = nr.find_mutations(seeds)
    assert got == want
    assert same_aligns(native_aligns(nr, reg), want_a)


@pytest.mark.parametrize("name", ["draft_partial", "ragged"])
def test_mutate_self(ctx, drv, name):
    """PSAlign.Mutate('self') loop: reps x (FindMutations, ScoreMutations, MakeMutations)."""
    reg = region(name)
    seeds = [ev.sequence for ev in reg.events[::2]]
    want_seq, want_nb, want_a = drv.mutate(reg, seeds, reps=4)
    pa = poreseqcpp.PSAlign()
    pa.sequence, pa.events, pa.params = reg.sequence, [e.copy() for e in reg.events], dict(reg.params)
    nb = pa.Mutate(reps=4)
    assert (nb, pa.sequence) == (want_nb, want_seq)
    assert same_aligns([(e.ref_align, e.ref_like) for e in pa.events], want_a)


def test_map_alignments_and_realign(ctx, drv):
    reg = region("draft_partial")
    rng = np.random.default_rng(3)
    newseq, _ = synth.corrupt_sequence(reg.sequence, 0.08, rng)
    want_a = drv.map_alignments(reg, newseq)
    nr = native(ctx, reg)
    nr.map_alignments(newseq)
    assert same_aligns(native_aligns(nr, reg), want_a)


@pytest.mark.parametrize("name", ["clean", "draft_partial", "ragged"])
def test_viterbi_best_path(ctx, drv, name):
    """nkeep=0: Viterbi liks/backptrs are IEEE-exact, so the best-path sequence is identical."""
    reg = region(name)
    want = drv.viterbi_mutate(reg, nkeep=0)
    got = native(ctx, reg).viterbi_mutate(0)
    assert got == want


@pytest.mark.parametrize("name", ["draft_partial", "ragged"])
def test_viterbi_samples(ctx, drv, name):
    """nkeep=16 forward-weighted samples on the same libc rand() stream (srand(1) before each side)."""
    reg = region(name)
    want = drv.viterbi_mutate(reg, nkeep=16, seed=1)
    drv.srand(1)
    got = native(ctx, reg).viterbi_mutate(16)
    assert len(got) == 16
    assert got == want


def test_mutate_viterbi(ctx, drv):
    """PSAlign.Mutate('viterbi') = ViterbiMutate seeds + Find/Score/Make loop (pyx:415-431)."""
    reg = region("draft_partial")
    seeds = drv.viterbi_mutate(reg, nkeep=16, seed=1)
    want_seq, want_nb, want_a = drv.mutate(reg, seeds, reps=4)
    pa = poreseqcpp.PSAlign()
    pa.sequence, pa.events, pa.params = reg.sequence, [e.copy() for e in reg.events], dict(reg.params)
    drv.srand(1)
    nb = pa.Mutate(seqs='viterbi')
    assert (nb, pa.sequence) == (want_nb, want_seq)
    assert same_aligns([(e.ref_align, e.ref_like) for e in pa.events], want_a)


def test_consensus_loop(ctx, drv):
    """The whole Mutate.py policy on the CUDA path vs the same policy driven through the reference:
    final consensus sequence identical, and better than the draft."""
    from poreseq_b200 import drivers
    reg = synth.make_region(500, 5, seed=21, draft_error=0.06, partial=0.2,
                            params=dict(realign_width=80, scoring_width=25, point_width=10, end_trim=20))
    # reference side: same call sequence through the checker
    import copy
    rr = copy.deepcopy(reg)

    def sync(al):
        for ev, (ra, rl) in zip(rr.events, al):
            ev.ref_align, ev.ref_like = ra, rl

    drv.srand(1)
    seq, _, al = drv.mutate(rr, [ev.sequence for ev in rr.events[::2]], reps=4)
    rr.sequence = seq; sync(al)
    for _ in range(4):
        seeds = drv.viterbi_mutate(rr, nkeep=16, seed=None)
        seq, _, al = drv.mutate(rr, seeds, reps=4)
        rr.sequence = seq; sync(al)
        seq, nb, al = drv.refine(rr)
        rr.sequence = seq; sync(al)
        if nb == 0:
            break
    want = rr.sequence[20:-20]
    drv.srand(1)
    pa = drivers.make_psalign(reg)
    got, acc = drivers.consensus(pa, refseq=reg.truth, reps=4)
    assert got == want
    assert acc > poreseqcpp.swalign(reg.sequence, reg.truth)[0]


def test_variant_modes(ctx, orc):
    from poreseq_b200 import drivers
    from poreseq_b200.Util import MutationInfo
    reg = region("draft_partial")
    pa = drivers.make_psalign(reg)
    st, og, mu = edge_mutations(reg.sequence, 3, count=40)
    muts = []
    for s, o, m in zip(st, og, mu):
        mi = MutationInfo(); mi.start, mi.orig, mi.mut = s + 1000, o, m
        muts.append(mi)
    got = drivers.variant(pa, muts=muts, region_start=1000)
    want, _ = orc.score_mutations(reg, st, og, mu)
    assert [g.score for g in got] == want.tolist()
    assert [g.start for g in got] == [s + 1000 for s in st]


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_event_sharded_partials(orc, precision):
    """Two event shards scored separately on the GPU (partial sums from 0) add up to the full score.  A shard's partial
    says nothing about the sign of the total, so partial sums are exact FP64 in both precision modes (FAST flags by
    totals): the two modes must give the same bits."""
    from poreseq_b200 import sharding
    reg = region("draft_partial")
    st, og, mu = edge_mutations(reg.sequence, 4, count=120)
    want, _ = orc.score_mutations(reg, st, og, mu)
    c2 = poreseqcpp.Context(0)
    try:
        c2.set_precision(precision)
        fn = sharding.cuda_partial(c2)
        parts = [fn(sharding.RegionShard(reg, rank, 2), st, og, mu) for rank in range(2)]
        got = -1e-6 + (parts[0] + parts[1])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
        assert np.array_equal(got >= 0, want >= 0)
        # each shard alone is the exact ordered FP64 sum over its events
        for rank in range(2):
            shard = sharding.RegionShard(reg, rank, 2)
            if shard.events:
                w, _ = orc.score_mutations(shard, st, og, mu)
                assert np.array_equal(parts[rank], w + 1e-6) or np.allclose(parts[rank], w + 1e-6, rtol=0, atol=1e-12)
    finally:
        c2.close()


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_fast_mode_decisions_exact(orc, name):
    """PS_PRECISION_FAST: FP32 scan + exact re-score.  Every score above -tau (in particular every
    accepted mutation) is bit-identical; the rest is within 1e-4 RELATIVE (BASELINE.json north_star; no
    absolute slack: tau = 0.5 x events is derived from the measured FP32 error, profiles/r2_fast_error.txt);
    Refine gives the identical sequence."""
    c2 = poreseqcpp.Context(0)
    c2.set_precision("fast")
    try:
        reg = region(name)
        want, want_a = orc.score_points(reg)
        w = np.array([x[3] for x in want])
        nr = poreseqcpp.NativeRegion(c2, reg.sequence, reg.events, reg.params, "point_width")
        st, og, mu, sc = nr.score_points()
        keep = w > -0.4 * len(reg.events)                     # inside the library's -tau = -0.5 x events
        assert np.array_equal(sc[keep], w[keep])
        assert np.array_equal(sc >= 0, w >= 0)
        assert np.all(np.abs(sc - w) <= REL_TOL * np.abs(w)), float(np.max(np.abs(sc - w) / np.abs(w)))
        assert same_aligns([nr.event_align(e) for e in range(len(reg.events))], want_a)
        st2, og2, mu2 = edge_mutations(reg.sequence, 11)
        w2, _ = orc.score_mutations(reg, st2, og2, mu2)
        nr = poreseqcpp.NativeRegion(c2, reg.sequence, reg.events, reg.params)
        s2 = nr.score_mutations(st2, og2, mu2)
        assert np.array_equal(s2 >= 0, w2 >= 0)
        assert np.all(np.abs(s2 - w2) <= REL_TOL * np.abs(w2)), float(np.max(np.abs(s2 - w2) / np.abs(w2)))
        nr = poreseqcpp.NativeRegion(c2, reg.sequence, reg.events, reg.params, "point_width")
        nb = nr.refine()
        seq, want_nb, want_a = orc.refine(reg)
        assert (nb, nr.sequence()) == (want_nb, seq)
        assert same_aligns([nr.event_align(e) for e in range(len(reg.events))], want_a)
    finally:
        c2.close()


# ---- launch classes, odd shapes, batching API, full-size configs ---------------------------------
@pytest.mark.parametrize("kw", [
    dict(length=401, coverage=3, seed=31, draft_error=0.03, params=dict(realign_width=60, scoring_width=15, point_width=8)),   # odd N
    dict(length=230, coverage=3, seed=32, draft_error=0.02, params=dict(realign_width=450, scoring_width=30, point_width=12)),  # band wider than the event
    dict(length=900, coverage=2, seed=33, draft_error=0.02, params=dict(realign_width=420, scoring_width=20, point_width=9)),   # wavefront wider than 160 threads
    dict(length=60, coverage=3, seed=34, params=dict(realign_width=7, scoring_width=3, point_width=2)),                        # tiny band
], ids=["odd", "band_gt_event", "wide_class", "tiny_band"])
def test_fill_shapes(ctx, orc, kw):
    """Odd column / row counts (2x2 tiles with a half-empty last strip or row pair), bands clipped by
    the event, the wide wavefront class and very narrow bands."""
    reg = synth.make_region(**kw)
    want_s, want_l, want_a = orc.score_alignments(reg, True)
    nr = native(ctx, reg)
    got_s, got_l = nr.score_alignments(True)
    assert np.array_equal(got_s, want_s)
    assert np.array_equal(got_l, want_l)
    assert same_aligns(native_aligns(nr, reg), want_a)
    want, want_a = orc.score_points(reg)
    nr = native(ctx, reg, "point_width")
    st, og, mu, sc = nr.score_points()
    assert np.array_equal(sc, np.array([w[3] for w in want]))
    assert same_aligns(native_aligns(nr, reg), want_a)


def test_short_events(ctx, orc):
    """Events of 1..6 levels (single row pairs, strips that end on their first step)."""
    reg = synth.make_region(120, 3, seed=35, params=dict(realign_width=10, scoring_width=4, point_width=3))
    for k, ev in enumerate(reg.events):
        n = 1 + k
        ev.mean, ev.stdv = ev.mean[:n].copy(), ev.stdv[:n].copy()
        ev.ref_align, ev.ref_like = ev.ref_align[:n].copy(), ev.ref_like[:n].copy()
    want_s, _, want_a = orc.score_alignments(reg)
    nr = native(ctx, reg)
    got_s, _ = nr.score_alignments()
    assert np.array_equal(got_s, want_s)
    assert same_aligns(native_aligns(nr, reg), want_a)
    want, _ = orc.score_points(reg)
    st, og, mu, sc = native(ctx, reg, "point_width").score_points()
    assert np.array_equal(sc, np.array([w[3] for w in want]))


def test_packed_and_async_batches(orc):
    """ps_region_add_events + ps_score_points_batch_begin/_end on two contexts: same scores as the
    per-event, synchronous path (and as the checker)."""
    regs = [region("clean"), region("draft_partial"), region("ragged")]
    packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
    want = [np.array([w[3] for w in orc.score_points(r)[0]]) for r in regs]
    ctxs = [poreseqcpp.Context(0), poreseqcpp.Context(0)]
    try:
        pend = []
        for c in ctxs:                       # two batches in flight, one per context
            nrs = [poreseqcpp.NativeRegion.from_packed(c, p, "point_width") for p in packs]
            pend.append(poreseqcpp.score_points_batch_begin(c, nrs))
        for p in pend:
            out = p.end()
            for (st, og, mu, sc), w in zip(out, want):
                assert np.array_equal(sc, w)
        # a second _begin on a busy context is refused, _end without _begin too
        nrs = [poreseqcpp.NativeRegion.from_packed(ctxs[0], packs[0], "point_width")]
        p = poreseqcpp.score_points_batch_begin(ctxs[0], nrs)
        with pytest.raises(RuntimeError):
            poreseqcpp.score_points_batch_begin(ctxs[0], nrs)
        p.end()
        with pytest.raises(RuntimeError):
            p.end()
    finally:
        for c in ctxs:
            c.close()


def test_config2_full_size(ctx, drv):
    """BASELINE.json configs[1] at full size (1 kb x 10x, default widths): every one of the 7968 point
    mutation scores bit-identical to the compiled reference (about 4 s of CPU)."""
    reg = synth.make_region(1000, 10, seed=101)
    want, want_a = drv.score_points(reg)
    nr = native(ctx, reg, "point_width")
    st, og, mu, sc = nr.score_points()
    assert len(sc) == 8 * (1000 - 4)
    assert np.array_equal(sc, np.array([w[3] for w in want]))
    assert same_aligns(native_aligns(nr, reg), want_a)


def test_config3_size_properties(drv):
    """BASELINE.json configs[2] size (10 kb, 30x = 60 events): ScoreEvents bit-identical to the compiled
    reference (one CPU pass, ~6 s); the full point scan (1157 s on one CPU core) is checked through
    properties: fast vs exact mode agree (decisions identical, scores within 1e-4 relative), 300
    sampled mutations match the reference bit for bit, and re-scoring the realigned region is stable."""
    reg = synth.make_region(10000, 30, seed=102, draft_error=0.01)
    want_s, _, want_a = drv.score_alignments(reg)
    cx, cf = poreseqcpp.Context(0), poreseqcpp.Context(0)
    cf.set_precision("fast")
    try:
        nr = native(cx, reg)
        got_s, _ = nr.score_alignments()
        assert np.array_equal(got_s, want_s)
        assert same_aligns(native_aligns(nr, reg), want_a)
        st, og, mu, sc = native(cx, reg, "point_width").score_points()
        st2, og2, mu2, sf = native(cf, reg, "point_width").score_points()
        assert len(sc) == 8 * (len(reg.sequence) - 4)
        assert np.array_equal(sc >= 0, sf >= 0)
        assert np.array_equal(sc[sc > -0.4 * len(reg.events)], sf[sc > -0.4 * len(reg.events)])
        assert np.all(np.abs(sf - sc) <= REL_TOL * np.abs(sc)), float(np.max(np.abs(sf - sc) / np.abs(sc)))
        rng = np.random.default_rng(5)
        pick = np.sort(rng.choice(len(sc), 300, replace=False))
        reg.params = dict(reg.params, scoring_width=reg.params["point_width"])
        w, _ = drv.score_mutations(reg, [int(st[i]) for i in pick], [chr(og[i]) if og[i] else "" for i in pick],
                                   [chr(mu[i]) if mu[i] else "" for i in pick])
        assert np.array_equal(sc[pick], w)
    finally:
        cx.close(); cf.close()


def test_oversized_job_is_split(orc, monkeypatch):
    """A job whose band matrices exceed the memory budget runs as consecutive sub-batches of whole
    regions (FindMutations on a 10 kb region at 30x needs ~190 GB otherwise): same results."""
    monkeypatch.setenv("PORESEQ_B200_BAND_BUDGET", "3e6")      # ~ one small region per sub-batch
    c = poreseqcpp.Context(0)
    try:
        regs = [region("clean"), region("draft_partial"), region("ragged")]
        nrs = [native(c, r, "point_width") for r in regs]
        out = poreseqcpp.score_points_batch(c, nrs)
        for r, nr, (st, og, mu, sc) in zip(regs, nrs, out):
            want, want_a = orc.score_points(r)
            assert np.array_equal(sc, np.array([w[3] for w in want]))
            assert same_aligns(native_aligns(nr, r), want_a)
    finally:
        c.close()


def test_swfull_device_matches_host(ctx, drv):
    """GPU Smith-Waterman (ps_sw.cu) vs the reference's swfull: score, accuracy and every aligned index
    pair identical, including tie-heavy low-complexity sequences and sizes that are not multiples of
    the thread strip."""
    rng = np.random.default_rng(17)
    cases = []
    for n in (1, 5, 37, 400, 1025, 3000):
        a = "".join(rng.choice(list("ACGT"), n))
        b, _ = synth.corrupt_sequence(a, 0.12, rng) if n > 4 else (a, None)
        cases.append((a, b))
    cases.append(("A" * 300, "A" * 280))                                   # every path ties
    cases.append(("ACACACACAC" * 40, "CACACACA" * 45))                     # periodic: many equal maxima
    cases.append(("".join(rng.choice(list("AC"), 700)), "".join(rng.choice(list("AC"), 650))))   # unrelated, two letters
    cases.append(("ACGT" * 50, "TTTT"))
    for a, b in cases:
        want_acc, want_score, want_pairs = drv.swfull(a, b)
        score, acc, pairs = poreseqcpp.swalign_device(ctx, a, b)
        assert [tuple(p) for p in pairs] == [tuple(p) for p in want_pairs], (len(a), len(b))
        assert score == want_score
        assert (acc == want_acc) or (np.isnan(acc) and np.isnan(want_acc))


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_long_replacements_warp_and_thread_forms(orc, monkeypatch, precision):
    """The exact mutation pass has two forms: one warp per (mutation, event) pair (k_mutscore_warp, chunks of 32
    narrow columns for long replacement strings) and one thread per pair (k_mutscore).  Both must give the
    checker's scores bit for bit, for replacement strings from 0 to 100 bases, also at region ends."""
    reg = region("draft_partial")
    rng = np.random.default_rng(5)
    L = len(reg.sequence)
    st, og, mu = [], [], []
    for n in (0, 1, 2, 7, 26, 27, 31, 32, 33, 58, 59, 64, 70, 100):
        for s in (0, 3, int(rng.integers(10, L - 120)), int(rng.integers(10, L - 120)), L - 30, L - 6):
            k = int(rng.integers(0, 4))
            st.append(s); og.append(reg.sequence[s:s + k]); mu.append("".join(rng.choice(list("ACGT"), n)))
    want, want_a = orc.score_mutations(reg, st, og, mu)
    for no_warp in ("", "1"):
        if no_warp:
            monkeypatch.setenv("PORESEQ_B200_NO_WARP", "1")
        else:
            monkeypatch.delenv("PORESEQ_B200_NO_WARP", raising=False)
        c = poreseqcpp.Context(0)
        c.set_precision(precision)
        nr = native(c, reg)
        got = nr.score_mutations(st, og, mu)
        if precision == "exact":
            bad = np.nonzero(got != want)[0]
        else:   # FP32 scan for the single-base edits that are clearly negative, exact for everything else
            bad = np.nonzero((got != want) & ~((np.array([len(m) for m in mu]) <= 1) & (np.abs(got - want) <= 1e-4 * np.abs(want))))[0]
        assert len(bad) == 0, (no_warp, [(int(i), st[i], og[i], mu[i], got[i], want[i]) for i in bad[:8]])
        assert same_aligns(native_aligns(nr, reg), want_a)


def _golden_paths():
    import glob
    import os
    return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g_*.npz")))


@pytest.mark.parametrize("path", _golden_paths(), ids=lambda p: p.split("/")[-1])
def test_native_matches_golden(ctx, path):
    """The CUDA path against the committed outputs of the reference's own C++ (tests/golden/g_*.npz for the DP entry
    points, d_*.npz for the drivers) with no CPU checker in the loop: every number bit for bit."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import aligned_copy, aligns_array, split_aligns, unpack_region
    import ctypes
    z, d = np.load(path), np.load(path.replace("g_", "d_"))
    reg = unpack_region(z)
    nr = native(ctx, reg)
    s, l = nr.score_alignments(True)
    assert np.array_equal(s, z["sa_scores"]) and np.array_equal(l, z["sa_likes"])
    assert np.array_equal(aligns_array(native_aligns(nr, reg)), z["sa_aligns"])
    nr = native(ctx, reg, "point_width")
    assert np.array_equal(nr.score_points()[3], z["sp_scores"])
    assert np.array_equal(aligns_array(native_aligns(nr, reg)), z["sp_aligns"])
    sm = native(ctx, reg).score_mutations(z["sm_start"].tolist(), z["sm_orig"].tolist(), z["sm_mut"].tolist())
    assert np.array_equal(sm, z["sm_scores"])
    nr = native(ctx, reg, "point_width")
    assert nr.refine() == int(z["rf_nbases"]) and nr.sequence() == str(z["rf_seq"])
    assert np.array_equal(aligns_array(native_aligns(nr, reg)), z["rf_aligns"])
    # drivers
    seeds = d["seeds"].tolist()
    score, acc, pairs = poreseqcpp.swalign_device(ctx, reg.sequence, seeds[-2])
    assert (acc, score) == (float(d["sw_acc"]), int(d["sw_score"])) and pairs == [tuple(p) for p in d["sw_pairs"].tolist()]
    assert poreseqcpp.swalign(reg.sequence, seeds[-2]) == (float(d["sw_acc"]), [tuple(p) for p in d["sw_pairs"].tolist()])
    nr = native(ctx, reg)
    nr.map_alignments(seeds[-2])
    assert np.array_equal(aligns_array(native_aligns(nr, reg)), d["ma_aligns"])
    nr = native(ctx, reg)
    assert nr.find_mutations(seeds) == list(zip(d["fm_start"].tolist(), d["fm_orig"].tolist(), d["fm_mut"].tolist()))
    assert np.array_equal(aligns_array(native_aligns(nr, reg)), d["fm_aligns"])
    nr = native(ctx, reg)
    assert nr.mutate(seeds, reps=3) == int(d["mu_nbases"]) and nr.sequence() == str(d["mu_seq"])
    assert np.array_equal(aligns_array(native_aligns(nr, reg)), d["mu_aligns"])
    rr = aligned_copy(reg, str(z["rf_seq"]), split_aligns(z["rf_aligns"], reg))
    assert native(ctx, rr).viterbi_mutate(0) == d["vit_best"].tolist()
    ctypes.CDLL(None).srand(ctypes.c_uint(1))          # ViterbiMutate draws from the process-global rand() stream
    assert native(ctx, rr).viterbi_mutate(4) == d["vit_samples"].tolist()


def test_event_pack_regions_score_like_in_memory_ones(ctx, orc, tmp_path):
    """Regions read back from an event-pack file (memory-mapped views, poreseq_b200/eventpack.py) go through the batched
    entry point and give the checker's scores."""
    from poreseq_b200 import eventpack
    regs = [region("clean"), region("draft_partial"), region("ragged")]
    path = str(tmp_path / "regions.psep")
    eventpack.write_pack(path, regs)
    nrs = poreseqcpp.native_regions_from_packed(ctx, list(eventpack.read_pack(path)), "point_width")
    out = poreseqcpp.score_points_batch(ctx, nrs)
    for reg, o in zip(regs, out):
        want, _ = orc.score_points(reg)
        assert np.array_equal(o[3], np.array([w[3] for w in want]))
    poreseqcpp.close_regions(nrs)


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_event_with_one_aligned_level(ctx, orc, precision):
    """An event whose ref_align holds ONE aligned level has a 0/0 slope in updaterefs (cpp/EventData.h:143-144): every other
    entry of its ref_index is NaN and the band centres are whatever std::lower_bound's probes make of that
    (cpp/EventData.h:172-183).  Found by scripts/gpu_sweep.py in the re-scoring round of Refine; the band planner's
    linear merge took a NaN array for a sorted one."""
    import copy
    reg = copy.deepcopy(region("draft_partial"))
    for e, lvl in ((1, 22), (2, 0), (3, len(reg.events[3].mean) - 1)):
        ra = np.zeros_like(reg.events[e].ref_align)
        ra[lvl] = float(min(lvl + 1, len(reg.sequence) - 4))
        reg.events[e].ref_align = ra
    c = poreseqcpp.Context(0)
    try:
        c.set_precision(precision)
        st, og, mu = edge_mutations(reg.sequence, 5, count=60)
        want, a = orc.score_mutations(reg, st, og, mu)
        nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params)
        got = nr.score_mutations(st, og, mu)
        if precision == "exact":
            assert np.array_equal(got, want)
        else:
            assert np.array_equal(got[want >= 0], want[want >= 0]) and np.all(np.abs(got - want) <= 1e-4 * np.abs(want))
        assert same_aligns([nr.event_align(e) for e in range(len(reg.events))], a)
        nr.close()
        seq, nb, a = orc.refine(reg)
        nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params, "point_width")
        assert nr.refine() == nb and nr.sequence() == seq
        assert same_aligns([nr.event_align(e) for e in range(len(reg.events))], a)
        nr.close()
    finally:
        c.close()


def test_train_loop_matches_reference_per_variant(drv):
    """`poreseq train` restated over a loaded region (drivers.train, poreseq/cmdline.py:246-267): the variants of an
    iteration run as regions in flight through ps_consensus_batch.  Every variant's consensus sequence -- hence its
    accuracy, the choice of the best variant and the parameters carried into the next iteration -- equals what the same
    loop gives when the checker (the reference's C++) runs each variant's consensus."""
    import random
    from poreseq_b200 import drivers
    from util import reference_consensus
    reg = synth.make_region(300, 6, seed=21, draft_error=0.08,
                            params=dict(realign_width=60, scoring_width=15, point_width=8, end_trim=10))
    base = dict(skip_t=0.141, skip_c=0.088, stay_t=0.043, stay_c=0.057, extend_t=0.072, extend_c=0.046,
                insert_t=0.020, insert_c=0.025, lik_offset=4.5)
    details = []
    best, hist = drivers.train(reg, base, iters=2, variants=4, in_flight=4, reps=2, seed=9, device=0, details=details)
    assert len(hist) == 2 and len(details) == 2
    assert set(best) == set(base) and best["lik_offset"] == 4.5
    # the same loop with the reference scoring every variant
    rng = random.Random(9)
    params = dict(base)
    for it in range(2):
        plist = drivers.vary_params(params, rng, 4)
        assert plist == details[it][0]
        accs = []
        for p in plist:
            seq = reference_consensus(drv, drivers.variant_region(reg, p), reps=2)
            accs.append(poreseqcpp.swalign(seq, reg.truth)[0])
        assert accs == details[it][1], (it, accs, details[it][1])
        k = int(np.argmax(accs))
        params = plist[k]
        assert hist[it] == accs[k]
    assert best == params


def test_native_pack_regions_score_like_in_memory_ones(ctx, orc, tmp_path):
    """The same pack opened below the C-ABI (ps_pack_open / ps_pack_regions_create, csrc/ps_pack.cu): regions built
    straight from the mapping give the checker's scores."""
    from poreseq_b200 import eventpack
    regs = [region("clean"), region("draft_partial"), region("ragged")]
    path = str(tmp_path / "regions.psep")
    eventpack.write_pack(path, regs)
    pack = eventpack.NativePack(path)
    nrs = pack.regions(ctx, width_key="point_width")
    out = poreseqcpp.score_points_batch(ctx, nrs)
    for reg, o in zip(regs, out):
        want, _ = orc.score_points(reg)
        assert np.array_equal(o[3], np.array([w[3] for w in want]))
    poreseqcpp.close_regions(nrs)
    pack.close()


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_baseline_config_3_shape_reduced(orc, precision):
    """BASELINE.json configs[3] (`poreseq variant -m`: known multi-base mutations at scoring_width 100 against deep
    coverage) at the size of the CPU test of the same name: 2 kb, 40 events, 600 random edits of up to 6 bases plus the
    boundary edits.  Exact mode: every score bit-identical; fast mode: scores >= 0 bit-identical, the rest within 1e-4
    relative, no absolute slack (as in test_fast_mode_decisions_exact)."""
    reg = synth.make_region(2000, 20, seed=4, draft_error=0.01, partial=0.3, params=dict(scoring_width=100))
    rng = np.random.default_rng(44)
    st, og, mu = synth.random_mutations(reg.sequence, 600, rng, max_len=6)
    e_st, e_og, e_mu = edge_mutations(reg.sequence, 45, count=0)
    st, og, mu = st + e_st, og + e_og, mu + e_mu
    want, want_a = orc.score_mutations(reg, st, og, mu)
    c2 = poreseqcpp.Context(0)
    try:
        c2.set_precision(precision)
        nr = native(c2, reg)
        got = nr.score_mutations(st, og, mu)
        if precision == "exact":
            bad = np.nonzero(got != want)[0]
            assert len(bad) == 0, [(int(i), st[i], og[i], mu[i], got[i], want[i]) for i in bad[:8]]
        else:
            assert np.array_equal(got[want >= 0], want[want >= 0]) and np.array_equal(got >= 0, want >= 0)
            assert np.all(np.abs(got - want) <= REL_TOL * np.abs(want)), float(np.max(np.abs(got - want) / np.abs(want)))
        assert same_aligns(native_aligns(nr, reg), want_a)
    finally:
        c2.close()


def test_cython_stub_is_a_drop_in(orc, tmp_path):
    """The Cython binding a PoreSeq maintainer would add (examples/cython_stub) run for real:
    PSAlign.ScoreEvents and PSAlign.Refine through the compiled stub -> C-ABI -> CUDA give the checker's scores,
    sequence, bases changed and per-level alignments, written back into the Python events in place (pyx:131-137)."""
    import sys
    from util import build_cython_stub
    where, log = build_cython_stub(tmp_path)
    if not where:
        pytest.skip("cannot build the Cython stub here: " + log[-300:])
    sys.path.insert(0, where)
    try:
        import poreseqcpp_b200 as m
    finally:
        sys.path.remove(where)
    reg = region("draft_partial")
    pa = m.PSAlign()
    pa.sequence, pa.events, pa.params = reg.sequence, [e.copy() for e in reg.events], dict(reg.params)
    assert np.array_equal(np.array(pa.ScoreEvents()), orc.score_alignments(reg)[0])
    want_seq, want_nb, want_a = orc.refine(reg)
    assert (pa.Refine(), pa.sequence) == (want_nb, want_seq)
    assert same_aligns([(e.ref_align, e.ref_like) for e in pa.events], want_a)


# ---------------------------------------------------------------------------------------------------------
# PSAlign.ScoreEvents (BASELINE.json configs[0]): scores only, the events keep their alignments

SCORE_EVENT_CASES = [c[0] for c in CASES] + ["1kb", "n_bases", "unaligned", "10kb"]


def score_events_region(name):
    if name == "1kb":
        return synth.make_region(1000, 10, seed=31, draft_error=0.02)             # configs[0]
    if name == "10kb":
        return synth.make_region(10000, 2, seed=32, draft_error=0.02, partial=0.5)
    if name == "n_bases":
        reg = synth.make_region(500, 3, seed=33, draft_error=0.02)
        s = list(reg.sequence)
        for q in (0, 17, 18, 250, 251, 252, 253, 254, 400, len(s) - 1):
            s[q] = "N"
        reg.sequence = "".join(s)
        return reg
    if name == "unaligned":
        reg = synth.make_region(400, 3, seed=34, params=dict(realign_width=60))
        reg.events[1].ref_align[:] = 0.0                                           # no alignment at all: unusable event
        reg.events[2].ref_align[5:] = 0.0                                          # a stub of an alignment
        return reg
    return region(name)


@pytest.mark.parametrize("name", SCORE_EVENT_CASES)
def test_score_events_exact(ctx, drv, name):
    """ps_score_events in EXACT mode = ScoreAlignments(data, NULL) bit for bit, and the handle's events are untouched."""
    reg = score_events_region(name)
    want, _, _ = drv.score_alignments(reg)
    nr = native(ctx, reg)
    before = native_aligns(nr, reg)
    got = nr.score_events()
    assert np.array_equal(got, want)
    assert same_aligns(native_aligns(nr, reg), before)
    nr.close()


@pytest.mark.parametrize("name", SCORE_EVENT_CASES)
def test_score_events_fast(drv, name):
    """FAST mode: the score-only FP32 fill (k_score_f32).  Scores within 1e-4 relative of the reference's (north_star);
    exactly 0 where the reference gives 0 (unusable events); the handle's events untouched; batch = one by one."""
    reg = score_events_region(name)
    want, _, _ = drv.score_alignments(reg)
    c2 = poreseqcpp.Context(0)
    try:
        c2.set_precision("fast")
        nr = native(c2, reg)
        before = native_aligns(nr, reg)
        got = nr.score_events()
        assert got.shape == want.shape
        assert np.all(np.abs(got - want) <= REL_TOL * np.abs(want)), (got, want)
        assert same_aligns(native_aligns(nr, reg), before)
        oreg = synth.make_region(300, 3, seed=35, draft_error=0.03, params=reg.params)   # (a batch shares the band widths)
        other = native(c2, oreg)
        twin = native(c2, reg)
        both = poreseqcpp.score_events_batch(c2, [nr, other, twin])
        assert np.array_equal(both[0], got) and np.array_equal(both[2], got)
        assert np.array_equal(both[1], other.score_events())
        with pytest.raises(RuntimeError):                              # a handle can be in a batch only once
            poreseqcpp.score_events_batch(c2, [nr, other, nr])
        nr.close(); other.close(); twin.close()
    finally:
        c2.close()


# ---------------------------------------------------------------------------------------------------------
# the consensus loop below the C-ABI (ps_consensus / ps_consensus_batch) against the Python policy over PSAlign

@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_native_consensus_equals_python_policy(precision):
    """drivers.consensus (a fresh native region per PSAlign method, like the reference) and ps_consensus (one handle for
    the whole loop) must walk through the same stages; a batch of regions in flight gives what one by one gives."""
    import ctypes
    from poreseq_b200 import drivers
    regs = [synth.make_region(500, 5, seed=60 + k, draft_error=0.08) for k in range(5)]
    regs.append(synth.make_region(300, 2, seed=70, draft_error=0.05))             # fewer than 5 events: left untouched
    c = poreseqcpp.Context(0)
    try:
        c.set_precision(precision)
        want = []
        for reg in regs:
            pa = drivers.make_psalign(reg)
            pa.ctx = c
            ctypes.CDLL("libc.so.6").srand(1)
            st = []
            seq, _ = drivers.consensus(pa, reps=4, stages=st)
            want.append((seq, [(s[0], s[1], s[2]) for s in st], [np.array(ev.ref_align) for ev in pa.events]))
        got = drivers.consensus_native(regs, ctx=c, reps=4, in_flight=4)
        for k, reg in enumerate(regs):
            assert got[k][2] == want[k][1], "stages of region %d differ" % k
            assert got[k][0] == want[k][0]
        one = poreseqcpp.NativeRegion(c, regs[0].sequence, regs[0].events, regs[0].params)
        assert one.consensus(4, int(regs[0].params["point_width"])) == want[0][1]
        assert all(np.array_equal(one.event_align(e)[0], want[0][2][e]) for e in range(len(regs[0].events)))
        one.close()
    finally:
        c.close()


@pytest.mark.parametrize("precision", ["exact", "fast"])
def test_score_points_direct(orc, precision):
    """ps_score_points_direct: PSAlign.ScorePoints from the caller's arrays without region handles -- the same scores as
    the handle path (bit-identical to the checker in exact mode), for a ragged batch; the inputs stay untouched."""
    regs = [region(c[0]) for c in CASES[:3]]                          # (one batch shares the band widths)
    packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
    keep = [(p.mean.copy(), p.ref_align.copy(), p.ref_like.copy()) for p in packs]
    c2 = poreseqcpp.Context(0)
    try:
        c2.set_precision(precision)
        got = poreseqcpp.score_points_direct(c2, packs)
        for k, reg in enumerate(regs):
            nr = native(c2, reg, "point_width")
            st, og, mu, sc = nr.score_points()
            assert np.array_equal(got[k][0], st) and got[k][1] == og and got[k][2] == mu
            assert np.array_equal(got[k][3], sc)
            if precision == "exact":
                want, _ = orc.score_points(reg)
                assert np.array_equal(got[k][3], np.array([w[3] for w in want]))
            assert np.array_equal(packs[k].mean, keep[k][0]) and np.array_equal(packs[k].ref_align, keep[k][1])
            assert np.array_equal(packs[k].ref_like, keep[k][2])
            nr.close()
    finally:
        c2.close()
