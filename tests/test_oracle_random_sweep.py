"""CPU tests (-m "not gpu"): randomised sweep of tiny and degenerate regions through the restatement
(oracle/ps_oracle.cpp) and the reference's own C++ (oracle/_ref/libps_ref.so), bit for bit.

The reference ships no tests; its edge behaviour (cpp/MakeMutations.cpp:46 start past the end,
cpp/Alignment.cpp:51-59 events without an alignment, cpp/Sequence.h:87 non-ACGT bases, bands narrower than the
event, one-level events, transition probabilities above 1, alignments jittered out of order) is pinned here on
several hundred seeded inputs of 6..90 bases, each run through every entry point of the path."""
import numpy as np
import pytest

from poreseq_b200 import synth
from util import edge_mutations, same_aligns


def _tiny_region(seed, lo, hi, rng):
    params = dict(realign_width=int(rng.integers(3, 40)), scoring_width=int(rng.integers(2, 15)),
                  point_width=int(rng.integers(1, 9)), lik_offset=float(rng.choice([0.0, 2.0, 4.5, 9.0])))
    return synth.make_region(int(rng.integers(lo, hi)), int(rng.integers(1, 4)), seed=seed + 1,
                             draft_error=float(rng.choice([0, 0.05, 0.2])), partial=float(rng.choice([0, 0.5, 1.0])),
                             p_unaligned=float(rng.choice([0, 0.3, 0.9])), jitter=int(rng.choice([0, 2, 6])), params=params)


@pytest.mark.parametrize("block", range(4))
def test_dp_entry_points_on_degenerate_regions(orc, ref, block):
    """ScoreAlignments (+ profile), ScorePoints, ScoreMutations with boundary edits, Refine."""
    for seed in range(block * 60, block * 60 + 60):
        rng = np.random.default_rng(1000 + seed)
        reg = _tiny_region(seed, 6, 60, rng)
        kind = seed % 5
        if kind == 1:
            reg.events[0].ref_align[:] = 0                          # an event without any aligned level
        if kind == 2 and len(reg.sequence) > 8:
            s = list(reg.sequence)
            s[int(rng.integers(0, len(s)))] = "N"                   # a non-ACGT base
            reg.sequence = "".join(s)
        if kind == 3:                                               # a one-level event
            ev = reg.events[-1]
            for f in ("mean", "stdv", "ref_align", "ref_like"):
                setattr(ev, f, np.ascontiguousarray(getattr(ev, f)[:1]))
        if kind == 4:
            reg.events[0].ref_align[:] = -1                         # every level an insertion
        s1, l1, a1 = ref.score_alignments(reg, True)
        s2, l2, a2 = orc.score_alignments(reg, True)
        assert np.array_equal(s1, s2) and np.array_equal(l1, l2) and same_aligns(a1, a2), seed
        p1, a1 = ref.score_points(reg)
        p2, a2 = orc.score_points(reg)
        assert p1 == p2 and same_aligns(a1, a2), seed
        st, og, mu = edge_mutations(reg.sequence, seed, count=30)
        m1, a1 = ref.score_mutations(reg, st, og, mu)
        m2, a2 = orc.score_mutations(reg, st, og, mu)
        assert np.array_equal(m1, m2) and same_aligns(a1, a2), seed
        r1, r2 = ref.refine(reg), orc.refine(reg)
        assert r1[0] == r2[0] and r1[1] == r2[1] and same_aligns(r1[2], r2[2]), seed


@pytest.mark.parametrize("block", range(4))
def test_driver_entry_points_on_degenerate_regions(orc, ref, block):
    """swfull, MapAlignments, FindMutations (with a repeated seed) and the Mutate loop."""
    for seed in range(block * 60, block * 60 + 60):
        rng = np.random.default_rng(5000 + seed)
        reg = _tiny_region(seed, 12, 90, rng)
        kind = seed % 4
        if kind == 1:
            reg.events[0].ref_align[:] = 0
        if kind == 2:                                               # transition probabilities at and above 1
            for ev in reg.events:
                ev.model.prob_skip = float(rng.choice([0.5, 1.0, 1.5]))
                ev.model.prob_insert = float(rng.choice([0.01, 1.2]))
        seeds = [ev.sequence for ev in reg.events[::2]][:3] + [synth.corrupt_sequence(reg.sequence, 0.1, rng)[0]]
        seeds.append(seeds[0])
        x, y = orc.swfull(reg.sequence, seeds[-2]), ref.swfull(reg.sequence, seeds[-2])
        assert x[1:] == y[1:], seed
        assert same_aligns(orc.map_alignments(reg, seeds[-2]), ref.map_alignments(reg, seeds[-2])), seed
        f1, a1 = ref.find_mutations(reg, seeds)
        f2, a2 = orc.find_mutations(reg, seeds)
        assert f1 == f2 and same_aligns(a1, a2), seed
        m1, m2 = ref.mutate(reg, seeds, reps=2), orc.mutate(reg, seeds, reps=2)
        assert m1[0] == m2[0] and m1[1] == m2[1] and same_aligns(m1[2], m2[2]), seed


def test_viterbi_on_small_regions(orc, ref):
    """ViterbiMutate (cpp/Viterbi.cpp:239-426) on short regions with few reads: the best path and 8 sampled
    paths on the same rand() stream, over a range of skip / stay / mutation-rate arguments."""
    import copy
    for seed in range(24):
        rng = np.random.default_rng(9000 + seed)
        reg = synth.make_region(int(rng.integers(20, 120)), int(rng.integers(1, 4)), seed=seed + 1,
                                draft_error=float(rng.choice([0, 0.05])), partial=float(rng.choice([0, 0.5])),
                                params=dict(realign_width=30, scoring_width=10, point_width=5))
        seq, _, al = ref.refine(reg)                                # aligned events
        if any(not (np.asarray(ra) > 0).any() for ra, _ in al):
            continue                                                # the reference dereferences an empty path there
        rr = copy.deepcopy(reg)
        rr.sequence = seq
        for ev, (ra, rl) in zip(rr.events, al):
            ev.ref_align, ev.ref_like = ra, rl
        kw = dict(skip=float(rng.choice([0.05, 0.2])), stay=float(rng.choice([0.01, 0.1])),
                  mut_min=float(rng.choice([0.0, 0.33])), mut_max=float(rng.choice([0.75, 1.0])))
        assert orc.viterbi_mutate(rr, nkeep=0, **kw) == ref.viterbi_mutate(rr, nkeep=0, **kw), seed
        assert orc.viterbi_mutate(rr, nkeep=8, seed=seed + 1, **kw) == ref.viterbi_mutate(rr, nkeep=8, seed=seed + 1, **kw), seed
