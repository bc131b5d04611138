"""CPU tests (-m "not gpu"): the restatement in oracle/ps_oracle.cpp is pinned against
(a) the committed fixtures generated from the reference's own C++ (tests/golden/*.npz) and
(b) the compiled reference itself (oracle/_ref/libps_ref.so) where it is available."""
import glob
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import aligned_copy, aligns_array, split_aligns, unpack_region  # noqa: E402
from util import CASES, edge_mutations, region, same_aligns  # noqa: E402

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_restatement_matches_golden(orc, path):
    z = np.load(path)
    reg = unpack_region(z)
    s, l, a = orc.score_alignments(reg, True)
    assert np.array_equal(s, z["sa_scores"]) and np.array_equal(l, z["sa_likes"])
    assert np.array_equal(aligns_array(a), z["sa_aligns"])
    pts, a = orc.score_points(reg)
    assert np.array_equal(np.array([p[3] for p in pts]), z["sp_scores"])
    assert np.array_equal(aligns_array(a), z["sp_aligns"])
    sm, _ = orc.score_mutations(reg, z["sm_start"].tolist(), z["sm_orig"].tolist(), z["sm_mut"].tolist())
    assert np.array_equal(sm, z["sm_scores"])
    seq, nb, a = orc.refine(reg)
    assert seq == str(z["rf_seq"]) and nb == int(z["rf_nbases"])
    assert np.array_equal(aligns_array(a), z["rf_aligns"])


@pytest.mark.parametrize("name", [c[0] for c in CASES[:3]])
def test_restatement_matches_reference(orc, ref, name):
    reg = region(name)
    s1, l1, a1 = ref.score_alignments(reg, True)
    s2, l2, a2 = orc.score_alignments(reg, True)
    assert np.array_equal(s1, s2) and np.array_equal(l1, l2) and same_aligns(a1, a2)
    p1, a1 = ref.score_points(reg)
    p2, a2 = orc.score_points(reg)
    assert p1 == p2 and same_aligns(a1, a2)
    st, og, mu = edge_mutations(reg.sequence, 11)
    m1, a1 = ref.score_mutations(reg, st, og, mu)
    m2, a2 = orc.score_mutations(reg, st, og, mu)
    assert np.array_equal(m1, m2) and same_aligns(a1, a2)
    r1, r2 = ref.refine(reg), orc.refine(reg)
    assert r1[0] == r2[0] and r1[1] == r2[1] and same_aligns(r1[2], r2[2])


def test_states_with_invalid_bases(orc, ref):
    for seq in ["ACGTACGTNACGTACGGTNNACGTTGCA", "NACGT", "ACGT", "ACGTN", "ACG-TACGTAC", "TTTTTTTTTT"]:
        assert orc.seq_to_states(seq).tolist() == ref.seq_to_states(seq).tolist()


def _noisy(seq, rng, rate):
    out = []
    for ch in seq:
        u = rng.random()
        if u < rate / 3:
            continue
        if u < 2 * rate / 3:
            out.append("ACGT"[rng.integers(4)])
            continue
        out.append(ch)
        if u < rate:
            out.append("ACGT"[rng.integers(4)])
    return "".join(out)


def test_swfull_restatement_matches_reference(orc, ref):
    """swfull (cpp/swlib.cpp:211-340): score, accuracy and the aligned index pairs, incl. empty overlaps, one-base
    sequences, homopolymers (ties between the gap moves and the pairing move) and unrelated sequences."""
    rng = np.random.default_rng(3)
    pairs = [("ACGT", "ACGT"), ("A", "A"), ("A", "C"), ("AAAAAAAAAA", "AAAAAAA"), ("ACGTACGT", "TTTT"),
             ("ACGTTTTTACGT", "ACGTTTACGT"), ("GATTACA", "GCATGCT"), ("ACACACACAC", "CACACACA")]
    for n in (30, 120, 400):
        base = "".join(rng.choice(list("ACGT"), n))
        pairs += [(base, _noisy(base, rng, 0.15)), (_noisy(base, rng, 0.3), base), (base, "".join(rng.choice(list("ACGT"), n)))]
    for a, b in pairs:
        x, y = orc.swfull(a, b), ref.swfull(a, b)
        assert x[1:] == y[1:] and (x[0] == y[0] or (np.isnan(x[0]) and np.isnan(y[0]))), (a, b)   # 0/0 on an empty alignment


@pytest.mark.parametrize("name", ["draft_partial", "ragged"])
def test_find_and_map_restatements_match_reference(orc, ref, name):
    """MapAlignments, FindMutations and the Mutate loop of the restatement against the reference's own C++: remapped
    alignments, candidate lists in order, final sequence, bases changed and every event's alignment."""
    reg = region(name)
    rng = np.random.default_rng(17)
    seeds = [ev.sequence for ev in reg.events[::2]][:4] + [_noisy(reg.sequence, rng, 0.08)]
    seeds.append(seeds[0])                                        # a repeated seed takes the cached profile
    assert same_aligns(orc.map_alignments(reg, seeds[-2]), ref.map_alignments(reg, seeds[-2]))
    f1, a1 = ref.find_mutations(reg, seeds)
    f2, a2 = orc.find_mutations(reg, seeds)
    assert f1 == f2 and len(f1) > 0 and same_aligns(a1, a2)
    m1, m2 = ref.mutate(reg, seeds, reps=3), orc.mutate(reg, seeds, reps=3)
    assert m1[0] == m2[0] and m1[1] == m2[1] and same_aligns(m1[2], m2[2])


@pytest.mark.parametrize("name", ["clean", "draft_partial", "ragged"])
def test_viterbi_restatement_matches_reference(orc, ref, name):
    """ViterbiMutate of the restatement against the reference's own C++: the best path (nkeep = 0) and sixteen
    sampled paths on the same rand() stream (both sides are reseeded before the call)."""
    reg = region(name)
    a, _, al = ref.refine(reg)                      # aligned events (ViterbiMutate needs every event aligned)
    import copy
    rr = copy.deepcopy(reg)
    rr.sequence = a
    for ev, (ra, rl) in zip(rr.events, al):
        ev.ref_align, ev.ref_like = ra, rl
    assert orc.viterbi_mutate(rr, nkeep=0) == ref.viterbi_mutate(rr, nkeep=0)
    got, want = orc.viterbi_mutate(rr, nkeep=16, seed=1), ref.viterbi_mutate(rr, nkeep=16, seed=1)
    assert got == want and len(want) == 16


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_restatement_matches_driver_golden(orc, path):
    """The driver-level restatements against the committed outputs of the reference (tests/golden/d_*.npz): swfull,
    MapAlignments, FindMutations (candidates in order), the Mutate loop, ViterbiMutate (best path and 4 samples)."""
    z = np.load(path)
    d = np.load(path.replace("g_", "d_"))
    reg = unpack_region(z)
    seeds = d["seeds"].tolist()
    acc, score, pairs = orc.swfull(reg.sequence, seeds[-2])
    assert (acc, score) == (float(d["sw_acc"]), int(d["sw_score"])) and pairs == [tuple(p) for p in d["sw_pairs"].tolist()]
    assert np.array_equal(aligns_array(orc.map_alignments(reg, seeds[-2])), d["ma_aligns"])
    fm, a = orc.find_mutations(reg, seeds)
    assert fm == list(zip(d["fm_start"].tolist(), d["fm_orig"].tolist(), d["fm_mut"].tolist()))
    assert np.array_equal(aligns_array(a), d["fm_aligns"])
    seq, nb, a = orc.mutate(reg, seeds, reps=3)
    assert seq == str(d["mu_seq"]) and nb == int(d["mu_nbases"]) and np.array_equal(aligns_array(a), d["mu_aligns"])
    rr = aligned_copy(reg, str(z["rf_seq"]), split_aligns(z["rf_aligns"], reg))
    assert orc.viterbi_mutate(rr, nkeep=0) == d["vit_best"].tolist()
    assert orc.viterbi_mutate(rr, nkeep=4, seed=1) == d["vit_samples"].tolist()


def test_baseline_configs_0_and_1_at_full_size(orc, ref):
    """BASELINE.json configs[0] (PSAlign.ScoreEvents, 1 kb x 10x) and configs[1] (the full single-base scan of the same
    region: 7968 edits x 20 events at the default widths) on the restatement and on the reference's own C++: event
    scores, likelihood profile, all 7968 mutation scores and every event's alignment, bit for bit."""
    from poreseq_b200 import synth
    reg = synth.make_region(1000, 10, seed=1)
    s1, l1, a1 = ref.score_alignments(reg, True)
    s2, l2, a2 = orc.score_alignments(reg, True)
    assert len(s1) == 20 and np.array_equal(s1, s2) and np.array_equal(l1, l2) and same_aligns(a1, a2)
    p1, a1 = ref.score_points(reg)
    p2, a2 = orc.score_points(reg)
    assert len(p1) == 7968 and p1 == p2 and same_aligns(a1, a2)


def test_baseline_config_3_shape_reduced(orc, ref):
    """BASELINE.json configs[3] (`poreseq variant -m`: known multi-base mutations scored at scoring_width 100 against
    deep coverage) at a size the CPU finishes in seconds: 2 kb, 40 events, 600 random sub / ins / del edits of up to 6
    bases plus the boundary edits, default widths."""
    from poreseq_b200 import synth
    reg = synth.make_region(2000, 20, seed=4, draft_error=0.01, partial=0.3, params=dict(scoring_width=100))
    rng = np.random.default_rng(44)
    st, og, mu = synth.random_mutations(reg.sequence, 600, rng, max_len=6)
    e_st, e_og, e_mu = edge_mutations(reg.sequence, 45, count=0)
    st, og, mu = st + e_st, og + e_og, mu + e_mu
    m1, a1 = ref.score_mutations(reg, st, og, mu)
    m2, a2 = orc.score_mutations(reg, st, og, mu)
    assert np.array_equal(m1, m2) and same_aligns(a1, a2) and (m1 > 0).any() and (m1 < -1).any()
