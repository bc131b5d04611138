"""Generates the committed known-answer fixtures from the REFERENCE's own C++.

    make -C oracle ref && python tests/golden/make_golden.py

Runs only where /root/reference exists (the compiled reference is oracle/_ref/libps_ref.so).  The
reference ships no tests or vectors of its own, so these files ARE the pin:
each .npz stores the complete seeded synthetic input (so no RNG has to reproduce) and the outputs of
the reference entry points on it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import binding  # noqa: E402
from poreseq_b200 import synth  # noqa: E402
from util import edge_mutations  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
P = dict(realign_width=40, scoring_width=12, point_width=6)
GOLDEN = [
    ("g_clean", dict(length=160, coverage=2, seed=101, params=P)),
    ("g_draft", dict(length=180, coverage=3, seed=102, draft_error=0.04, partial=0.4, params=P)),
    ("g_ragged", dict(length=150, coverage=2, seed=103, draft_error=0.06, partial=0.5, p_unaligned=0.3, jitter=2, params=P)),
]


def pack_region(reg):
    d = {"sequence": np.array(reg.sequence), "truth": np.array(reg.truth),
         "param_keys": np.array(sorted(reg.params)), "param_vals": np.array([float(reg.params[k]) for k in sorted(reg.params)]),
         "n_events": np.array(len(reg.events))}
    for i, ev in enumerate(reg.events):
        d["ev%d_levels" % i] = np.stack([ev.mean, ev.stdv, ev.ref_align, ev.ref_like])
        m = ev.model
        d["ev%d_model" % i] = np.stack([m.level_mean, m.level_stdv, m.sd_mean, m.sd_stdv])
        d["ev%d_trans" % i] = np.array([m.prob_skip, m.prob_stay, m.prob_extend, m.prob_insert, float(m.complement)])
        d["ev%d_seq" % i] = np.array(ev.sequence)
    return d


def unpack_region(z):
    reg = synth.SynthRegion()
    reg.sequence = str(z["sequence"])
    reg.truth = str(z["truth"])
    reg.params = {k: (int(v) if float(v).is_integer() and k != "lik_offset" else float(v))
                  for k, v in zip(z["param_keys"].tolist(), z["param_vals"].tolist())}
    for i in range(int(z["n_events"])):
        ev = synth.SynthEvent()
        lv = z["ev%d_levels" % i]
        ev.mean, ev.stdv, ev.ref_align, ev.ref_like = [np.ascontiguousarray(x) for x in lv]
        md = z["ev%d_model" % i]
        m = ev.model
        m.level_mean, m.level_stdv, m.sd_mean, m.sd_stdv = [np.ascontiguousarray(x) for x in md]
        t = z["ev%d_trans" % i]
        m.prob_skip, m.prob_stay, m.prob_extend, m.prob_insert = [float(x) for x in t[:4]]
        m.complement = bool(t[4])
        ev.sequence = str(z["ev%d_seq" % i])
        reg.events.append(ev)
    return reg


def aligns_array(al):
    return np.concatenate([np.stack(a) for a in al], axis=1) if al else np.zeros((2, 0))


def noisy(seq, rng, rate):
    """Deterministic corruption of a sequence (substitutions, deletions, insertions at rate/3 each)."""
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


def aligned_copy(reg, seq, al):
    """The region with a new sequence and the given per-event alignments."""
    import copy
    rr = copy.deepcopy(reg)
    rr.sequence = seq
    for ev, (ra, rl) in zip(rr.events, al):
        ev.ref_align, ev.ref_like = ra, rl
    return rr


def driver_fixture(ref, reg):
    """Outputs of the reference's driver-level entry points (swfull, MapAlignments, FindMutations, the Mutate loop,
    ViterbiMutate) on a region of the g_*.npz files: the d_*.npz files."""
    rng = np.random.default_rng(23)
    seeds = [ev.sequence for ev in reg.events[::2]] + [noisy(reg.sequence, rng, 0.08)]
    seeds.append(seeds[0])
    d = {"seeds": np.array(seeds)}
    acc, score, pairs = ref.swfull(reg.sequence, seeds[-2])
    d["sw_acc"], d["sw_score"], d["sw_pairs"] = np.array(acc), np.array(score), np.array(pairs, dtype=np.int32).reshape(-1, 2)
    d["ma_aligns"] = aligns_array(ref.map_alignments(reg, seeds[-2]))
    fm, a = ref.find_mutations(reg, seeds)
    d["fm_start"] = np.array([m[0] for m in fm], dtype=np.int32)
    d["fm_orig"], d["fm_mut"] = np.array([m[1] for m in fm]), np.array([m[2] for m in fm])
    d["fm_aligns"] = aligns_array(a)
    seq, nb, a = ref.mutate(reg, seeds, reps=3)
    d["mu_seq"], d["mu_nbases"], d["mu_aligns"] = np.array(seq), np.array(nb), aligns_array(a)
    rseq, _, ral = ref.refine(reg)                    # ViterbiMutate needs every event aligned
    rr = aligned_copy(reg, rseq, ral)
    d["vit_best"] = np.array(ref.viterbi_mutate(rr, nkeep=0))
    d["vit_samples"] = np.array(ref.viterbi_mutate(rr, nkeep=4, seed=1))
    return d


def split_aligns(arr, reg):
    """Inverse of aligns_array for a region: [(ref_align, ref_like)] per event."""
    out, at = [], 0
    for ev in reg.events:
        n = len(ev.mean)
        out.append((arr[0, at:at + n], arr[1, at:at + n]))
        at += n
    return out


def main():
    ref = binding.load("ref")
    for name, kw in GOLDEN:
        reg = synth.make_region(**kw)
        d = pack_region(reg)
        s, l, a = ref.score_alignments(reg, True)
        d["sa_scores"], d["sa_likes"], d["sa_aligns"] = s, l, aligns_array(a)
        pts, a = ref.score_points(reg)
        d["sp_scores"] = np.array([p[3] for p in pts])
        d["sp_aligns"] = aligns_array(a)
        st, og, mu = edge_mutations(reg.sequence, 7, count=60, max_len=4)
        sm, a = ref.score_mutations(reg, st, og, mu)
        d["sm_start"], d["sm_orig"], d["sm_mut"], d["sm_scores"] = np.array(st), np.array(og), np.array(mu), sm
        seq, nb, a = ref.refine(reg)
        d["rf_seq"], d["rf_nbases"], d["rf_aligns"] = np.array(seq), np.array(nb), aligns_array(a)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "events", len(reg.events), "points", len(pts), "refine", nb)
        dd = driver_fixture(ref, reg)
        np.savez_compressed(os.path.join(HERE, name.replace("g_", "d_") + ".npz"), **dd)
        print("  drivers: found", len(dd["fm_start"]), "mutate", int(dd["mu_nbases"]), "viterbi", len(str(dd["vit_best"][0])))


if __name__ == "__main__":
    main()
