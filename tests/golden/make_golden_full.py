"""Full-size known answers from the REFERENCE's own C++ (oracle/_ref) for the BASELINE.json configs the small
fixtures do not reach: the whole consensus loop at 10 kb (configs[2]: 30x coverage; configs[4]: one region at 50x) and
one `poreseq variant` region (configs[3]: 10 kb, 100x coverage, 1 k multi-base edits at scoring_width 100).

    make -C oracle ref && python tests/golden/make_golden_full.py c3 c2 c4 [c2s]

CPU-hours on one core each (the reference pays O(L) per (mutation, event) pair), so the cases run as
separate processes.  The inputs are NOT stored (18-30 MB of float64 per region): poreseq_b200.synth is seeded numpy,
the fixture keeps a checksum of every input array so that a drifted generator is detected instead of silently
comparing different problems, plus every stage's outputs (sequence after each stage of the Mutate.py loop, bases
changed, a checksum of all events' alignments; for variant scoring the scores themselves).
"""
import copy
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import binding  # noqa: E402
from poreseq_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (kind, make_region kwargs)
CASES = {
    "c2": ("loop", dict(length=10000, coverage=30, seed=7, draft_error=0.10)),       # configs[2]
    "c4": ("loop", dict(length=10000, coverage=50, seed=11, draft_error=0.10)),      # one region of configs[4]
    "c2s": ("loop", dict(length=2000, coverage=30, seed=13, draft_error=0.10)),      # the same loop, minutes instead of hours
    "c3": ("variant", dict(length=10000, coverage=100, seed=17)),                    # one region of configs[3]
}
N_VARIANT_MUTS = 1200


def input_digest(reg):
    h = hashlib.sha256()
    h.update(reg.sequence.encode())
    for ev in reg.events:
        for a in (ev.mean, ev.stdv, ev.ref_align):
            h.update(np.ascontiguousarray(a, dtype="f8").tobytes())
        h.update(ev.sequence.encode())
        m = ev.model
        for a in (m.level_mean, m.level_stdv, m.sd_mean, m.sd_stdv):
            h.update(np.ascontiguousarray(a, dtype="f8").tobytes())
    return h.hexdigest()


def align_digest(events_or_aligns):
    """sha256 over every event's ref_align (integers) -- ref_like is compared by value where it matters."""
    h = hashlib.sha256()
    for it in events_or_aligns:
        ra = it[0] if isinstance(it, tuple) else it.ref_align
        h.update(np.ascontiguousarray(ra, dtype="f8").tobytes())
    return h.hexdigest()


def variant_mutations(reg, count=N_VARIANT_MUTS):
    rng = np.random.default_rng(4242)
    return synth.random_mutations(reg.sequence, count, rng, max_len=4)


def reference_loop(ref, reg, reps=4, log=None):
    """poreseq/Mutate.py:47-99 through the reference's C++: returns [(stage, sequence, nbases, align digest)]."""
    rr = copy.deepcopy(reg)
    libc = __import__("ctypes").CDLL("libc.so.6")
    libc.srand(1)
    stages = []

    def took(name, seq, nb, al):
        rr.sequence = seq
        for ev, (ra, rl) in zip(rr.events, al):
            ev.ref_align, ev.ref_like = ra, rl
        stages.append((name, seq, int(nb), align_digest(al)))
        if log:
            log("%s: %d bases changed, length %d" % (name, nb, len(seq)))

    seq, nb, al = ref.mutate(rr, [ev.sequence for ev in rr.events[::2]], reps=reps)
    took("mutate_self", seq, nb, al)
    for k in range(reps):
        seeds = ref.viterbi_mutate(rr, nkeep=16, seed=None)
        seq, nb, al = ref.mutate(rr, seeds, reps=reps)
        took("mutate_viterbi_%d" % k, seq, nb, al)
        seq, nb, al = ref.refine(rr)
        took("refine_%d" % k, seq, nb, al)
        if nb == 0:
            break
    return stages


def main():
    ref = binding.load("ref")
    for name in sys.argv[1:]:
        kind, kw = CASES[name]
        t0 = time.time()
        reg = synth.make_region(**kw)
        d = {"kw_keys": np.array(sorted(kw)), "kw_vals": np.array([float(kw[k]) for k in sorted(kw)]),
             "input_sha256": np.array(input_digest(reg)), "kind": np.array(kind)}
        log = lambda s: (sys.stderr.write("[%s %.0fs] %s\n" % (name, time.time() - t0, s)), sys.stderr.flush())
        if kind == "loop":
            stages = reference_loop(ref, reg, log=log)
            d["stage_names"] = np.array([s[0] for s in stages])
            d["stage_seqs"] = np.array([s[1] for s in stages])
            d["stage_nbases"] = np.array([s[2] for s in stages])
            d["stage_aligns"] = np.array([s[3] for s in stages])
        else:
            st, og, mu = variant_mutations(reg)
            scores, al = ref.score_mutations(reg, st, og, mu)
            d["scores"] = scores
            d["aligns_sha256"] = np.array(align_digest(al))
            d["ref_like_sum"] = np.array([float(np.sum(a[1])) for a in al])
            ev_scores, _, _ = ref.score_alignments(reg)
            d["event_scores"] = ev_scores
        d["seconds"] = np.array(time.time() - t0)
        np.savez_compressed(os.path.join(HERE, "f_%s.npz" % name), **d)
        log("done")


if __name__ == "__main__":
    main()
