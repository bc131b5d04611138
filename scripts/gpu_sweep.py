"""GPU: the randomised degenerate-region sweep of tests/test_oracle_random_sweep.py run through the CUDA path
(C-ABI) against the CPU checker: tiny regions (6-60 bases, 1-3 reads), narrow bands, events without alignment, with one
level, with non-ACGT bases; ScoreAlignments, ScorePoints, ScoreMutations and Refine in both precisions.

    gpurun --timeout 600 -- 'timeout 500 python scripts/gpu_sweep.py 600 > gpurun_out/gpu_sweep.log 2>&1'

Its first run on a B200 found two bugs the fixed cases had not (a zero-sized grid for a batch of events without levels;
the band planner taking the NaN ref_index of an event with ONE aligned level for a sorted array); 600 seeds x 2
precisions are clean since.  tests/test_gpu_random_sweep.py runs a part of it with the GPU suite.
Prints every mismatching seed with the entry point that differed; exit code 1 if there was one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from oracle import binding  # noqa: E402
from poreseq_b200 import poreseqcpp, synth  # noqa: E402
from util import edge_mutations, same_aligns  # noqa: E402


def native(ctx, reg, width_key=None):
    return poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params, width_key)


def aligns(nr, reg):
    return [nr.event_align(e) for e in range(len(reg.events))]


def tiny_region(seed, lo, hi, rng):
    params = dict(realign_width=int(rng.integers(3, 40)), scoring_width=int(rng.integers(2, 15)),
                  point_width=int(rng.integers(1, 9)), lik_offset=float(rng.choice([0.0, 2.0, 4.5, 9.0])))
    return synth.make_region(int(rng.integers(lo, hi)), int(rng.integers(1, 4)), seed=seed + 1,
                             draft_error=float(rng.choice([0, 0.05, 0.2])), partial=float(rng.choice([0, 0.5, 1.0])),
                             p_unaligned=float(rng.choice([0, 0.3, 0.9])), jitter=int(rng.choice([0, 2, 6])), params=params)


def main(n=None, first=0):
    if n is None:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 240
    binding.build("oracle")
    orc = binding.load("oracle")
    ctx = poreseqcpp.Context(0)
    bad = 0
    for precision in ("exact", "fast"):
        ctx.set_precision(precision)
        for seed in range(first, first + n):
            rng = np.random.default_rng(1000 + seed)
            reg = tiny_region(seed, 6, 60, rng)
            kind = seed % 5
            if kind == 1:
                reg.events[0].ref_align[:] = 0
            if kind == 2 and len(reg.sequence) > 8:
                s = list(reg.sequence)
                s[int(rng.integers(0, len(s)))] = "N"
                reg.sequence = "".join(s)
            if kind == 3:
                ev = reg.events[-1]
                for f in ("mean", "stdv", "ref_align", "ref_like"):
                    setattr(ev, f, np.ascontiguousarray(getattr(ev, f)[:1]))
            if kind == 4:
                reg.events[0].ref_align[:] = -1
            what = []
            try:
                if len(reg.sequence) < 5:
                    continue                                        # documented limit of the library
                s, l, a = orc.score_alignments(reg, True)
                nr = native(ctx, reg)
                gs, gl = nr.score_alignments(True)
                if not (np.array_equal(gs, s) and np.array_equal(gl, l) and same_aligns(aligns(nr, reg), a)):
                    what.append("score_alignments")
                want, a = orc.score_points(reg)
                nr = native(ctx, reg, "point_width")
                st, og, mu, sc = nr.score_points()
                w = np.array([x[3] for x in want])
                same = np.array_equal(sc, w) if precision == "exact" else (
                    len(sc) == len(w) and np.array_equal(sc[w >= 0], w[w >= 0]) and bool(np.all(np.abs(sc - w) <= 1e-4 * np.abs(w) + 1e-3)))
                if not (same and same_aligns(aligns(nr, reg), a)):
                    what.append("score_points")
                st, og, mu = edge_mutations(reg.sequence, seed, count=30)
                want, a = orc.score_mutations(reg, st, og, mu)
                nr = native(ctx, reg)
                got = nr.score_mutations(st, og, mu)
                same = np.array_equal(got, want) if precision == "exact" else (
                    np.array_equal(got[want >= 0], want[want >= 0]) and bool(np.all(np.abs(got - want) <= 1e-4 * np.abs(want) + 1e-3)))
                if not (same and same_aligns(aligns(nr, reg), a)):
                    what.append("score_mutations")
                seq, nb, a = orc.refine(reg)
                nr = native(ctx, reg, "point_width")
                if not (nr.refine() == nb and nr.sequence() == seq and same_aligns(aligns(nr, reg), a)):
                    what.append("refine")
            except Exception as e:                                  # noqa: BLE001
                what.append("exception %r" % (e,))
            if what:
                bad += 1
                print("MISMATCH precision=%s seed=%d kind=%d len=%d events=%d params=%s: %s"
                      % (precision, seed, kind, len(reg.sequence), len(reg.events), reg.params, ", ".join(what)), flush=True)
    print("gpu_sweep: %d regions x 2 precisions, %d mismatching" % (n, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
