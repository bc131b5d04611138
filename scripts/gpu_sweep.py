"""GPU: the randomised degenerate-region sweep of tests/test_oracle_random_sweep.py run through the CUDA path
(C-ABI) against the CPU checker: tiny regions (6-60 bases, 1-3 reads), narrow bands, events without alignment, with one
level, with non-ACGT bases; ScoreAlignments, ScorePoints, ScoreMutations and Refine in both precisions.

    gpurun --timeout 600 -- 'timeout 500 python scripts/gpu_sweep.py 600 > gpurun_out/gpu_sweep.log 2>&1'

Its first run on a B200 found two bugs the fixed cases had not (a zero-sized grid for a batch of events without levels;
the band planner taking the NaN ref_index of an event with ONE aligned level for a sorted array), then a third and a
fourth (ViterbiMutate stopped at the last refend although extrapolated ref_index values still matched positions beyond
it; the FP32 scan took a real diagonal out of a seed column whose band ends right above the first narrow row);
5400 + 1100 seeds x 2 precisions are clean since.  tests/test_gpu_random_sweep.py runs a part of it with the GPU suite.
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
                    len(sc) == len(w) and np.array_equal(sc[w >= 0], w[w >= 0]) and bool(np.all(np.abs(sc - w) <= 1e-4 * np.abs(w))))
                if not (same and same_aligns(aligns(nr, reg), a)):
                    what.append("score_points")
                st, og, mu = edge_mutations(reg.sequence, seed, count=30)
                want, a = orc.score_mutations(reg, st, og, mu)
                nr = native(ctx, reg)
                got = nr.score_mutations(st, og, mu)
                same = np.array_equal(got, want) if precision == "exact" else (
                    np.array_equal(got[want >= 0], want[want >= 0]) and bool(np.all(np.abs(got - want) <= 1e-4 * np.abs(want))))
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


def main_drivers(n=None, first=0):
    """The driver-level half (tests/test_oracle_random_sweep.py: test_driver_entry_points_on_degenerate_regions and
    test_viterbi_on_small_regions) through the CUDA path: swfull on the device, MapAlignments, FindMutations with a
    repeated seed, the Mutate loop, ViterbiMutate (best path and 8 sampled walks on the same rand() stream)."""
    import copy
    import ctypes
    if n is None:
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    binding.build("oracle")
    orc = binding.load("oracle")
    ctx = poreseqcpp.Context(0)
    libc = ctypes.CDLL("libc.so.6")
    bad = 0
    for precision in ("exact", "fast"):
        ctx.set_precision(precision)
        for seed in range(first, first + n):
            rng = np.random.default_rng(5000 + seed)
            reg = tiny_region(seed, 12, 90, rng)
            kind = seed % 4
            if kind == 1:
                reg.events[0].ref_align[:] = 0
            if kind == 2:
                for ev in reg.events:
                    ev.model.prob_skip = float(rng.choice([0.5, 1.0, 1.5]))
                    ev.model.prob_insert = float(rng.choice([0.01, 1.2]))
            seeds = [ev.sequence for ev in reg.events[::2]][:3] + [synth.corrupt_sequence(reg.sequence, 0.1, rng)[0]]
            seeds.append(seeds[0])
            what = []
            try:
                x = orc.swfull(reg.sequence, seeds[-2])
                y = poreseqcpp.swalign_device(ctx, reg.sequence, seeds[-2])
                if not (x[1] == y[0] and x[0] == y[1] and x[2] == [tuple(p) for p in y[2]]):
                    what.append("swfull")
                nr = native(ctx, reg)
                nr.map_alignments(seeds[-2])
                if not same_aligns(aligns(nr, reg), orc.map_alignments(reg, seeds[-2])):
                    what.append("map_alignments")
                f, a = orc.find_mutations(reg, seeds)
                nr = native(ctx, reg)
                if not (nr.find_mutations(seeds) == f and same_aligns(aligns(nr, reg), a)):
                    what.append("find_mutations")
                m = orc.mutate(reg, seeds, reps=2)
                nr = native(ctx, reg)
                nb = nr.mutate(seeds, reps=2)
                if not (nr.sequence() == m[0] and nb == m[1] and same_aligns(aligns(nr, reg), m[2])):
                    what.append("mutate")
                # ViterbiMutate needs every event aligned (the reference dereferences an empty path otherwise)
                seq, _, al = orc.refine(reg)
                if len(seq) >= 5 and all((np.asarray(ra) > 0).any() for ra, _ in al):
                    rr = copy.deepcopy(reg)
                    rr.sequence = seq
                    for ev, (ra, rl) in zip(rr.events, al):
                        ev.ref_align, ev.ref_like = ra, rl
                    kw = dict(skip=float(rng.choice([0.05, 0.2])), stay=float(rng.choice([0.01, 0.1])),
                              mut_min=float(rng.choice([0.0, 0.33])), mut_max=float(rng.choice([0.75, 1.0])))
                    want = orc.viterbi_mutate(rr, nkeep=8, seed=seed + 1, **kw)
                    nr = native(ctx, rr)
                    libc.srand(seed + 1)
                    if nr.viterbi_mutate(nkeep=8, **kw) != want:
                        what.append("viterbi_mutate")
            except Exception as e:                                  # noqa: BLE001
                what.append("exception %r" % (e,))
            if what:
                bad += 1
                print("MISMATCH(drivers) precision=%s seed=%d kind=%d len=%d events=%d params=%s: %s"
                      % (precision, seed, kind, len(reg.sequence), len(reg.events), reg.params, ", ".join(what)), flush=True)
    print("gpu_sweep drivers: %d regions x 2 precisions, %d mismatching" % (n, bad))
    return 1 if bad else 0


def main_mid(n=None, first=0):
    """Mid-sized regions (150-500 bases, 1-4 reads, realign_width 20-300): the wavefront classes of the wide fill, the
    batched strip switches, interior and masked tiles, the batched and the handle-less entry points and the FP32
    score-only fill, with jittered / partial / missing alignments."""
    if n is None:
        n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    binding.build("oracle")
    orc = binding.load("oracle")
    ctx = poreseqcpp.Context(0)
    bad = 0
    for precision in ("exact", "fast"):
        ctx.set_precision(precision)
        for seed in range(first, first + n):
            rng = np.random.default_rng(7000 + seed)
            params = dict(realign_width=int(rng.choice([20, 60, 150, 300])), scoring_width=int(rng.integers(5, 40)),
                          point_width=int(rng.integers(3, 21)), lik_offset=float(rng.choice([2.0, 4.5])))
            regs = [synth.make_region(int(rng.integers(150, 500)), int(rng.integers(1, 5)), seed=10 * seed + k + 1,
                                      draft_error=float(rng.choice([0, 0.05, 0.15])), partial=float(rng.choice([0, 0.3, 1.0])),
                                      p_unaligned=float(rng.choice([0, 0.2])), jitter=int(rng.choice([0, 3, 10])), params=params)
                    for k in range(2)]
            if seed % 3 == 1:
                s = list(regs[0].sequence)
                s[int(rng.integers(0, len(s)))] = "N"
                regs[0].sequence = "".join(s)
            what = []
            try:
                wants = [orc.score_points(r) for r in regs]
                nrs = [native(ctx, r, "point_width") for r in regs]
                outs = poreseqcpp.score_points_batch(ctx, nrs)
                packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
                direct = poreseqcpp.score_points_direct(ctx, packs)
                for k, r in enumerate(regs):
                    w = np.array([x[3] for x in wants[k][0]])
                    for name, sc in (("batch", outs[k][3]), ("direct", direct[k][3])):
                        same = np.array_equal(sc, w) if precision == "exact" else (
                            len(sc) == len(w) and np.array_equal(sc[w >= 0], w[w >= 0]) and bool(np.all(np.abs(sc - w) <= 1e-4 * np.abs(w))))
                        if not same:
                            what.append("score_points(%s)[%d]" % (name, k))
                    if not same_aligns(aligns(nrs[k], r), wants[k][1]):
                        what.append("score_points aligns[%d]" % k)
                poreseqcpp.close_regions(nrs)
                for k, r in enumerate(regs):
                    sa, _, _ = orc.score_alignments(r, True)
                    nr = native(ctx, r)
                    se = nr.score_events()
                    ok = np.array_equal(se, sa) if precision == "exact" else bool(np.all(np.abs(se - sa) <= 1e-4 * np.abs(sa)))
                    if not ok:
                        what.append("score_events[%d]" % k)
                    seq, nb, a = orc.refine(r)
                    nr = native(ctx, r, "point_width")
                    if not (nr.refine() == nb and nr.sequence() == seq and same_aligns(aligns(nr, r), a)):
                        what.append("refine[%d]" % k)
                    # multi-base edits at scoring_width (poreseq variant), seed-driven candidate search, the Mutate loop
                    st, og, mu = edge_mutations(r.sequence, seed, count=80, max_len=6)
                    want, a = orc.score_mutations(r, st, og, mu)
                    nr = native(ctx, r)
                    got = nr.score_mutations(st, og, mu)
                    same = np.array_equal(got, want) if precision == "exact" else (
                        np.array_equal(got[want >= 0], want[want >= 0]) and bool(np.all(np.abs(got - want) <= 1e-4 * np.abs(want))))
                    if not (same and same_aligns(aligns(nr, r), a)):
                        what.append("score_mutations[%d]" % k)
                    if k == 0 and len(r.events) >= 2:
                        seeds = [ev.sequence for ev in r.events[::2]][:3] + [synth.corrupt_sequence(r.sequence.replace("N", "A"), 0.08, rng)[0]]
                        f, a = orc.find_mutations(r, seeds)
                        nr = native(ctx, r)
                        if not (nr.find_mutations(seeds) == f and same_aligns(aligns(nr, r), a)):
                            what.append("find_mutations")
                        m = orc.mutate(r, seeds, reps=2)
                        nr = native(ctx, r)
                        nb = nr.mutate(seeds, reps=2)
                        if not (nr.sequence() == m[0] and nb == m[1] and same_aligns(aligns(nr, r), m[2])):
                            what.append("mutate")
            except Exception as e:                                  # noqa: BLE001
                what.append("exception %r" % (e,))
            if what:
                bad += 1
                print("MISMATCH(mid) precision=%s seed=%d lens=%s events=%s params=%s: %s"
                      % (precision, seed, [len(r.sequence) for r in regs], [len(r.events) for r in regs], params, ", ".join(what)), flush=True)
    print("gpu_sweep mid: %d x 2 regions x 2 precisions, %d mismatching" % (n, bad))
    return 1 if bad else 0


def main_large(n=None, first=0):
    """Long events (2-5 kb, 1-2 reads, realign_width 100 / 300): many strips per thread in the wide fill, every wavefront
    class, backtrace walks over thousands of levels, long updaterefs scans -- ScoreAlignments (+ profile), ScorePoints
    and Refine against the checker in both precisions (the checker runs once per region: ~20 s of CPU each)."""
    if n is None:
        n = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    binding.build("oracle")
    orc = binding.load("oracle")
    ctx = poreseqcpp.Context(0)
    bad = 0
    for seed in range(first, first + n):
        rng = np.random.default_rng(13000 + seed)
        reg = synth.make_region(int(rng.integers(2000, 5000)), int(rng.integers(1, 3)), seed=seed + 1,
                                draft_error=float(rng.choice([0.02, 0.1])), partial=float(rng.choice([0, 0.3])),
                                p_unaligned=float(rng.choice([0, 0.1])), jitter=int(rng.choice([0, 5, 40])),
                                params=dict(realign_width=int(rng.choice([100, 300])), scoring_width=30,
                                            point_width=int(rng.integers(5, 21))))
        s, l, a_sa = orc.score_alignments(reg, True)
        want, a_sp = orc.score_points(reg)
        w = np.array([x[3] for x in want])
        seq, nb, a_rf = orc.refine(reg)
        for precision in ("exact", "fast"):
            ctx.set_precision(precision)
            what = []
            try:
                nr = native(ctx, reg)
                gs, gl = nr.score_alignments(True)
                if not (np.array_equal(gs, s) and np.array_equal(gl, l) and same_aligns(aligns(nr, reg), a_sa)):
                    what.append("score_alignments")
                nr = native(ctx, reg, "point_width")
                sc = nr.score_points()[3]
                same = np.array_equal(sc, w) if precision == "exact" else (
                    len(sc) == len(w) and np.array_equal(sc[w >= 0], w[w >= 0]) and bool(np.all(np.abs(sc - w) <= 1e-4 * np.abs(w))))
                if not (same and same_aligns(aligns(nr, reg), a_sp)):
                    what.append("score_points")
                nr = native(ctx, reg, "point_width")
                if not (nr.refine() == nb and nr.sequence() == seq and same_aligns(aligns(nr, reg), a_rf)):
                    what.append("refine")
            except Exception as e:                                  # noqa: BLE001
                what.append("exception %r" % (e,))
            if what:
                bad += 1
                print("MISMATCH(large) precision=%s seed=%d len=%d events=%d params=%s: %s"
                      % (precision, seed, len(reg.sequence), len(reg.events), reg.params, ", ".join(what)), flush=True)
    print("gpu_sweep large: %d regions x 2 precisions, %d mismatching" % (n, bad))
    return 1 if bad else 0


def main_psalign(n=None, first=0):
    """The PSAlign mirror (poreseq_b200/poreseqcpp.py, pyx:189-472) on tiny regions: ScoreMutations -> ApplyMuts (the
    MakeMutations recursion on a caller-scored list of multi-base edits), Copy independence -- against the checker's
    score_mutations / make_mutations (RealignTo is Python on both sides: PSEvent.mapaligns)."""
    from poreseq_b200 import drivers
    from poreseq_b200.Util import MutationInfo
    if n is None:
        n = 100
    binding.build("oracle")
    orc = binding.load("oracle")
    ctx = poreseqcpp.Context(0)
    bad = 0
    for precision in ("exact", "fast"):
        ctx.set_precision(precision)
        for seed in range(first, first + n):
            rng = np.random.default_rng(17000 + seed)
            reg = tiny_region(seed, 12, 90, rng)
            what = []
            try:
                pa = drivers.make_psalign(reg)
                pa.ctx = ctx
                st, og, mu = synth.random_mutations(reg.sequence, 40, rng, 4)
                want_sc, a = orc.score_mutations(reg, st, og, mu)
                infos = []
                for s_, o_, m_ in zip(st, og, mu):
                    mi = MutationInfo(); mi.start, mi.orig, mi.mut = s_, o_, m_
                    infos.append(mi)
                keep = pa.Copy()
                scored = pa.ScoreMutations(infos)
                got_sc = np.array([x.score for x in scored])
                if not np.array_equal(got_sc[want_sc >= 0], want_sc[want_sc >= 0]):
                    what.append("ScoreMutations")
                # ApplyMuts with the reference's scores on both sides; PSAlign.ScoreMutations does not write the
                # realignment back (pyx:310-345), so both start from the region's own alignments
                seq, nb, a2 = orc.make_mutations(reg, st, og, mu, want_sc)
                for x, w in zip(scored, want_sc):
                    x.score = float(w)
                pa.ApplyMuts(scored)
                if not (pa.sequence == seq and same_aligns([(ev.ref_align, ev.ref_like) for ev in pa.events], a2)):
                    what.append("ApplyMuts")
                if keep.sequence != reg.sequence or any(not np.array_equal(x.ref_align, y.ref_align) for x, y in zip(keep.events, reg.events)):
                    what.append("Copy")
            except Exception as e:                                  # noqa: BLE001
                what.append("exception %r" % (e,))
            if what:
                bad += 1
                print("MISMATCH(psalign) precision=%s seed=%d len=%d events=%d: %s"
                      % (precision, seed, len(reg.sequence), len(reg.events), ", ".join(what)), flush=True)
    print("gpu_sweep psalign: %d regions x 2 precisions, %d mismatching" % (n, bad))
    return 1 if bad else 0


def main_consensus(n=None, first=0):
    """The whole Mutate.py loop below the C-ABI (ps_consensus_batch, regions in lockstep, region-private rand() streams)
    against the same loop driven through the checker from a fresh rand() stream, on small regions of 3-6 reads."""
    from poreseq_b200 import drivers
    from util import reference_consensus
    if n is None:
        n = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    binding.build("oracle")
    chk = binding.load("ref") if binding.available("ref") else binding.load("oracle")
    ctx = poreseqcpp.Context(0)
    bad = 0
    for precision in ("exact", "fast"):
        ctx.set_precision(precision)
        for base in range(first, first + n, 4):
            regs = []
            for seed in range(base, min(base + 4, first + n)):
                rng = np.random.default_rng(11000 + seed)
                regs.append(synth.make_region(int(rng.integers(80, 260)), int(rng.integers(3, 7)), seed=seed + 1,
                                              draft_error=float(rng.choice([0.03, 0.08, 0.15])), partial=float(rng.choice([0, 0.3])),
                                              params=dict(realign_width=60, scoring_width=int(rng.integers(8, 25)),
                                                          point_width=int(rng.integers(4, 12)), end_trim=int(rng.choice([0, 10])))))
            what = []
            try:
                got = drivers.consensus_native(regs, ctx=ctx, reps=3, in_flight=4)
                for k, r in enumerate(regs):
                    if got[k][0] != reference_consensus(chk, r, reps=3):
                        what.append("consensus[%d]" % (base + k))
            except Exception as e:                                  # noqa: BLE001
                what.append("exception %r" % (e,))
            if what:
                bad += 1
                print("MISMATCH(consensus) precision=%s seeds %d..: %s" % (precision, base, ", ".join(what)), flush=True)
    print("gpu_sweep consensus: %d regions x 2 precisions, %d mismatching batches" % (n, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    rc = main()
    rc = main_drivers() or rc
    rc = main_mid() or rc
    sys.exit(main_consensus() or rc)
