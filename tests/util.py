"""Shared helpers for the parity tests."""
import numpy as np

from poreseq_b200 import synth

SMALL = dict(realign_width=60, scoring_width=15, point_width=8)

# (name, kwargs for synth.make_region) -- small enough for the CPU checkers to finish in seconds
CASES = [
    ("clean", dict(length=400, coverage=4, seed=1, params=SMALL)),
    ("draft_partial", dict(length=400, coverage=4, seed=2, draft_error=0.03, partial=0.3, params=SMALL)),
    ("ragged", dict(length=400, coverage=4, seed=3, draft_error=0.05, partial=0.5, p_unaligned=0.3, jitter=3, params=SMALL)),
    ("default_widths", dict(length=700, coverage=3, seed=4, draft_error=0.02)),
]


def region(name):
    return synth.make_region(**dict(CASES)[name])


def edge_mutations(seq, seed, count=200, max_len=5):
    """Random multi-base edits plus the boundary cases of cpp/MakeMutations.cpp:46 and
    cpp/Sequence.h:41-46 (start at / past the end, long deletions at 0, empty regions)."""
    rng = np.random.default_rng(seed)
    st, og, mu = synth.random_mutations(seq, count, rng, max_len)
    L = len(seq)
    extra = [(L, "", "A"), (L - 1, seq[-1:], ""), (L + 1, "", "C"), (0, "", "ACGTACGT"), (0, seq[:5], ""),
             (L + 7, "", "G"), (L - 4, seq[-4:], ""), (L - 5, seq[-5:-2], "TT"), (1, seq[1:2], "GGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGG"),
             (3, seq[3:30], ""), (L - 2, "", "ACGT"), (2, "", "N"), (5, seq[5:6], "N")]
    for s, o, m in extra:
        st.append(s); og.append(o); mu.append(m)
    return st, og, mu


def same_aligns(a, b):
    return all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a, b))


def build_cython_stub(tmp_dir):
    """Builds examples/cython_stub (the binding of INTEGRATION.md section 2) in a scratch directory; returns the
    directory to put on sys.path, or None when the toolchain refuses (reported by the caller)."""
    import os
    import shutil
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = os.path.join(str(tmp_dir), "cython_stub")
    shutil.copytree(os.path.join(root, "examples", "cython_stub"), dst)
    env = dict(os.environ, PORESEQ_B200_ROOT=root)
    proc = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=dst, env=env, capture_output=True, text=True)
    if proc.returncode != 0:
        return None, proc.stdout[-2000:] + proc.stderr[-2000:]
    return dst, ""


def reference_consensus(drv, reg, reps=4):
    """poreseq/Mutate.py:47-99 driven through the checker (`drv`: the reference's own C++ when oracle/_ref is built,
    else the restatement), from a fresh rand() stream: the end-trimmed consensus sequence."""
    import copy
    rr = copy.deepcopy(reg)

    def sync(al):
        for ev, (ra, rl) in zip(rr.events, al):
            ev.ref_align, ev.ref_like = ra, rl

    if len(rr.events) < 5:
        return rr.sequence
    drv.srand(1)
    seq, _, al = drv.mutate(rr, [ev.sequence for ev in rr.events[::2]], reps=reps)
    rr.sequence = seq; sync(al)
    for _ in range(reps):
        seeds = drv.viterbi_mutate(rr, nkeep=16, seed=None)
        seq, _, al = drv.mutate(rr, seeds, reps=4)          # Mutate.py:76: pa.Mutate(seqs='viterbi'), reps defaults to 4
        rr.sequence = seq; sync(al)
        seq, nb, al = drv.refine(rr)
        rr.sequence = seq; sync(al)
        if nb == 0:
            break
    seq = rr.sequence
    if "end_trim" in rr.params and len(seq) > 2 * rr.params["end_trim"]:      # Mutate.py:85-86 as written: an end_trim of 0
        t = int(rr.params["end_trim"])                                        # slices [0:-0] = the empty string
        seq = seq[t:-t]
    return seq
