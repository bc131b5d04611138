"""Measures the error of the FAST mode's FP32 scan against the exact FP64 kernels on the GPU: per configuration the
largest |FP32 total - exact total| over all single-base edits (tau forced to 0 so that nothing is re-scored), the
same per event of the region (the per-pair bound PS_FAST_PAIR_ERR in csrc/ps_host.cu must stay above it), how many
edits the derived threshold sends to the exact pass, and the largest RELATIVE error the shipped FAST mode leaves.

    python scripts/fast_error.py > profiles/r2_fast_error.txt
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from poreseq_b200 import poreseqcpp, synth  # noqa: E402

PAIR_ERR = 5e-5
CASES = [
    ("1kb x10 W=20 clean", dict(length=1000, coverage=10, seed=1), "point_width"),
    ("1kb x10 W=20 draft 3%", dict(length=1000, coverage=10, seed=2, draft_error=0.03), "point_width"),
    ("1kb x10 W=20 partial", dict(length=1000, coverage=10, seed=3, partial=0.4, draft_error=0.01), "point_width"),
    ("1kb x10 W=100", dict(length=1000, coverage=10, seed=4, draft_error=0.01), "scoring_width"),
    ("3kb x30 W=20", dict(length=3000, coverage=30, seed=5, draft_error=0.02), "point_width"),
    ("10kb x15 W=20", dict(length=10000, coverage=15, seed=6, draft_error=0.01), "point_width"),
    ("2kb x100 W=100", dict(length=2000, coverage=100, seed=7), "scoring_width"),
]


def main():
    exact = poreseqcpp.Context(0)
    fast = poreseqcpp.Context(0)
    fast.set_precision("fast")
    os.environ["PORESEQ_B200_TAU"] = "0"
    raw = poreseqcpp.Context(0)                  # FP32 totals as they are, nothing re-scored
    raw.set_precision("fast")
    del os.environ["PORESEQ_B200_TAU"]
    worst_pair, worst_rel = 0.0, 0.0
    for name, kw, wk in CASES:
        reg = synth.make_region(**kw)
        E = len(reg.events)
        out = []
        for ctx in (exact, raw, fast):
            nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params, wk)
            t0 = time.perf_counter()
            st, og, mu, sc = nr.score_points()
            out.append((sc.copy(), time.perf_counter() - t0, ctx.last_timing()["mutscore"]))
            nr.close()
        ex, rw, fs = out[0][0], out[1][0], out[2][0]
        err = np.abs(rw - ex)
        tau = PAIR_ERR * 1e4 * E
        kept = ex <= -tau                                   # edits that keep their FP32 value in the shipped mode
        rel = np.abs(fs - ex) / np.maximum(np.abs(ex), 1e-300)
        same = fs[~kept] == ex[~kept]
        worst_pair = max(worst_pair, err.max() / E)
        worst_rel = max(worst_rel, rel.max())
        print("%-24s E=%3d edits=%6d | raw FP32: max|err| %.3e (per event %.3e), 99.9%% %.3e | tau=%.2f flags %5.2f%% "
              "(exact==: %s) | shipped FAST max rel err %.3e | mutscore ms exact %.2f raw %.2f fast %.2f"
              % (name, E, len(ex), err.max(), err.max() / E, np.quantile(err, 0.999), tau, 100.0 * (~kept).mean(),
                 bool(same.all()), rel.max(), out[0][2], out[1][2], out[2][2]))
        for t in (0.1, 0.5, 1, 2, 5, 10):
            sys.stdout.write("    score > -%g: %.2f%%" % (t, 100.0 * (ex > -t).mean()))
        sys.stdout.write("\n")
    print("worst per-event error %.3e (bound PS_FAST_PAIR_ERR = %.1e); worst relative error of the shipped FAST mode %.3e"
          % (worst_pair, PAIR_ERR, worst_rel))


if __name__ == "__main__":
    main()
