"""One seed of scripts/gpu_sweep.py (main sweep) in detail.   usage: sweep_one.py seed [exact|fast]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np
from oracle import binding
from poreseq_b200 import poreseqcpp
import gpu_sweep

seed = int(sys.argv[1]); precision = sys.argv[2] if len(sys.argv) > 2 else "exact"
binding.build("oracle"); orc = binding.load("oracle")
ctx = poreseqcpp.Context(0); ctx.set_precision(precision)
rng = np.random.default_rng(1000 + seed)
reg = gpu_sweep.tiny_region(seed, 6, 60, rng)
kind = seed % 5
if kind == 1: reg.events[0].ref_align[:] = 0
if kind == 2 and len(reg.sequence) > 8:
    s = list(reg.sequence); s[int(rng.integers(0, len(s)))] = "N"; reg.sequence = "".join(s)
if kind == 3:
    ev = reg.events[-1]
    for f in ("mean", "stdv", "ref_align", "ref_like"): setattr(ev, f, np.ascontiguousarray(getattr(ev, f)[:1]))
if kind == 4: reg.events[0].ref_align[:] = -1
print("len", len(reg.sequence), "events", [len(e.mean) for e in reg.events], {k: reg.params[k] for k in ("realign_width", "scoring_width", "point_width", "lik_offset")})
want, a = orc.score_points(reg)
nr = gpu_sweep.native(ctx, reg, "point_width")
st, og, mu, sc = nr.score_points()
w = np.array([x[3] for x in want])
print("n", len(sc), len(w), "aligns same", gpu_sweep.same_aligns(gpu_sweep.aligns(nr, reg), a))
print(">=0 identical", np.array_equal(sc[w >= 0], w[w >= 0]), "sign flips", int(((sc >= 0) != (w >= 0)).sum()))
err = np.abs(sc - w); rel = err / np.maximum(np.abs(w), 1e-300)
bad = np.nonzero(err > 1e-4 * np.abs(w))[0]
print("worst rel", float(rel.max()), "violations of pure 1e-4 rel:", len(bad))
for i in bad[:10]: print("  mut", i, want[i][:3], "gpu", sc[i], "ref", w[i], "abs err", err[i])
