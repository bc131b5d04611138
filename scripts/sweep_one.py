"""One seed of scripts/gpu_sweep.py in detail: which output of Refine differs from the checker.
usage: sweep_one.py seed [exact|fast]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np
from oracle import binding
from poreseq_b200 import poreseqcpp
import gpu_sweep

seed = int(sys.argv[1]); precision = sys.argv[2] if len(sys.argv) > 2 else "exact"
binding.build("oracle"); orc = binding.load("oracle")
ref = binding.load("ref") if binding.available("ref") else None
ctx = poreseqcpp.Context(0); ctx.set_precision(precision)
rng = np.random.default_rng(1000 + seed)
reg = gpu_sweep.tiny_region(seed, 6, 60, rng)
print("len", len(reg.sequence), "events", [len(e.mean) for e in reg.events], reg.params["realign_width"], reg.params["point_width"])
seq, nb, a = orc.refine(reg)
if ref is not None:
    seq2, nb2, a2 = ref.refine(reg)
    print("oracle == compiled reference:", seq == seq2, nb == nb2, gpu_sweep.same_aligns(a, a2))
nr = gpu_sweep.native(ctx, reg, "point_width")
gnb = nr.refine(); gseq = nr.sequence(); ga = gpu_sweep.aligns(nr, reg)
print("nb", gnb, nb, "seq same", gseq == seq)
if gseq != seq: print(gseq); print(seq)
for e, (x, y) in enumerate(zip(ga, a)):
    for k, nm in enumerate(("ref_align", "ref_like")):
        if not np.array_equal(np.asarray(x[k]), np.asarray(y[k])):
            d = np.nonzero(np.asarray(x[k]) != np.asarray(y[k]))[0]
            print("event", e, nm, "differs at", d[:10], np.asarray(x[k])[d[:6]], np.asarray(y[k])[d[:6]])
want, _ = orc.score_points(reg)
st, og, mu, sc = gpu_sweep.native(ctx, reg, "point_width").score_points()
w = np.array([x[3] for x in want])
print("score_points equal", np.array_equal(sc, w), "n>=0", int((w >= 0).sum()), "ties among accepted", len(w[w >= 0]) - len(set(w[w >= 0].tolist())))
