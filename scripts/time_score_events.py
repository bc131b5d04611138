"""ScoreEvents timing (BASELINE.json configs[0]): usage: time_score_events.py L coverage regions [fast|exact]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from poreseq_b200 import poreseqcpp, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cov = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nreg = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = poreseqcpp.Context(0)
if len(sys.argv) > 4:
    ctx.set_precision(sys.argv[4])
regs = [synth.make_region(L, cov, seed=s + 1) for s in range(nreg)]
nrs = [poreseqcpp.NativeRegion(ctx, r.sequence, r.events, r.params) for r in regs]
for it in range(5):
    t0 = time.time()
    out = poreseqcpp.score_events_batch(ctx, nrs)
    t1 = time.time()
    tm = ctx.last_timing()
    w, n = ctx.last_cells()
    print("iter", it, "call %.2f ms" % ((t1 - t0) * 1e3), "forward=%.3f total=%.3f" % (tm["forward"], tm["total"]),
          "wide=%.4g GCUPS(kernel)=%.1f GCUPS(call)=%.1f" % (w, w / tm["forward"] / 1e6, w / (t1 - t0) / 1e9))
print("sum", float(np.sum([o.sum() for o in out])))
