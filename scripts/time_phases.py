"""Quick phase timing of ScorePoints on synthetic regions (run under gpurun).
usage: time_phases.py L coverage regions [fast|exact] [packed]"""
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
packed = len(sys.argv) > 5 and sys.argv[5] == "packed"
regs = [synth.make_region(L, cov, seed=s + 1) for s in range(nreg)]
packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
for it in range(4):
    t0 = time.time()
    if packed:
        nrs = [poreseqcpp.NativeRegion.from_packed(ctx, p, "point_width") for p in packs]
    else:
        nrs = [poreseqcpp.NativeRegion(ctx, r.sequence, r.events, r.params, "point_width") for r in regs]
    t1 = time.time()
    out = poreseqcpp.score_points_batch(ctx, nrs)
    t2 = time.time()
    tm = ctx.last_timing()
    w, n = ctx.last_cells()
    print("iter", it, "marshal %.1f ms  call %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3),
          " ".join("%s=%.3f" % (k, v) for k, v in tm.items()), "wide=%.3g narrow=%.3g GCUPS(dev)=%.2f" % (w, n, (w + n) / tm["total"] / 1e6))
    for x in nrs: x.close()
print("positive", sum(int((o[3] >= 0).sum()) for o in out))
