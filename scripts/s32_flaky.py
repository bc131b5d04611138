import sys, os
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from poreseq_b200 import poreseqcpp, synth
reg = synth.make_region(10000, 2, seed=32, draft_error=0.02, partial=0.5)
oreg = synth.make_region(300, 3, seed=35, draft_error=0.03, params=reg.params)
cx = poreseqcpp.Context(0)
want = poreseqcpp.NativeRegion(cx, reg.sequence, reg.events, reg.params).score_events()
wanto = poreseqcpp.NativeRegion(cx, oreg.sequence, oreg.events, oreg.params).score_events()
def err(o, w): return "%.1e" % float(np.max(np.abs(o - w) / w))
for rep in range(3):
    c = poreseqcpp.Context(0); c.set_precision("fast")
    nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params)
    got = nr.score_events()
    ot = poreseqcpp.NativeRegion(c, oreg.sequence, oreg.events, oreg.params)
    outs = poreseqcpp.score_events_batch(c, [nr, ot, nr])
    print(rep, "single", err(got, want), "batch", [err(o, w) for o, w in zip(outs, [want, wanto, want])], outs[0], outs[2])
    outs = poreseqcpp.score_events_batch(c, [nr, ot, nr])
    print(rep, "again batch", [err(o, w) for o, w in zip(outs, [want, wanto, want])])
    c.close()
