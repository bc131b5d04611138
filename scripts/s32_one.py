import sys, os
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from poreseq_b200 import poreseqcpp, synth
mode = sys.argv[1]
reg = synth.make_region(10000, 2, seed=32, draft_error=0.02, partial=0.5)
oreg = synth.make_region(300, 3, seed=35, draft_error=0.03, params=reg.params)
c = poreseqcpp.Context(0); c.set_precision("fast")
nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params)
ot = poreseqcpp.NativeRegion(c, oreg.sequence, oreg.events, oreg.params)
want = np.array([19934.36523438, 19432.02929688, 9009.1328125, 9063.38964844])
if "s" in mode:
    got = nr.score_events()
lst = {"a": [nr, ot, nr], "b": [nr, nr], "c": [nr, nr, nr, nr], "d": [ot, nr, nr]}[mode[0]]
outs = poreseqcpp.score_events_batch(c, lst)
print(mode, [bool(np.allclose(o, want, rtol=1e-6)) for o in outs if len(o) == 4])
