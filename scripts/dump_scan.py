"""Dump the FAST-mode ScorePoints scores of three fixed synthetic regions (regression check of the FP32 scan:
compare the dump of a new build with the dump of the previous one, they must be bit-identical)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from poreseq_b200 import poreseqcpp, synth

out = sys.argv[1]
ctx = poreseqcpp.Context(0)
ctx.set_precision("fast")
res = []
for seed, kw in ((1, {}), (2, dict(draft_error=0.03)), (3, dict(partial=0.3))):
    reg = synth.make_region(600, 6, seed=seed, **kw)
    nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params, "point_width")
    st, og, mu, sc = nr.score_points()
    res.append(sc.copy())
    nr.close()
np.save(out, np.concatenate(res))
if len(sys.argv) > 2:
    ref = np.load(sys.argv[2])
    print("identical to %s:" % sys.argv[2], np.array_equal(ref, np.concatenate(res)))
print(out, len(np.concatenate(res)), ctx.last_timing()["mutscore"])
