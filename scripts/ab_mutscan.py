"""A/B check of the two FP32 mutation scans: run with and without PORESEQ_B200_MUT_OLD=1, compare dumps."""
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
print(out, len(np.concatenate(res)), ctx.last_timing()["mutscore"])
