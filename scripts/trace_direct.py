import sys, time
sys.path.insert(0, ".")
from poreseq_b200 import poreseqcpp, synth
ctx = poreseqcpp.Context(0)
ctx.set_precision("fast")
regs = [synth.make_region(1000, 10, seed=s + 1) for s in range(44)]
packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
for it in range(4):
    t0 = time.perf_counter()
    p = poreseqcpp.PendingDirect(ctx, packs)
    t1 = time.perf_counter()
    out = p.end()
    t2 = time.perf_counter()
    print("iter %d begin %.2f ms end %.2f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr)
