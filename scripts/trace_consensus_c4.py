import sys, time
sys.path.insert(0, ".")
from poreseq_b200 import drivers, poreseqcpp, synth
ctx = poreseqcpp.Context(0)
ctx.set_precision("fast")
reg = synth.make_region(10000, 50, seed=11, draft_error=0.10)
for it in range(2):
    t0 = time.perf_counter()
    out = drivers.consensus_native([reg], ctx=ctx, in_flight=1)
    print("PASS %d: %.2f s" % (it, time.perf_counter() - t0), file=sys.stderr)
