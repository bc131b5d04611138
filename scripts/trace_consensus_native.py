import sys, time, os
sys.path.insert(0, ".")
from poreseq_b200 import drivers, poreseqcpp, synth
ctx = poreseqcpp.Context(0)
ctx.set_precision("fast")
regs = [synth.make_region(1000, 10, seed=500 + k, draft_error=0.10) for k in range(2)]
drivers.consensus_native(regs[:1], ctx=ctx, in_flight=1)
t0 = time.perf_counter()
n0 = ctx.launch_count()
drivers.consensus_native(regs[1:], ctx=ctx, in_flight=1)
print("one region: %.1f ms, %d launches" % ((time.perf_counter() - t0) * 1e3, ctx.launch_count() - n0), file=sys.stderr)
