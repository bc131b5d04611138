"""Host-only: time the marshalling of a bench step (22 regions) per variant / worker-thread count."""
import sys, time
sys.path.insert(0, ".")
from poreseq_b200 import poreseqcpp, synth
ctx = poreseqcpp.Context(0)
regs = [synth.make_region(1000, 10, seed=s + 1) for s in range(22)]
packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
def best(fn):
    ts = []
    for it in range(8):
        t0 = time.perf_counter(); nrs = fn(); t1 = time.perf_counter()
        for x in nrs: x.close()
        ts.append((t1 - t0) * 1e3)
    return min(ts)
print("per-region from_packed loop: %.2f ms" % best(lambda: [poreseqcpp.NativeRegion.from_packed(ctx, p, "point_width") for p in packs]))
print("ps_regions_create (bulk):    %.2f ms" % best(lambda: poreseqcpp.native_regions_from_packed(ctx, packs, "point_width")))
