"""GPU box: where the wall time of a pipelined bench step goes on the host (create / begin / end), per context count."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from poreseq_b200 import poreseqcpp, synth

nreg = int(sys.argv[1]) if len(sys.argv) > 1 else 44
regs = [synth.make_region(1000, 10, seed=s + 1) for s in range(nreg)]
packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
for nctx in (4,):
    ctxs = [poreseqcpp.Context(0) for _ in range(nctx)]
    for c in ctxs:
        c.set_precision("fast")
    T = {"create": 0.0, "begin": 0.0, "end": 0.0, "close": 0.0}
    def begin(c):
        t0 = time.perf_counter()
        nrs = poreseqcpp.native_regions_from_packed(c, packs, "point_width")
        t1 = time.perf_counter()
        p = poreseqcpp.score_points_batch_begin(c, nrs)
        t2 = time.perf_counter()
        T["create"] += t1 - t0; T["begin"] += t2 - t1
        return p
    def end(p):
        t0 = time.perf_counter()
        out = p.end()
        t1 = time.perf_counter()
        poreseqcpp.close_regions(p.regions)
        t2 = time.perf_counter()
        T["end"] += t1 - t0; T["close"] += t2 - t1
        return out
    def run(count):
        infl = []
        for k in range(count):
            infl.append(begin(ctxs[k % nctx]))
            if len(infl) == nctx:
                end(infl.pop(0))
        while infl:
            end(infl.pop(0))
    run(6)
    for k in T: T[k] = 0.0
    steps = 30
    c0 = time.process_time(); t0 = time.perf_counter(); run(steps); wall = time.perf_counter() - t0; cpu = time.process_time() - c0
    print("%d contexts: %.2f ms/step wall, %.1f ms cpu/step; host per step: " % (nctx, wall / steps * 1e3, cpu / steps * 1e3) +
          ", ".join("%s %.2f" % (k, v / steps * 1e3) for k, v in T.items()), flush=True)
    print("   last timing", {k: round(v, 3) for k, v in ctxs[0].last_timing().items()}, flush=True)
    del ctxs
