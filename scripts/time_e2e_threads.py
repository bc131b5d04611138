"""GPU box: one host thread per context, each driving whole steps; per-phase host time per step (summed over threads /
steps) shows which phase stops scaling when several threads run it at once.  usage: time_e2e_threads.py [regions] [threads]"""
import sys, time, threading
sys.path.insert(0, ".")
from poreseq_b200 import poreseqcpp, synth

nreg = int(sys.argv[1]) if len(sys.argv) > 1 else 44
nthr = int(sys.argv[2]) if len(sys.argv) > 2 else 4
regs = [synth.make_region(1000, 10, seed=s + 1) for s in range(nreg)]
packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
ctxs = [poreseqcpp.Context(0) for _ in range(nthr)]
for c in ctxs:
    c.set_precision("fast")
T = {"create": 0.0, "begin": 0.0, "end": 0.0, "close": 0.0}
lock = threading.Lock()

def drive(c, n):
    loc = dict.fromkeys(T, 0.0)
    for _ in range(n):
        t0 = time.perf_counter()
        nrs = poreseqcpp.native_regions_from_packed(c, packs, "point_width")
        t1 = time.perf_counter()
        p = poreseqcpp.score_points_batch_begin(c, nrs)
        t2 = time.perf_counter()
        p.end()
        t3 = time.perf_counter()
        poreseqcpp.close_regions(p.regions)
        t4 = time.perf_counter()
        loc["create"] += t1 - t0; loc["begin"] += t2 - t1; loc["end"] += t3 - t2; loc["close"] += t4 - t3
    with lock:
        for k in T:
            T[k] += loc[k]

def run(n):
    ths = [threading.Thread(target=drive, args=(c, n)) for c in ctxs]
    for t in ths: t.start()
    for t in ths: t.join()

run(3)
for k in T: T[k] = 0.0
per = 10
c0 = time.process_time(); t0 = time.perf_counter(); run(per); wall = time.perf_counter() - t0; cpu = time.process_time() - c0
steps = per * nthr
print("%d driver threads: %.2f ms/step wall, %.1f ms cpu/step; per step inside a thread: " % (nthr, wall / steps * 1e3, cpu / steps * 1e3) +
      ", ".join("%s %.2f" % (k, v / steps * 1e3) for k, v in T.items()), flush=True)
