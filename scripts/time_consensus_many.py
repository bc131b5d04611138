"""Consensus loops of several regions at once: one host thread + one context (stream) per region in flight.
usage: time_consensus_many.py L coverage n_regions n_threads"""
import sys, time, threading, queue
sys.path.insert(0, ".")
from poreseq_b200 import drivers, poreseqcpp, synth

L, cov, nreg, nth = (int(x) for x in sys.argv[1:5])
regs = [synth.make_region(L, cov, seed=100 + k, draft_error=0.10) for k in range(nreg)]
q = queue.Queue()
for r in regs:
    q.put(r)
out = []

ctxs = []
def worker(k):
    ctx = ctxs[k]
    while True:
        try:
            reg = q.get_nowait()
        except queue.Empty:
            break
        pa = drivers.make_psalign(reg)
        pa.ctx = ctx
        seq, acc = drivers.consensus(pa, refseq=reg.truth, reps=4)
        out.append(acc)

# warm-up (CUDA init)
w = poreseqcpp.Context(0); pa = drivers.make_psalign(synth.make_region(300, 5, seed=1, draft_error=0.05)); pa.ctx = w
drivers.consensus(pa, reps=1)
# warm contexts: every thread's context runs one region before the timed pass (buffer growth, pinned allocations)
for k in range(nth):
    c = poreseqcpp.Context(0); c.set_precision("fast"); ctxs.append(c)
wq = q; q = queue.Queue()
for k in range(nth):
    q.put(synth.make_region(L, cov, seed=900 + k, draft_error=0.10))
ths = [threading.Thread(target=worker, args=(k,)) for k in range(nth)]
for t in ths: t.start()
for t in ths: t.join()
out.clear(); q = wq
c0 = time.process_time()
t0 = time.time()
ths = [threading.Thread(target=worker, args=(k,)) for k in range(nth)]
for t in ths: t.start()
for t in ths: t.join()
dt = time.time() - t0
cpu = time.process_time() - c0
print("L=%d cov=%d regions=%d threads=%d: %.2f s  %.3f kb/s  mean accuracy %.2f%%  cpu %.2f s (%.1f cores busy, %.0f ms cpu per region)" % (L, cov, nreg, nth, dt, nreg * L / 1000.0 / dt, sum(out) / len(out), cpu, cpu / dt, cpu / nreg * 1e3))
