"""Consensus throughput below the C-ABI: usage: time_consensus_native.py regions in_flight [L coverage]"""
import sys, time
sys.path.insert(0, ".")
from poreseq_b200 import drivers, poreseqcpp, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
fl = int(sys.argv[2]) if len(sys.argv) > 2 else 16
L = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
cov = int(sys.argv[4]) if len(sys.argv) > 4 else 10
ctx = poreseqcpp.Context(0)
ctx.set_precision("fast")
regs = [synth.make_region(L, cov, seed=500 + k, draft_error=0.10) for k in range(n)]
drivers.consensus_native(regs[:min(n, fl)], ctx=ctx, in_flight=fl)          # contexts allocate their buffers
for it in range(3):
    t0 = time.perf_counter()
    out = drivers.consensus_native(regs, ctx=ctx, in_flight=fl, refseqs=None)
    dt = time.perf_counter() - t0
    print("iter", it, "%d regions of %d b x %dx, %d in flight: %.3f s = %.1f kb/s" % (n, L, cov, fl, dt, n * L / 1000.0 / dt))
accs = [poreseqcpp.swalign(o[0], r.truth)[0] for o, r in zip(out, regs)]
print("mean accuracy %.2f%%" % (sum(accs) / len(accs)))
