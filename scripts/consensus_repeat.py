import sys, time
sys.path.insert(0, ".")
import bench
from poreseq_b200 import poreseqcpp
ctxs = [poreseqcpp.Context(0) for _ in range(8)]
for c in ctxs: c.set_precision("fast")
bench.consensus_throughput(ctxs, 8, 1000, 10, seed0=9000)
for n in (32, 32, 32, 64, 64):
    dt, acc = bench.consensus_throughput(ctxs, n, 1000, 10, seed0=500)
    print(n, "regions: %.2f s  %.2f kb/s  acc %.2f" % (dt, n / dt, acc), flush=True)
