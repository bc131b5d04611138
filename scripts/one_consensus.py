"""One warm consensus loop of a 1 kb x 10x region (for ncu launch lists: the second region is the measured one)."""
import sys, time
sys.path.insert(0, ".")
from poreseq_b200 import drivers, poreseqcpp, synth
poreseqcpp.default_context().set_precision("fast")
for seed in (7, 8):
    reg = synth.make_region(1000, 10, seed=seed, draft_error=0.10)
    pa = drivers.make_psalign(reg)
    n0 = poreseqcpp.default_context().launch_count()
    t0 = time.time(); drivers.consensus(pa, refseq=reg.truth, reps=4)
    print("region seed %d: %.3f s, %d launches" % (seed, time.time() - t0, poreseqcpp.default_context().launch_count() - n0))
