import sys, time, cProfile, pstats
sys.path.insert(0, ".")
from poreseq_b200 import drivers, poreseqcpp, synth
poreseqcpp.default_context().set_precision("fast")
w = drivers.make_psalign(synth.make_region(300, 5, seed=1, draft_error=0.05)); drivers.consensus(w, reps=1)
reg = synth.make_region(1000, 10, seed=7, draft_error=0.10)
pa = drivers.make_psalign(reg)
t0=time.time(); drivers.consensus(pa, refseq=reg.truth, reps=4); print("warm run", time.time()-t0)
reg = synth.make_region(1000, 10, seed=8, draft_error=0.10)
pa = drivers.make_psalign(reg)
pr = cProfile.Profile(); pr.enable()
t0=time.time(); drivers.consensus(pa, refseq=reg.truth, reps=4); dt=time.time()-t0
pr.disable()
print("profiled run", dt)
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
