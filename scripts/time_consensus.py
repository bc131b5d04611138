"""Consensus loop (poreseq/Mutate.py policy) on one synthetic region: wall time per phase, kb/s, accuracy.
usage: time_consensus.py L coverage [fast|exact] [draft_error]"""
import sys, time
sys.path.insert(0, ".")
from poreseq_b200 import drivers, poreseqcpp, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cov = int(sys.argv[2]) if len(sys.argv) > 2 else 10
mode = sys.argv[3] if len(sys.argv) > 3 else "exact"
err = float(sys.argv[4]) if len(sys.argv) > 4 else 0.10
reg = synth.make_region(L, cov, seed=7, draft_error=err)
poreseqcpp.default_context().set_precision(mode)
pa = drivers.make_psalign(reg)
acc0 = poreseqcpp.swalign(reg.sequence, reg.truth)[0]
T = {}
def timed(name, fn, *a, **k):
    t = time.time(); r = fn(*a, **k); T[name] = T.get(name, 0) + time.time() - t; return r
t0 = time.time()
timed("mutate_self", pa.Mutate, reps=4)
for _ in range(4):
    timed("mutate_viterbi", pa.Mutate, seqs='viterbi')
    nb = timed("refine", pa.Refine)
    if nb == 0:
        break
dt = time.time() - t0
acc = poreseqcpp.swalign(pa.sequence, reg.truth)[0]
print("L=%d cov=%d mode=%s: %.2f s  %.3f kb/s  accuracy %.2f%% -> %.2f%%  launches %d  phases %s" %
      (L, cov, mode, dt, L / 1000.0 / dt, acc0, acc, poreseqcpp.default_context().launch_count(),
       {k: round(v, 2) for k, v in T.items()}))
