"""Small pass through the round-2 kernels for compute-sanitizer (memcheck): k_score_f32 (staged, persistent, inv), the
cluster Viterbi chain, the step-major Smith-Waterman, the lockstep consensus, the direct ScorePoints path."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from poreseq_b200 import drivers, poreseqcpp, synth
ctx = poreseqcpp.Context(0)
ctx.set_precision("fast")
regs = [synth.make_region(300, 3, seed=40 + k, draft_error=0.05) for k in range(6)]
s = list(regs[1].sequence); s[10] = "N"; s[150] = "N"; regs[1].sequence = "".join(s)
nrs = [poreseqcpp.NativeRegion(ctx, r.sequence, r.events, r.params) for r in regs]
print("score_events", [float(x.sum()) for x in poreseqcpp.score_events_batch(ctx, nrs)][:2])
big = synth.make_region(2500, 2, seed=50, draft_error=0.02)
print("long", poreseqcpp.NativeRegion(ctx, big.sequence, big.events, big.params).score_events())
packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
print("direct", float(poreseqcpp.score_points_direct(ctx, packs)[0][3].sum()))
out = drivers.consensus_native(regs, ctx=ctx, in_flight=4)
print("consensus", [len(o[0]) for o in out])
print("swalign_device", poreseqcpp.swalign_device(ctx, regs[0].sequence, regs[0].truth)[0])
