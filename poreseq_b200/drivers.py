"""The iteration policies the reference keeps in Python (poreseq/Mutate.py, poreseq/Variant.py),
restated over an already loaded PSAlign so they can run on synthetic input (fast5/BAM loading,
poreseq/LoadData.py, is outside the hot path).  Nothing here computes: every call goes through
the PSAlign mirror into the CUDA library."""
import sys

import numpy as np

from . import poreseqcpp
from .Util import MutationInfo  # noqa: F401


def make_psalign(region):
    """PSAlign from any object with .sequence/.events/.params (what LoadAlignedEvents returns,
    poreseq/LoadData.py:47-50)."""
    pa = poreseqcpp.PSAlign()
    pa.sequence = region.sequence
    pa.events = [ev.copy() for ev in region.events]
    pa.params = dict(region.params)
    return pa


def consensus(pa, refseq=None, reps=4, verbose=0, log=sys.stderr, stages=None):
    """Consensus error correction of one region: Mutate('self') then up to `reps` rounds of
    (Mutate('viterbi'), Refine) until Refine changes nothing, then end trimming
    (poreseq/Mutate.py:47-99).  Returns (sequence, accuracy vs refseq or None).
    `stages`: a list that receives (stage name, sequence, bases changed, [ref_align per event]) after every stage
    (the parity tests compare them with the reference's)."""
    params = pa.params
    if len(pa.events) < 5:                       # Mutate.py:50-53
        return pa.sequence, 100

    def note(name, nb):
        if stages is not None:
            stages.append((name, pa.sequence, int(nb), [np.array(ev.ref_align, dtype="f8") for ev in pa.events]))

    note("mutate_self", pa.Mutate(reps=reps))
    if verbose and refseq:
        log.write("Accuracy: %.1f%%\n" % poreseqcpp.swalign(pa.sequence, refseq)[0])
    for k in range(reps):
        note("mutate_viterbi_%d" % k, pa.Mutate(seqs='viterbi'))
        nbases = pa.Refine()
        note("refine_%d" % k, nbases)
        if verbose and refseq:
            log.write("Accuracy: %.1f%%\n" % poreseqcpp.swalign(pa.sequence, refseq)[0])
        if nbases == 0:
            break
    if 'end_trim' in params and len(pa.sequence) > 2 * params['end_trim']:
        t = int(params['end_trim'])
        pa.sequence = pa.sequence[t:-t]
    acc = poreseqcpp.swalign(pa.sequence, refseq)[0] if refseq else None
    return pa.sequence, acc


def consensus_native(regions, ctx=None, reps=4, in_flight=16, refseqs=None):
    """The same policy run below the C-ABI (ps_consensus_batch): one native region per input region for the whole loop,
    `in_flight` regions side by side on the GPU.  `regions`: objects with .sequence/.events/.params (PSAlign, synthetic
    regions) or poreseqcpp.PackedRegion views of an event-pack file.  Returns [(sequence after end_trim, accuracy or None,
    stages)] per region."""
    ctx = ctx or poreseqcpp.default_context()
    packs = [r if isinstance(r, poreseqcpp.PackedRegion) else poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regions]
    nrs = poreseqcpp.native_regions_from_packed(ctx, packs)
    try:
        pw = int(packs[0].params.get("point_width", packs[0].params.get("scoring_width", 20))) if packs else 20
        poreseqcpp.consensus_batch(ctx, nrs, reps=reps, point_width=pw, in_flight=in_flight)
        out = []
        for k, nr in enumerate(nrs):
            seq = nr.sequence()
            params = packs[k].params
            if 'end_trim' in params and len(seq) > 2 * params['end_trim']:
                t = int(params['end_trim'])
                seq = seq[t:-t]
            acc = poreseqcpp.swalign(seq, refseqs[k])[0] if refseqs else None
            out.append((seq, acc, nr.stages()))
        return out
    finally:
        poreseqcpp.close_regions(nrs)


def variant(pa, muts=None, var_seqs=None, region_start=0):
    """Variant scoring of one region (poreseq/Variant.py:44-95).

    var_seqs: dict name -> sequence  => {name: sum(ScoreEvents on the realigned copy) - base}
    muts: list of MutationInfo (region-relative after subtracting region_start); empty list
    means every point mutation (ScorePoints, at point_width)."""
    if var_seqs is not None:
        base = float(np.sum(pa.ScoreEvents()))
        out = {}
        for name, seq in var_seqs.items():
            pav = pa.Copy()
            pav.RealignTo(seq)
            out[name] = float(np.sum(pav.ScoreEvents())) - base
        return out
    for m in muts:
        m.start -= region_start
    scores = pa.ScoreMutations(muts) if len(muts) > 0 else pa.ScorePoints()
    for ms in scores:
        ms.start += region_start
    return scores


# ---------------------------------------------------------------------------------------------------------
# `poreseq train` (poreseq/cmdline.py:246-267, poreseq/Params.py:31-60): every iteration tries 16
# random variations of the transition parameters, runs the consensus loop with each, keeps the most accurate.  The
# reference gives every variant to a pool of worker processes; here the variants of an iteration are regions in flight
# on one GPU through ps_consensus_batch, like the consensus throughput mode.

def vary_params(params, rng, count=16):
    """VaryParams (Params.py:31-60): `count` copies of params, each with 3 randomly chosen `_t` / `_c` parameters multiplied
    by gauss(1, 0.15).  `rng` is a random.Random (the reference uses the module-level generator)."""
    names = [k for k in params.keys() if k[-2:] == '_t' or k[-2:] == '_c']
    out = []
    for _ in range(count):
        p = dict(params)
        for k in rng.sample(names, 3):
            p[k] *= rng.gauss(1.0, 0.15)
        out.append(p)
    return out


def set_params(events, params):
    """PSEvent.setparams (poreseq/EventData.py:286-312): `skip_t` sets model.prob_skip of the template events, `skip_c` of
    the complement ones, likewise stay / extend / insert."""
    for ev in events:
        for k, v in params.items():
            name = 'prob_' + k[:-2]
            if not hasattr(ev.model, name):
                continue
            if (k[-2:] == '_t' and not ev.model.complement) or (k[-2:] == '_c' and ev.model.complement):
                setattr(ev.model, name, v)


def variant_region(region, params):
    """A copy of `region` with one parameter set applied: the transition probabilities of its events' models
    (set_params) and the region-level lik_offset -- what Mutate(params=...) loads for one training variant."""
    import copy
    reg = copy.deepcopy(region)
    set_params(reg.events, params)
    reg.params = dict(reg.params, **{q: v for q, v in params.items() if q in ("lik_offset",)})
    return reg


def train(region, params, iters=1, variants=16, in_flight=8, reps=4, seed=None, device=None, log=None, details=None):
    """The training loop on one loaded region with known truth (`region.truth`): returns (best params, [best accuracy per
    iteration]).  Every variant starts from the region's draft sequence and seed alignments and runs the whole consensus
    loop below the C-ABI (ps_consensus_batch: the variants of an iteration are regions in flight on one GPU, each with
    the rand() stream a freshly started process draws -- what the reference's worker pool gives a variant).
    `details`: a list that receives (parameter sets, accuracies) of every iteration."""
    import random
    rng = random.Random(seed)
    params = dict(params)
    ctx = poreseqcpp.Context(device if device is not None else poreseqcpp.default_context().device)
    history = []
    try:
        for it in range(iters):
            plist = vary_params(params, rng, variants)
            regs = [variant_region(region, p) for p in plist]
            out = consensus_native(regs, ctx=ctx, reps=reps, in_flight=in_flight, refseqs=[region.truth] * len(regs))
            accs = [o[1] for o in out]
            best = int(np.argmax(accs))                    # cmdline.py:263: first maximum
            params = plist[best]
            history.append(accs[best])
            if details is not None:
                details.append((plist, accs))
            if log is not None:
                log.write('Best at iter {}: {}\n'.format(it + 1, accs[best]))
    finally:
        ctx.close()
    return params, history
