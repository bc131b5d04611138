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


def consensus(pa, refseq=None, reps=4, verbose=0, log=sys.stderr):
    """Consensus error correction of one region: Mutate('self') then up to `reps` rounds of
    (Mutate('viterbi'), Refine) until Refine changes nothing, then end trimming
    (poreseq/Mutate.py:47-99).  Returns (sequence, accuracy vs refseq or None)."""
    params = pa.params
    if len(pa.events) < 5:                       # Mutate.py:50-53
        return pa.sequence, 100
    pa.Mutate(reps=reps)
    if verbose and refseq:
        log.write("Accuracy: %.1f%%\n" % poreseqcpp.swalign(pa.sequence, refseq)[0])
    for _ in range(reps):
        pa.Mutate(seqs='viterbi')
        nbases = pa.Refine()
        if verbose and refseq:
            log.write("Accuracy: %.1f%%\n" % poreseqcpp.swalign(pa.sequence, refseq)[0])
        if nbases == 0:
            break
    if 'end_trim' in params and len(pa.sequence) > 2 * params['end_trim']:
        t = int(params['end_trim'])
        pa.sequence = pa.sequence[t:-t]
    acc = poreseqcpp.swalign(pa.sequence, refseq)[0] if refseq else None
    return pa.sequence, acc


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
