"""Seeded synthetic squiggle data for the PoreSeq scoring path.

fast5 / BAM inputs are not available offline, so every test and benchmark feeds the hot path with
events simulated from a random reference and a synthetic 5-mer pore model with injected skips and
stays.  The objects produced here are duck-typed stand-ins for the reference's ``PSEvent`` /
``PSModel`` (poreseq/EventData.py:46-100): float64 1-D arrays ``mean, stdv, ref_align, ref_like``,
a ``sequence`` string, ``makecontiguous()``, ``mapaligns(pairs)`` and a ``model`` with float64[1024]
``level_mean, level_stdv, sd_mean, sd_stdv`` plus ``prob_skip/stay/extend/insert`` and ``complement``.
That is exactly what the boundary marshals (poreseq/_poreseqcpp.pyx:99-129).

Pure numpy: no GPU and no native code involved.
"""
import copy

import numpy as np

N_STATES = 1024
BASES = "ACGT"

# defaults.conf:1-19 of the reference (the .conf keys are part of the drop-in contract)
DEFAULT_PARAMS = {
    "realign_width": 300, "scoring_width": 100, "point_width": 20,
    "min_coverage": 0, "max_coverage": 30, "min_overlap": 500, "max_length": 10000, "end_trim": 150,
    "lik_offset": 4.5,
    "skip_t": 0.141, "skip_c": 0.088, "stay_t": 0.043, "stay_c": 0.057,
    "extend_t": 0.072, "extend_c": 0.046, "insert_t": 0.020, "insert_c": 0.025,
    "verbose": 0,
}


class SynthModel(object):
    """Stand-in for PSModel (poreseq/EventData.py:46-77)."""

    def __init__(self):
        self.level_mean = np.zeros(N_STATES)
        self.level_stdv = np.ones(N_STATES)
        self.sd_mean = np.ones(N_STATES)
        self.sd_stdv = np.ones(N_STATES)
        self.prob_skip = 0.1
        self.prob_stay = 0.1
        self.prob_extend = 0.1
        self.prob_insert = 0.01
        self.name = ""
        self.complement = False


def _contig(obj):
    for k, v in list(vars(obj).items()):
        if isinstance(v, np.ndarray):
            setattr(obj, k, np.ascontiguousarray(v, dtype="f8"))


class SynthEvent(object):
    """Stand-in for PSEvent (poreseq/EventData.py:79-100)."""

    def __init__(self):
        self.mean = np.zeros(0)
        self.stdv = np.zeros(0)
        self.ref_align = np.zeros(0)
        self.ref_like = np.zeros(0)
        self.model = SynthModel()
        self.sequence = ""
        self.flipped = False

    def copy(self):
        return copy.deepcopy(self)

    def makecontiguous(self):
        _contig(self)
        _contig(self.model)

    def mapaligns(self, pairs):
        """Re-map ref_align through aligned index pairs; same contract as PSEvent.mapaligns
        (poreseq/EventData.py:207-236): unique in x, linear interpolation, 0 outside, rounded."""
        pairs = np.asarray(pairs)
        refal = self.ref_align
        aligned = refal > 0
        out = np.zeros_like(refal)
        _, first = np.unique(pairs[:, 0], return_index=True)
        pairs = pairs[first, :]
        out[aligned] = np.round(np.interp(refal[aligned], pairs[:, 0], pairs[:, 1], 0, 0))
        self.ref_align = out
        self.makecontiguous()

    def setparams(self, params):
        """params dict -> model.prob_* per strand (poreseq/EventData.py:288-312)."""
        for k, v in params.items():
            name = "prob_" + k[:-2]
            if not hasattr(self.model, name):
                continue
            if (k.endswith("_t") and not self.model.complement) or (k.endswith("_c") and self.model.complement):
                setattr(self.model, name, v)


def random_sequence(length, rng):
    return "".join(BASES[i] for i in rng.integers(0, 4, size=length))


def seq_to_states(seq):
    """5-mer state index per position (cpp/Sequence.h:69-100), valid ACGT input only."""
    code = np.array([BASES.index(c) for c in seq], dtype=np.int64)
    n = len(seq) - 4
    if n <= 0:
        return np.zeros(0, dtype=np.int64)
    st = np.zeros(n, dtype=np.int64)
    for k in range(5):
        st = st * 4 + code[k:k + n]
    return st


def make_models(rng, params=None):
    """One template and one complement synthetic pore model."""
    params = DEFAULT_PARAMS if params is None else params
    models = []
    for comp in (False, True):
        m = SynthModel()
        m.level_mean = rng.uniform(40.0, 80.0, N_STATES)
        m.level_stdv = rng.uniform(0.8, 1.6, N_STATES)
        m.sd_mean = rng.uniform(0.7, 1.3, N_STATES)
        m.sd_stdv = rng.uniform(0.2, 0.4, N_STATES)
        m.complement = comp
        m.name = "synthetic_" + ("complement" if comp else "template")
        suffix = "_c" if comp else "_t"
        m.prob_skip = params["skip" + suffix]
        m.prob_stay = params["stay" + suffix]
        m.prob_extend = params["extend" + suffix]
        m.prob_insert = params["insert" + suffix]
        models.append(m)
    return models


def corrupt_sequence(seq, rate, rng):
    """Inject substitutions / insertions / deletions at the given total rate.

    Returns (new_seq, pos_map) where pos_map[i] is the index in new_seq that truth base i maps to
    (nondecreasing; a deleted base maps to the next surviving one, clipped to the end)."""
    out = []
    pos_map = np.zeros(len(seq), dtype=np.int64)
    for i, c in enumerate(seq):
        pos_map[i] = len(out)
        u = rng.random()
        if u < rate / 3.0:
            out.append(BASES[(BASES.index(c) + 1 + rng.integers(0, 3)) % 4])
        elif u < 2.0 * rate / 3.0:
            out.append(c)
            out.append(BASES[rng.integers(0, 4)])
        elif u < rate:
            pass
        else:
            out.append(c)
    new = "".join(out)
    pos_map = np.minimum(pos_map, max(len(new) - 1, 0))
    return new, pos_map


def simulate_event(truth_states, model, rng, first=0, last=None, p_skip=0.10, p_stay=0.05,
                   p_unaligned=0.0, jitter=0):
    """Walk truth_states[first:last]; skip a state w.p. p_skip, emit 1+Geom levels (stays).

    ref_align (1-based state index into the truth sequence) is the seed alignment; a fraction
    p_unaligned of levels is zeroed and +-jitter noise added to exercise updaterefs()."""
    last = len(truth_states) if last is None else last
    means, stdvs, aligns = [], [], []
    for k in range(first, last):
        if rng.random() < p_skip:
            continue
        s = truth_states[k]
        n = 1
        while rng.random() < p_stay:
            n += 1
        for _ in range(n):
            means.append(rng.normal(model.level_mean[s], model.level_stdv[s]))
            mu = model.sd_mean[s]
            lam = mu ** 3 / model.sd_stdv[s] ** 2
            stdvs.append(max(rng.wald(mu, lam), 1e-3))
            aligns.append(k + 1)
    ev = SynthEvent()
    ev.mean = np.array(means, dtype="f8")
    ev.stdv = np.array(stdvs, dtype="f8")
    ra = np.array(aligns, dtype="f8")
    if jitter > 0 and len(ra):
        ra = np.maximum(1.0, ra + rng.integers(-jitter, jitter + 1, size=len(ra)))
        ra = np.maximum.accumulate(ra)
    if p_unaligned > 0 and len(ra) > 4:
        drop = rng.random(len(ra)) < p_unaligned
        drop[0] = drop[-1] = False
        ra[drop] = 0.0
    ev.ref_align = ra
    ev.ref_like = np.zeros(len(ra))
    ev.model = copy.deepcopy(model)
    return ev


class SynthRegion(object):
    """A region's worth of synthetic input: truth, draft sequence, events, params."""

    def __init__(self):
        self.truth = ""
        self.sequence = ""
        self.events = []
        self.params = dict(DEFAULT_PARAMS)


def make_region(length=1000, coverage=10, seed=1, draft_error=0.0, read_error=0.12, partial=0.0,
                p_unaligned=0.0, jitter=0, params=None):
    """Build a synthetic region: `coverage` reads -> 2*coverage events (template + complement),
    mirroring LoadData.py:140-148.  `draft_error` > 0 makes the region sequence an erroneous copy
    of the truth (events' seed alignments are mapped onto it); `partial` is the fraction of reads
    that cover only part of the region."""
    rng = np.random.default_rng(seed)
    reg = SynthRegion()
    if params is not None:
        reg.params.update(params)
    reg.truth = random_sequence(length, rng)
    models = make_models(rng, reg.params)
    truth_states = seq_to_states(reg.truth)
    n_states = len(truth_states)
    if draft_error > 0:
        reg.sequence, pos_map = corrupt_sequence(reg.truth, draft_error, rng)
    else:
        reg.sequence, pos_map = reg.truth, np.arange(length)
    n_draft_states = max(len(reg.sequence) - 4, 1)
    for _ in range(coverage):
        first, last = 0, n_states
        if partial > 0 and rng.random() < partial:
            span = int(rng.integers(n_states // 3, max(n_states // 3 + 1, (2 * n_states) // 3)))
            first = int(rng.integers(0, n_states - span + 1))
            last = first + span
        read_seq, _ = corrupt_sequence(reg.truth[first:last + 4], read_error, rng)
        for model in models:
            ev = simulate_event(truth_states, model, rng, first, last, p_unaligned=p_unaligned, jitter=jitter)
            al = ev.ref_align > 0
            idx = np.clip((ev.ref_align[al] - 1).astype(np.int64), 0, len(pos_map) - 1)   # jitter may step past the end
            ev.ref_align[al] = np.minimum(pos_map[idx] + 1, n_draft_states).astype("f8")
            ev.sequence = read_seq
            ev.makecontiguous()
            reg.events.append(ev)
    return reg


def point_mutations(seq):
    """All single-base del / sub / ins candidates in FindPointMutations order
    (cpp/FindMutations.cpp:191-234): per position i < len-4: del, subs (ACGT order, skipping the
    same base), 4 insertions.  Returns (start list, orig list, mut list)."""
    starts, origs, muts = [], [], []
    for i in range(max(len(seq) - 4, 0)):
        b = seq[i]
        starts.append(i); origs.append(b); muts.append("")
        for c in BASES:
            if c != b:
                starts.append(i); origs.append(b); muts.append(c)
        for c in BASES:
            starts.append(i); origs.append(""); muts.append(c)
    return starts, origs, muts


def random_mutations(seq, count, rng, max_len=4):
    """Random single/multi-base candidate edits sorted by start (config 4 style)."""
    starts = np.sort(rng.integers(0, len(seq), size=count))
    out_s, out_o, out_m = [], [], []
    for s in starts:
        lo = int(rng.integers(0, max_len + 1))
        lm = int(rng.integers(0, max_len + 1))
        if lo == 0 and lm == 0:
            lm = 1
        orig = seq[s:s + lo]
        mut = random_sequence(lm, rng) if lm else ""
        out_s.append(int(s)); out_o.append(orig); out_m.append(mut)
    return out_s, out_o, out_m
