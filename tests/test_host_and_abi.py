"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol the
header declares, host-only helpers match the checker, and compute calls fail loudly without a GPU
(no silent CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from poreseq_b200 import build, poreseqcpp, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native_lib():
    build.build()
    return ctypes.CDLL(build.LIB)


def test_every_declared_symbol_is_exported(native_lib):
    header = open(os.path.join(ROOT, "include", "poreseq_b200.h")).read()
    names = set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 20
    missing = [n for n in sorted(names) if not hasattr(native_lib, n)]
    assert not missing, missing


def test_seq_to_states_matches_checker(orc):
    rng = np.random.default_rng(3)
    for seq in [synth.random_sequence(50, rng), "ACGTACGTNACGTACGGTNNACGTTGCA", "NACGT", "ACGT", "ACG-TACGTAC"]:
        assert poreseqcpp.seqtostates(seq) == orc.seq_to_states(seq).tolist()


def test_point_mutation_enumeration_is_host_only(orc):
    """FindPointMutations order (cpp/FindMutations.cpp:191-234) without touching the GPU."""
    reg = synth.make_region(60, 1, seed=9)
    c = poreseqcpp.Context(0)
    nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params)
    cap = 8 * len(reg.sequence)
    n = ctypes.c_int(0)
    st = np.zeros(cap, dtype=np.int32)
    og = ctypes.create_string_buffer(cap)
    mu = ctypes.create_string_buffer(cap)
    c.check(c.lib.ps_find_point_mutations(nr.handle, cap, ctypes.byref(n), st.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), og, mu))
    s2, o2, m2 = synth.point_mutations(reg.sequence)
    assert n.value == len(s2) == 8 * (len(reg.sequence) - 4)
    assert st[:n.value].tolist() == s2
    assert [chr(b) if b else "" for b in og.raw[:n.value]] == o2
    assert [chr(b) if b else "" for b in mu.raw[:n.value]] == m2


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    reg = synth.make_region(60, 1, seed=9)
    pa = poreseqcpp.PSAlign()
    pa.sequence, pa.events, pa.params = reg.sequence, reg.events, reg.params
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pa.ScoreEvents()


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under poreseq_b200/ may mention it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "poreseq_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def test_swalign_matches_reference(ref):
    """Host-side integer Smith-Waterman (cpp/swlib.cpp:211-340): pairs, accuracy, tie-breaking."""
    rng = np.random.default_rng(5)
    for n, rate in [(1, 0.0), (8, 0.3), (60, 0.2), (300, 0.15), (700, 0.4), (500, 0.02)]:
        a = synth.random_sequence(n, rng)
        b, _ = synth.corrupt_sequence(a, rate, rng)
        if not b:
            b = "A"
        acc, pairs = poreseqcpp.swalign(a, b)
        want_acc, _, want_pairs = ref.swfull(a, b)
        assert pairs == want_pairs
        assert acc == want_acc or (np.isnan(acc) and np.isnan(want_acc))
    # low-complexity sequences exercise the tie rules
    acc, pairs = poreseqcpp.swalign("AAAAAAAAAACCCCCCCCCC", "AAAAACCCCCCCCCCCCAAAAA")
    want_acc, _, want_pairs = ref.swfull("AAAAAAAAAACCCCCCCCCC", "AAAAACCCCCCCCCCCCAAAAA")
    assert pairs == want_pairs and acc == want_acc


def test_packed_marshalling_equals_per_event():
    """ps_region_add_events (one call per region) builds the same native region as one
    ps_region_add_event per event: same sequence, same per-event alignments, same event count.
    No compute call, so this runs without a GPU."""
    import numpy as np
    from poreseq_b200 import poreseqcpp, synth
    reg = synth.make_region(300, 3, seed=9, draft_error=0.05, partial=0.4)
    ctx = poreseqcpp.Context(0)
    a = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params, "point_width")
    pack = poreseqcpp.PackedRegion(reg.sequence, reg.events, reg.params)
    b = poreseqcpp.NativeRegion.from_packed(ctx, pack, "point_width")
    assert a.sequence() == b.sequence() == reg.sequence
    assert ctx.lib.ps_region_num_events(a.handle) == ctx.lib.ps_region_num_events(b.handle) == len(reg.events)
    for e in range(len(reg.events)):
        ra, rl = a.event_align(e)
        rb, lb = b.event_align(e)
        assert np.array_equal(ra, rb) and np.array_equal(rl, lb)
        assert np.array_equal(ra, reg.events[e].ref_align)
    # a model index outside the table is refused and leaves the region unchanged
    bad = poreseqcpp.PackedRegion(reg.sequence, reg.events, reg.params)
    bad.model_index = bad.model_index + 7
    with pytest.raises(RuntimeError):
        poreseqcpp.NativeRegion.from_packed(ctx, bad, "point_width")
    assert pack.nbytes() > 0


def test_bulk_region_creation_equals_per_region():
    """ps_regions_create (one call, worker threads) == ps_region_create + ps_region_add_events per region."""
    import numpy as np
    from poreseq_b200 import poreseqcpp, synth
    regs = [synth.make_region(200 + 37 * k, 2 + k % 3, seed=40 + k, draft_error=0.04, partial=0.3) for k in range(9)]
    ctx = poreseqcpp.Context(0)
    packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
    many = poreseqcpp.native_regions_from_packed(ctx, packs, "point_width")
    assert len(many) == len(regs)
    for r, p, m in zip(regs, packs, many):
        one = poreseqcpp.NativeRegion.from_packed(ctx, p, "point_width")
        assert m.sequence() == one.sequence() == r.sequence
        assert m.n_levels == one.n_levels
        for e in range(len(r.events)):
            a, b = m.event_align(e), one.event_align(e)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # one bad descriptor refuses the whole batch
    bad = poreseqcpp.PackedRegion(regs[0].sequence, regs[0].events, regs[0].params)
    bad.model_index = bad.model_index + 3
    with pytest.raises(RuntimeError):
        poreseqcpp.native_regions_from_packed(ctx, packs[:2] + [bad], "point_width")


def test_bulk_create_and_destroy_without_gpu():
    """ps_regions_create / ps_regions_destroy are host-only (no CUDA call): a batch of regions can be marshalled and
    released on a box without a GPU; NULL entries and empty batches are ignored."""
    import ctypes as C
    from poreseq_b200 import poreseqcpp, synth
    ctx = poreseqcpp.Context(0)
    regs = [synth.make_region(200, 3, seed=s + 1) for s in range(6)]
    packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
    nrs = poreseqcpp.native_regions_from_packed(ctx, packs, "point_width")
    assert [ctx.lib.ps_region_num_events(n.handle) for n in nrs] == [len(r.events) for r in regs]
    assert [n.sequence() for n in nrs] == [r.sequence for r in regs]
    poreseqcpp.close_regions(nrs)
    assert all(n.handle is None for n in nrs)
    poreseqcpp.close_regions(nrs)                      # already released: nothing to do
    ctx.lib.ps_regions_destroy((C.c_void_p * 2)(None, None), 2)
    ctx.lib.ps_regions_destroy(None, 0)
    h, d = ctx.last_bytes()
    assert (h, d) == (0, 0)                            # no batch has run on this context


def test_event_pack_round_trip(tmp_path):
    """Event-pack files (poreseq_b200/eventpack.py): regions written to one flat file come back as PackedRegion views with
    identical contents, and marshal into native regions like the in-memory ones (host only)."""
    from poreseq_b200 import eventpack, poreseqcpp, synth
    regs = [synth.make_region(150 + 30 * k, 2 + k, seed=40 + k, draft_error=0.05 * k, partial=0.2 * k,
                              params=dict(realign_width=40, scoring_width=12, point_width=6, lik_offset=4.5)) for k in range(3)]
    path = str(tmp_path / "regions.psep")
    eventpack.write_pack(path, regs)
    pack = eventpack.read_pack(path)
    assert len(pack) == 3
    for reg, got in zip(regs, pack):
        want = poreseqcpp.PackedRegion(reg.sequence, reg.events, reg.params)
        assert got.sequence == want.sequence and got.params == {k: want.params[k] for k in want.params}
        for name in ("n0", "mean", "stdv", "ref_align", "ref_like", "model_index", "models", "probs", "complement"):
            a, b = getattr(got, name), getattr(want, name)
            assert a.dtype == b.dtype and np.array_equal(a, b), name
        assert got.seq2d == want.seq2d
    ctx = poreseqcpp.Context(0)
    nrs = poreseqcpp.native_regions_from_packed(ctx, list(pack), "point_width")
    assert [n.sequence() for n in nrs] == [r.sequence for r in regs]
    assert [ctx.lib.ps_region_num_events(n.handle) for n in nrs] == [len(r.events) for r in regs]
    poreseqcpp.close_regions(nrs)
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(ValueError):
        eventpack.read_pack(path)


def test_train_parameter_variation():
    """vary_params / set_params of the training driver (poreseq/Params.py:31-60, poreseq/EventData.py:286-312)."""
    import random
    from poreseq_b200 import drivers, synth
    base = dict(realign_width=300, lik_offset=4.5, skip_t=0.141, skip_c=0.088, stay_t=0.043, stay_c=0.057,
                extend_t=0.072, extend_c=0.046, insert_t=0.020, insert_c=0.025)
    vs = drivers.vary_params(base, random.Random(5), 16)
    assert len(vs) == 16
    for v in vs:
        changed = [k for k in base if v[k] != base[k]]
        assert len(changed) == 3 and all(k[-2:] in ("_t", "_c") for k in changed)
        assert v["lik_offset"] == 4.5 and v["realign_width"] == 300
    assert vs == drivers.vary_params(base, random.Random(5), 16)          # reproducible from the generator
    reg = synth.make_region(120, 2, seed=3)
    drivers.set_params(reg.events, dict(skip_t=0.3, skip_c=0.4, insert_c=0.5, lik_offset=9.0))
    for ev in reg.events:
        assert ev.model.prob_skip == (0.4 if ev.model.complement else 0.3)
        if ev.model.complement:
            assert ev.model.prob_insert == 0.5


def test_bad_arguments_are_reported_not_dereferenced():
    """Error convention of the C-ABI (include/poreseq_b200.h): null handles, null arrays, negative sizes and
    too-small output buffers return PS_E_ARG / PS_E_CAPACITY with a message naming the entry point; nothing is
    dereferenced.  (The reference never reports errors from C++; these are the boundary's own.)"""
    C = ctypes
    L = poreseqcpp.lib()
    c = poreseqcpp.Context(0)
    reg = synth.make_region(30, 1, seed=4)
    nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params)
    vp = C.c_void_p

    def call(name, *args):
        fn = getattr(L, name)
        saved = fn.argtypes
        fn.argtypes = None                                 # raw call: let None through as NULL
        try:
            return fn(*args)
        finally:
            fn.argtypes = saved

    h = vp(nr.handle) if not isinstance(nr.handle, vp) else nr.handle
    assert call("ps_score_alignments", h, None, None) == -1
    assert b"ps_score_alignments" in L.ps_last_error(c.handle)
    assert call("ps_score_mutations", h, C.c_int(3), None, None, None, None) == -1
    assert b"ps_score_mutations" in L.ps_last_error(c.handle)
    assert call("ps_make_mutations", h, C.c_int(-1), None, None, None, None, None) == -1
    assert call("ps_mutate", h, C.c_int(2), None, C.c_int(1), None) == -1
    assert call("ps_find_mutations", h, C.c_int(2), None, None) == -1
    assert call("ps_map_alignments", h, None) == -1
    assert call("ps_region_get_event_align", h, C.c_int(99), None, None) == -1
    buf = C.create_string_buffer(4)
    assert call("ps_region_get_sequence", h, buf, C.c_int(4)) == -3
    n = C.c_int(0)
    assert call("ps_find_point_mutations", h, C.c_int(3), C.byref(n), None, None, None) == -3 and n.value == 8 * 26
    for name, args in [("ps_score_alignments", (None, None, None)), ("ps_refine", (None, None)),
                       ("ps_region_get_sequence", (None, None, C.c_int(0))), ("ps_mutate", (None, C.c_int(0), None, C.c_int(1), None)),
                       ("ps_viterbi_mutate", (None, C.c_int(0), C.c_double(.1), C.c_double(.1), C.c_double(.1), C.c_double(.1), None)),
                       ("ps_seq_to_states", (None, C.c_int(5), None)), ("ps_last_timing", (None, None)),
                       ("ps_score_points_batch", (None, C.c_int(2), C.c_int(10), None, None, None, None, None, None))]:
        assert call(name, *args) == -1, name
    call("ps_region_destroy", None)
    call("ps_destroy", None)
    nr.close()


def test_recycled_level_arrays_keep_their_contents():
    """The per-event level arrays come from a recycling allocator (csrc/ps_internal.h PoolAlloc): batches of regions
    created, read back and destroyed over many cycles, from several threads at once, always hold the caller's data."""
    import threading
    from poreseq_b200 import poreseqcpp, synth
    regs = [synth.make_region(120 + 40 * s, 1 + s % 3, seed=60 + s, partial=0.3, p_unaligned=0.2) for s in range(8)]
    packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
    errors = []

    def cycle(tid):
        try:
            ctx = poreseqcpp.Context(0)
            for it in range(12):
                pick = [(tid + it + k) % len(regs) for k in range(1 + (tid + it) % 5)]
                nrs = poreseqcpp.native_regions_from_packed(ctx, [packs[i] for i in pick], "point_width")
                for i, nr in zip(pick, nrs):
                    for e, ev in enumerate(regs[i].events):
                        ra, rl = nr.event_align(e)
                        if not (np.array_equal(ra, ev.ref_align) and np.array_equal(rl, ev.ref_like)):
                            errors.append((tid, it, i, e))
                if it % 2:
                    poreseqcpp.close_regions(nrs)
                else:
                    for nr in nrs:
                        nr.close()
            ctx.close()
        except Exception as ex:                        # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=cycle, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:4]


def test_native_event_pack_reader(tmp_path):
    """ps_pack_* (csrc/ps_pack.cu): the library maps an event-pack file itself; every block of every region is where the
    Python reader finds it, parameters and 2D sequences come back, regions marshal straight from the mapping."""
    import ctypes as C
    from poreseq_b200 import eventpack
    regs = [synth.make_region(150 + 30 * k, 2 + k, seed=40 + k, draft_error=0.05 * k, partial=0.2 * k,
                              params=dict(realign_width=40, scoring_width=12, point_width=6, lik_offset=4.5)) for k in range(3)]
    regs.append(synth.make_region(40, 1, seed=50))
    regs[-1].events = []                                   # a region without events
    path = str(tmp_path / "regions.psep")
    eventpack.write_pack(path, regs)
    views = eventpack.read_pack(path)
    pack = eventpack.NativePack(path)
    assert len(pack) == len(regs) == len(views)

    def arr(ptr, ctype, n):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).copy() if n else np.zeros(0)

    for k, (reg, v) in enumerate(zip(regs, views)):
        d = pack.desc(k, "point_width")
        assert C.string_at(d.bases, d.len) == v.sequence and d.n_events == len(v.n0) and d.n_models == len(v.models)
        assert (d.params.lik_offset, d.params.scoring_width, d.params.realign_width) == (
            reg.params["lik_offset"], reg.params["point_width"], reg.params["realign_width"])
        assert pack.desc(k).params.scoring_width == reg.params["scoring_width"]
        n_lev = int(v.n0.sum())
        assert np.array_equal(arr(d.n0, C.c_int, d.n_events), v.n0)
        assert np.array_equal(arr(d.model_index, C.c_int, d.n_events), v.model_index)
        assert np.array_equal(arr(d.complement, C.c_int, d.n_events), v.complement)
        for name in ("mean", "stdv", "ref_align", "ref_like"):
            assert np.array_equal(arr(getattr(d, name), C.c_double, n_lev), getattr(v, name)), name
        assert np.array_equal(arr(d.models, C.c_double, d.n_models * 4096), v.models.ravel())
        assert np.array_equal(arr(d.probs, C.c_double, d.n_models * 4), v.probs.ravel())
        assert pack.param(k, "point_width") == reg.params["point_width"] and pack.param(k, "no_such_key") is None
        assert [pack.event_sequence(k, e) for e in range(d.n_events)] == v.seq2d
    ctx = poreseqcpp.Context(0)
    nrs = pack.regions(ctx, first=1, count=3, width_key="point_width")
    for reg, nr in zip(regs[1:], nrs):
        assert nr.sequence() == reg.sequence and ctx.lib.ps_region_num_events(nr.handle) == len(reg.events)
        for e, ev in enumerate(reg.events):
            ra, rl = nr.event_align(e)
            assert np.array_equal(ra, ev.ref_align) and np.array_equal(rl, ev.ref_like)
    poreseqcpp.close_regions(nrs)
    with pytest.raises(RuntimeError, match="ps_pack_regions_create"):
        pack.regions(ctx, first=3, count=2)
    with pytest.raises(IndexError):
        pack.desc(9)
    pack.close()


def test_native_event_pack_reader_refuses_damaged_files(tmp_path):
    """Truncated, re-labelled or internally inconsistent packs are refused at ps_pack_open with a reason (nothing is
    dereferenced outside the mapping)."""
    import struct
    from poreseq_b200 import eventpack
    regs = [synth.make_region(80, 2, seed=70), synth.make_region(90, 1, seed=71)]
    good = str(tmp_path / "good.psep")
    eventpack.write_pack(good, regs)
    data = open(good, "rb").read()
    n, index_at = struct.unpack("<QQ", data[8:24])
    off1 = struct.unpack("<Q", data[index_at + 16:index_at + 24])[0]

    def patched(at, raw):
        b = bytearray(data)
        b[at:at + len(raw)] = raw
        return bytes(b)

    damaged = {
        "missing": None,
        "empty": b"",
        "magic": b"PSEP0002" + data[8:],
        "truncated": data[:len(data) // 2],
        "no_index": data[:index_at + 8],
        "count": patched(8, struct.pack("<Q", 1 << 40)),
        "index_offset": patched(16, struct.pack("<Q", len(data) + 64)),
        "region_offset": patched(index_at, struct.pack("<Q", len(data) - 8)),
        "region_size": patched(index_at + 8, struct.pack("<Q", 1 << 50)),
        "levels": patched(off1 + 16, struct.pack("<I", 7)),                 # n_levels disagrees with sum(n0)
        "events": patched(off1 + 4, struct.pack("<I", 1 << 30)),            # n_events far beyond the block
        "model_index": patched(off1 + 32 + 24 * len(regs[1].params) + ((len(regs[1].sequence) + 7) & ~7) + 4 * 2 + 4, struct.pack("<i", 5)),
    }
    for name, raw in damaged.items():
        path = str(tmp_path / (name + ".psep"))
        if raw is not None:
            open(path, "wb").write(raw)
        with pytest.raises(ValueError, match="ps_pack_open"):
            eventpack.NativePack(path)
    assert len(eventpack.NativePack(good)) == 2


def test_swalign_fuzz_against_reference(ref):
    """The anti-diagonal host swfull (csrc/ps_swhost.cpp) against cpp/swlib.cpp:211-340 on ~900 pairs: small alphabets
    (ties everywhere), shifted / truncated / unrelated sequences, one-base inputs, both argument orders."""
    rng = np.random.default_rng(11)

    def same(a, b):
        x, y = poreseqcpp.swalign(a, b), ref.swfull(a, b)
        return (x[0] == y[0] or (np.isnan(x[0]) and np.isnan(y[0]))) and [tuple(p) for p in x[1]] == [tuple(p) for p in y[2]]

    fixed = [("A", "A"), ("A", "C"), ("AAAAAAAAAA", "AAAAAAA"), ("ACGTACGT", "TTTT"), ("A", "AAAAAAAA"), ("ACGT" * 50, "ACGT" * 37),
             ("AC" * 100, "CA" * 100), ("A" * 300, "A" * 299), ("ACGTN", "ACGTN"), ("NNNN", "NNNN")]
    for a, b in fixed:
        assert same(a, b) and same(b, a), (a, b)
    alph = [list("ACGT"), list("AC"), list("A")]
    for it in range(800):
        k = int(rng.integers(0, 3))
        la, lb = int(rng.integers(1, 120)), int(rng.integers(1, 120))
        a = "".join(rng.choice(alph[k], la))
        mode = it % 4
        if mode == 0:
            b = "".join(rng.choice(alph[k], lb))
        elif mode == 1:
            b = synth.corrupt_sequence(a, float(rng.choice([0.02, 0.1, 0.3])), rng)[0] or "A"
        elif mode == 2:
            b = (a[int(rng.integers(0, la)):] + "".join(rng.choice(alph[k], int(rng.integers(0, 20))))) or "C"
        else:
            b = ("".join(rng.choice(alph[k], int(rng.integers(0, 30)))) + a)[:max(1, lb)]
        assert same(a, b), (a, b)
    a = synth.random_sequence(2500, rng)
    assert same(a, synth.corrupt_sequence(a, 0.1, rng)[0]) and same(synth.random_sequence(2000, rng), a)


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/poreseq_b200.h is a C header (gcc -std=c99 -pedantic), and examples/score_points.c -- a caller with no
    Python and no C++ in it -- builds against it, opens an event pack and marshals its regions through the library."""
    import subprocess
    from poreseq_b200 import eventpack
    build.build()
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(inc, "poreseq_b200.h")], check=True)
    exe = str(tmp_path / "score_points")
    libdir = os.path.dirname(build.LIB)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + inc, os.path.join(ROOT, "examples", "score_points.c"),
                    "-L" + libdir, "-lporeseq_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    regs = [synth.make_region(120, 2, seed=80 + k) for k in range(3)]
    path = str(tmp_path / "regions.psep")
    eventpack.write_pack(path, regs)
    out = subprocess.run([exe, path], capture_output=True, text=True)
    assert "3 regions" in out.stdout, out.stdout + out.stderr
    import torch
    if not torch.cuda.is_available():
        assert out.returncode == 0 and "no CPU fallback" in out.stdout


def test_cython_stub_of_integration_md_builds_and_binds(tmp_path):
    """The reference-side binding of INTEGRATION.md section 2 is a real module (examples/cython_stub): Cython compiles it
    against include/poreseq_b200.h, it links to the library, its host-only entry points give the product's answers, the
    PSAlign marshalling reaches the library, and a compute call without a GPU raises the 'no CPU fallback' error."""
    import subprocess
    import sys
    from util import build_cython_stub
    build.build()
    where, log = build_cython_stub(tmp_path)
    assert where, log
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import poreseqcpp_b200 as m
from poreseq_b200 import synth, poreseqcpp
for a, b in [("ACGTTTTTACGT", "ACGTTTACGT"), ("GATTACA", "GCATGCT"), ("AAAAAAAAAA", "AAAAAAA")]:
    x, y = m.swalign(a, b), poreseqcpp.swalign(a, b)
    assert x[0] == y[0] and [tuple(p) for p in x[1]] == [tuple(p) for p in y[1]], (a, b)
assert m.seqtostates("ACGTACGTNACGT") == poreseqcpp.seqtostates("ACGTACGTNACGT")
reg = synth.make_region(80, 2, seed=3)
pa = m.PSAlign(); pa.sequence, pa.events, pa.params = reg.sequence, reg.events, reg.params
assert pa.NumEvents() == 4
import torch
if not torch.cuda.is_available():
    try:
        pa.Refine()
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Refine returned without a GPU")
print("stub ok")
""" % (where, ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "stub ok" in out.stdout, out.stdout + out.stderr


def test_make_mutations_accept_loop_matches_reference(ref):
    """The product's host-side MakeMutations (cpp/MakeMutations.cpp:74-146: std::sort by score with its tie placement,
    drop the negative tail, apply in order, invalidate neighbours within mutspc = 10, shift later starts) against the
    reference's own C++ on crafted score lists: random positives, exact ties, all-equal scores, -1e-6 entries, edits at
    and past the end.  Lists that defer more than 10 edits need the GPU scorer for the recursion and are skipped here
    (they run in the GPU tests through Refine / Mutate)."""
    ctx = poreseqcpp.Context(0)
    compared = 0
    for seed in range(240):
        rng = np.random.default_rng(seed)
        reg = synth.make_region(int(rng.integers(30, 300)), 1, seed=seed + 1, params=dict(realign_width=20, scoring_width=8, point_width=4))
        L = len(reg.sequence)
        st, og, mu = synth.point_mutations(reg.sequence)
        s2, o2, m2 = synth.random_mutations(reg.sequence, int(rng.integers(0, 40)), rng, max_len=6)
        st, og, mu = st + s2 + [L, L + 3, 0], og + o2 + ["", "", reg.sequence[:7]], mu + m2 + ["AC", "G", ""]
        n = len(st)
        sc = -np.abs(rng.normal(3, 2, n)) - 0.01
        pos = rng.choice(n, size=int(rng.integers(0, 18)), replace=False)
        mode = seed % 4
        if mode == 0:
            sc[pos] = rng.uniform(0, 5, len(pos))
        elif mode == 1:
            sc[pos] = rng.integers(0, 3, len(pos)).astype(float)       # exact ties, zeros included
        elif mode == 2:
            sc[pos] = 1.0
        else:
            sc[pos] = rng.uniform(0, 5, len(pos))
            sc[rng.choice(n, 5)] = -1e-6
        want_seq, want_nb, _ = ref.make_mutations(reg, st, og, mu, sc.tolist())
        nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params, "point_width")
        try:
            nb = nr.make_mutations(st, og, mu, sc.tolist())
        except RuntimeError as e:
            assert "no CPU fallback" in str(e)
            continue
        finally:
            seq = nr.sequence()
            nr.close()
        compared += 1
        assert (seq, nb) == (want_seq, want_nb), seed
    import torch
    assert compared >= (240 if torch.cuda.is_available() else 120)


def test_map_alignments_host_path_matches_reference(ref):
    """ps_map_alignments is host-only (cpp/EventUtil.cpp:12-55: swfull + fillinds, then every level's ref_align carried
    over through lower_bound): against the reference's own C++ on 150 random regions -- drafts with 0-30 % errors, partial
    and unaligned reads, jittered seed alignments, new sequences that are shifted, truncated or unrelated."""
    ctx = poreseqcpp.Context(0)
    for seed in range(150):
        rng = np.random.default_rng(300 + seed)
        reg = synth.make_region(int(rng.integers(12, 260)), int(rng.integers(1, 4)), seed=seed + 1,
                                draft_error=float(rng.choice([0, 0.05, 0.2])), partial=float(rng.choice([0, 0.5])),
                                p_unaligned=float(rng.choice([0, 0.3])), jitter=int(rng.choice([0, 3])))
        mode = seed % 4
        if mode == 0:
            newseq = synth.corrupt_sequence(reg.sequence, float(rng.choice([0.02, 0.1, 0.3])), rng)[0]
        elif mode == 1:
            newseq = reg.sequence[int(rng.integers(0, len(reg.sequence) // 2)):] + synth.random_sequence(int(rng.integers(0, 30)), rng)
        elif mode == 2:
            newseq = synth.random_sequence(int(rng.integers(0, 20)), rng) + reg.sequence[:int(rng.integers(6, len(reg.sequence) + 1))]
        else:
            newseq = synth.random_sequence(int(rng.integers(5, 200)), rng)
        if len(newseq) < 5:
            continue
        want = ref.map_alignments(reg, newseq)
        nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params)
        nr.map_alignments(newseq)
        got = [nr.event_align(e) for e in range(len(reg.events))]
        assert nr.sequence() == newseq and all(np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]) for g, w in zip(got, want)), seed
        nr.close()


def test_find_mutations_host_half_matches_reference(ref):
    """ps_pick_candidates = the host-only second half of the product's FindMutations (swfull per seed, CUSUM of the profile
    differences, greedy peak picking; cpp/FindMutations.cpp:51-186).  Fed with the likelihood profiles of the reference's
    own ScoreAlignments (realign, MapAlignments per seed, realign the mapped copy -- what the GPU half computes), it must
    return the reference's FindMutations candidate list, in order, incl. a repeated seed."""
    import copy
    ctx = poreseqcpp.Context(0)

    def aligned(reg, seq, aligns):
        rr = copy.deepcopy(reg)
        rr.sequence = seq
        for ev, (ra, rl) in zip(rr.events, aligns):
            ev.ref_align, ev.ref_like = ra, rl
        return rr

    nonempty = 0
    for seed in range(40):
        rng = np.random.default_rng(seed)
        reg = synth.make_region(int(rng.integers(40, 400)), int(rng.integers(1, 4)), seed=seed + 1,
                                draft_error=float(rng.choice([0.03, 0.1])), partial=float(rng.choice([0, 0.4])),
                                params=dict(realign_width=40, scoring_width=12, point_width=6))
        seeds = [ev.sequence for ev in reg.events[::2]] + [synth.corrupt_sequence(reg.sequence, 0.08, rng)[0]]
        seeds.append(seeds[0])
        want, _ = ref.find_mutations(reg, seeds)
        _, base, a1 = ref.score_alignments(reg, True)
        reg1 = aligned(reg, reg.sequence, a1)
        profs = [ref.score_alignments(aligned(reg1, sd, ref.map_alignments(reg1, sd)), True)[1] if len(sd) >= 5 else np.zeros(len(sd))
                 for sd in seeds]
        nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params)
        assert nr.pick_candidates(seeds, base, profs) == want, seed
        nonempty += len(want) > 0
        nr.close()
    assert nonempty >= 30


def test_round2_entry_points_refuse_bad_arguments_and_have_no_cpu_path():
    """The entry points added in round 2 (ScoreEvents, the direct ScorePoints path, the consensus loop, the event-shard
    communicator): bad arguments come back as PS_E_ARG with a message, and on a machine without a GPU the compute calls
    fail with PS_E_CUDA -- there is no CPU path behind them either."""
    import ctypes as C
    L = poreseqcpp.lib()
    assert L.ps_score_events(None, None) == -1
    assert L.ps_score_events_batch(None, 3, None) == -1
    assert L.ps_consensus(None, 4, 20, None) == -1
    assert L.ps_consensus_batch(None, None, 1, 4, 20, 4) == -1
    assert L.ps_score_points_direct_begin(None, 1, None, 0, None, None, None, None, None) == -1
    assert L.ps_comm_init(None, None, 0, 0, 1, 1) == -1
    assert L.ps_region_get_stage(None, 0, None, 0, None, 0, None) == -1
    ctx = poreseqcpp.Context(0)
    try:
        reg = synth.make_region(80, 3, seed=11)
        nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params)
        assert L.ps_region_num_stages(nr.handle) == 0
        # a sharded call without a communicator is an argument error, whatever the machine
        st = np.zeros(1, dtype=np.int32)
        out = np.zeros(1)
        o = (C.c_char_p * 1)(b"A"); m = (C.c_char_p * 1)(b"C")
        rc = L.ps_score_mutations_sharded(nr.handle, 1, st.ctypes.data_as(C.POINTER(C.c_int)), o, m, out.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == -1 and b"communicator" in L.ps_last_error(ctx.handle)
        # the same handle twice in one batch
        import torch
        if not torch.cuda.is_available():
            for call in (lambda: nr.score_events(), lambda: nr.consensus(1, 8),
                         lambda: poreseqcpp.score_points_direct(ctx, [poreseqcpp.PackedRegion(reg.sequence, reg.events, reg.params)])):
                with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
                    call()
        nr.close()
    finally:
        ctx.close()


def _updaterefs(ra):
    """cpp/EventData.h:110-169 restated in Python (IEEE doubles, the reference's operation order)."""
    n = len(ra)
    a = 0
    while a < n and not ra[a] > 0:
        a += 1
    z = n - 1
    while z >= 0 and not ra[z] > 0:
        z -= 1
    if a == n or z < 0:
        return None
    ri = np.array(ra, dtype="f8")
    with np.errstate(all="ignore"):
        m = np.float64(ra[z] - ra[a]) / np.float64(z - a)
        b = ra[a] - m * a
        last = -1
        for i in range(n):
            if i < a or i > z:
                ri[i] = m * i + b
            elif ra[i] > 0:
                if last > 0:
                    mm = np.float64(ra[i] - ra[last]) / np.float64(i - last)
                    for j in range(last + 1, i):
                        ri[j] = mm * (j - last) + ra[last]
                last = i
    return ri


def _lower_bound(a, v):
    """std::lower_bound as libstdc++ probes it (cpp/EventData.h:172-183) -- on an array with NaNs the probes decide."""
    first, length = 0, len(a)
    while length > 0:
        half = length >> 1
        if a[first + half] < v:
            first += half + 1
            length -= half + 1
        else:
            length = half
    return first


def test_band_centres_follow_the_reference_binary_search():
    """ps_band_centres (host only): ref_index and getrefstate(c) of events with ordinary, jittered (unsorted), sparse,
    single-level and empty alignments against updaterefs + std::lower_bound restated above.  The single aligned level
    (0/0 slope, ref_index of NaNs) is the case the GPU sweep found: a linear merge gives other centres than the
    reference's binary search there."""
    rng = np.random.default_rng(5)
    ctx = poreseqcpp.Context(0)
    try:
        for trial in range(60):
            reg = synth.make_region(int(rng.integers(30, 200)), 2, seed=300 + trial, partial=float(rng.choice([0, 0.5])),
                                    jitter=int(rng.choice([0, 0, 4])))
            n_cols = len(reg.sequence) + 8
            for e, ev in enumerate(reg.events):
                ra = np.array(ev.ref_align, dtype="f8")
                kind = (trial + e) % 6
                if kind == 1:                                   # one aligned level: somewhere, first, last
                    keep = [int(rng.integers(0, len(ra))), 0, len(ra) - 1][trial % 3]
                    v = max(float(ra[keep]), 1.0)
                    ra[:] = 0
                    ra[keep] = v
                elif kind == 2:
                    ra[:] = 0                                   # no alignment at all
                elif kind == 3:
                    ra[rng.random(len(ra)) < 0.8] = -1          # mostly insertions
                elif kind == 4:
                    ra[:] = 0
                    ra[0] = 3.0
                    ra[-1] = 3.0 + len(ra)                      # aligned ends only; an aligned level 0 (the `last > 0` quirk)
                ev.ref_align = ra
            nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params)
            for e, ev in enumerate(reg.events):
                cen, ri, mono = nr.band_centres(e, n_cols)
                want_ri = _updaterefs(np.asarray(ev.ref_align, dtype="f8"))
                if want_ri is None:
                    assert ri is None and np.all(cen == 1), (trial, e)
                    continue
                assert ri is not None and np.array_equal(ri, want_ri, equal_nan=True), (trial, e)
                want = np.array([_lower_bound(want_ri, float(c)) for c in range(n_cols)])
                assert np.array_equal(cen, want), (trial, e, (trial + e) % 6)
                assert mono == bool(np.all(np.diff(want) >= 0)) or mono, (trial, e)
            nr.close()
    finally:
        ctx.close()


def test_viterbi_positions_follow_the_reference_loop():
    """ps_viterbi_positions (host only): the positions ViterbiMutate keeps, against cpp/Viterbi.cpp:262-325 restated with
    std::find over ref_index (getrefstates, cpp/EventData.h:187-204).  Sparse alignments make the extrapolated ends of
    ref_index integer-valued, so reads "sit" on positions beyond every read's refend -- where the library used to stop."""
    rng = np.random.default_rng(9)
    ctx = poreseqcpp.Context(0)
    try:
        seen_beyond = 0
        for trial in range(60):
            reg = synth.make_region(int(rng.integers(20, 120)), int(rng.integers(1, 4)), seed=700 + trial,
                                    partial=float(rng.choice([0, 0.5])))
            for e, ev in enumerate(reg.events):
                ra = np.array(ev.ref_align, dtype="f8")
                kind = (trial + e) % 4
                if kind == 1:                                   # two aligned levels one base apart: slope 1, integer tails
                    k = int(rng.integers(0, len(ra) - 1))
                    v = max(float(ra[k]), 1.0)
                    ra[:] = 0
                    ra[k], ra[k + 1] = v, v + 1
                elif kind == 2:
                    keep = rng.random(len(ra)) < 0.3
                    keep[int(rng.integers(0, len(ra)))] = True
                    ra[~keep] = 0
                ev.ref_align = ra
            nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params)
            refs = [_updaterefs(np.asarray(ev.ref_align, dtype="f8")) for ev in reg.events]
            if any(r is None for r in refs):
                with pytest.raises(RuntimeError):
                    nr.viterbi_positions()
                nr.close()
                continue
            spans = []
            for ev in reg.events:
                ra = np.asarray(ev.ref_align)
                al = ra[ra > 0]
                spans.append((int(al[0]), int(al[-1])))
            want = []
            refind = min(s[0] for s in spans)
            while True:
                nlik = sum(1 for ri in refs if np.any(ri == refind))
                nal = sum(1 for s in spans if s[0] <= refind <= s[1])
                if nlik <= nal * 0.2:
                    if nal == 0:
                        break
                    refind += 1
                    continue
                want.append((refind, nlik))
                refind += 1
            assert nr.viterbi_positions() == want, trial
            seen_beyond += bool(want) and want[-1][0] > max(s[1] for s in spans) + 1
            nr.close()
        assert seen_beyond > 0, "no trial reached past the last refend: the test lost its point"
    finally:
        ctx.close()
