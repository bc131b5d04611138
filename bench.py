#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (BASELINE.json metric: FindMutations / ScoreMutations GCUPS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--regions R] [--impl reference]

Workload (BASELINE.json configs[1]): the full single-base sub/ins/del scan (`PSAlign.ScorePoints` =
FindPointMutations + ScoreMutations, point_width 20, realign_width 300) of 1 kb regions at 10x
coverage (20 events of ~940 levels each, synthetic 5-mer pore model, injected skips/stays).  One
"step" is one pass of that scan over a batch of R independent regions per GPU, submitted through the
C-ABI in one call; with N GPUs every rank scans its own R regions (weak scaling, no data-path
collective: regions are independent).

Reported numbers
  value       GCUPS with inputs resident in HBM: algorithmic DP cells (wide fill both
              directions + (|mut|+5) x band rows per (mutation, event) pair) / device time of the kernel
              sequence, CUDA events on the library's stream, max over ranks
  e2e         same metric through the public API with HOST buffers: region marshalling, H2D, kernels,
              D2H all inside the timed region (wall clock bracketed by barrier + synchronize)
  roofline    dominant kernel against the FP32-issue roofline SURVEY.md 8d defines
              (24 lane-ops per cell, peak = SMs x 128 lanes x measured SM clock), plus HBM GB/s
  cpu_baseline  the reference's own C++ (oracle/_ref) on one host core, one region, same run

`--impl reference` times the reference's CPU implementation on all host cores (one process per
region, the reference's own scaling model, README.md:48-54).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the consensus legs keep 16 regions (streams) in flight per GPU: with the default 8 hardware queues streams share a
# connection and serialise behind each other (25 -> 40 kb/s); must be set before CUDA initialises
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from poreseq_b200 import synth  # noqa: E402

METRIC = "FindMutations GCUPS (ScorePoints full single-base scan)"
UNIT = "GCUPS"
OPS_PER_CELL = 24            # 8 emission + 8 add + 8 max lane-ops per cell
REGION_LEN, COVERAGE = 1000, 10


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p.get("hbm_gbs", 6650.0)), float(p.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe): ONE nvidia-smi
    process looping with -lms 200, started before the timed region and stopped after it (a fresh nvidia-smi
    per sample re-initialises NVML every time, which holds up the CUDA calls of every process on the box)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            out = ""
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                self.samples.append(f)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def make_regions(count, seed0):
    return [synth.make_region(REGION_LEN, COVERAGE, seed=seed0 + i) for i in range(count)]


def region_bytes(reg):
    """Host bytes marshalled per region: 4 level arrays + 4x1024 model doubles + 4 probs per event + bases."""
    return sum(4 * 8 * len(ev.mean) + 4 * 1024 * 8 + 4 * 8 for ev in reg.events) + len(reg.sequence)


def algorithmic_cells(reg):
    """SURVEY.md 8d cell count for ScorePoints on one region (same formula the library reports)."""
    rw, w = int(reg.params["realign_width"]), int(reg.params["point_width"])
    n_states = len(reg.sequence) - 4
    wide = 0.0
    for ev in reg.events:
        n0 = len(ev.mean)
        idx = np.asarray(ev.ref_align)          # generator emits a dense monotone seed alignment
        mid = np.clip(np.searchsorted(idx, np.arange(1, n_states + 1), side="left"), 1, n0)
        wide += float(np.sum(np.minimum(n0, mid + rw) - np.maximum(1, mid - rw) + 1))
    narrow = 0.0
    for ev in reg.events:
        rows = min(len(ev.mean), 2 * w + 1)
        narrow += n_states * (5 + 3 * 6 + 4 * 6) * rows
    return 2 * wide, narrow


# ------------------------------------------------------------------------------------------------
def cpu_score_points_worker(args):
    """One region through the reference's ScorePoints on this core.  `args` = (checker, seed) or (checker, region): the
    reference arm hands over regions generated before its timed region starts."""
    which, what = args
    from oracle import binding
    chk = binding.load(which)
    reg = synth.make_region(REGION_LEN, COVERAGE, seed=what) if isinstance(what, int) else what
    t0 = time.perf_counter()
    chk.score_points(reg)
    return time.perf_counter() - t0


def cpu_checker_kind():
    from oracle import binding
    if binding.available("ref"):
        return "ref", "reference"
    binding.build("oracle")
    return "oracle", "port"


def run_cpu_baseline(n_regions=4):
    """A few of the step's regions on one host core, timed beside the GPU run (rank 0, N=1 only)."""
    which, kind = cpu_checker_kind()
    cells, dt = 0.0, 0.0
    for seed in range(1, n_regions + 1):
        wide, narrow = algorithmic_cells(synth.make_region(REGION_LEN, COVERAGE, seed=seed))
        cells += wide + narrow
        dt += cpu_score_points_worker((which, seed))
    return {"value": cells / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "ScorePoints on %d of the step's 1 kb x 10x regions, one after the other on one core "
                      "(%.0f M cells, %.1f s)" % (n_regions, cells / 1e6, dt)}


def libc_srand(seed):
    import ctypes
    ctypes.CDLL("libc.so.6").srand(seed)


CONS_REGIONS = 64          # per GPU (32 left the tail of the last regions in flight visible: 17-23 kb/s run to run)
CONS_IN_FLIGHT = 32        # lanes (library threads + streams) per GPU: 16 lockstep groups of 4 regions with 2 lanes each


def consensus_cpu(length, coverage, seed):
    """The same policy driven through the reference's own C++ (oracle/_ref) on one host core."""
    import copy
    from oracle import binding
    ref = binding.load("ref")
    rr = copy.deepcopy(synth.make_region(length, coverage, seed=seed, draft_error=0.10))

    def sync(al):
        for ev, (ra, rl) in zip(rr.events, al):
            ev.ref_align, ev.ref_like = ra, rl

    libc_srand(1)
    t0 = time.perf_counter()
    seq, _, al = ref.mutate(rr, [ev.sequence for ev in rr.events[::2]], reps=4)
    rr.sequence = seq; sync(al)
    for _ in range(4):
        seeds = ref.viterbi_mutate(rr, nkeep=16, seed=None)
        seq, _, al = ref.mutate(rr, seeds, reps=4)
        rr.sequence = seq; sync(al)
        seq, nb, al = ref.refine(rr)
        rr.sequence = seq; sync(al)
        if nb == 0:
            break
    dt = time.perf_counter() - t0
    t = int(rr.params.get("end_trim", 0))
    out = rr.sequence[t:-t] if t and len(rr.sequence) > 2 * t else rr.sequence
    return {"value": length / 1000.0 / dt, "unit": "kb/s", "seconds": dt, "cores": 1, "kind": "reference"}, out


def golden(name):
    """tests/golden/f_<name>.npz (outputs of the reference's own C++ at full size), or None."""
    path = os.path.join(ROOT, "tests", "golden", "f_%s.npz" % name)
    return np.load(path) if os.path.exists(path) else None


def fast_mode_error(ctx_fast, device):
    """Largest relative error the FAST mode leaves against the EXACT (bit-identical) mode, measured in this run on one
    of the step's regions; scores >= -tau are re-scored exactly, so they must be equal."""
    from poreseq_b200 import poreseqcpp
    reg = synth.make_region(REGION_LEN, COVERAGE, seed=1, draft_error=0.03)
    cx = poreseqcpp.Context(device)
    try:
        ex = poreseqcpp.NativeRegion(cx, reg.sequence, reg.events, reg.params, "point_width").score_points()[3].copy()
        fa = poreseqcpp.NativeRegion(ctx_fast, reg.sequence, reg.events, reg.params, "point_width").score_points()[3].copy()
    finally:
        cx.close()
    rel = np.abs(fa - ex) / np.maximum(np.abs(ex), 1e-300)
    return {"max_rel_err": float(rel.max()), "tolerance": 1e-4, "edits": int(len(ex)),
            "accept_reject_identical": bool(np.array_equal(fa >= 0, ex >= 0)),
            "bit_identical_above_threshold": bool(np.array_equal(fa[ex > -0.4 * len(reg.events)], ex[ex > -0.4 * len(reg.events)])),
            "sample": "ScorePoints of one 1 kb x 10x region with a 3% draft error, FAST vs EXACT"}


def leg_score_events(device, sms, sm_max_mhz, cpu=True):
    """BASELINE.json configs[0]: PSAlign.ScoreEvents on a 1 kb region at 10x -- FAST mode = k_score_f32 (score-only FP32
    fill).  One region (latency) and 44 regions per call (throughput, the kernel's roofline)."""
    from poreseq_b200 import poreseqcpp
    ctx = poreseqcpp.Context(device)
    ctx.set_precision("fast")
    cx = poreseqcpp.Context(device)
    try:
        regs = [synth.make_region(REGION_LEN, COVERAGE, seed=1 + k) for k in range(44)]
        nrs = [poreseqcpp.NativeRegion(ctx, r.sequence, r.events, r.params) for r in regs]
        one = nrs[:1]
        exact = poreseqcpp.NativeRegion(cx, regs[0].sequence, regs[0].events, regs[0].params).score_events()
        for _ in range(3):
            got = poreseqcpp.score_events_batch(ctx, one)[0]
            poreseqcpp.score_events_batch(ctx, nrs)
        t0 = time.perf_counter()
        kern = []
        for _ in range(20):
            poreseqcpp.score_events_batch(ctx, one)
            kern.append(ctx.last_timing()["forward"])
        lat = (time.perf_counter() - t0) / 20
        cells1, _ = ctx.last_cells()
        kb = []
        t0 = time.perf_counter()
        for _ in range(20):
            poreseqcpp.score_events_batch(ctx, nrs)
            kb.append(ctx.last_timing()["forward"])
        wall44 = (time.perf_counter() - t0) / 20
        cells44, _ = ctx.last_cells()
        kb.sort(); kern.sort()
        k44 = kb[len(kb) // 2] * 1e-3
        peak = sms * 128 * sm_max_mhz * 1e6 / 1e12
        ach = cells44 * OPS_PER_CELL / k44 / 1e12
        out = {"config": "PSAlign.ScoreEvents, 1 kb region x 10x coverage (BASELINE.json configs[0]), fast precision: score-only FP32 fill",
               "one_region": {"seconds_per_call": lat, "kernel_ms": kern[len(kern) // 2], "gcups_call": cells1 / lat / 1e9,
                              "gcups_kernel": cells1 / (kern[len(kern) // 2] * 1e-3) / 1e9, "cells": cells1},
               "batch_44_regions": {"kernel_ms": k44 * 1e3, "gcups_kernel": cells44 / k44 / 1e9, "gcups_call": cells44 / wall44 / 1e9,
                                    "cells": cells44},
               "roofline": {"bound": "fp32-issue", "kernel": "k_score_f32", "achieved": ach, "peak": peak, "unit": "Tlane-op/s",
                            "frac": ach / peak, "ops_per_cell": OPS_PER_CELL, "traffic": None,
                            "note": "44 regions per launch; nothing is stored: HBM traffic is the 16 B level records, read once per CTA"},
               "max_rel_err_vs_exact": float(np.max(np.abs(got - exact) / exact)), "tolerance": 1e-4}
        if cpu:
            from oracle import binding
            which, kind = cpu_checker_kind()
            chk = binding.load(which)
            t0 = time.perf_counter()
            want, _, _ = chk.score_alignments(regs[0])
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": cells1 / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
                                   "sample": "the same region, ScoreAlignments of the reference on one core",
                                   "exact_mode_bit_identical": bool(np.array_equal(want, exact))}
        poreseqcpp.close_regions(nrs)
        return out
    finally:
        ctx.close(); cx.close()


def leg_variant(rank, world, device, dist, torch):
    """BASELINE.json configs[3], ONE of its six regions: ScoreMutations (poreseq variant -m) of 1200 single / multi-base
    edits at scoring_width 100 against a 10 kb region at 100x coverage (200 events).  With N GPUs the region's events are
    split across the ranks and the per-mutation sums are combined over NCCL inside the library, in event order
    (ps_score_mutations_sharded: bit-identical to one GPU).  This is the problem of tests/golden/f_c3.npz."""
    from poreseq_b200 import poreseqcpp, sharding
    kw = dict(length=10000, coverage=100, seed=17)
    reg = synth.make_region(**kw)
    st, og, mu = synth.random_mutations(reg.sequence, 1200, np.random.default_rng(4242), max_len=4)
    ctx = poreseqcpp.Context(device)
    ctx.set_precision("fast")
    try:
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid = torch.frombuffer(bytearray(poreseqcpp.comm_unique_id()), dtype=torch.uint8).cuda()
            dist.broadcast(uid, 0)
            ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world, ordered=True)
            shard = sharding.RegionShard(reg, rank, world)
            nr = poreseqcpp.NativeRegion(ctx, shard.sequence, shard.events, shard.params)
            call = lambda: nr.score_mutations_sharded(st, og, mu)
        else:
            nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params)
            call = lambda: nr.score_mutations(st, og, mu)
        call()                                               # buffers grow on the first call
        nr.close()
        if world > 1:
            nr = poreseqcpp.NativeRegion(ctx, shard.sequence, shard.events, shard.params)
        else:
            nr = poreseqcpp.NativeRegion(ctx, reg.sequence, reg.events, reg.params)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = call()
        dt = time.perf_counter() - t0
        wide, narrow = ctx.last_cells()
        t = torch.tensor([dt, wide + narrow], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t[0:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(t[1:2], op=dist.ReduceOp.SUM)
        out = {"config": "poreseq variant -m: ScoreMutations of 1200 edits (<= 4 bases) at scoring_width 100, one 10 kb region x 100x "
                         "(200 events) = one of the six regions of BASELINE.json configs[3]; events split over %d GPU(s), sums "
                         "combined in event order over NCCL inside the library" % world,
               "seconds": t[0].item(), "mutations_per_s": 1200 / t[0].item(), "gcups": t[1].item() / t[0].item() / 1e9,
               "n_gpus": world, "events_per_rank": len(nr.n_levels), "precision": "fast"}
        z = golden("c3")
        if z is not None:
            want = z["scores"]
            rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
            out["parity_vs_reference"] = {"accept_reject_identical": bool(np.array_equal(got >= 0, want >= 0)),
                                          "scores_ge0_bit_identical": bool(np.array_equal(got[want >= 0], want[want >= 0])),
                                          "max_rel_err": float(rel.max()), "tolerance": 1e-4,
                                          "golden": "tests/golden/f_c3.npz (reference C++, %.0f s on one core incl. input generation)" % float(z["seconds"])}
            out["cpu_baseline"] = {"seconds": float(z["seconds"]), "cores": 1, "kind": "reference",
                                   "sample": "recorded when the golden was generated (not timed in this run): the reference's ScoreMutations + ScoreAlignments on the same region"}
        nr.close()
        if world > 1:
            ctx.comm_destroy()
        return out
    finally:
        ctx.close()


def leg_polish(rank, world, device, dist, torch, per_rank=2, in_flight=2):
    """BASELINE.json configs[4] in small: consensus polish of 10 kb regions at 50x coverage (100 events each) read from an
    event-pack file, `per_rank` regions per GPU through ps_consensus_batch (the whole Mutate.py loop below the C-ABI),
    regions dealt out to the ranks round robin.  Region 0 is the problem of tests/golden/f_c4.npz."""
    import tempfile
    from poreseq_b200 import drivers, eventpack, poreseqcpp
    seeds = [11 + 100 * k for k in range(world * per_rank)][rank::world]
    regs = [synth.make_region(10000, 50, seed=sd, draft_error=0.10) for sd in seeds]
    path = os.path.join(tempfile.gettempdir(), "poreseq_b200_polish_%d_%d.pack" % (os.getpid(), rank))
    eventpack.write_pack(path, regs)
    ctx = poreseqcpp.Context(device)
    ctx.set_precision("fast")
    try:
        warm = synth.make_region(2000, 50, seed=5, draft_error=0.05)
        drivers.consensus_native([warm] * 1, ctx=ctx, in_flight=1)
        packs = list(eventpack.read_pack(path))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = drivers.consensus_native(packs, ctx=ctx, reps=4, in_flight=in_flight)
        dt = time.perf_counter() - t0
        accs = [poreseqcpp.swalign(r[0], g.truth[150:-150])[0] for r, g in zip(res, regs)]
        t = torch.tensor([dt, float(sum(accs)), float(len(regs))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t[0:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(t[1:3], op=dist.ReduceOp.SUM)
        out = {"config": "consensus polish (Mutate.py loop, ps_consensus_batch) of %d regions of 10 kb x 50x per GPU from an event-pack "
                         "file, %d in flight, fast precision (BASELINE.json configs[4] is ~512 such regions)" % (per_rank, in_flight),
               "value": 10.0 * t[2].item() / t[0].item(), "unit": "kb/s", "seconds": t[0].item(), "regions": int(t[2].item()),
               "n_gpus": world, "mean_accuracy_pct": t[1].item() / t[2].item()}
        z = golden("c4")
        if z is not None and rank == 0:
            want = str(z["stage_seqs"][-1])
            out["parity_vs_reference"] = {"region": "seed 11", "identical_final_sequence": bool(res[0][2][-1][1] == want),
                                          "identical_stages": bool([s[1] for s in res[0][2]] == z["stage_seqs"].tolist())}
            out["cpu_baseline"] = {"value": 10.0 / float(z["seconds"]), "unit": "kb/s", "seconds": float(z["seconds"]), "cores": 1,
                                   "kind": "reference", "sample": "recorded when the golden was generated (not timed in this run): the same loop on region seed 11"}
        return out
    finally:
        ctx.close()
        try:
            os.remove(path)
        except OSError:
            pass


def recorded_traffic(kernel, regions):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full
    capture (profiles/r2_traffic.json), valid only for the batch size it was captured at."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            rec = json.load(f)
        k = rec.get(kernel)
        if k and int(k.get("regions", -1)) == int(regions):
            return float(k["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def run_reference_arm(args, rank, world, out=sys.stdout):
    """--impl reference: the reference CPU path on all host cores, one process per region."""
    if rank != 0:
        return
    import multiprocessing as mp
    which, kind = cpu_checker_kind()
    cores = max(1, min(os.cpu_count() or 1, 64))
    reg = synth.make_region(REGION_LEN, COVERAGE, seed=1)
    wide, narrow = algorithmic_cells(reg)
    cells_per_region = wide + narrow
    # the inputs of every step exist before the clock starts (the workers get them pickled: ~2 MB per region)
    steps_in = [[(which, synth.make_region(REGION_LEN, COVERAGE, seed=2000 + k * cores + i)) for i in range(cores)]
                for k in range(args.steps_ref)]
    with mp.get_context("spawn").Pool(cores) as pool:
        for w in range(args.warmup_ref):
            pool.map(cpu_score_points_worker, [(which, 1000 + i) for i in range(cores)])
        t0 = time.perf_counter()
        for k in range(args.steps_ref):
            pool.map(cpu_score_points_worker, steps_in[k])
        dt = time.perf_counter() - t0
    value = cells_per_region * cores * args.steps_ref / dt / 1e9
    sample = "%d regions of 1 kb x 10x per step, one process per region on %d host cores" % (cores, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": dt / args.steps_ref * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "ScorePoints 1 kb region x 10x coverage, point_width 20, realign_width 300",
                       "regions_per_step": cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------------
def claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on the
    first communicator when NCCL_DEBUG asks for it): point fd 1 at stderr for the whole run and keep the real stdout
    for the line itself."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--min-seconds", type=float, default=2.0, help="each timed region repeats its block of --steps steps until this much time has been measured")
    ap.add_argument("--regions", type=int, default=44, help="1 kb regions per GPU per step (44 x 40 fill CTAs ~ 4 waves of 148 SMs x 3)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-consensus", action="store_true", help="skip the secondary consensus kb/s measurement")
    ap.add_argument("--e2e-path", default="direct", choices=["direct", "handles"],
                    help="direct: ps_score_points_direct from the host buffers (scores only, like PSAlign.ScorePoints); handles: native region objects per step")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[0] / configs[3] / configs[4] legs")
    ap.add_argument("--driver", default="threads", choices=["threads", "pipeline"],
                    help="e2e region: --drivers host threads with two contexts each (threads) or one host thread pipelining "
                         "--contexts contexts (pipeline)")
    ap.add_argument("--drivers", type=int, default=3, help="host threads of the e2e region (driver threads)")
    ap.add_argument("--contexts", type=int, default=4, help="library contexts (streams + staging areas) per GPU in the e2e pipeline")
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"],
                    help="fast: FP32 mutation scan + exact FP64 re-score of every candidate (decisions and accepted scores "
                         "bit-identical, other scores within 1e-4 relative); exact: everything FP64 bit-identical")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        # bounded: each step is `cores` regions (~5 s); cap the step count so the run ends in minutes
        args.steps_ref = max(1, min(args.steps, 6))
        args.warmup_ref = max(0, min(args.warmup, 1))
        run_reference_arm(args, rank, world, real_stdout)
        return

    # the library's host worker threads (marshalling, per-level logs, band planning): share the box's cores between
    # the ranks of this node instead of 8 per rank (PORESEQ_B200_THREADS is the library's own knob)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if "PORESEQ_B200_THREADS" not in os.environ:
        share = (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)) // max(local_world, 1)
        if args.driver == "threads":
            share -= max(1, args.drivers) - 1                 # the driver threads marshal too
        os.environ["PORESEQ_B200_THREADS"] = str(max(1, min(8, share)))

    import torch
    import torch.distributed as dist
    from poreseq_b200 import build, poreseqcpp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scoring path has no CPU fallback")
    build.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    # several contexts (stream + staging / device buffers each) on this rank's GPU: while the GPU works on step k
    # the host marshals and stages the next steps (ps_score_points_batch_begin / _end)
    n_ctx = 2 * max(1, args.drivers) if args.driver == "threads" else max(2, args.contexts)
    ctxs = [poreseqcpp.Context(local_rank) for _ in range(n_ctx)]
    for c in ctxs:
        c.set_precision(args.precision)
    ctx = ctxs[0]
    regions = make_regions(args.regions, seed0=1 + rank * args.regions)
    cells = [algorithmic_cells(r) for r in regions]
    wide_cells = sum(c[0] for c in cells)
    narrow_cells = sum(c[1] for c in cells)
    step_cells = wide_cells + narrow_cells
    h2d = sum(region_bytes(r) for r in regions)      # host bytes marshalled per step (the H2D payload is the same data re-laid out)
    phase = {}

    # the step's inputs as host buffers (one set of concatenated level arrays + model table per region);
    # every step marshals them into fresh native regions: 1 ps_region_create + 1 ps_region_add_events each
    packs = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regions]

    # --e2e-path direct (default): ps_score_points_direct -- PSAlign.ScorePoints straight from the host buffers, scores out
    # (the reference drops ScorePoints' realignment too, pyx:278-308); handles: ps_regions_create + ps_score_points_batch
    # + ps_regions_destroy per step (native region objects, realigned events copied back into them)
    def begin(c):
        if args.e2e_path == "direct":
            return poreseqcpp.PendingDirect(c, packs, "point_width")
        nrs = poreseqcpp.native_regions_from_packed(c, packs, "point_width")
        return poreseqcpp.score_points_batch_begin(c, nrs)

    def end(p, record):
        out = p.end()
        if record:
            for k, v in p.ctx.last_timing().items():
                phase[k] = phase.get(k, 0.0) + v
        if args.e2e_path != "direct":
            poreseqcpp.close_regions(p.regions)
        return out

    def run_steps(count, record):
        out = None
        inflight = []
        for k in range(count):
            inflight.append(begin(ctxs[k % len(ctxs)]))
            if len(inflight) == len(ctxs):
                out = end(inflight.pop(0), record)
        while inflight:
            out = end(inflight.pop(0), record)
        return out

    def run_steps_threads(count, record):
        """A few host threads, each pipelining whole steps (create, begin ... end, destroy) over its own two contexts:
        the host work of different steps runs on different cores (ctypes releases the GIL inside the library) and
        every thread keeps one batch queued behind the one it waits for, so the GPU never runs dry while the
        threads marshal (with one batch per thread the threads fall into lock-step: all wait, then all marshal)."""
        nxt = [0]
        lock = threading.Lock()
        last = [None]

        def take():
            with lock:
                k = nxt[0]
                nxt[0] += 1
            return k

        def drive(mine):
            pending = None                                   # (step index, batch in flight)
            turn = 0
            while True:
                k = take()
                cur = (k, begin(mine[turn])) if k < count else None
                turn ^= 1
                if pending is not None:
                    o = end(pending[1], record)
                    if pending[0] == count - 1:
                        last[0] = o
                pending = cur
                if cur is None:
                    return

        ths = [threading.Thread(target=drive, args=(ctxs[2 * t:2 * t + 2],)) for t in range(len(ctxs) // 2)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return last[0]

    def run_steps_serial(count, record):
        out = None
        for k in range(count):
            out = end(begin(ctxs[0]), record)
        return out

    # priming (not counted as warm-up): two steps on every context -- each context owns its staging and device buffers,
    # which grow on first use; then exactly --warmup warm-up steps
    n_prime = 2 * len(ctxs)
    out = run_steps(n_prime, False)
    n_warm = max(args.warmup, 0)
    if n_warm:
        out = run_steps(n_warm, False)
    # bytes the library copied for one step: staged level records, band centres, mutation tables, models ... up;
    # realigned events (ref_align, ref_like, ref_index) and scores down
    host_in = h2d
    h2d, d2h = ctxs[0].last_bytes()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # Both timed regions run BLOCKS of exactly K steps, each block bracketed by barrier + synchronize, until at least
    # --min-seconds have been timed (at most 25 blocks); the reported step time is the MEDIAN block, the spread is kept.
    # timed region 1 (`value`): K steps one after the other on one context; the kernel phases are timed with
    # CUDA events on the library's stream, nothing else runs on the GPU, the inputs of a phase are in HBM
    kernel_keys = ["centres", "forward", "backward", "backtrace", "join", "mutscore", "reduce"]
    dev_blocks, phase_blocks = [], []

    def blocks_wanted(first_seconds):
        """How many blocks every rank runs in all: enough for --min-seconds, the same number everywhere (max over ranks)."""
        n = int(min(25, max(1, -(-args.min_seconds // max(first_seconds, 1e-4)))))
        if world > 1:
            t = torch.tensor([n], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n = int(t.item())
        return n

    def device_block():
        phase.clear()
        barrier()
        t0 = time.perf_counter()
        run_steps_serial(args.steps, True)
        barrier()
        dev_blocks.append(sum(phase[k] for k in kernel_keys) / args.steps)
        phase_blocks.append(dict(phase))
        return time.perf_counter() - t0

    for _ in range(blocks_wanted(device_block()) - 1):
        device_block()
    # timed region 2 (`e2e`): K steps through the C-ABI from host buffers, several contexts in flight so that the
    # host staging and H2D of step k+1 overlap the kernels of step k; wall clock around barrier + synchronize
    wall_blocks = []
    launches0 = sum(c.launch_count() for c in ctxs)

    def e2e_block():
        barrier()
        t0 = time.perf_counter()
        (run_steps_threads if args.driver == "threads" else run_steps)(args.steps, False)
        barrier()
        wall_blocks.append((time.perf_counter() - t0) / args.steps * 1e3)
        return time.perf_counter() - t0

    for _ in range(blocks_wanted(e2e_block()) - 1):
        e2e_block()
    launches = (sum(c.launch_count() for c in ctxs) - launches0) // len(wall_blocks)
    sampler.stop()

    def median(v):
        v = sorted(v)
        return v[len(v) // 2]

    mid = sorted(range(len(dev_blocks)), key=lambda k: dev_blocks[k])[len(dev_blocks) // 2]
    phase = phase_blocks[mid]
    dev_ms = dev_blocks[mid]
    wall_ms = median(wall_blocks)
    times = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = times.tolist()
    total_cells = step_cells * world

    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    clocks = sampler.summary()
    props = torch.cuda.get_device_properties(local_rank)
    sms = props.multi_processor_count
    # dominant kernel: whichever phase took longest; its algorithmic cells / its own CUDA-event time
    # (the phase events sit on the library's stream right around the phase's launches)
    phase_ms = {k: phase[k] / args.steps for k in kernel_keys}
    dom = max(phase_ms, key=phase_ms.get)
    dom_kernel = {"forward": "k_fill", "mutscore": "k_mutscore_rows_f32" if args.precision == "fast" else "k_mutscore"}.get(dom, dom)
    dom_cells = {"forward": wide_cells, "mutscore": narrow_cells}.get(dom, 0.0)
    dom_s = phase_ms[dom] * 1e-3
    clock_mhz = clocks["sm_mhz"] or sm_max_mhz
    peak_ops = sms * 128 * sm_max_mhz * 1e6 / 1e12                  # T lane-ops/s at max clock
    achieved_ops = dom_cells * OPS_PER_CELL / dom_s / 1e12 if dom_s > 0 else 0.0
    # algorithmic bytes of the dominant kernel: the wide fill writes main + stay matrix and a step
    # byte per forward cell (16.5 B) and the main matrix per reverse cell (8 B): 12.25 B per cell over both directions;
    # the mutation kernel reads 8 B seed + 8 B reverse cell per band row
    dom_bytes = {"forward": wide_cells * 12.25, "mutscore": narrow_cells / 5.875 * 16.0}.get(dom, 0.0)
    # the fill is FP64 (exact): 45 FP64-pipe instructions per cell against the FP64 pipe of the SMs
    # (64 lanes per SM; profiles/r1_fp64_ubench.txt measures 0.43 warp-instructions per cycle per scheduler)
    fp64_peak = sms * 64 * sm_max_mhz * 1e6 / 1e12
    fp64_per_cell = {"forward": 45, "mutscore": 0 if args.precision == "fast" else 46}.get(dom, 0)
    fp64_achieved = dom_cells * fp64_per_cell / dom_s / 1e12 if dom_s > 0 else 0.0
    line = {
        "metric": METRIC, "value": total_cells / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": n_warm, "ms_per_step": wall_ms, "higher_is_better": True,
        "timing": {"blocks_of_steps": {"device": len(dev_blocks), "e2e": len(wall_blocks)}, "priming_steps": n_prime,
                   "device_ms_per_step": {"median": dev_ms, "min": min(dev_blocks), "max": max(dev_blocks)},
                   "e2e_ms_per_step": {"median": wall_ms, "min": min(wall_blocks), "max": max(wall_blocks)}},
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64" if args.precision == "fast" else "f64", "data": "synthetic",
        "config": {"workload": "ScorePoints (FindPointMutations+ScoreMutations) on 1 kb regions x 10x coverage, "
                               "point_width 20, realign_width 300 (BASELINE.json configs[1])",
                   "regions_per_gpu_per_step": args.regions, "events_per_region": 2 * COVERAGE,
                   "mutations_per_region": 8 * (REGION_LEN - 4), "cells_per_step_per_gpu": step_cells,
                   "l2": "band working set %.0f MB per step exceeds the 126 MB L2" % (wide_cells * 12.25 / 1e6),
                   "host_threads_per_rank": int(os.environ["PORESEQ_B200_THREADS"]),
                   "entry_point": "ps_score_points_direct_begin/_end (host arrays in, scores out; no region handles)" if args.e2e_path == "direct" else "ps_regions_create + ps_score_points_batch_begin/_end + ps_regions_destroy",
                   "pipelining": ("%d host threads x 2 contexts: host staging and H2D of the other steps overlap the kernels of step k" % (len(ctxs) // 2)
                                  if args.driver == "threads" else
                                  "%d contexts, one host thread: host staging and H2D of the next steps overlap the kernels of step k" % len(ctxs)),
                   "precision": ("fp32 mutation scan + exact fp64 re-score of all candidates > -tau; wide fills/backtrace fp64"
                                 if args.precision == "fast" else "fp64 exact (bit-identical to the reference)")},
        "e2e": {"value": total_cells / (wall_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "host_input_bytes_per_step": host_in},
        "gpu_launches": launches,
        "clocks": clocks,
        "phase_ms": phase_ms,
        "kernels": {"k_fill (forward+reverse wide fill, fp64)": {"ms": phase_ms["forward"], "cells": wide_cells,
                                                                "gcups": wide_cells / phase_ms["forward"] / 1e6},
                    dom_kernel if dom == "mutscore" else ("k_mutscore_rows_f32 + exact re-score" if args.precision == "fast" else "k_mutscore"):
                        {"ms": phase_ms["mutscore"], "cells": narrow_cells, "gcups": narrow_cells / phase_ms["mutscore"] / 1e6},
                    "k_join": {"ms": phase_ms["join"]}, "k_backtrace": {"ms": phase_ms["backtrace"]}},
        "roofline": {"bound": "fp32-issue", "kernel": dom_kernel,
                     "achieved": achieved_ops, "peak": peak_ops, "unit": "Tlane-op/s", "frac": achieved_ops / peak_ops,
                     "ops_per_cell": OPS_PER_CELL, "clock_mhz_under_load": clock_mhz, "peak_source": peak_src,
                     "traffic": recorded_traffic(dom_kernel, args.regions),
                     "traffic_source": "recorded: dram__bytes_read+write per launch from the committed ncu --set full capture of this batch size (profiles/r2_traffic.json), not sampled in this run",
                     "note": "SURVEY.md 8d definition (24 FP32 lane-ops per cell); the dominant kernel computes in FP64, see roofline_fp64"},
        "roofline_fp64": {"bound": "fp64-pipe", "kernel": dom_kernel, "achieved": fp64_achieved, "peak": fp64_peak,
                          "unit": "T fp64-op/s", "frac": fp64_achieved / fp64_peak if fp64_peak else 0.0,
                          "ops_per_cell": fp64_per_cell},
        "roofline_hbm": {"bound": "hbm", "kernel": dom_kernel, "achieved": dom_bytes / dom_s / 1e9 if dom_s > 0 else 0.0,
                         "peak": hbm_peak, "unit": "GB/s", "frac": dom_bytes / dom_s / 1e9 / hbm_peak if dom_s > 0 else 0.0,
                         "peak_source": peak_src, "traffic": recorded_traffic(dom_kernel, args.regions)},
    }
    for c in ctxs[1:]:
        c.close()
    extras_errors = {}

    def attempt(name, fn):
        try:
            return fn()
        except Exception as ex:                                   # an extra leg must not cost the headline line
            extras_errors[name] = "%s: %s" % (type(ex).__name__, ex)
            return None

    if rank == 0:
        line["fast_mode_parity"] = attempt("fast_mode_parity", lambda: fast_mode_error(ctxs[0], local_rank))
    ctxs[0].close()
    if rank == 0 and not args.no_extras:
        line["score_events"] = attempt("score_events", lambda: leg_score_events(local_rank, sms, sm_max_mhz, cpu=not args.no_cpu_baseline))
    if not args.no_consensus:
        from poreseq_b200 import drivers
        # consensus kb/s (BASELINE.json's second headline): the Mutate.py loop below the C-ABI (ps_consensus_batch),
        # CONS_REGIONS regions of 1 kb x 10x per GPU, CONS_IN_FLIGHT of them side by side on every rank
        def cons_throughput():
            cctx = poreseqcpp.Context(local_rank)
            cctx.set_precision("fast")
            try:
                warm = [synth.make_region(1000, 10, seed=9000 + k, draft_error=0.10) for k in range(CONS_REGIONS)]
                drivers.consensus_native(warm, ctx=cctx, in_flight=CONS_IN_FLIGHT)          # untimed: every group and lane allocates its buffers
                regs = [synth.make_region(1000, 10, seed=500 + CONS_REGIONS * rank + k, draft_error=0.10) for k in range(CONS_REGIONS)]
                # the inputs as host buffers (what an event-pack file maps to): marshalling into native regions is timed,
                # turning Python event objects into flat arrays is not
                packed = [poreseqcpp.PackedRegion(r.sequence, r.events, r.params) for r in regs]
                times = []
                for _ in range(3):
                    barrier()
                    t0 = time.perf_counter()
                    res = drivers.consensus_native(packed, ctx=cctx, in_flight=CONS_IN_FLIGHT)
                    barrier()
                    times.append(time.perf_counter() - t0)
                acc = sum(poreseqcpp.swalign(r[0], g.truth)[0] for r, g in zip(res, regs)) / len(regs)
                return times, acc
            finally:
                cctx.close()
        got = attempt("consensus_throughput", cons_throughput)
        if got is not None:
            times, acc = got
            tt = torch.tensor(times + [acc], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt[0:3], op=dist.ReduceOp.MAX)
                dist.all_reduce(tt[3:4], op=dist.ReduceOp.SUM)
            tl = sorted(tt[0:3].tolist())
            line["consensus"] = {"throughput": {"value": float(CONS_REGIONS) * world / tl[1], "unit": "kb/s", "n_gpus": world,
                                                "seconds": {"median": tl[1], "min": tl[0], "max": tl[2]},
                                                "mean_accuracy_pct": tt[3].item() / world,
                                                "config": "consensus loop (Mutate.py policy below the C-ABI, ps_consensus_batch) on %d regions of 1 kb x 10x "
                                                          "per GPU (drafts with 10%% errors), %d regions in flight per GPU, fast precision; "
                                                          "3 passes, max over ranks each" % (CONS_REGIONS, CONS_IN_FLIGHT)}}
        else:
            line["consensus"] = {}
        if rank == 0 and world == 1:
            def cons_single(length, coverage, seed, gold):
                cctx = poreseqcpp.Context(local_rank)
                cctx.set_precision("fast")
                try:
                    drivers.consensus_native([synth.make_region(300, 5, seed=99, draft_error=0.05)], ctx=cctx, in_flight=1)
                    reg = synth.make_region(length, coverage, seed=seed, draft_error=0.10)
                    t0 = time.perf_counter()
                    drivers.consensus_native([reg], ctx=cctx, in_flight=1)
                    cold = time.perf_counter() - t0                 # first job of this size on the context: its buffers grow
                    t0 = time.perf_counter()
                    seq, acc, stages = drivers.consensus_native([reg], ctx=cctx, in_flight=1, refseqs=[reg.truth])[0]
                    dt = time.perf_counter() - t0
                    out = {"value": length / 1000.0 / dt, "unit": "kb/s", "seconds": dt, "seconds_first_call_on_a_fresh_context": cold,
                           "accuracy_pct": acc,
                           "draft_accuracy_pct": poreseqcpp.swalign(reg.sequence, reg.truth)[0],
                           "config": "consensus loop (ps_consensus) on a %d b region at %dx coverage (draft with 10%% errors), fast precision" % (length, coverage)}
                    z = golden(gold) if gold else None
                    if z is not None:
                        out["parity_vs_reference"] = {"identical_stages": bool([st[1] for st in stages] == z["stage_seqs"].tolist()),
                                                      "golden": "tests/golden/f_%s.npz" % gold}
                        out["cpu_baseline"] = {"value": length / 1000.0 / float(z["seconds"]), "unit": "kb/s", "seconds": float(z["seconds"]),
                                               "cores": 1, "kind": "reference",
                                               "sample": "recorded when the golden was generated (not timed in this run): the same loop, same region"}
                    return out, seq
                finally:
                    cctx.close()
            big = attempt("consensus_10kb", lambda: cons_single(10000, 30, 7, "c2"))
            small = attempt("consensus_1kb", lambda: cons_single(1000, 10, 7, None))
            if big is not None:
                line["consensus"]["configs[2] 10 kb x 30x"] = big[0]
            if small is not None:
                line["consensus"]["1 kb x 10x"] = small[0]
                if not args.no_cpu_baseline:
                    from oracle import binding
                    if binding.available("ref"):
                        def cpu_small():
                            cpu, seq_cpu = consensus_cpu(1000, 10, seed=7)
                            cpu["identical_consensus"] = bool(seq_cpu == small[1])
                            return cpu
                        line["consensus"]["1 kb x 10x"]["cpu_baseline"] = attempt("consensus_1kb_cpu", cpu_small)
    if not args.no_extras:
        v = attempt("variant", lambda: leg_variant(rank, world, local_rank, dist, torch))
        pl = attempt("polish", lambda: leg_polish(rank, world, local_rank, dist, torch))
        if rank == 0:
            line["variant_configs3"] = v
            line["polish_configs4"] = pl
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = attempt("cpu_baseline", run_cpu_baseline)
        if extras_errors:
            line["extras_errors"] = extras_errors
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
